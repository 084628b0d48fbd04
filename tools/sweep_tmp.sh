timeout 150 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or block0 or host_route_equals" 2>&1 | tail -3
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "LAT parity failed or hung"; fi
VASR_TC_LAT=0 timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -k "golden and f16x3" 2>&1 | tail -2
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('LAT on :', d['ms_per_step'], d['latency_b1'])"
VASR_TC_LAT=0 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('LAT off:', d['ms_per_step'], d['latency_b1'])"
