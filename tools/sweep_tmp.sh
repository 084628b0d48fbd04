export VASR_TC_ALT=1
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -k "golden and en15x5 and rand and f16x3" 2>&1 | tail -2
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "ALT parity failed or hung"; exit 0; fi
timeout 120 python tools/prof_encoder.py 256 5 2>&1 | tail -1
VASR_TC_PROF=1 timeout 120 python tools/prof_encoder.py 256 2 2>&1 | grep -E "TCSEG.*items=(7680|11520)" | tail -2
unset VASR_TC_ALT
timeout 120 python tools/prof_encoder.py 256 5 2>&1 | tail -1
