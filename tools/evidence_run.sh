#!/bin/bash
# Round-end evidence on one B200: GPU test suite, bench (both arms), ncu launch list of the bench command, DRAM traffic
# of the encoder kernels.  Usage (from the repo root): bash tools/evidence_run.sh TAG   -> gpurun_out/TAG_*
TAG=${1:-rX}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_gpu.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"segment_kernel|subblock_kernel" --csv --log-file gpurun_out/${TAG}_dram.csv \
    python tools/prof_encoder.py 256 2 > gpurun_out/${TAG}_ncu_dram.log 2>&1
python tools/traffic_from_ncu.py gpurun_out/${TAG}_dram.csv 2 gpurun_out/${TAG}_traffic.json > /dev/null 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log; head -c 600 gpurun_out/${TAG}_bench_default.json; echo; head -c 300 gpurun_out/${TAG}_bench_reference.json; echo
tail -2 gpurun_out/${TAG}_ncu_dram.log; head -c 400 gpurun_out/${TAG}_traffic.json
echo "--- two-group variant (VASR_TC_ALT=1), role counters"
VASR_TC_ALT=1 VASR_TC_PROF=1 timeout 120 python tools/prof_encoder.py 256 2 2>&1 | grep -E "TCSEG.*items=(7680|11520)|encoder ms" | tee gpurun_out/${TAG}_alt_roles.log
