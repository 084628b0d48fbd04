#!/bin/bash
# Round-end evidence on one B200: GPU test suite, bench (both arms), ncu launch list of the bench command, ncu counters of the
# encoder kernels, --set full capture of the dominant launch.  Usage (from the repo root): bash tools/evidence_run.sh TAG
TAG=${1:-rX}
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__issue_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg"
timeout 600 python -m pytest tests -m gpu -q --timeout 150 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_gpu.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 400 ncu --metrics $M --clock-control none -k regex:"segment|subblock" --csv --log-file gpurun_out/${TAG}_encoder_metrics.csv \
    python tools/prof_encoder.py 256 1 > gpurun_out/${TAG}_ncu_metrics.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:segment_pair -s 2 -c 1 -f -o gpurun_out/${TAG}_pair512 \
    python tools/prof_encoder.py 256 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log; head -c 600 gpurun_out/${TAG}_bench_c3.json; echo; head -c 300 gpurun_out/${TAG}_bench_reference.json; echo
