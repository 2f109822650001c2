#!/bin/bash
# Run on the GPU box (gpurun): final evidence for profiles/.  Two calls, one ncu invocation each, each after the same
# command exited 0 without ncu:   tools/capture_profiles.sh a   (tests, both bench arms, launch list)
#                                 tools/capture_profiles.sh b   (full-set capture of the hot kernels)
set -x
mkdir -p gpurun_out/final
if [ "$1" != "b" ]; then
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/final/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final/pytest_gpu.log
python bench.py --impl reference > gpurun_out/final/bench_reference.json 2> gpurun_out/final/bench_reference.err
python bench.py > gpurun_out/final/bench_b200.json 2> gpurun_out/final/bench_b200.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/final/nvidia_smi.csv
# launch list of a short bench run (cold-cache, serialised: compare SHARES)
python bench.py --steps 1 --warmup 1 --n-seq 1024 --scan-mbases 200 --no-cpu-baseline > gpurun_out/final/plain_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final/launches.csv \
    python bench.py --steps 1 --warmup 1 --n-seq 1024 --scan-mbases 200 --no-cpu-baseline > gpurun_out/final/ncu_launches.log 2>&1
else
# full-set capture of every hot kernel on a small fixed workload
python tools/prof_run.py 256 100 > gpurun_out/final/prof_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"nw_wave|mt_|scan_cov|hits_|dist_flags|dist_epilogue|cv_|prefilter|bd_|lk_" -c 120 -o gpurun_out/final/full \
    python tools/prof_run.py 256 100 > gpurun_out/final/ncu_full.log 2>&1
# gpurun brings back at most 64 MiB: keep the raw-page CSV, drop the report when it is too big
ncu -i gpurun_out/final/full.ncu-rep --page raw --csv > gpurun_out/final/full_raw.csv 2>/dev/null
if [ $(stat -c %s gpurun_out/final/full.ncu-rep) -gt 40000000 ]; then rm -f gpurun_out/final/full.ncu-rep; fi
fi
echo done
