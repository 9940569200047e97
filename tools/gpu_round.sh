#!/bin/bash
# One gpurun call: GPU tests, bench (both arms), ncu launch list, ncu --set full of the scan and dfs kernels.
# usage (from the repo root on the GPU box): bash tools/gpu_round.sh <tag> [skip-tests]
TAG=${1:-rXX}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
  echo "pytest rc=$?" >> $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
fi
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?"; cat $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
echo "ref rc=$?"; cat $OUT/bench_ref.json
# launch list (cold-cache, serialised; shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
# one full capture of the lookup (scan) kernel and the dfs kernel (first device step)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rcx_scan_thread_kernel|rcx_dfs_kernel|rcx_finalize_thread_kernel' -c 3 \
  -o $OUT/full_pipeline -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT
