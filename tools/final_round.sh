#!/bin/bash
# Short closing call: bench line (with the CPU baseline leg), ncu launch list, ncu --set full of scan / dfs / finalize.
TAG=${1:-rXX}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
timeout 170 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?"; cut -c1-300 $OUT/bench.json
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'rcx_scan_thread_kernel|rcx_dfs_kernel|rcx_finalize_thread_kernel' -c 3 \
  -o $OUT/full_pipeline -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT
