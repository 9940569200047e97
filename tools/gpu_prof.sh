#!/bin/bash
# ncu --set full of one launch of each pipeline kernel on a reduced batch. usage: bash tools/gpu_prof.sh <tag> [kernel-regex] [reads]
TAG=${1:-prof}; RE=${2:-rcx_scan_kernel|rcx_prep_kernel|rcx_dfs_kernel|rcx_finalize_kernel}; READS=${3:-4000000}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$RE" -c 6 -o $OUT/full -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --reads $READS > $OUT/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/ncu.log; ls -la $OUT
