#!/bin/bash
# Round-2 opening call (1 GPU): GPU suite, the GPU-vs-oracle campaigns, both bench arms (reference first, so our arm's parity
# check meets the reference's own bytes), the configs[3] table sweep.   usage (GPU box): bash tools/r02_open.sh <tag>
TAG=${1:-r02_a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt; free -g | head -2 >> $OUT/nproc.txt; df -h /dev/shm | tail -1 >> $OUT/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cut -c1-200 $OUT/bench_ref.json
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-400 $OUT/bench.json; tail -5 $OUT/bench.err
timeout 300 python tools/fuzz_gpu_vs_oracle.py --cases 120 --seed 1 > $OUT/fuzz_gpu.log 2>&1; echo "gpu fuzz rc=$?"; tail -2 $OUT/fuzz_gpu.log
timeout 240 python tools/fuzz_gpu_vs_oracle.py --cases 60 --seed 3 --wild-quals > $OUT/fuzz_gpu_quals.log 2>&1; echo "gpu fuzz quals rc=$?"; tail -2 $OUT/fuzz_gpu_quals.log
timeout 240 python tools/fuzz_gpu_vs_oracle.py --cases 30 --seed 2 --long > $OUT/fuzz_gpu_long.log 2>&1; echo "gpu fuzz long rc=$?"; tail -2 $OUT/fuzz_gpu_long.log
timeout 600 python tools/cfg4_sweep.py --kmers 64,256,1000 --out $OUT/cfg4_sweep.json > /dev/null 2> $OUT/cfg4.err; echo "cfg4 rc=$?"; tail -4 $OUT/cfg4.err | cut -c1-600
