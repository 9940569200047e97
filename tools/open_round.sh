#!/bin/bash
# Opening gpurun call of a session: GPU test suite, the GPU-vs-oracle campaign (short and long reads), both bench arms,
# the other read shapes.  About 6-7 minutes of box time.   usage (repo root, GPU box): bash tools/open_round.sh <tag>
TAG=${1:-rXX}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 600 python tools/fuzz_gpu_vs_oracle.py --cases 200 --seed 1 > $OUT/fuzz_gpu.log 2>&1; echo "gpu fuzz rc=$?"; tail -3 $OUT/fuzz_gpu.log
timeout 600 python tools/fuzz_gpu_vs_oracle.py --cases 100 --seed 3 --wild-quals > $OUT/fuzz_gpu_quals.log 2>&1; echo "gpu fuzz quals rc=$?"; tail -3 $OUT/fuzz_gpu_quals.log
timeout 600 python tools/fuzz_gpu_vs_oracle.py --cases 40 --seed 2 --long > $OUT/fuzz_gpu_long.log 2>&1; echo "gpu fuzz long rc=$?"; tail -3 $OUT/fuzz_gpu_long.log
timeout 300 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-260 $OUT/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cut -c1-200 $OUT/bench_ref.json
bash tools/shapes_round.sh $TAG
