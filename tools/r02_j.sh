#!/bin/bash
# validation call: GPU suite, extended GPU-vs-oracle campaigns (IUPAC letters, tiny chunks, k-mer counting), memcheck of the new kernels
TAG=${1:-r02_j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
timeout 300 python tools/fuzz_gpu_vs_oracle.py --cases 120 --seed 11 --iupac --chunks > $OUT/fuzz_iupac_chunks.log 2>&1; echo "fuzz iupac+chunks rc=$?"; tail -2 $OUT/fuzz_iupac_chunks.log
timeout 300 python tools/fuzz_gpu_vs_oracle.py --cases 120 --seed 12 --count --chunks > $OUT/fuzz_count_chunks.log 2>&1; echo "fuzz count+chunks rc=$?"; tail -2 $OUT/fuzz_count_chunks.log
timeout 300 python tools/fuzz_gpu_vs_oracle.py --cases 60 --seed 13 --wild-quals --iupac > $OUT/fuzz_quals_iupac.log 2>&1; echo "fuzz quals+iupac rc=$?"; tail -2 $OUT/fuzz_quals_iupac.log
timeout 300 python tools/fuzz_gpu_vs_oracle.py --cases 20 --seed 14 --long --iupac > $OUT/fuzz_long_iupac.log 2>&1; echo "fuzz long+iupac rc=$?"; tail -2 $OUT/fuzz_long_iupac.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_count.py -m gpu -q -x \
  -k "s100 or k31deep or many_accepted or spill or counts_of_2 or chunked or counts_equal or grows" > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 $OUT/memcheck.log
