# Compare variant libraries in one call.  Build them first, e.g.:
#   cd rcorrector_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DDFS_MINBLOCKS=7 -c rcx_dfs.cu -o ../../build/rcx_dfs_mb7.o
#   nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/librcx_mb7.so <the other .o files> ../../build/rcx_dfs_mb7.o -lcudart
# (r01_y: 7 / 8 search blocks per SM measured equal / slower than 6)
OUT=gpurun_out/r01_y; mkdir -p $OUT
A="--no-cpu-baseline --steps 3 --warmup 3"
Q="--reads 4000000 --len 150 --k 31 --err 0.04 --maxcork 8 --tx-bases 50000000 --no-cpu-baseline --steps 3 --warmup 3"
for mb in 7 8; do
  RCX_LIB=$PWD/build/librcx_mb$mb.so timeout 200 python bench.py $A > $OUT/bench_100_mb$mb.json 2> $OUT/bench_100_mb$mb.err
  RCX_LIB=$PWD/build/librcx_mb$mb.so timeout 200 python bench.py $Q > $OUT/bench_k31_mb$mb.json 2> $OUT/bench_k31_mb$mb.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.1f ms/step"%d["ms_per_step"], "e2e %.1f"%d["e2e"]["ms_per_step"], {k:round(v["ms"],1) for k,v in d["kernels"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
