#!/bin/bash
# Round-2 second call: GPU suite + campaigns after the 16-bit edit lists, effect of heavy-first / paired re-probes, ncu captures.
TAG=${1:-r02_b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
timeout 300 python tools/fuzz_gpu_vs_oracle.py --cases 150 --seed 1 > $OUT/fuzz_gpu.log 2>&1; echo "gpu fuzz rc=$?"; tail -2 $OUT/fuzz_gpu.log
timeout 240 python tools/fuzz_gpu_vs_oracle.py --cases 80 --seed 3 --wild-quals > $OUT/fuzz_gpu_quals.log 2>&1; echo "gpu fuzz quals rc=$?"; tail -2 $OUT/fuzz_gpu_quals.log
B="python bench.py --no-cpu-baseline --no-parity --no-scaling-base --steps 5 --warmup 3"
for hw in 46 0 30 70; do
  RCX_HEAVY_WINDOWS=$hw timeout 200 $B > $OUT/bench_hw$hw.json 2> $OUT/bench_hw$hw.err; echo "hw=$hw rc=$?"
  python - $OUT/bench_hw$hw.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print(" ms/step %.2f e2e %.2f | " % (d["ms_per_step"], d["e2e"]["ms_per_step"]) + " ".join("%s %.2f" % (k.replace("rcx_","").replace("_kernel",""), v["ms"]) for k,v in d["kernels"].items()))
P
done
for cfg in 3 5; do
  timeout 300 $B --config $cfg > $OUT/bench_cfg$cfg.json 2> $OUT/bench_cfg$cfg.err; echo "cfg=$cfg rc=$?"
  python - $OUT/bench_cfg$cfg.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print(" ms/step %.2f e2e %.2f | " % (d["ms_per_step"], d["e2e"]["ms_per_step"]) + " ".join("%s %.2f" % (k.replace("rcx_","").replace("_kernel",""), v["ms"]) for k,v in d["kernels"].items()))
P
done
# ncu: launch list + full capture of the three big kernels of one device step (configs[1])
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rcx_(scan_thread|dfs|finalize_thread)" -s 4 -c 4 -f -o $OUT/full \
  python bench.py --no-cpu-baseline --no-parity --no-scaling-base --steps 1 --warmup 1 > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT | tail -20
