#!/bin/bash
# 24-bit count rows with escape in the scan kernel: GPU suite, rows24 vs rows32 on the configs[2] shape, then that bench with parity
TAG=${1:-r02_u}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 200 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
timeout 90 python tools/row24_timing.py --out $OUT/row24_timing.json > /dev/null 2> $OUT/row24_timing.err; grep -h "^rows" $OUT/row24_timing.err
timeout 150 python bench.py --config 3 --no-cpu-baseline --no-scaling-base --steps 3 --warmup 2 > $OUT/bench_cfg3.json 2> $OUT/bench_cfg3.err
python - $OUT/bench_cfg3.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print("cfg3 value %.1f M reads/s ms %.2f | e2e %.1f M ms %.2f | parity %s |" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], (d.get("parity") or {}).get("identical")),
          " ".join("%s %.2f (%.3f)" % (k.replace("rcx_","").replace("_kernel",""), v["ms"], v["frac_of_random_request_peak"]) for k,v in d["kernels"].items()))
except Exception as e:
    print("cfg3 no line:", e)
P
tail -2 $OUT/bench_cfg3.err
