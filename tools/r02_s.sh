#!/bin/bash
# 24-bit count rows for the 64-read tiles: GPU suite, then the configs[2] shape with the rows on / off / scan only, then the default bench
TAG=${1:-r02_s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 220 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
show() { python - "$1" "$2" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(sys.argv[2], "value %.1f M reads/s ms %.2f | e2e %.1f M ms %.2f | parity %s |" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], (d.get("parity") or {}).get("identical")),
          " ".join("%s %.2f (%.3f)" % (k.replace("rcx_","").replace("_kernel",""), v["ms"], v["frac_of_random_request_peak"]) for k,v in d["kernels"].items()), "unsupported", d.get("stats",{}).get("unsupported"))
except Exception as e:
    print(sys.argv[2], "no line:", e)
P
}
RCX_VERBOSE=1 timeout 260 python bench.py --config 3 --no-cpu-baseline --no-scaling-base --steps 3 --warmup 2 > $OUT/cfg3_row24.json 2> $OUT/cfg3_row24.err; show $OUT/cfg3_row24.json cfg3/rows24; grep -h "\[rcx\]" $OUT/cfg3_row24.err | head -2
RCX_ROW24=0 timeout 120 python bench.py --config 3 --no-cpu-baseline --no-scaling-base --no-parity --steps 3 --warmup 2 > $OUT/cfg3_row32.json 2> $OUT/cfg3_row32.err; show $OUT/cfg3_row32.json cfg3/rows32
RCX_ROW24=1 timeout 120 python bench.py --config 3 --no-cpu-baseline --no-scaling-base --no-parity --steps 3 --warmup 2 > $OUT/cfg3_scan24.json 2> $OUT/cfg3_scan24.err; show $OUT/cfg3_scan24.json cfg3/scan24
timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 --no-scaling-base > $OUT/bench.json 2> $OUT/bench.err; show $OUT/bench.json cfg2
tail -2 $OUT/bench.err
