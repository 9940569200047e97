#!/bin/bash
# closing 1-GPU call: GPU suite, both bench arms exactly as the driver runs them
TAG=${1:-r02_h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; tail -5 $OUT/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
s=$(date +%s); timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$? $(( $(date +%s) - s )) s"; cut -c1-160 $OUT/bench_ref.json
s=$(date +%s); timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$? $(( $(date +%s) - s )) s"
python - $OUT/bench.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print(" value %.1f M reads/s ms %.2f | e2e %.1f M ms %.2f | parity %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d.get("parity")))
print(" kernels " + " ".join("%s %.2f (%.2f of req peak)" % (k.replace("rcx_","").replace("_kernel",""), v["ms"], v["frac_of_random_request_peak"]) for k,v in d["kernels"].items()))
print(" roofline", {k:d["roofline"][k] for k in ("kernel","achieved","frac","traffic","frac_of_random_request_peak")}, "cpu", d.get("cpu_baseline",{}).get("value"))
print(" cfg3 base", {k:(v if not isinstance(v,dict) else "...") for k,v in d.get("scaling_base_cfg3",{}).items()})
P
tail -3 $OUT/bench.err
timeout 200 python bench.py --config 5 --no-cpu-baseline --no-parity --no-scaling-base --steps 5 --warmup 3 > $OUT/bench_cfg5.json 2>/dev/null; python -c "
import json;d=json.loads(open('$OUT/bench_cfg5.json').read());print('cfg5', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), {k:round(v['ms'],2) for k,v in d['kernels'].items()})"
