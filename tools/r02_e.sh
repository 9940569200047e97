#!/bin/bash
TAG=${1:-r02_e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python tools/debug_compact.py > $OUT/debug_compact.log 2>&1; grep -c " ok" $OUT/debug_compact.log; grep -v " ok" $OUT/debug_compact.log | head -5
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; tail -6 $OUT/pytest_gpu.log
B="python bench.py --no-cpu-baseline --no-parity --no-scaling-base --steps 5 --warmup 3"
for cfg in 2 3 5; do
  timeout 300 $B --config $cfg > $OUT/bench_cfg$cfg.json 2> $OUT/bench_cfg$cfg.err; echo "cfg=$cfg rc=$?"
  python - $OUT/bench_cfg$cfg.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print(" ms/step %.2f e2e %.2f | " % (d["ms_per_step"], d["e2e"]["ms_per_step"]) + " ".join("%s %.2f" % (k.replace("rcx_","").replace("_kernel",""), v["ms"]) for k,v in d["kernels"].items()))
P
done
timeout 600 python tools/cfg4_sweep.py --kmers 64,256,1000 --out $OUT/cfg4_sweep.json > /dev/null 2> $OUT/cfg4.err; echo "cfg4 rc=$?"; tail -3 $OUT/cfg4.err | cut -c1-400
# DRAM bytes of the lookup kernel per table size (one launch each)
for mk in 64 1000; do
  timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:rcx_scan_thread -c 1 --csv --log-file $OUT/cfg4_ncu_$mk.csv \
    python tools/cfg4_sweep.py --kmers $mk --steps 1 --warmup 0 > /dev/null 2> $OUT/cfg4_ncu_$mk.err; echo "ncu cfg4 $mk rc=$?"; tail -4 $OUT/cfg4_ncu_$mk.csv | cut -c1-300
done
