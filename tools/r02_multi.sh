#!/bin/bash
# Multi-GPU call (gpurun --gpus N): the scaling bench on the configs[2] shape, the drop-in binary with -gpus N against the
# reference's bytes, the -gpus 2 GPU test; N = 8 adds the replicated configs[3] sweep.   usage: bash tools/r02_multi.sh N <tag> [steps]
N=${1:-2}
TAG=${2:-r02_n$N}
STEPS=${3:-5}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1; nproc >> $OUT/topo.txt; free -g | head -2 >> $OUT/topo.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $STEPS --warmup 3 \
  > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench N=$N rc=$?"
python - $OUT/bench_n$N.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(" value %.1f M reads/s  ms/step %.2f | e2e %.1f M reads/s ms %.2f | table %s | parity %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["table"], d.get("parity")))
    print(" kernels " + " ".join("%s %.2f" % (k.replace("rcx_","").replace("_kernel",""), v["ms"]) for k,v in d["kernels"].items()), d["e2e"]["chunk_sums_ms"], d["e2e"]["link_gbs_per_gpu"])
except Exception as e:
    print("unreadable", e)
P
tail -3 $OUT/bench_n$N.err
timeout 300 python -m pytest tests/test_gpu_golden.py -q -k "gpus_2" > $OUT/pytest_gpus2.log 2>&1; echo "pytest -gpus 2 rc=$?"; tail -2 $OUT/pytest_gpus2.log
RCX_TIMING=1 timeout 600 python tools/binary_round.py ${BR_PAIRS:-3000000} ${BR_NTX:-20000} 150 1 -- -gpus 1 -- -gpus $N > $OUT/binary_gpus$N.json 2> $OUT/binary_gpus$N.err; echo "binary rc=$?"
python - $OUT/binary_gpus$N.json <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(" identical", d["identical_output"], "| ref correct %.1fs whole %.1fs" % (d["reference"]["correct_s"], d["reference"]["exit"]))
    for n in ("ours","ours_1"):
        m=d[n]; print(" ", m["args"], "table %.2fs rate %.2fs correct %.2fs whole %.2fs -> %.1f M reads/s in phase" % (m["load_table_s"], m["error_rate_s"], m["correct_s"], m["exit"], m["reads_per_s_correct_phase"]/1e6))
except Exception as e:
    print("unreadable", e)
P
if [ "$N" == "8" ]; then
  for g in 0 1 2 3 4 5 6 7; do
    CUDA_VISIBLE_DEVICES=$g timeout 500 python tools/cfg4_sweep.py --kmers 1000 --steps 2 --out $OUT/cfg4_gpu$g.json > /dev/null 2> $OUT/cfg4_gpu$g.err &
  done
  wait
  echo "cfg4 x8 done"; grep -h "reads_per_s" $OUT/cfg4_gpu0.json | head -3
fi
