#!/bin/bash
TAG=${1:-r02_g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="python bench.py --no-cpu-baseline --no-parity --no-scaling-base --steps 5 --warmup 3"
show() { python - "$1" "$2" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(sys.argv[2], "ms/step %.2f e2e %.2f | " % (d["ms_per_step"], d["e2e"]["ms_per_step"]) + " ".join("%s %.2f" % (k.replace("rcx_","").replace("_kernel",""), v["ms"]) for k,v in d["kernels"].items()), d["e2e"].get("chunk_sums_ms"))
except Exception as e:
    print(sys.argv[2], "unreadable", e)
P
}
for cb in 104e6 156e6 208e6 312e6; do
  RCX_CHUNK_BASES=$cb timeout 200 $B --config 2 > $OUT/chunk_c2_$cb.json 2> $OUT/chunk_c2_$cb.err; show $OUT/chunk_c2_$cb.json "cfg2 chunk $cb"
done
for cb in 156e6 208e6 312e6 450e6; do
  RCX_CHUNK_BASES=$cb timeout 300 $B --config 3 > $OUT/chunk_c3_$cb.json 2> $OUT/chunk_c3_$cb.err; show $OUT/chunk_c3_$cb.json "cfg3 chunk $cb"
done
# the configs[2] shape on one GPU, both arms (the reference first: our arm's parity check then meets the reference's own bytes)
timeout 900 python bench.py --impl reference --config 3 --steps 3 --warmup 1 > $OUT/bench_cfg3_ref.json 2> $OUT/bench_cfg3_ref.err; echo "cfg3 ref rc=$?"; cut -c1-220 $OUT/bench_cfg3_ref.json
timeout 600 python bench.py --config 3 --no-cpu-baseline --steps 5 --warmup 3 > $OUT/bench_cfg3.json 2> $OUT/bench_cfg3.err; echo "cfg3 rc=$?"; show $OUT/bench_cfg3.json "cfg3"
python -c "
import json;d=json.loads(open('$OUT/bench_cfg3.json').read());print(d.get('parity'));print(d['e2e']['value']/1e6, d['value']/1e6)"
# -countkmers against -c on the same files
D=/dev/shm/rcx_ck; mkdir -p $D
tools/bin/rcx_synth -seed 3 -k 23 -ntx 7000 -txmin 300 -txmax 2500 -polya 0.3 -sigma 2.6 -reads 4000000 -len 100 -err 0.01 -nrate 0.002 -short 0.02 -paired -fq1 $D/r_1.fq -fq2 $D/r_2.fq 2>/dev/null
s=$(date +%s.%N); tools/bin/rcx_kmercount -k 23 -o $D/exact.dump $D/r_1.fq $D/r_2.fq; e=$(date +%s.%N); echo "rcx_kmercount (CPU exact counter, Jellyfish stand-in): $(echo "$e - $s" | bc) s"
for mode in c countkmers; do
  arg="-countkmers"; [ $mode == c ] && arg="-c $D/exact.dump"
  s=$(date +%s.%N); RCX_TIMING=1 rcorrector_b200/bin/rcorrector -p $D/r_1.fq $D/r_2.fq $arg -k 23 -od $D/out_$mode -t 16 2> $OUT/countkmers_$mode.err; e=$(date +%s.%N)
  echo "mode $mode: $(echo "$e - $s" | bc) s wall"; grep -h "Stored\|Weak\|Processed\|counted\|table built\|dump parsed" $OUT/countkmers_$mode.err | tr '\n' ';'; echo; md5sum $D/out_$mode/*.cor.fq | cut -c1-32 | tr '\n' ' '; echo
done
# ncu of the closing state: launch list of two steps, full capture of the three big kernels
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 24 --csv --log-file $OUT/launches.csv python bench.py --no-cpu-baseline --no-parity --no-scaling-base --steps 2 --warmup 2 > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rcx_(scan_thread|dfs|finalize_thread)" -s 4 -c 4 -f -o $OUT/full \
  python bench.py --no-cpu-baseline --no-parity --no-scaling-base --steps 1 --warmup 1 > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
