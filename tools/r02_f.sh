#!/bin/bash
TAG=${1:-r02_f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; tail -6 $OUT/pytest_gpu.log
RCX_TIMING=1 timeout 600 python tools/binary_round.py 4000000 7000 100 0 > $OUT/binary_single_4M.json 2> $OUT/binary_single_4M.err; echo "binary single rc=$?"
RCX_TIMING=1 timeout 600 python tools/binary_round.py 4000000 7000 100 1 > $OUT/binary_paired_8M.json 2> $OUT/binary_paired_8M.err; echo "binary paired rc=$?"
python - $OUT/binary_single_4M.json $OUT/binary_paired_8M.json <<'P'
import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read())
        m=d["ours"]; r=d["reference"]
        print(f.split("/")[-1], "identical", d["identical_output"], "| ref: table %.1f rate %.1f correct %.1f whole %.1f | ours: table %.2f rate %.2f correct %.2f whole %.2f -> %.1f M reads/s in phase; speedup phase %.1f whole %.1f" % (
            r["load_table_s"], r["error_rate_s"], r["correct_s"], r["exit"], m["load_table_s"], m["error_rate_s"], m["correct_s"], m["exit"], m["reads_per_s_correct_phase"]/1e6, d["speedup_correct_phase"], d["speedup_whole_run"]))
    except Exception as e:
        print(f, "unreadable", e)
P
grep -h "RCX_TIMING" $OUT/binary_single_4M.err | tail -12
# -countkmers against -c on the same files
D=/dev/shm/rcx_binary_round
for mode in "-c $D/t.dump" "-countkmers"; do
  /usr/bin/time -f "%e s wall" env RCX_TIMING=1 rcorrector_b200/bin/rcorrector -p $D/r_1.fq $D/r_2.fq $mode -k 23 -od $D/out_ck -t 16 2> $OUT/countkmers_$(echo $mode | cut -c2-4).err; tail -4 $OUT/countkmers_$(echo $mode | cut -c2-4).err | head -3; md5sum $D/out_ck/*.cor.fq | cut -c1-32 | tr '\n' ' '; echo
done
grep -h "RCX_TIMING" $OUT/countkmers_cou.err | head -8
