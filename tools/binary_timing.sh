d=/dev/shm/rcx_bt; mkdir -p $d
tools/bin/rcx_synth -seed 3 -k 23 -ntx 7000 -txmin 300 -txmax 2500 -polya 0.3 -sigma 2.6 -reads 4000000 -len 100 -err 0.01 -nrate 0.002 -short 0.02 -fq1 $d/r_1.fq -dump $d/t.dump > /dev/null 2>&1
for l in 0.2 0.4; do echo "== load $l"; RCX_TABLE_LOAD=$l RCX_TIMING=1 rcorrector_b200/bin/rcorrector -r $d/r_1.fq -c $d/t.dump -k 23 -od $d/o -t 16 2>&1 | grep -E "RCX_TIMING|Processed"; done
