import sys, os, json, numpy as np
sys.path.insert(0, '/root/repo' if os.path.exists('/root/repo/bench.py') else '.')
import bench
from rcorrector_b200 import Corrector, SINGLE
class A: pass
a = bench.parse_args()
d = bench.gen_table(a, 1); bench.gen_reads(a, 1)
codes=np.fromfile(d+'/table.codes',dtype=np.uint64); counts=np.fromfile(d+'/table.counts',dtype=np.int32)
c=Corrector(0); c.build_table(a.k,codes,counts); c.set_params(a.k,0.01,48,4)
seq=np.fromfile(d+'/reads.seq',dtype=np.uint8); qual=np.fromfile(d+'/reads.qual',dtype=np.uint8); off=np.fromfile(d+'/reads.off',dtype=np.uint32)
for i in range(2):
    c.correct(SINGLE, seq, qual, off)
st=c.stats(); print(st)
dc=c.debug_counters(29)
names=['scan_probes','search_probes','rescan_probes','nodes','work','nwork','work2','nwork2','wx0','wx1','wx2','wx3','wx4','rounds','ff','regular','ph0 fetch','ph1 planes+loads','ph2 thr/masks/islands','ph3 search','ph4 after-search','ph5 postfilter','ph6 lmh','ph7 tail','jumps','jump cycles','rescue steps','round cycles','round levels']
tot=float(sum(dc[16:24]))
for n,v in zip(names,dc): print('%-24s %15d %s'%(n,v, '%.1f%%'%(100*v/tot) if (n.startswith('ph') or 'cycles' in n) else ''))
