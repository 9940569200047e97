#!/usr/bin/env python
"""CPU differential campaign for the dump ingest of the drop-in binary: random Jellyfish-style dumps with
awkward records — counts 0 / 1 / 2^24-1 / 2^24 / 2^29 / 2^31-1, k-mers with N, reverse complements and duplicates with
other counts (last writer wins), k-mers one base short or long, blank lines, blanks or CRLF as separators, no final
newline — fed to the UNMODIFIED reference binary and to rcorrector_main.cpp linked against the librcx test double.
Compares return code, stderr (Stored / rate lines) and the corrected files.  Test infrastructure only.

usage: python tools/fuzz_dump_vs_ref.py [seed] [cases]
"""
import sys,os,subprocess,shutil,random,gzip
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,os.path.join(ROOT,'tests')); sys.path.insert(0,os.path.join(ROOT,'tools'))
import rcx_testlib as T
import fuzz_binary_vs_ref as F
import tempfile
top=tempfile.mkdtemp(prefix="rcxdumpfuzz_")
exe=F.build_double(top)
meta,fq,dump,_=T.load_golden("paired_k23",top)
k=23
toks=open(dump,"rb").read().split()
recs=[(toks[i],toks[i+1]) for i in range(0,len(toks)-1,2)]
rng=random.Random(int(sys.argv[1]) if len(sys.argv)>1 else 1)
def rc(s):
    return bytes({65:84,67:71,71:67,84:65}.get(c,c) for c in reversed(s))
bad=0
N=int(sys.argv[2]) if len(sys.argv)>2 else 40
for case in range(N):
    wd=os.path.join(top,"c%03d"%case); os.makedirs(wd)
    rs=list(recs)
    rng.shuffle(rs)
    rs=rs[:rng.choice([50,500,len(rs)])]
    out=[]
    feats=[]
    for (c,km) in rs:
        r=rng.random()
        if r<0.02: c=b">1"; feats.append("count1")
        elif r<0.04: c=b">0"; feats.append("count0")
        elif r<0.05: c=b">%d"%rng.choice([16777215,16777216,2**29,2**31-1]); feats.append("big")
        elif r<0.06: km=km[:5]+b"N"+km[6:]; feats.append("N")
        elif r<0.07: km=rc(km); feats.append("rc")
        elif r<0.08: out.append((c,km)); c=b">%d"%rng.choice([2,7,5000]); feats.append("dup")
        elif r<0.085: km=km[:-1]; feats.append("short")
        elif r<0.09: km=km+b"A"; feats.append("long")
        elif r<0.095: c=b">1500"; feats.append("c1500")
        out.append((c,km))
    sep=rng.choice(["nl","nl","nl","space","crlf","blank"])
    if sep=="nl": data=b"".join(c+b"\n"+km+b"\n" for c,km in out)
    elif sep=="space": data=b" ".join(c+b" "+km for c,km in out)+b"\n"
    elif sep=="crlf": data=b"".join(c+b"\r\n"+km+b"\r\n" for c,km in out)
    else: data=b"".join(c+b"\n"+km+b"\n"+(b"\n" if rng.random()<0.1 else b"") for c,km in out)
    if rng.random()<0.2 and data.endswith(b"\n"): data=data[:-1]; feats.append("nofinalnl")
    dp=os.path.join(wd,"d.dump"); open(dp,"wb").write(data)
    t=rng.choice(["1","2","4"])
    res=[]
    for tag,b in (("ref",T.REF_BIN),("ours",exe)):
        od=os.path.join(wd,"o_"+tag)
        r=subprocess.run([b,"-p",fq[0],fq[1],"-c",dp,"-k",str(k),"-od",od,"-t",t],capture_output=True)
        outs={f:open(os.path.join(od,f),"rb").read() for f in sorted(os.listdir(od))} if os.path.isdir(od) else {}
        res.append((r.returncode,r.stderr,outs))
    if res[0]!=res[1]:
        bad+=1
        print("MISMATCH",case,sep,sorted(set(feats)),"t",t)
        print("  ref :",res[0][0],res[0][1][:200]); print("  ours:",res[1][0],res[1][1][:200])
    else:
        shutil.rmtree(wd)
print("dump fuzz:",N,"cases",bad,"mismatching")
shutil.rmtree(top,ignore_errors=True)
sys.exit(1 if bad else 0)
