#!/bin/bash
# ThreadSanitizer and ASan/UBSan runs of the drop-in binary's HOST side (no GPU): rcorrector_main.cpp against the librcx
# test double, over the reader / pipeline / sharding / gzip paths.  Test infrastructure.   usage: bash tools/host_sanitize.sh
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d /tmp/rcx_san.XXXXXX)
make -s -C $ROOT/oracle librcx_oracle.so
LINK="$ROOT/rcorrector_b200/csrc/rcorrector_main.cpp $ROOT/tests/harness/rcx_double.cpp -L$ROOT/oracle -l:librcx_oracle.so -lz -Wl,-rpath,$ROOT/oracle"
g++ -O1 -g -fsanitize=thread -std=c++17 -pthread -o $W/tsan $LINK
g++ -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -std=c++17 -pthread -o $W/asan $LINK
python - "$ROOT" "$W" <<'PY'
import gzip, os, shutil, subprocess, sys
root, w = sys.argv[1], sys.argv[2]
sys.path.insert(0, os.path.join(root, "tests"))
import rcx_testlib as T
m, fq, dump, exp = T.load_golden("paired_k23", w)
big1, big2 = os.path.join(w, "b_1.fq"), os.path.join(w, "b_2.fq")
open(big1, "wb").write(open(fq[0], "rb").read() * 30)
open(big2, "wb").write(open(fq[1], "rb").read() * 30)
gz1 = big1 + ".gz"
open(gz1, "wb").write(gzip.compress(open(big1, "rb").read(), 1))
L = open(fq[0], "rb").read().split(b"\n")
L[7] = L[7][:-3]
L[203] += b"II"
rag = os.path.join(w, "rag_1.fq")
open(rag, "wb").write(b"\n".join(L))
cases = [["-p", big1, big2, "-t", "4", "-batch", "2048", "-gpus", "3"], ["-r", gz1, "-t", "4", "-batch", "4096"],
         ["-p", big1, big2, "-t", "3", "-stdout"], ["-p", fq[0], fq[1], "-t", "2"], ["-p", rag, fq[1], "-t", "2", "-batch", "64"],
         ["-i", fq[0], "-t", "2"], ["-r", os.path.join(w, "nope")]]
bad = 0
for exe, env in (("tsan", {"RCX_DOUBLE_PASSTHROUGH": "1"}), ("asan", {"ASAN_OPTIONS": "detect_leaks=0"})):
    for args in cases:
        od = os.path.join(w, "o")
        shutil.rmtree(od, ignore_errors=True)
        r = subprocess.run([os.path.join(w, exe)] + args + ["-c", dump, "-k", "23", "-od", od], capture_output=True, env=dict(os.environ, **env))
        err = r.stderr.decode(errors="replace")
        found = "ThreadSanitizer" in err or "AddressSanitizer" in err or "runtime error" in err
        bad += found
        print("%s %-60s rc %d %s" % (exe, " ".join(os.path.basename(a) for a in args)[:60], r.returncode, "FINDINGS" if found else "clean"))
        if found:
            print(err[:4000])
sys.exit(1 if bad else 0)
PY
rm -rf $W
