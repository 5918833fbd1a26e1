"""Blackwell-specific SASS mnemonics per kernel of the shipped library (cuobjdump works without a GPU).
   python tools/sass_mnemonics.py > profiles/r02_sass_mnemonics.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "gspn_b200", "libgspn_b200.so")
WANT = ["UTCHMMA", "LDTM", "STTM", "UTCATOMSWS", "UTCBAR", "UTMASTG", "UBLKCP", "STAS", "SYNCS", "UGETNEXTWORKID", "UCGABAR_ARV", "CREDUX", "REDUX",
        "FFMA2", "FADD2", "FMUL2", "F2FP", "ELECT", "LDG.E.ENL2.256", "STG.E.ENL2.256", "HMMA"]
pat = re.compile(r"^\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
per, cur = collections.OrderedDict(), None
for line in out.splitlines():
    if "Function :" in line:
        cur = line.split("Function :")[1].strip()
        per[cur] = collections.Counter()
        continue
    m = pat.match(line)
    if m and cur:
        op = m.group(1)
        for w in WANT:
            if op == w or op.startswith(w + "."):
                per[cur][w] += 1
                break
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print("# cuobjdump -sass gspn_b200/libgspn_b200.so : Blackwell-specific mnemonics per kernel (count of SASS instructions); tools/sass_mnemonics.py")
print("# UTCHMMA = tcgen05.mma (kind::f16), LDTM/STTM = tcgen05.ld/st, UTCATOMSWS = tcgen05.alloc/dealloc, UTCBAR = tcgen05.commit, UTMASTG = TMA tensor store,")
print("# UBLKCP = cp.async.bulk, STAS = st.async (DSMEM), SYNCS = mbarrier ops, UGETNEXTWORKID = clusterlaunchcontrol.try_cancel (work-stealing tiles),")
print("# UCGABAR = barrier.cluster, CREDUX/REDUX = redux.sync, FFMA2/FADD2/FMUL2 = packed fp32x2 math, ELECT = elect.sync, *.ENL2.256 = 256-bit global access")
print("total: " + ", ".join("%s=%d" % (w, tot[w]) for w in WANT))
print()
for k, c in per.items():
    if sum(c.values()):
        print(k)
        print("    " + ", ".join("%s=%d" % (w, c[w]) for w in WANT if c[w]))
