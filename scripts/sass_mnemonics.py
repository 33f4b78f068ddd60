"""SASS mnemonic counts per kernel of the built library (cuobjdump -sass; no GPU needed):
   python scripts/sass_mnemonics.py > profiles/<tag>_sass_mnemonics.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "clearconverse_b200", "libresep_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
it = iter(names)
per = collections.OrderedDict()
cur = None
KEEP = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "MUFU", "FFMA2", "FADD2", "FMUL2", "F2FP", "HFMA2", "LDGSTS", "SYNCS", "LDSM")
for line in out.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"\(.*", "", next(it)).replace("resep::", "").replace("void ", "")
        per.setdefault(cur, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op = m.group(1)
        if op in KEEP:
            per[cur][op + (".2CTA" if ".2CTA" in m.group(2) else "")] += 1
print("SASS mnemonic counts per kernel of clearconverse_b200/libresep_b200.so (cuobjdump -sass, sm_100a)")
print("UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store,")
print("UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc, HMMA = legacy mma.sync, FADD2/FFMA2 = packed fp32x2, SYNCS = mbarrier\n")
for k, c in per.items():
    if c:
        print(k)
        print("    " + ", ".join(f"{op}: {n}" for op, n in sorted(c.items())))
