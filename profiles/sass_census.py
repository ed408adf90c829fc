"""SASS census of libcogaim_b200.so: which kernels carry the Blackwell-native instructions (B200_PROFILING.md, "What
proves a Blackwell-native kernel").  Run here (no GPU needed):  python profiles/sass_census.py > profiles/sass_census.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(__file__).resolve().parent.parent / "cognitive_aim_depth_estimation_b200" / "libcogaim_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "UTCATOMSWS",
             "USETMAXREG", "FFMA2", "FADD2", "MUFU.EX2", "HMMA", "HGMMA"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "").replace("void ", "")
        cur = re.sub(r"\(.*", "", cur).replace("ca::", "")
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        per[cur]["_total"] += 1
        for mn in MNEMONICS:
            if op.startswith(mn):
                per[cur][mn] += 1
                if mn == "UTCHMMA" and ".2CTA" in op:
                    per[cur]["UTCHMMA.2CTA"] += 1
                if mn == "UTMALDG" and ".2CTA" in op:
                    per[cur]["UTMALDG.2CTA"] += 1
cols = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMALDG.2CTA", "UTMAREDG", "SYNCS", "FFMA2",
        "MUFU.EX2", "HMMA", "HGMMA"]
print(f"# cuobjdump -sass {lib.name}: instruction counts per kernel (static SASS)")
print("# tcgen05.mma -> UTCHMMA (.2CTA = cta_group::2), tcgen05.ld/st -> LDTM/STTM, TMA load -> UTMALDG, TMA reduce-add ->")
print("# UTMAREDG, tcgen05.commit -> UTCBAR, mbarrier -> SYNCS; HMMA (mma.sync) and HGMMA (wgmma) must be absent")
print(f"{'kernel':64s} {'instrs':>7s} " + " ".join(f"{c:>12s}" for c in cols))
for k, c in per.items():
    print(f"{k[:64]:64s} {c['_total']:7d} " + " ".join(f"{c[x]:12d}" for x in cols))
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print(f"{'TOTAL':64s} {tot['_total']:7d} " + " ".join(f"{tot[x]:12d}" for x in cols))
assert tot["HMMA"] == 0 and tot["HGMMA"] == 0, "legacy tensor path found"
