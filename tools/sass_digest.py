#!/usr/bin/env python
"""Per-kernel count of the SASS mnemonics that prove (or disprove) a Blackwell-native path, from the in-tree library:
UTCHMMA/UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG (cp.async.bulk.tensor), UBLKCP / UBLKRED
(cp.async.bulk / cp.reduce.async.bulk), SYNCS (mbarrier), HMMA (legacy mma.sync - must be 0).
usage: tools/sass_digest.py [lib.so] > profiles/rN_sass_digest.txt"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "lsfa_b200", "lib", "liblsfa_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
PAT = OrderedDict([("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
                   ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"), ("UBLKRED", r"\bUBLKRED"), ("SYNCS", r"\bSYNCS"),
                   ("UTCBAR", r"\bUTCBAR"), ("HMMA", r"\bHMMA"), ("LDGSTS", r"\bLDGSTS")])
tot = Counter()
rows = []
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0].strip()
    c = {k: len(re.findall(p, f)) for k, p in PAT.items()}
    if any(c.values()):
        dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        cut = dem.rfind(">(")
        dem = dem[:cut + 1] if cut > 0 else dem.split("(")[0]
        dem = dem.replace("(int)", "").replace("void ", "").replace("lsfa::", "")
        rows.append((dem, c))
        tot.update(c)
print("# SASS digest of %s (cuobjdump -sass, sm_100a)" % os.path.relpath(lib, ROOT))
print("# %-64s " % "kernel" + " ".join("%8s" % k for k in PAT))
for dem, c in sorted(rows, key=lambda t: t[0]):
    print("  %-64s " % dem[:64] + " ".join("%8d" % c[k] for k in PAT))
print("  %-64s " % "TOTAL" + " ".join("%8d" % tot[k] for k in PAT))
