#!/usr/bin/env python
"""SASS mnemonic counts per tensor-core kernel of libcodae_b200.so (cuobjdump -sass, no GPU needed):
    python tools/sass_summary.py > profiles/r02_sass_summary.txt
UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA bulk tensor load / store, UTCBAR = tcgen05.commit,
SYNCS = mbarrier operations, UCGABAR = cluster barrier arrive / wait."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "mui-deepautoencoder_b200", "codae", "_lib", "libcodae_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
parts = re.split(r"\n\s*Function : ", txt)
print(__doc__.strip().replace("\n", "\n# ").join(["# ", ""]))
print("%-70s %8s %6s %8s %8s %7s %6s %8s" % ("kernel", "UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "UCGABAR"))
for p in parts[1:]:
    name = p.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", "-p", name], capture_output=True, text=True).stdout.strip()
    dem = dem.replace("(anonymous namespace)::", "")
    c = collections.Counter(re.findall(r"\b(UTCHMMA|LDTM|UTMALDG|UTMASTG|UTCBAR|SYNCS|UCGABAR_ARV|UCGABAR_WAIT)\b", p))
    if c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"]:
        print("%-70s %8d %6d %8d %8d %7d %6d %8d" % (dem[:70], c["UTCHMMA"], c["LDTM"], c["UTMALDG"], c["UTMASTG"], c["UTCBAR"],
                                                     c["SYNCS"], c["UCGABAR_ARV"] + c["UCGABAR_WAIT"]))
