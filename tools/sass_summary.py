"""Per-kernel counts of the Blackwell-specific SASS mnemonics in libmaua_b200.so (cuobjdump -sass):
    python tools/sass_summary.py > profiles/sass_summary.txt
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops,
ACQBULK / UBLKCP = bulk copies, FFMA2 = packed fp32 FMA."""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(__file__).resolve().parent.parent / "maua_style_b200" / "libmaua_b200.so"
out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
names = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "FFMA2", "FFMA", "DADD", "HMMA", "STG", "LDG"]
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"maua::\(anonymous namespace\)::", "", cur).split("(")[0].replace("void ", "")
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    base = op.split(".")[0]
    if base in names:
        counts[cur][base] += 1
    if op.startswith("UTCHMMA") and ".2CTA" in op:
        counts[cur]["UTCHMMA.2CTA"] += 1
print(f"{'kernel':58s} " + " ".join(f"{n:>12s}" for n in names))
for k, c in counts.items():
    if sum(c.values()) == 0:
        continue
    print(f"{k[:58]:58s} " + " ".join(f"{c.get(n, 0):12d}" for n in names))
