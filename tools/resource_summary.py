"""Registers / stack / static shared memory of every kernel in libmaua_b200.so (cuobjdump --dump-resource-usage):
    python tools/resource_summary.py > profiles/resource_usage.txt
STACK > 0 means local-memory frames (spills or indexed local arrays); dynamic shared memory is chosen at launch and not listed."""
import re
import subprocess
from pathlib import Path

lib = Path(__file__).resolve().parent.parent / "maua_style_b200" / "libmaua_b200.so"
out = subprocess.run(["cuobjdump", "--dump-resource-usage", str(lib)], capture_output=True, text=True).stdout
rows, cur = [], None
for line in out.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"maua::\(anonymous namespace\)::", "", cur).split("(")[0].replace("void ", "")
        continue
    m = re.match(r"\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and cur:
        rows.append((cur, *map(int, m.groups())))
        cur = None
print(f"{'kernel':72s} {'regs':>5s} {'stack':>6s} {'smem(static)':>13s} {'local':>6s}")
for name, reg, stack, smem, local in rows:
    print(f"{name[:72]:72s} {reg:5d} {stack:6d} {smem:13d} {local:6d}")
print(f"\n{len(rows)} kernels; with a stack frame: {sum(1 for r in rows if r[2] > 0)}; max registers: {max(r[1] for r in rows)}")
