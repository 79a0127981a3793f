"""Per-kernel histogram of the SASS mnemonics that show which hardware path a kernel uses (tcgen05 = UTCHMMA / UTCBAR,
TMA = UTMALDG, tensor memory = LDTM / STTM, legacy tensor cores = HMMA, MUFU ...).  CPU only: reads the built library.
Usage: python scripts/sass_opcodes.py [librdeic_b200.so] > profiles/r02_sass_opcodes.txt"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

lib = sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).resolve().parents[1] / "rdeic_b200" / "librdeic_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
dem = {}
kernels = OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        kernels[cur][m.group(1).split(".")[0]] += 1
names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
WATCH = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMAPF", "LDTM", "STTM", "UTCATOM", "HMMA", "MUFU", "LDGSTS", "LDSM", "SYNCS", "FFMA2", "DFMA", "DADD"]
print(f"# {lib}: {len(kernels)} kernels; columns = instruction counts in the SASS of each kernel")
print(f"{'kernel':110s} " + " ".join(f"{w:>7s}" for w in WATCH) + "   total")
for (mangled, c), name in zip(kernels.items(), names):
    short = re.sub(r"\(.*", "", name).replace("void rdeic::", "").replace("rdeic::", "")
    print(f"{short[:110]:110s} " + " ".join(f"{c.get(w, 0):7d}" for w in WATCH) + f" {sum(c.values()):7d}")
tot = Counter()
for c in kernels.values():
    tot.update(c)
print(f"{'ALL':110s} " + " ".join(f"{tot.get(w, 0):7d}" for w in WATCH) + f" {sum(tot.values()):7d}")
