"""SASS opcode histogram per kernel of the in-tree library (no GPU needed: cuobjdump -sass on the built .so), so the
tcgen05 / TMEM / TMA claim is checkable without rebuilding:

    python tools/sass_histogram.py [path/to/lib.so] > profiles/r2_sass_histogram.md

Per kernel: instruction count and the counts of the mnemonics that prove a Blackwell-native kernel
(UTCHMMA = tcgen05.mma kind::f16, UTMALDG/UTMASTG = TMA tensor load/store, UTMAPF = TMA prefetch, LDTM/STTM =
tcgen05.ld/st, UTCBAR = tcgen05.commit, SYNCS = mbarrier, MUFU.EX2) and of the legacy tensor path (HMMA must be 0).
R2UR / BRA.U.ANY: vector -> uniform register moves and the loops ptxas wraps around them when a tcgen05 / TMA instruction
is issued from divergent code (session 4): 4-7 R2UR per UTCHMMA before, < 1 in the converted roles."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "ecad_b200" / "libecad_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKPF", "LDTM", "STTM", "UTCBAR", "SYNCS", "MUFU.EX2", "MUFU.TANH",
        "HMMA", "LDG", "STG", "LDS", "STS", "FFMA2", "BAR", "R2UR", "BRA.U.ANY"]
funcs: dict[str, collections.Counter] = {}
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = funcs.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["_total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR") and op.startswith(k)):
                cur[k] += 1
print(f"# SASS opcode histogram of `{lib.relative_to(ROOT) if lib.is_relative_to(ROOT) else lib}` "
      f"({lib.stat().st_size} bytes, sm_100a; `python tools/sass_histogram.py`)\n")
print("| kernel | SASS instr | " + " | ".join(KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
tot = collections.Counter()
for name in sorted(funcs, key=lambda n: -funcs[n]["_total"]):
    c = funcs[name]
    tot.update(c)
    short = re.sub(r"CUtensorMap_st(, )?", "", demangle(name)).replace("ecadk::", "")
    short = re.sub(r"\(.*", "", short).replace("void ", "")
    print(f"| `{short}` | {c['_total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")
print(f"| **all {len(funcs)} kernels** | {tot['_total']} | " + " | ".join(str(tot[k]) for k in KEYS) + " |")
print(f"\nLegacy tensor path (`HMMA` = mma.sync / wmma): **{tot['HMMA']}** instructions.")
