"""Top stall-sample SASS lines of one kernel in an .ncu-rep (source page), to see where warps wait.

    python tools/ncu_hot_sass.py REPORT KERNEL_REGEX [launch_skip] [top_n]
"""
import csv
import io
import subprocess
import sys

rep, regex = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{regex}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
lines = raw.splitlines()
print(lines[0][:160])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
i_src, i_samp, i_exec = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_")]
data = []
for k, r in enumerate(rows[1:]):
    try:
        data.append((int(r[i_samp]), k, r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print(f"total samples {tot}, {len(data)} SASS lines")
for s, k, r in sorted(data, reverse=True)[:top]:
    reasons = sorted(((int(r[i]) if r[i].isdigit() else 0, h) for i, h in stall_cols), reverse=True)[:2]
    rs = ", ".join(f"{h[6:]}={v}" for v, h in reasons if v)
    print(f"{100*s/tot:5.1f}%  line {k:5d}  exec={r[i_exec]:>8}  {r[i_src].strip()[:70]:70s} {rs}")
