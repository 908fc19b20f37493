"""Summarise an .ncu-rep (ncu --set full) into a small CSV + markdown table for profiles/.

    python tools/summarize_ncu.py gpurun_out/prof_r1_kernels.ncu-rep profiles/r1_kernels
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("lts__t_bytes.sum", "l2_MB"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__cycles_active.avg", "cycles"),
]


def to_float(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def main(rep, out_prefix):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(header)}
    name_i = idx["Kernel Name"]
    out_rows = []
    for r in data:
        rec = {"kernel": r[name_i][:90]}
        for key, short in KEYS:
            cands = [h for h in header if h.startswith(key)]
            if not cands:
                rec[short] = ""
                continue
            i = idx[cands[0]]
            v = to_float(r[i])
            u = units[i]
            if v is None:
                rec[short] = r[i]
                continue
            if short == "dur_us":
                v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
            if short.endswith("_MB"):
                scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                v = v * scale
            rec[short] = round(v, 2)
        out_rows.append(rec)
    cols = ["kernel"] + [s for _, s in KEYS]
    with open(out_prefix + ".csv", "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=cols)
        w.writeheader()
        w.writerows(out_rows)
    with open(out_prefix + ".md", "w") as f:
        f.write("| " + " | ".join(cols) + " |\n|" + "---|" * len(cols) + "\n")
        for r in out_rows:
            f.write("| " + " | ".join(str(r[c]) for c in cols) + " |\n")
    print(open(out_prefix + ".md").read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
