"""Summarise `ncu --page raw --csv` exports (gpurun_out/*.csv) into compact tables under profiles/.

    python tools/summarize_ncu_csv.py gpurun_out/r02_infer_b4.csv profiles/r02_ncu_infer_b4.txt "title"
"""
import csv
import re
import sys

COLS = [
    ("us", "gpu__time_duration.sum", 1e-3),
    ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("GB/s", "dram__bytes.sum.per_second", 1e-9),
    ("rdMB", "dram__bytes_read.sum", 1e-6),
    ("wrMB", "dram__bytes_write.sum", 1e-6),
    ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1),
    ("tc%", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", 1),
    ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
    ("alu%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1),
    ("fma%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1),
    ("xu%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1),
    ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1),
    ("l1wf%", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1),
    ("L2hit%", "lts__t_sector_hit_rate.pct", 1),
    ("regs", "launch__registers_per_thread", 1),
    ("smemKB", "launch__shared_mem_per_block_dynamic", 1e-3),
    ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
]


def short(name):
    m = re.search(r"(\w+_kernel)(<.*>)?", name)
    if not m:
        return name[:70]
    args = (m.group(2) or "").replace("(int)", "").replace("(bool)", "")
    return (m.group(1).replace("w2s::", "") + args)[:84]


def unit_scale(unit):
    u = unit.lower()
    return {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9,
            "byte/s": 1.0, "kbyte/s": 1e3, "mbyte/s": 1e6, "gbyte/s": 1e9, "tbyte/s": 1e12}.get(u, 1.0)


def main():
    src, dst = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else src
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = [f"# {title}", "# one line per profiled launch; % columns are ncu pct_of_peak_sustained (tensor% = sm__pipe_tensor_cycles_active)",
           "", f"{'kernel':86s} {'grid':>10s} " + " ".join(f"{c[0]:>8s}" for c in COLS)]
    for r in rows[2:]:
        vals = []
        for label, key, mul in COLS:
            if key in idx and r[idx[key]] not in ("", "n/a"):
                v = float(r[idx[key]].replace(",", "")) * unit_scale(units[idx[key]]) * mul
                vals.append(f"{v:8.1f}")
            else:
                vals.append(f"{'-':>8s}")
        out.append(f"{short(r[idx['Kernel Name']]):86s} {r[idx['Grid Size']]:>10s} " + " ".join(vals))
    open(dst, "w").write("\n".join(out) + "\n")
    print("wrote", dst, len(rows) - 2, "launches")


if __name__ == "__main__":
    main()
