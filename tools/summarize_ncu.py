"""Summarise ncu artefacts from gpurun_out/ into small text files under profiles/ (tracked).

    python tools/summarize_ncu.py r01
"""
import csv
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "profiles"
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "smsp__inst_issued.sum",
    "lts__t_sector_hit_rate.pct",
]
STALLS = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active.ratio")


def summarize_report(rep: Path, out: Path):
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full summary of {rep.name}", ""]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        lines.append(f"## kernel: {d.get('Kernel Name', '?')}")
        lines.append(f"grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d:
                lines.append(f"{k:85s} {d[k]:>18s} {u[k]}")
        st = sorted(((float(d[h]), STALLS.match(h).group(1)) for h in hdr if STALLS.match(h) and d[h]), reverse=True)
        lines.append("stall reasons (warps per issue-active cycle): " + ", ".join(f"{n}={v:.2f}" for v, n in st[:8]))
        lines.append("")
    out.write_text("\n".join(lines))
    print("wrote", out)


def summarize_launches(csvf: Path, out: Path, last=102):
    lines = [l for l in csvf.read_text().splitlines() if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    seq = [(r["Kernel Name"], r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1000) for r in rows][-last:]
    tot = sum(t for *_, t in seq)
    agg = {}
    for k, g, b, t in seq:
        m = re.search(r"(conv_stream_kernel|conv_igemm_kernel|epoch_mixer_kernel|seq_mixer_kernel|first_conv_kernel|argmax_kernel)(<[^>]*>)?", k)
        name = (m.group(1) + (m.group(2) or "")).replace("(int)", "").replace("(bool)", "") if m else k[:60]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    txt = [f"# ncu launch list of one bench step ({len(seq)} launches, {tot:.1f} us total, cold-cache serialised replays)",
           "# columns: kernel<template args> | launches | total us | share of step", ""]
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        txt.append(f"{name:75s} {n:4d} {t:10.1f} {100 * t / tot:6.2f}%")
    txt += ["", "# every launch in order: kernel | grid | block | us"]
    for k, g, b, t in seq:
        m = re.search(r"(conv_stream_kernel|conv_igemm_kernel|epoch_mixer_kernel|seq_mixer_kernel|first_conv_kernel|argmax_kernel)(<[^>]*>)?", k)
        name = (m.group(1) + (m.group(2) or "")).replace("(int)", "").replace("(bool)", "") if m else k[:60]
        txt.append(f"{name:75s} {g:16s} {b:12s} {t:9.1f}")
    out.write_text("\n".join(txt))
    print("wrote", out)


OUT.mkdir(exist_ok=True)
g = ROOT / "gpurun_out"
if (g / "launches.csv").exists():
    last = 102
    log = g / "ncu_bench.log"
    if log.exists():  # one step's launch count as bench.py itself counted it
        import json
        for line in log.read_text().splitlines():
            if line.startswith("{") and "gpu_launches" in line:
                d = json.loads(line)
                last = d["gpu_launches"] // max(d["steps"], 1)
    n_rows = sum(1 for l in (g / "launches.csv").read_text().splitlines() if l.startswith('"')) - 1
    if not log.exists():
        last = n_rows  # captured with --profile-from-start off around exactly one step
    summarize_launches(g / "launches.csv", OUT / f"{tag}_launch_list_bench_step.txt", last=last)
for rep in sorted(g.glob(f"{tag}_prof_*.ncu-rep")):
    summarize_report(rep, OUT / f"{tag}_{rep.stem}.txt")
