#!/usr/bin/env python
"""Summarise gpurun_out/ profiling artefacts into profiles/ (tracked).
  python tools/summarize_profile.py <tag>
Reads gpurun_out/launches_<tag>.csv (ncu --metrics gpu__time_duration.sum launch list) and, when present,
gpurun_out/prof_*_<tag>.ncu-rep (ncu --set full), writes profiles/<tag>_launches.md and profiles/<tag>_ncu_<name>.md."""
import collections
import csv
import glob
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma_type_fp16.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second"]


def launches(tag):
    path = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = row["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        name = row["Kernel Name"].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list `{tag}` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / tot:.3f} | {v[1] / v[0]:.1f} |\n")
        f.write(f"\ntotal {tot:.1f} us over {sum(v[0] for v in agg.values())} launches\n")
    print("wrote", f"profiles/{tag}_launches.md")


def dram(tag):
    """gpurun_out/dram_<tag>.csv: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum over whole
    forwards at the bench's micro-batch -> profiles/<tag>_dram.md and profiles/traffic.json (read by bench.py's roofline)."""
    import json
    path = os.path.join(ROOT, "gpurun_out", f"dram_{tag}.csv")
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.defaultdict(lambda: collections.defaultdict(float))   # kernel -> metric -> sum (bytes / us)
    cnt = collections.Counter()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u, m = row["Metric Unit"], row["Metric Name"]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1.0)
        name = row["Kernel Name"].split("(")[0]
        per[name][m] += v * scale
        if m == "gpu__time_duration.sum":
            cnt[name] += 1
    conv = [k for k in per if "conv_halo_kernel" in k or "conv_tc_kernel" in k]
    with open(os.path.join(ROOT, "profiles", f"{tag}_dram.md"), "w") as f:
        f.write(f"# ncu DRAM traffic per kernel `{tag}` (dram__bytes_read.sum + dram__bytes_write.sum; whole forwards at micro-batch 64)\n\n")
        f.write("| kernel | launches | read MB / launch | write MB / launch | avg us | GB/s |\n|---|---:|---:|---:|---:|---:|\n")
        for k in sorted(per, key=lambda k: -per[k]["gpu__time_duration.sum"]):
            n = max(cnt[k], 1)
            r, w, t = per[k]["dram__bytes_read.sum"] / n, per[k]["dram__bytes_write.sum"] / n, per[k]["gpu__time_duration.sum"] / n
            f.write(f"| `{k}` | {n} | {r / 1e6:.1f} | {w / 1e6:.1f} | {t:.1f} | {(r + w) / max(t, 1e-9) / 1e3:.0f} |\n")
    nconv = sum(cnt[k] for k in conv)
    if nconv:
        tot = sum(per[k]["dram__bytes_read.sum"] + per[k]["dram__bytes_write.sum"] for k in conv)
        with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
            json.dump({"conv_tcgen05": {"dram_bytes_per_launch": tot / nconv, "launches": nconv,
                                        "note": f"ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over {nconv} tcgen05 conv launches of whole forwards "
                                                f"at micro-batch 64 (profiles/{tag}_dram.md)"}}, f, indent=1)
    print("wrote", f"profiles/{tag}_dram.md")


def full(tag):
    for rep in glob.glob(os.path.join(ROOT, "gpurun_out", f"prof_*_{tag}.ncu-rep")):
        name = os.path.basename(rep)[len("prof_"):-len(f"_{tag}.ncu-rep")]
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_{name}.md"), "w") as f:
            f.write(f"# ncu --set full `{name}` ({tag}); one block per captured launch\n")
            for row in rows[2:]:
                d = dict(zip(hdr, row))
                f.write(f"\n## {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}\n\n| metric | unit | value |\n|---|---|---:|\n")
                for i, h in enumerate(hdr):
                    if h in KEEP:
                        f.write(f"| {h} | {units[i]} | {row[i]} |\n")
        print("wrote", f"profiles/{tag}_ncu_{name}.md")


if __name__ == "__main__":
    t = sys.argv[1]
    launches(t)
    dram(t)
    full(t)
    b = os.path.join(ROOT, "gpurun_out", f"bench_{t}.json")
    if os.path.exists(b) and os.path.getsize(b):
        import shutil
        shutil.copy(b, os.path.join(ROOT, "profiles", f"{t}_bench.json"))
        print("copied bench json")
