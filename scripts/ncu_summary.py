"""Summarise ncu outputs (run here, no GPU needed) into small tracked files under profiles/.

  python scripts/ncu_summary.py launches gpurun_out/launches.csv profiles/rN_launches.md
  python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/rN_ncu_full.csv
"""
import collections
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def short(name):
    return name.split("(")[0].replace("void ", "").replace("eqvio::", "")


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if hdr is None:
            if r and r[0] == "ID":
                hdr = r
            continue
        if len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        v = float(d["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(d["Metric Unit"], v)
        a = agg.setdefault(short(d["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    mine = {k: v for k, v in agg.items() if not (k.startswith("at::") or k.startswith("cutlass") or "cublas" in k.lower())}
    tot_m = sum(v[1] for v in mine.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` -- per-launch times are cold-cache and serialised;\n"
                "compare SHARES, not absolutes.  Share(own) excludes torch / cuBLAS kernels (L2 flush, DGEMM peak calibration).\n\n")
        f.write("| kernel | launches | total us | avg us | share(all) | share(own) |\n|---|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            own = f"{100 * v[1] / tot_m:.1f}%" if k in mine else "-"
            f.write(f"| {k} | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f}% | {own} |\n")
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in KEEP if c in idx]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{c} [{units[idx[c]]}]" for c in cols])
        for r in rows[2:]:
            w.writerow([short(r[idx["Kernel Name"]])] + [r[idx[c]] for c in cols])
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
