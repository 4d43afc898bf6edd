"""Per-CUDA-source-line stall samples of one kernel from an ncu report (needs -lineinfo + --import-source on).
   python scripts/ncu_lines.py gpurun_out/prof.ncu-rep <kernel regex> [top]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kern}",
                      "--launch-skip", "0", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[hi]
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
cols = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_mio", "stall_wait", "stall_math", "stall_lg", "stall_branch_resolving", "stall_no_inst"]
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[0]:
        continue
    try:
        s = int(r[ix["# Samples"]])
    except ValueError:
        continue
    data.append((s, r[0], r[1].strip()[:100], [r[ix[c]] for c in cols], r[ix["Instructions Executed"]]))
tot = sum(d[0] for d in data) or 1
print("total samples", tot, "| columns:", " ".join(c.replace("stall_", "") for c in cols))
for d in sorted(data, reverse=True)[:top]:
    print(f"{100 * d[0] / tot:5.1f}% L{d[1]:>5s} {' '.join(f'{v:>5s}' for v in d[3])} inst={d[4]:>8s} | {d[2]}")
