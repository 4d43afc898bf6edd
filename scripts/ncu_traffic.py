"""profiles/traffic.json from ncu --set full captures (run here, no GPU needed):

    python scripts/ncu_traffic.py 64=gpurun_out/prof_n64.ncu-rep 256=gpurun_out/prof_n256.ncu-rep 1024=gpurun_out/prof_n1024.ncu-rep

Per kernel class of bench.py's roofline object (prop_ll, chunk_factor, downdate, chol_trail): dram__bytes_read.sum +
dram__bytes_write.sum per launch, mean over the captured launches of that kernel in the report for that N."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASSES = {"prop_ll_kernel": "prop_ll", "chunk_factor_kernel": "chunk_factor", "chunk_downdate_kernel": "downdate",
           "chunk_downdate_tc_kernel": "downdate_tc", "gemm_nt_sub_kernel": "chol_trail", "observer_fused_kernel": "observer_fused",
           "bc_diag_kernel": "bc_diag", "bc_panel_kernel": "bc_panel", "bc_trail_kernel": "bc_trail", "bc_next_kernel": "bc_next",
           "bc_build_kernel": "bc_build"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def per_launch_bytes(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    kcol = names.index("Kernel Name")
    cols = [i for i, n in enumerate(names) if n in ("dram__bytes_read.sum", "dram__bytes_write.sum")]
    acc = collections.defaultdict(list)
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        k = r[kcol].split("(")[0].replace("void ", "").replace("eqvio::", "").split("<")[0]
        if k in CLASSES:
            acc[CLASSES[k]].append(sum(float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0) for i in cols))
    return {k: (sum(v) / len(v), len(v)) for k, v in acc.items()}


def main():
    traffic = collections.defaultdict(dict)
    sources = {}
    for arg in sys.argv[1:]:
        n, rep = arg.split("=", 1)
        for k, (b, cnt) in per_launch_bytes(rep).items():
            traffic[k][n] = b
            sources.setdefault(n, {})[k] = cnt
    traffic = dict(traffic)
    traffic["_source"] = ("ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the "
                          "captured launches (counts: %s); reports: %s; summaries under profiles/ with the same stem"
                          % (json.dumps(sources), ", ".join(os.path.basename(a.split("=", 1)[1]) for a in sys.argv[1:])))
    json.dump(traffic, open(os.environ.get("EQVIO_TRAFFIC_OUT", os.path.join(ROOT, "profiles", "traffic.json")), "w"), indent=1)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
