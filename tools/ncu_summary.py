#!/usr/bin/env python
"""Summarise .ncu-rep captures (ncu --set full) into the text kept under profiles/: per kernel launch, the metrics the
roofline argument uses (duration, DRAM bytes, DRAM / tensor-pipe utilisation, occupancy limits).

    python tools/ncu_summary.py gpurun_out/prof_r1_*.ncu-rep > profiles/r01_ncu_full_metrics.txt
"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_tensor_op_gen5.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg",
           "sm__cycles_elapsed.max", "launch__occupancy_limit_shared_mem"]


def summarise(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
    if out.returncode != 0:
        print("# %s: ncu -i failed: %s" % (path, out.stderr.strip()[:200]))
        return
    rows = list(csv.reader(io.StringIO(out.stdout)))
    header, units, body = rows[0], rows[1], rows[2:]
    col = {name: i for i, name in enumerate(header)}
    print("# %s" % path)
    for r in body:
        print("  - %s" % r[col["Kernel Name"]][:60])
        for m in METRICS:
            if m in col:
                print("      %-66s %s %s" % (m, r[col[m]], units[col[m]]))


def traffic(pairs):
    """--traffic key=report.ncu-rep[:kernel-substring][#i,j] ... -> JSON {key: {bytes_per_launch, launches, kernels, read, write, source}}:
    mean dram__bytes_read.sum + dram__bytes_write.sum per captured launch (bench.py reads it as roofline.traffic)."""
    import json
    out = {}
    for pair in pairs:
        key, rest = pair.split("=", 1)
        rest, _, picks = rest.partition("#")                  # optional "#0,2": launch indices (capture order) to average over
        path, _, sub = rest.partition(":")
        picks = [int(x) for x in picks.split(",")] if picks else None
        res = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
        if res.returncode != 0:
            continue
        rows = list(csv.reader(io.StringIO(res.stdout)))
        header, units, body = rows[0], rows[1], rows[2:]
        col = {name: i for i, name in enumerate(header)}
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd, wr, names = [], [], []
        for li, r in enumerate(body):
            if (sub and sub not in r[col["Kernel Name"]]) or (picks is not None and li not in picks):
                continue
            ur, uw = units[col["dram__bytes_read.sum"]], units[col["dram__bytes_write.sum"]]
            rd.append(float(r[col["dram__bytes_read.sum"]].replace(",", "")) * scale.get(ur, 1.0))
            wr.append(float(r[col["dram__bytes_write.sum"]].replace(",", "")) * scale.get(uw, 1.0))
            names.append(r[col["Kernel Name"]][:48])
        if rd:
            out[key] = {"bytes_per_launch": (sum(rd) + sum(wr)) / len(rd), "launches": len(rd), "read": sum(rd) / len(rd),
                        "write": sum(wr) / len(wr), "kernels": sorted(set(names)), "source": path.split("/")[-1]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--traffic":
        traffic(sys.argv[2:])
    else:
        for p in sys.argv[1:]:
            summarise(p)
