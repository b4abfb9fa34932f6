"""Turn ncu artefacts brought back in gpurun_out/ into the small, diffable extracts kept under profiles/.

    python tools/ncu_extract.py metrics gpurun_out/X.ncu-rep profiles/X_metrics.csv     # selected metrics of every kernel in a report
    python tools/ncu_extract.py launches gpurun_out/L.csv                               # per-kernel share table of a launch list
    python tools/ncu_extract.py traffic gpurun_out/T.csv profiles/T.json tiles          # summed dram bytes of a traffic list
"""
import collections
import csv
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma_type_fp16.avg.pct_of_peak_sustained_active", "lts__t_sectors_srcunit_tex_op_read.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def metrics(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, u = rows[0], rows[1]
    cols = [i for i, n in enumerate(h) if n in ("ID", "Kernel Name") or n in KEEP]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([h[i] for i in cols])
        w.writerow([u[i] for i in cols])
        for r in rows[2:]:
            w.writerow([r[i] for i in cols])
    print("wrote", out, len(rows) - 2, "kernels")


def _load(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi, mi, ii = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name"), h.index("ID")
    return [(r[ii], r[ki], r[mi], float(r[vi].replace(",", ""))) for r in rows[hdr + 1:]]


def launches(path, marker="stem_s2d"):
    d = [x for x in _load(path) if x[2] == "gpu__time_duration.sum"]
    idx = [i for i, x in enumerate(d) if marker in x[1]]
    a, b = (idx[-2], idx[-1]) if len(idx) >= 2 else (0, len(d))
    agg, tot = collections.OrderedDict(), 0.0
    for _, k, _, v in d[a:b]:
        name = k.split("(")[0][:90]
        agg.setdefault(name, [0.0, 0])
        agg[name][0] += v
        agg[name][1] += 1
        tot += v
    for k, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
        print("%-92s %4d %10.1f us %5.1f%%" % (k, c, v / 1e3, 100 * v / tot))
    print("total %.1f us over %d launches (one step)" % (tot / 1e3, b - a))


def traffic(path, out, tiles):
    d = _load(path)
    rd = sum(v for _, _, m, v in d if m == "dram__bytes_read.sum")
    wr = sum(v for _, _, m, v in d if m == "dram__bytes_write.sum")
    t = sum(v for _, _, m, v in d if m == "gpu__time_duration.sum")
    n = len([1 for _, _, m, _ in d if m == "gpu__time_duration.sum"])
    # ncu prints bytes with a unit column we dropped: values are in the unit of the first row; keep raw numbers + note
    json.dump({"tiles": int(tiles), "launches": n, "dram_bytes_read": rd, "dram_bytes_write": wr, "time_ns_sum": t,
               "source": path}, open(out, "w"), indent=1)
    print("wrote", out, n, "launches", rd, wr)


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "metrics":
        metrics(sys.argv[2], sys.argv[3])
    elif cmd == "launches":
        launches(*sys.argv[2:])
    elif cmd == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
