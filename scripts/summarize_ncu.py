"""Turns the ncu artefacts a gpurun call brought back into the committed summaries under profiles/.

    python scripts/summarize_ncu.py <tag> <launches.csv> [<report.ncu-rep> ...]
writes profiles/<tag>_launches.md (per-kernel share of the step) and profiles/<tag>_<report>.md / .csv
(selected raw metrics of every captured launch)."""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"]


def launches(tag, path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e6 if r["Metric Unit"] in ("ns", "nsecond") else v / 1e3 if r["Metric Unit"] in ("us", "usecond") else v
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1; t[1] += v
    total = sum(v[1] for v in tot.values())
    with open(os.path.join(OUT, tag + "_launches.md"), "w") as f:
        f.write("# %s: every kernel launch of `bench.py --profile-run --steps 1` under ncu\n\n" % tag)
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (1 warm-up step + 1 step captured; times are\n"
                "cold-cache and serialised, so compare SHARES, not absolutes). %d launches, %.2f ms summed.\n\n" % (len(rows), total))
        f.write("| kernel | launches | ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f%% |\n" % (k[:90], v[0], v[1], 100 * v[1] / total))
    print("wrote", tag + "_launches.md")


def report(tag, path):
    base = os.path.splitext(os.path.basename(path))[0]
    if base.startswith(tag + "_"):
        base = base[len(tag) + 1:]
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    keep = [k for k in KEEP if k in idx]
    with open(os.path.join(OUT, "%s_%s.csv" % (tag, base)), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(keep); w.writerow([units[idx[k]] for k in keep])
        for r in rows[2:]:
            w.writerow([r[idx[k]] for k in keep])
    with open(os.path.join(OUT, "%s_%s.md" % (tag, base)), "w") as f:
        f.write("# %s / %s  (`ncu --set full --clock-control none --import-source on`)\n\n" % (tag, base))
        for r in rows[2:]:
            f.write("## %s  grid %s block %s\n\n" % (r[idx["Kernel Name"]][:100], r[idx["Grid Size"]], r[idx["Block Size"]]))
            for k in keep[3:]:
                f.write("- `%s` = %s %s\n" % (k, r[idx[k]], units[idx[k]]))
            f.write("\n")
    print("wrote", "%s_%s.md/.csv" % (tag, base))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    tag = sys.argv[1]
    launches(tag, sys.argv[2])
    for p in sys.argv[3:]:
        report(tag, p)
