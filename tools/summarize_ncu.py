"""Turn gpurun_out ncu artefacts into the small, tracked summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_v0.csv profiles/r01_v0_launches.txt
    python tools/summarize_ncu.py full gpurun_out/prof_conv_v0.ncu-rep profiles/r01_v0_conv_generic_full.txt
"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "launch__shared_mem_per_block_dynamic",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
           "lts__t_bytes.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg, rows, tot = collections.OrderedDict(), [], 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1.0)
        key = row["Kernel Name"].split("(")[0].replace("void ", "").replace("unnamed>::", "")
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        rows.append((v, key, row.get("Grid Size", ""), row.get("Block Size", "")))
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none : one step (cold-cache, serialised launches)\n")
        f.write("# source: %s ; total %.3f ms over %d launches\n" % (src, tot, len(rows)))
        f.write("%-44s %5s %10s %7s\n" % ("kernel", "n", "ms", "share"))
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-44s %5d %10.3f %6.1f%%\n" % (k[:44], n, ms, 100 * ms / tot))
        f.write("\n# 15 longest launches\n")
        for v, k, g, b in sorted(rows, reverse=True)[:15]:
            f.write("%9.3f ms  %-40s grid %s block %s\n" % (v, k[:40], g, b))
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(out.splitlines()))
    hdr, units = rd[0], rd[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none ; source: %s\n" % src)
        for row in rd[2:]:
            f.write("\nkernel: %s\n" % row[hdr.index("Kernel Name")])
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write("  %-72s %s %s\n" % (m, row[i], units[i]))
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
