"""Turn gpurun_out ncu artefacts into the small, tracked summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_v0.csv profiles/r01_v0_launches.txt
    python tools/summarize_ncu.py full gpurun_out/prof_conv_v0.ncu-rep profiles/r01_v0_conv_generic_full.txt
    python tools/summarize_ncu.py traffic gpurun_out/traffic.csv profiles/r02_traffic.json
        (traffic.csv: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv
         over ONE step of tools/prof_step.py --steps 1: DRAM bytes of every launch, summed per kernel family)
    python tools/summarize_ncu.py sass eabnet_b200/libeabnet_b200.so profiles/r02_sass_opcodes.txt
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


FAMILY = [("stft_stage_kernel", "stft"), ("conv_raw_kernel", "conv2d"), ("conv_tma_kernel", "conv2d"), ("stage_kernel", "conv2d"), ("conv_umma_kernel", "conv2d"),
          ("conv_generic_kernel", "conv2d"), ("combine_kernel", "conv2d"), ("tcm_chain_kernel", "tcm"), ("lstm_umma_kernel", "head"),
          ("lstm_kernel", "head"), ("head_fused_kernel", "head"), ("beam_", "head"), ("istft_kernel", "istft"), ("stft_kernel", "stft")]


def traffic(src, dst):
    """per-family DRAM bytes of one step from a per-launch ncu CSV (see the module docstring)"""
    import json
    lines = [l for l in open(src) if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = per.setdefault(row["ID"], {"name": row["Kernel Name"]})
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        if row["Metric Name"].startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        else:
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1.0)
        d[row["Metric Name"]] = v
    fam = collections.OrderedDict()
    stft_left = 0
    for d in per.values():
        name = d["name"]
        f = next((fm for key, fm in FAMILY if key in name), None)
        if f is None:                            # not one of the library's kernels (torch's own launches)
            continue
        if "stft_stage_kernel" in name:
            stft_left = 3                       # the three DFT-GEMM launches of the STFT follow its stage kernel
        elif "conv_tma_kernel" in name and stft_left > 0:
            f, stft_left = "stft", stft_left - 1
        a = fam.setdefault(f, {"dram_bytes_per_step": 0.0, "read": 0.0, "write": 0.0, "ms_under_ncu": 0.0, "launches": 0})
        a["read"] += d.get("dram__bytes_read.sum", 0.0)
        a["write"] += d.get("dram__bytes_write.sum", 0.0)
        a["dram_bytes_per_step"] = a["read"] + a["write"]
        a["ms_under_ncu"] += d.get("gpu__time_duration.sum", 0.0)
        a["launches"] += 1
    out = {"_source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none over one "
                      "64 x 6 s step (%s); per-family sums over every launch of the step" % src}
    for f, a in fam.items():
        a["note"] = "%d launches, %.3f GB read + %.3f GB written, %.3f ms under ncu (cold-cache, serialised)" % (
            a["launches"], a["read"] / 1e9, a["write"] / 1e9, a["ms_under_ncu"])
        out[f] = a
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


def sass(src, dst):
    """tensor-core / TMA opcode counts per kernel of the built library (cuobjdump -sass)"""
    txt = subprocess.run(["cuobjdump", "-sass", src], capture_output=True, text=True).stdout
    ops = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "USETMAXREG", "SYNCS", "FFMA", "MUFU", "STL", "LDL"]
    cur, counts = None, collections.OrderedDict()
    for line in txt.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            counts[cur] = collections.Counter()
        elif cur:
            for o in ops:
                if (" " + o) in line or ("\t" + o) in line:
                    counts[cur][o] += 1
    with open(dst, "w") as f:
        f.write("# cuobjdump -sass %s : opcode counts per kernel (UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk,\n" % src)
        f.write("# UTMALDG/UTMASTG = tensor-map TMA, USETMAXREG = setmaxnreg, SYNCS = mbarrier ops, STL/LDL = local-memory spills)\n")
        f.write("%-90s %s\n" % ("kernel", " ".join("%10s" % o for o in ops)))
        for k, c in counts.items():
            short = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip() or k
            short = short.replace("eab::(anonymous namespace)::", "").split("(")[0]
            f.write("%-90s %s\n" % (short[:90], " ".join("%10d" % c[o] for o in ops)))
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic, "sass": sass}[sys.argv[1]](sys.argv[2], sys.argv[3])
