"""Warp-stall samples of an `ncu --set full --import-source on` capture aggregated per CUDA source line (top lines) and, for
conv_raw_kernel, per warp role (line ranges of csrc/conv_raw.cu).  The .ncu-rep stays in the scratch gpurun_out/; the text
this prints is what goes under profiles/.

    python tools/ncu_stalls.py gpurun_out/r02_v3_raw_l48.ncu-rep [top_n] [extra ncu import filters, e.g. --launch-count 1]
"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
extra = sys.argv[3:]
out = subprocess.run(["ncu", "-i", rep, *extra, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur_file = cur_fn = hdr = None
agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), ""])
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall = [i for i, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
        continue
    if hdr is None or r[0] == "":
        continue
    off = len(r) - len(hdr)                    # source text with quotes / commas widens the row
    a = agg[(cur_fn, cur_file, int(r[0]))]
    a[0] += int(r[iS + off])
    a[1] += int(r[iI + off])
    a[3] = r[1].strip()
    for i in stall:
        v = r[i + off]
        if v not in ("", "-", "0"):
            a[2][hdr[i]] += int(v)
fns = sorted({k[0] for k in agg})
for fn in fns:
    items = [(k, v) for k, v in agg.items() if k[0] == fn]
    tot = sum(v[0] for _, v in items)
    if tot == 0:
        continue
    print("kernel: %s\ntotal warp-stall samples %d" % (fn, tot))
    for k, v in sorted(items, key=lambda kv: -kv[1][0])[:topn]:
        print("%6d %5.1f%%  %s:%d  %-72s %s" % (v[0], 100 * v[0] / tot, k[1], k[2], v[3][:72], dict(v[2].most_common(2))))
    if "conv_raw_kernel" in fn:
        src = open("eabnet_b200/csrc/conv_raw.cu").read().splitlines()
        marks = [(i + 1, m.group(1)) for i, l in enumerate(src) for m in [re.search(r"={20,} (raw loader|MMA issuer|weight loader|epilogue|transform warps)", l)] if m]
        first_xf = next(i + 1 for i, l in enumerate(src) if "auto xf_tile" in l)

        def role(f, ln):
            if f == "umma.cuh":
                return "mbarrier waits / tcgen05 wrappers (all roles)"
            if f != "conv_raw.cu":
                return "inlined headers (" + f + ")"
            if ln < first_xf and ln >= 85 and ln < 240:
                return "transform helpers (xf4 / xf_items)" if ln < 209 else "epilogue helpers"
            if first_xf <= ln < marks[0][0]:
                return "transform (xf_tile)"
            name = "prologue / epilogue of the kernel"
            for start, nm in marks:
                if ln >= start:
                    name = nm
            return name
        reg = collections.Counter()
        for k, v in items:
            reg[role(k[1], k[2])] += v[0]
        print("  by role:")
        for nm, n in reg.most_common():
            print("  %6d %5.1f%%  %s" % (n, 100 * n / tot, nm))
    print()
