import csv, glob, collections, sys
for f in sorted(glob.glob(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/split_*.csv")):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]; ik = hdr.index("Kernel Name"); im = hdr.index("Metric Name"); iv = hdr.index("Metric Value"); iid = hdr.index("ID")
    d = {}
    for r in rows[1:]:
        d.setdefault((r[iid], r[ik]), {})[r[im]] = float(r[iv].replace(",", ""))
    agg = collections.defaultdict(lambda: collections.defaultdict(float)); cnt = collections.Counter()
    for (i, k), m in d.items():
        kk = k.split("::")[-1][:34]
        cnt[kk] += 1
        for a, b in m.items(): agg[kk][a] += b
    print(f)
    for k, m in agg.items():
        n = cnt[k]
        if m["gpu__time_duration.sum"] / n < 50e3: continue
        print("  %-36s n=%3d  %9.1f us  fp64 util %.2f  dram R %.0f MB W %.0f MB  L2 %.0f MB" % (k, n, m["gpu__time_duration.sum"] / n / 1e3, m["sm__inst_executed_pipe_fp64.sum"] / (148 * 2 * m["sm__cycles_elapsed.max"]), m["dram__bytes_read.sum"] / n / 1e6, m["dram__bytes_write.sum"] / n / 1e6, m["lts__t_bytes.sum"] / n / 1e6))
