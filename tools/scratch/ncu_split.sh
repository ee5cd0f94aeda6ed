export SASSENA_SELF_PATH=split
cat > tools/scratch/p_split.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import sassena_b200
from sassena_b200 import synth
NF, NA, NM = int(sys.argv[1]), int(sys.argv[2]), 200
ctx = sassena_b200.ScatterContext(0)
d = ctx.device_alloc(NA * NF * 12)
ctx.synth_trajectory(d, NF, 30000, 70.0, 0.05, 3, layout=1, NA_out=NA)
h = __import__("numpy").empty((NA, NF, 3), dtype="float32"); ctx.memcpy_d2h(h, d); ctx.stage_atoms(h)
ctx.set_factors(synth.factors(NA))
q = 1.0 * synth.unit_vectors(NM, 4)
ctx.compute_self_vectors(q)
PY
for shape in "10000 1024" "50000 256"; do
  ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,sm__cycles_elapsed.max,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -c 40 --csv --log-file gpurun_out/split_$(echo $shape | tr ' ' '_').csv python tools/scratch/p_split.py $shape > /dev/null 2>&1
done
python - <<'PY'
import csv, glob
for f in sorted(glob.glob("gpurun_out/split_*.csv")):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]; ik = hdr.index("Kernel Name"); im = hdr.index("Metric Name"); iv = hdr.index("Metric Value"); iid = hdr.index("ID")
    d = {}
    for r in rows[1:]:
        d.setdefault((r[iid], r[ik]), {})[r[im]] = r[iv]
    print(f)
    for (i, k), m in d.items():
        if "self_" in k or "sf_" in k:
            print("  %-40s %s" % (k[:40], {a: b for a, b in m.items()}))
PY
