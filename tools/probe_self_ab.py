"""Split self path (v2) against the fused residue kernel as the independent check; owned buffer (decimated order).
("v1" selected the first design of kernel A while both existed: profiles/r02_self_split_v3_ncu_summary.txt.)
Usage: python tools/probe_self_ab.py [NF] [NA] [variants...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sassena_b200
from sassena_b200 import synth

NF = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
NA = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
variants = sys.argv[3:] or ["fused", "v2"]
NM = 200
q = 1.0 * synth.unit_vectors(NM, 4)
out = {}
c0 = sassena_b200.ScatterContext(0)
d = c0.device_alloc(NA * NF * 12)
c0.synth_trajectory(d, NF, 30000, 70.0, 0.05, 3, layout=1, NA_out=NA)
xa = np.empty((NA, NF, 3), dtype=np.float32); c0.memcpy_d2h(xa, d); c0.device_free(d); c0.close()
for v in variants:
    os.environ["SASSENA_SELF_PATH"] = "fused" if v == "fused" else "split"
    if v == "v1":
        raise SystemExit("v1 (the first design of kernel A) was removed; see profiles/r02_self_split_v3_ncu_summary.txt")
    ctx = sassena_b200.ScatterContext(0)
    ctx.stage_atoms(xa)
    ctx.set_factors(synth.factors(NA))
    for it in range(3):
        ctx.synchronize(); t0 = time.time()
        r = ctx.compute_self_vectors(q)
        dt = time.time() - t0
    out[v] = r
    tl = NA * NM
    print(f"{v}: {NA} atoms x {NF} frames x {NM} q: {dt*1e3:.1f} ms (kernels {ctx.last_amplitude_ms():.1f} ms) -> {tl/dt:.3e} timelines/s, {tl*NF/dt:.3e} evals/s", flush=True)
    ctx.close()
ref = out[variants[0]]
for v in variants[1:]:
    e = np.max(np.abs(out[v][0] - ref[0])) / np.max(np.abs(ref[0]))
    print(f"{v} vs {variants[0]} rel.err fqt: {e:.3e}  fq: {abs(out[v][1]-ref[1])/abs(ref[1]):.3e}")
