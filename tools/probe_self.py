"""Fused vs split self kernels on config-2 / config-5 shaped slices, owned buffer (so that the split path may keep the
frames in its decimated order).  Usage: python tools/probe_self.py [NF] [NA]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sassena_b200
from sassena_b200 import synth

NF = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
NA = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
NM = 200
q = 1.0 * synth.unit_vectors(NM, 4)
out = {}
c0 = sassena_b200.ScatterContext(0)
d = c0.device_alloc(NA * NF * 12)
c0.synth_trajectory(d, NF, 30000, 70.0, 0.05, 3, layout=1, NA_out=NA)
xa = np.empty((NA, NF, 3), dtype=np.float32); c0.memcpy_d2h(xa, d); c0.device_free(d); c0.close()
for path in ("fused", "split"):
    os.environ["SASSENA_SELF_PATH"] = path
    ctx = sassena_b200.ScatterContext(0)
    ctx.stage_atoms(xa)
    ctx.set_factors(synth.factors(NA))
    for it in range(3):
        ctx.synchronize(); t0 = time.time()
        r = ctx.compute_self_vectors(q)
        dt = time.time() - t0
    out[path] = r
    tl = NA * NM
    print(f"{path}: {NA} atoms x {NF} frames x {NM} q: {dt*1e3:.1f} ms (kernels {ctx.last_amplitude_ms():.1f} ms) -> {tl/dt:.3e} timelines/s, {tl*NF/dt:.3e} evals/s")
    ctx.close()
e = np.max(np.abs(out["split"][0] - out["fused"][0])) / np.max(np.abs(out["fused"][0]))
print("split vs fused rel.err fqt:", e)
