"""Probe of the streamed self path's pieces on a config-5 shaped sample (NF = 50k): wave staging rate and the fused kernel
per (wave, |q|), against the same atoms staged directly."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sassena_b200
from sassena_b200 import synth

ctx = sassena_b200.ScatterContext(0)
NA, NF, NM = int(os.environ.get("NA", 2048)), int(os.environ.get("NF", 50000)), 200
c = synth.CONFIGS["C5"]
d = ctx.device_alloc(NA * NF * 12)
ctx.synth_trajectory(d, NF, c["NA"], c["box"], c["sigma"], c["seed"], layout=0, NA_out=NA)
pin = ctx.pinned((NF, NA, 3)); frames = pin.array
ctx.memcpy_d2h(frames, d); ctx.device_free(d)
q = 1.0 * synth.unit_vectors(NM, 4)
b = synth.factors(NA)

def timed(f):
    ctx.synchronize(); t0 = time.perf_counter(); r = f(); ctx.synchronize(); return time.perf_counter() - t0, r

for it in range(2):
    dt, _ = timed(lambda: ctx.stage_atoms_wave(frames, 0, 1, NA))
    print(f"stage_atoms_wave {NA} atoms x {NF} frames: {dt*1e3:.1f} ms = {frames.nbytes/dt/1e9:.1f} GB/s")
ctx.set_factors(b)
plen = ctx.partial_len("autocorrelate")
dp = ctx.device_alloc(plen * 8)
for it in range(3):
    dt, _ = timed(lambda: ctx.compute_self_vectors_partial(q, dp))
    print(f"wave-staged: compute_self_vectors_partial {NA*NM} timelines: {dt*1e3:.1f} ms (kernel {ctx.last_amplitude_ms():.1f} ms) -> {NA*NM/dt:.3e} timelines/s")
# the same atoms, atom-major on the host
xa = np.ascontiguousarray(frames.transpose(1, 0, 2))
dt, _ = timed(lambda: ctx.stage_atoms(xa))
print(f"stage_atoms (atom-major pageable host): {dt*1e3:.1f} ms")
ctx.set_factors(b)
for it in range(2):
    dt, _ = timed(lambda: ctx.compute_self_vectors_partial(q, dp))
    print(f"direct-staged: {dt*1e3:.1f} ms -> {NA*NM/dt:.3e} timelines/s")
