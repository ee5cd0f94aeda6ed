"""Where does the end-to-end step go?  Times staging alone, compute alone and staged+compute on the config-3 shape."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sassena_b200
from sassena_b200 import synth

NF = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
NA, NM = 100000, 500
ctx = sassena_b200.ScatterContext(0)
d = ctx.device_alloc(NF * NA * 12)
ctx.synth_trajectory(d, NF, NA, 100.0, 0.05, 5)
host = ctx.pinned((NF, NA, 3), np.float32)
ctx.memcpy_d2h(host.array, d)
b = synth.factors(NA)
q = 2.0 * synth.unit_vectors(NM, 6)
plen = None

def t(f, n=3):
    best = 1e9
    for _ in range(n):
        ctx.synchronize(); t0 = time.perf_counter(); f(); ctx.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e3

ctx.stage_frames_device(d, NF, NA); ctx.set_factors(b)
print(f"resident compute: {t(lambda: ctx.compute_all_vectors(q)):.1f} ms")
def stage_only():
    ctx.stage_frames(host.array)
print(f"stage only (H2D {NF*NA*12/1e9:.1f} GB): {t(stage_only):.1f} ms")
def stage_call_only():
    t0 = time.perf_counter(); ctx.stage_frames(host.array); return time.perf_counter() - t0
ctx.synchronize(); print(f"stage_frames host call returns after {stage_call_only()*1e3:.1f} ms"); ctx.synchronize()
def both():
    ctx.stage_frames(host.array); ctx.set_factors(b); ctx.compute_all_vectors(q)
print(f"stage + compute: {t(both):.1f} ms")

def staged_then_compute():
    ctx.stage_frames(host.array); ctx.synchronize(); ctx.set_factors(b)
    t0 = time.perf_counter(); ctx.compute_all_vectors(q); return (time.perf_counter() - t0) * 1e3
print(f"compute after staging finished (chunked launches if SASSENA_FORCE_CHUNKED): {min(staged_then_compute() for _ in range(3)):.1f} ms")
