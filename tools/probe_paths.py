"""Device probe of the self / multipole paths and the H2D staging rate on config-shaped slices (not the bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sassena_b200
from sassena_b200 import synth

ctx = sassena_b200.ScatterContext(0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"

if which in ("all", "self"):
    NA, NF, NM = int(os.environ.get("SELF_NA", 512)), int(os.environ.get("SELF_NF", 10000)), 200
    d = ctx.device_alloc(NA * NF * 12)
    ctx.synth_trajectory(d, NF, 30000, 70.0, 0.05, 3, layout=1, NA_out=NA)
    ctx.stage_atoms_device(d, NA, NF)
    ctx.set_factors(synth.factors(NA))
    q = 1.0 * synth.unit_vectors(NM, 4)
    for it in range(2):
        ctx.synchronize(); t0 = time.time()
        fqt, fq, fq2 = ctx.compute_self_vectors(q)
        dt = time.time() - t0
    tl = NA * NM
    print(f"self: {NA} atoms x {NF} frames x {NM} q: {dt*1e3:.1f} ms -> {tl/dt:.3e} timelines/s, {tl*NF/dt:.3e} evals/s; "
          f"cfg2 (1.2e8 timelines) would take {1.2e8/(tl/dt):.1f} s for all 20 |q|")
    ctx.device_free(d)

if which in ("all", "mp", "mpbatch"):
    NA, NF, L = 1000000, int(os.environ.get("MP_NF", 16)), 20
    d = ctx.device_alloc(NA * NF * 12)
    ctx.synth_trajectory(d, NF, NA, 220.0, 0.05, 7, offset=-110.0)
    ctx.stage_frames_device(d, NF, NA)
    # adopt + convert: need an owned buffer for in-place conversion -> copy through host for the probe
    h = np.empty((NF, NA, 3), dtype=np.float32); ctx.memcpy_d2h(h, d); ctx.device_free(d)
    ctx.stage_frames(h); ctx.frames_to_spherical()
    ctx.set_factors(synth.factors(NA))
    mom = np.array([(0, 0)] + [(l, m) for l in range(1, L + 1) for m in range(-l, l + 1)])
    for ql in (0.01, 0.25, 0.5):
        for it in range(2):
            ctx.synchronize(); t0 = time.time()
            ctx.compute_mpsphere(ql, mom, dsp="square")
            dt = time.time() - t0
        print(f"mpsphere |q|={ql}: {NA} atoms x {NF} frames x {len(mom)} moments: {dt*1e3:.1f} ms (amp {ctx.last_amplitude_ms():.1f} ms) -> "
              f"{NA*NF*len(mom)/dt:.3e} moment-evals/s; cfg4 (1000 frames x 200 |q|) would take {dt/NF*1000*200:.0f} s")

    nq8 = int(os.environ.get("MP_NQ", 8))
    ql8 = np.linspace(0.01, 0.5, nq8)
    for it in range(2):
        ctx.synchronize(); t0 = time.time()
        ctx.compute_mpsphere_batch(ql8, mom, dsp="square")
        dt = time.time() - t0
    print(f"mpsphere batch of {nq8} |q|: {dt*1e3:.1f} ms -> {nq8*NA*NF*len(mom)/dt:.3e} moment-evals/s; cfg4 would take {dt/NF*1000*200/nq8:.0f} s")

if which in ("all", "mpcyl"):
    # multipole cylinder (K6) on a config-4 shaped slice: 1M atoms, moments (0,0) + (l, 0..3) for l <= L
    NA, NF, L = 1000000, int(os.environ.get("MP_NF", 16)), int(os.environ.get("CYL_L", 10))
    d = ctx.device_alloc(NA * NF * 12)
    ctx.synth_trajectory(d, NF, NA, 220.0, 0.05, 7, offset=-110.0)
    h = np.empty((NF, NA, 3), dtype=np.float32); ctx.memcpy_d2h(h, d); ctx.device_free(d)
    axis = (0.0, 0.0, 1.0)
    ctx.stage_frames(h); ctx.frames_to_cylindrical(axis)
    ctx.set_factors(synth.factors(NA))
    mom = np.array([(0, 0)] + [(l, m) for l in range(1, L + 1) for m in range(4)])
    for q in ((0.05, 0.02, 0.03), (0.3, -0.2, 0.25)):
        for it in range(2):
            ctx.synchronize(); t0 = time.time()
            ctx.compute_mpcylinder(q, axis, mom, dsp="square")
            dt = time.time() - t0
        print(f"mpcylinder q={q}: {NA} atoms x {NF} frames x {len(mom)} moments (orders <= {2*L}): {dt*1e3:.1f} ms (amp {ctx.last_amplitude_ms():.1f} ms) -> "
              f"{NA*NF*len(mom)/dt:.3e} moment-evals/s, {NA*NF/dt:.3e} atom-frames/s")

if which in ("all", "h2d"):
    n = 1 << 30
    host = ctx.pinned((n,), np.uint8)
    d = ctx.device_alloc(n)
    for it in range(3):
        t0 = time.time(); ctx.memcpy_h2d(d, host.array); dt = time.time() - t0
    print(f"H2D pinned 1 GiB: {n/dt/1e9:.1f} GB/s")
    pag = np.zeros(n, dtype=np.uint8)
    t0 = time.time(); ctx.memcpy_h2d(d, pag); dt = time.time() - t0
    print(f"H2D pageable 1 GiB: {n/dt/1e9:.1f} GB/s")
