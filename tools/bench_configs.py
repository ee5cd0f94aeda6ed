"""Measures the BASELINE.json configurations that are NOT the bench.py headline (C1 full; C2, C4 on stated sub-samples with
linear extrapolation, labelled as such) and the CPU oracle on small samples of the same shapes.  One JSON line per config.
C5 (self, trajectory larger than the coordinate budget) runs through the host layer's streamed stager on a sample of atoms.
Usage: python tools/bench_configs.py [C1 C2 C4 C5]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sassena_b200
from sassena_b200 import synth
from oracle import oracle as o

which = sys.argv[1:] or ["C1", "C2", "C4", "C5"]
ctx = sassena_b200.ScatterContext(0)
cores = o.max_threads()


def best(f, n=3):
    b = 1e9
    for _ in range(n):
        ctx.synchronize(); t0 = time.perf_counter(); r = f(); ctx.synchronize(); b = min(b, time.perf_counter() - t0)
    return b, r


if "C1" in which:
    c = synth.CONFIGS["C1"]
    xyz = synth.trajectory(c["NF"], c["NA"], c["box"], c["sigma"], c["seed"])
    b = synth.factors(c["NA"]); u = synth.unit_vectors(c["NM"], c["vseed"]); qls = synth.qlengths(*c["q"])
    ctx.stage_frames(xyz); ctx.set_factors(b)
    def run():
        return [ctx.compute_all_vectors(ql * u) for ql in qls]
    dt, res = best(run)
    t0 = time.perf_counter(); ref = [o.compute_all_vectors(xyz, b, ql * u, nthreads=4) for ql in qls]; cpu = time.perf_counter() - t0
    err = max(float(np.max(np.abs(r[0] - rr[0])) / np.max(np.abs(rr[0]))) for r, rr in zip(res, ref))
    ev = c["NA"] * c["NF"] * c["NM"] * len(qls)
    print(json.dumps({"config": "C1 coherent 1k atoms x 100 frames x 10|q| x 100 vectors (full)", "gpu_s": dt, "gpu_evals_per_s": ev / dt,
                      "cpu_s": cpu, "cpu_threads": 4, "cpu_evals_per_s": ev / cpu, "max_rel_err_fqt": err}))

if "C2" in which:
    c = synth.CONFIGS["C2"]
    NA_s, NF, NM = 2048, c["NF"], c["NM"]
    d = ctx.device_alloc(NA_s * NF * 12)
    ctx.synth_trajectory(d, NF, c["NA"], c["box"], c["sigma"], c["seed"], layout=1, NA_out=NA_s)
    ctx.stage_atoms_device(d, NA_s, NF); ctx.set_factors(synth.factors(c["NA"])[:NA_s])
    u = synth.unit_vectors(NM, c["vseed"]); ql = synth.qlengths(*c["q"])[10]
    dt, res = best(lambda: ctx.compute_self_vectors(ql * u), 2)
    full = dt * c["NA"] / NA_s * c["q"][2]
    # CPU oracle on 8 atoms x 16 vectors of the same trajectory
    na_c, nm_c = 2 * cores // 2, 16
    xa = np.empty((na_c, NF, 3), dtype=np.float32); ctx.memcpy_d2h(xa, d)
    t0 = time.perf_counter(); ref = o.compute_self_vectors(xa, synth.factors(c["NA"])[:na_c], ql * u[:nm_c], nthreads=cores); cpu = time.perf_counter() - t0
    ctx.stage_atoms(xa); ctx.set_factors(synth.factors(c["NA"])[:na_c]); got = ctx.compute_self_vectors(ql * u[:nm_c])
    err = float(np.max(np.abs(got[0] - ref[0])) / np.max(np.abs(ref[0])))
    ctx.device_free(d)
    print(json.dumps({"config": "C2 self 30k atoms x 10k frames x 20|q| x 200 vectors", "sample": f"{NA_s} of 30000 atoms, one |q| (extrapolated linearly)",
                      "gpu_s_sample": dt, "gpu_timelines_per_s": NA_s * NM / dt, "gpu_evals_per_s": NA_s * NM * NF / dt, "gpu_s_full_extrapolated": full,
                      "cpu_sample": f"{na_c} atoms x {nm_c} vectors, {cores} threads", "cpu_s_sample": cpu, "cpu_timelines_per_s": na_c * nm_c / cpu,
                      "cpu_s_full_extrapolated": cpu * (c["NA"] * NM * c["q"][2]) / (na_c * nm_c), "max_rel_err_fqt": err}))

if "C4" in which:
    c = synth.CONFIGS["C4"]
    NF_s, NA, NQ_s = 296, c["NA"], 8  # two frames per SM; the device batches 8 |q| per pass (Y_lm tables shared)
    d = ctx.device_alloc(NA * NF_s * 12)
    ctx.synth_trajectory(d, NF_s, NA, c["box"], c["sigma"], c["seed"], offset=c["offset"])
    h = np.empty((NF_s, NA, 3), dtype=np.float32); ctx.memcpy_d2h(h, d); ctx.device_free(d)
    ctx.stage_frames(h); ctx.frames_to_spherical(); bf = synth.factors(NA); ctx.set_factors(bf)
    mom = o.moments_sphere(c["L"]); ql = 0.25
    qbatch = np.linspace(0.2, 0.3, NQ_s)
    dt, res = best(lambda: ctx.compute_mpsphere_batch(qbatch, mom, dsp="square"), 2)
    full = dt / (NF_s * NQ_s) * c["NF"] * c["q"][2]
    # CPU oracle: 2000 atoms x 2 frames x all moments
    na_c = 2000
    sph = o.cart_to_spherical(h[:2, :na_c])
    t0 = time.perf_counter(); ref = o.compute_mpsphere(sph, bf[:na_c], ql, mom, dsp="square", nthreads=cores); cpu = time.perf_counter() - t0
    ctx.stage_frames(sph, repr=1); ctx.set_factors(bf[:na_c]); got = ctx.compute_mpsphere(ql, mom, dsp="square")
    err = float(np.max(np.abs(got[0] - ref[0])) / np.max(np.abs(ref[0])))
    me = NA * NF_s * NQ_s * len(mom)
    print(json.dumps({"config": "C4 multipole sphere 1M atoms x 1k frames x 200|q| x 441 moments", "sample": f"{NF_s} of 1000 frames, one batch of {NQ_s} |q| (extrapolated linearly)",
                      "gpu_s_sample": dt, "gpu_moment_evals_per_s": me / dt, "gpu_s_full_extrapolated": full,
                      "cpu_sample": f"{na_c} atoms x 2 frames x {len(mom)} moments, {cores} threads", "cpu_s_sample": cpu,
                      "cpu_moment_evals_per_s": na_c * 2 * len(mom) / cpu, "cpu_s_full_extrapolated": cpu * (NA * c["NF"] * c["q"][2]) / (na_c * 2),
                      "max_rel_err_fqt": err}))

if "C5" in which:
    # streamed self scattering: the sample's coordinates exceed limits.stage.memory.data, so the host layer stages the
    # atoms in waves (frame-major pinned host buffer -> chunked async H2D -> GPU transpose) and evaluates every wave for
    # all |q|.  Sample: NA_s atoms x all 50k frames x 2 |q| x 200 vectors, budget = a third of the sample -> 3 waves.
    from sassena_b200 import host
    c = synth.CONFIGS["C5"]
    NA_s, NF, NM, NQ_s = 6144, c["NF"], c["NM"], 2
    d = ctx.device_alloc(NA_s * NF * 12)
    ctx.synth_trajectory(d, NF, c["NA"], c["box"], c["sigma"], c["seed"], layout=0, NA_out=NA_s)  # frame-major [NF][NA_s][3]
    pin = ctx.pinned((NF, NA_s, 3)); frames = pin.array  # the stager's pinned host buffer
    ctx.memcpy_d2h(frames, d); ctx.device_free(d)
    bf = synth.factors(c["NA"])[:NA_s]
    qls = synth.qlengths(*c["q"])[9:9 + NQ_s]
    qv = np.array([[ql, 0.0, 0.0] for ql in qls])
    p = host.Params().set("scattering.type", "self").set("scattering.average.orientation.type", "vectors")
    p.set("scattering.average.orientation.vectors.type", "file").set_vectors(synth.unit_vectors(NM, c["vseed"])).create()
    p.set("limits.stage.memory.data", (NA_s // 3) * NF * 12)
    t0 = time.perf_counter(); recs, _, tm = host.run_scatter(p, frames, qv, b=bf, ctx=ctx); dt = time.perf_counter() - t0
    waves = tm["sd:compute"][1]
    tl = NA_s * NM * NQ_s
    full1 = tm["sd:compute"][0] / tl * (c["NA"] * NM * c["q"][2])
    # CPU oracle: a few atoms x 8 vectors of the same trajectory, one |q|
    na_c, nm_c = max(2, cores // 4), 8
    xa = np.ascontiguousarray(frames[:, :na_c].transpose(1, 0, 2))
    u = p.init_subvectors(qv[0])
    t0 = time.perf_counter(); ref = o.compute_self_vectors(xa, bf[:na_c], u[:nm_c], nthreads=cores); cpu = time.perf_counter() - t0
    ctx.stage_atoms(xa); ctx.set_factors(bf[:na_c]); got = ctx.compute_self_vectors(u[:nm_c])
    err = float(np.max(np.abs(got[0] - ref[0])) / np.max(np.abs(ref[0])))
    print(json.dumps({"config": "C5 self streamed 500k atoms x 50k frames x 20|q| x 200 vectors",
                      "sample": f"{NA_s} of 500000 atoms x all frames x {NQ_s} |q|, coordinate budget = 1/3 of the sample -> {waves} waves through the streamed stager",
                      "gpu_s_sample": dt, "stage_s": tm["sd:stage"][0], "compute_s": tm["sd:compute"][0],
                      "h2d_GBps_staging": frames.nbytes * waves / max(tm["sd:stage"][0], 1e-9) / 1e9,
                      "gpu_timelines_per_s": tl / tm["sd:compute"][0], "gpu_evals_per_s": tl * NF / tm["sd:compute"][0],
                      "gpu_s_full_extrapolated_1gpu": full1, "gpu_s_full_extrapolated_8gpu": full1 / 8,
                      "cpu_sample": f"{na_c} atoms x {nm_c} vectors, {cores} threads", "cpu_s_sample": cpu,
                      "cpu_timelines_per_s": na_c * nm_c / cpu,
                      "cpu_s_full_extrapolated": cpu * (c["NA"] * NM * c["q"][2]) / (na_c * nm_c), "max_rel_err_fqt": err}))
