#!/usr/bin/env python
"""bench.py — headline benchmark of the Sassena scattering hot path on B200.

Metric (BASELINE.json): amplitude evaluations/s, one evaluation = one (atom, frame, q-vector) triple, plus the
F(q,t) wall time it implies.  The headline workload is BASELINE configs[2] "coherent F(q,t): 100k atoms x 10k frames,
50 |q| x 500 sphere vectors" — the configuration north_star's Target sentence is quoted on; it fits one GPU (12 GB of
coordinates).

A STEP is the whole job's hot path: all 50 |q| x 500 orientation vectors over the full trajectory — amplitudes -> FFT
autocorrelation -> orientational average -> fqt/fq/fq2 for every |q| (2.5e13 evaluations).  The |q| list is what the
reference's own scan generator produces (`synth.qlengths` = ScatteringVectorsParameters::create_from_scans,
parameters.cpp:1125-1189: float-rounded fractions), so the library plans the corrected symmetric scan kernel for it
(DESIGN.md "K1s").  The same scan on exactly equally spaced |q| (np.linspace; plain symmetric kernel) is reported beside it
under `workloads.C3_equally_spaced`; `--mode per-q` times the general kernel instead (one |q| per step).

The default invocation also measures, each with its own roofline / e2e / parity / cpu_baseline under `workloads`:
  C2   BASELINE configs[1], incoherent self scattering 30k atoms x 10k frames, at full size: a step is one |q| with its 200
       vectors over all atoms (6e6 per-atom timelines: amplitudes + FFT autocorrelation in the self kernels);
  C4   BASELINE configs[3], multipole sphere averaging 1M atoms x 1k frames, at full size: a step is one pass of the batched
       multipole kernel (8 |q| x 441 moments over all atoms and frames);
  C5s  a bounded sample of BASELINE configs[4] (stager-streamed self scattering, 50k frames): C5S_ATOMS atoms stream from
       pinned host memory through the double-buffered wave stager while the previous wave is evaluated.
`--workload C2|C4|C5s|C3|C1` runs one of them alone as the top-level line; `--workload C5` runs config 5 at full size
(500k atoms x 50k frames, `--nq` |q| values; needs 8 GPUs and ~300 GB of host memory).

N GPUs: one process per GPU (torchrun), total work fixed ("strong").  Coherent: the FRAMES are sharded with DivAssignment (the
reference's own decomposition, all_vectors_scatter_device.cpp:61,248,408), the per-rank amplitude blocks are exchanged over
NVLink so that every rank correlates a DivAssignment block of the timelines, and the packed partials are summed with one
small all-reduce (`--shard vectors`: replicated coordinates, sharded subvectors, no amplitude exchange).  Self: atoms sharded
by ModAssignment, one all-reduce of the packed partial (self_vectors_scatter_device.cpp:50,213-221).  Multipole: atoms sharded
by DivAssignment, amplitudes all-reduced before the DSP.

`--impl reference` times the reference's CPU implementation on ALL host cores (len(os.sched_getaffinity(0)), never
OMP_NUM_THREADS — torchrun sets that to 1) on a bounded sample of the same workload: for the coherent workload the reference's
OWN AllVectorsScatterDevice (oracle/_ref: its sources compiled where they lie over shim headers, since its build system
cannot be used in this image); for the self and multipole workloads the oracle port, which reproduces the reference's devices
bit for bit (their oracle/_ref builds run over the oracle's DFT / special functions instead of FFTW / Boost.Math, so timing
them would not be timing the reference).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FLOP_PER_EVAL = 45.0        # algorithmic FP64 flop per amplitude evaluation
FP64_INSTR_PER_EVAL = 21.0  # what the kernel executes (sincos_qt.cuh)
# |q|-scan kernel (DESIGN.md "K1s"): per (atom, direction, pass) a setup of one dot product, two phase scalings, two
# sincos, the b scaling and one complex rotation; per |q| of the pass one 3-term recurrence step and one accumulation on
# both components
SCAN_SETUP_FLOP, SCAN_STEP_FLOP = 65.0, 6.0    # algorithmic FP64 flop
SCAN_SETUP_INSTR, SCAN_STEP_INSTR = 49.0, 4.0  # executed FP64-pipe instructions
SCAN_MAX_PASS, SCAN_MAX_PASS_CORR = 28, 16     # amplitude.cu amplitude_scan_max_pass defaults


def scan_passes(nq, max_b=SCAN_MAX_PASS):
    """pass sizes (kernel template B) sgpu_capi.cu plan_scan uses for nq |q| values: even shares, multiples of 4"""
    npass = (nq + max_b - 1) // max_b
    out, n0 = [], 0
    for p in range(npass):
        want = (nq - n0 + (npass - p) - 1) // (npass - p)
        L = min(min((want + 3) // 4 * 4, max_b), nq - n0)
        out.append((L + 3) // 4 * 4)
        n0 += L
    return out

WORKLOADS = {
    # name: (config key in sassena_b200.synth.CONFIGS, description)
    "C3": "coherent F(q,t): 100k atoms x 10k frames, 50 |q| x 500 sphere vectors",
    "C1": "synthetic 1k-atom box, 100 frames, 10 |q| x 100 sphere vectors, coherent",
    "C2": "incoherent self F_s(q,t): 30k atoms x 10k frames, 20 |q| x 200 vectors, per-atom FFT correlation (one |q| per step)",
    "C4": "static SAXS via multipole sphere averaging (MPSphereScatterDevice): 1M atoms, 1k frames, 200 |q|, moments l <= 20 "
          "(one batch of 8 |q| per step)",
}


def div_assignment(NN, rank, N):
    """DivAssignment (reference src/decomposition/assignment.cpp:27-35)."""
    first = (rank * N) // NN
    nxt = ((rank + 1) * N) // NN
    return first, nxt - first


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.p = None
        self.path = f"/tmp/bench_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                       "100", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        try:
            self.p.terminate()
            self.p.wait(timeout=5)
            self.f.close()
            sm, smax, power = [], [], []
            reasons = set()
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1]))
                    smax.append(float(c[2]))
                    power.append(float(c[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            if sm:
                out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                       "samples": len(sm), "power_w_max": max(power)}
            os.unlink(self.path)
        except Exception:
            pass
        return out


def cpu_sample_shape(cfg, cores, target_core_seconds, evals_per_core_s=2.4e7):
    """bounded CPU sample of the workload: all atoms, first NF_s frames, NM_s subvectors of one |q|"""
    NA = cfg["NA"]
    NM_s = min(cfg["NM"], 4 * cores)
    NM_s = max(cores, (NM_s // cores) * cores)
    evals = evals_per_core_s * cores * target_core_seconds
    NF_s = int(max(16, min(cfg["NF"], evals / (NA * NM_s))))
    return NF_s, NM_s


def cpu_reference_kind():
    """"reference": oracle/_ref/libsmath_ref.so (the reference's OWN scatter devices, stagers and DSP built for one rank over the
    shims in oracle/shim; it travels with the repository snapshot) is there; "port": only the oracle restatement is."""
    from oracle import oracle as o
    return "reference" if (o.have_ref_smath() and not _REF_BUILD_FAILED) else "port"


_REF_BUILD_FAILED = False  # set when oracle/_ref is present but could not be loaded / run on this box: the port is timed, and says so


def run_cpu_oracle(cfg, NF_s, NM_s, ql, threads, coords=None):
    """coherent CPU leg on the bounded sample: the reference's own AllVectorsScatterDevice with `threads` worker threads
    (limits.computation.threads) when its build is present -- timed by the reference's own "sd:runner" timer, i.e. compute + write
    without staging -- else the oracle port with `threads` OpenMP threads"""
    from oracle import oracle as o
    from sassena_b200 import synth
    o.build()
    if coords is None:
        coords = synth.trajectory(NF_s, cfg["NA"], cfg["box"], cfg["sigma"], cfg["seed"])
    b = synth.factors(cfg["NA"])
    u = synth.unit_vectors(cfg["NM"], cfg["vseed"])[:NM_s]
    if cpu_reference_kind() == "reference":
        try:
            _, fqt, fq, fq2 = o.ref_scatter_run("all", coords, b, [[ql, 0.0, 0.0]], orient=u, vectors_type="file", threads=threads)
            return o.ref_timer_seconds("sd:runner"), (fqt[0], fq[0], fq2[0]), coords
        except (OSError, AttributeError, RuntimeError) as e:  # e.g. a stale .so from another toolchain
            global _REF_BUILD_FAILED
            _REF_BUILD_FAILED = True
            print(f"bench: oracle/_ref present but unusable ({e!r}); timing the oracle port instead", file=sys.stderr)
    q = ql * u
    t0 = time.perf_counter()
    res = o.compute_all_vectors(coords, b, q, nthreads=threads)
    dt = time.perf_counter() - t0
    return dt, res, coords


def run_cpu_self(xa, b, q_unit, ql, threads):
    """self CPU leg: the oracle port with `threads` OpenMP threads over atoms -- the reference's rank parallelism (ModAssignment
    over MPI ranks).  The reference's own SelfVectorsScatterDevice builds too (oracle/_ref) and the port reproduces it bit for
    bit (tests/test_reference_devices.py), but inside one rank it runs the per-timeline FFT autocorrelation serially on the main
    thread, and in that build the FFT behind FFTW's API is the oracle's DFT: timing it would not be the reference with FFTW.
    xa float32 [NA_s][NF][3] atom-major"""
    from oracle import oracle as o
    t0 = time.perf_counter()
    res = o.compute_self_vectors(xa, b, ql * q_unit, nthreads=threads)
    return time.perf_counter() - t0, res


def cpu_sample_note(threads, kind=None):
    if (kind or cpu_reference_kind()) == "reference":
        return (f"the reference's own scatter device (oracle/_ref build, one rank, limits.computation.threads = {threads} worker "
                "threads; its sd:runner timer: compute + write, staging excluded)")
    return f"oracle port, {threads} OpenMP threads"


def run_reference(args):
    """--impl reference: the reference algorithm on the host cores (oracle port; rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from oracle import oracle as o
    from sassena_b200 import synth
    cfg = dict(synth.CONFIGS[args.workload])
    if args.frames:
        cfg["NF"] = args.frames
    if args.atoms:
        cfg["NA"] = args.atoms
    cores = o.max_threads()
    NF_s, NM_s = cpu_sample_shape(cfg, cores, args.cpu_seconds)
    qls = synth.qlengths(*cfg["q"])
    coords = synth.trajectory(NF_s, cfg["NA"], cfg["box"], cfg["sigma"], cfg["seed"])
    times = []
    for i in range(args.warmup + args.steps):
        dt, _, _ = run_cpu_oracle(cfg, NF_s, NM_s, qls[i % len(qls)], cores, coords)
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    evals = float(cfg["NA"]) * NF_s * NM_s * len(times)
    value = evals / total
    sample = (f"{cfg['NA']} atoms x first {NF_s} frames x {NM_s} of {cfg['NM']} subvectors of one |q| per step "
              f"(amplitudes + FFT autocorrelation + store); " + cpu_sample_note(cores))
    line = {
        "impl": "reference", "metric": "amplitude evals/s (atom*frame*q-vector)", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "NA": cfg["NA"], "NF": cfg["NF"], "NM": cfg["NM"],
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": cpu_reference_kind(), "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    # exactly ONE line may reach stdout (the JSON line): NCCL / torchrun banners go to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        from sassena_b200 import synth
        if synth.CONFIGS[args.workload]["kind"] == "self":
            return _run_self(args, saved_stdout)
        if synth.CONFIGS[args.workload]["kind"] == "mpsphere":
            return _run_mpsphere(args, saved_stdout)
        return _run_ours(args, saved_stdout)
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)


def _run_ours(args, json_fd):
    import torch
    import torch.distributed as dist
    import sassena_b200
    from sassena_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch multi-GPU runs with torchrun (one process per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = dict(synth.CONFIGS[args.workload])
    if args.frames:
        cfg["NF"] = args.frames
    if args.atoms:
        cfg["NA"] = args.atoms
    NA, NF, NM = cfg["NA"], cfg["NF"], cfg["NM"]
    scan = args.mode in ("scan", "scan-rounded")
    if args.mode == "scan":
        qls = np.linspace(cfg["q"][0], cfg["q"][1], cfg["q"][2])  # equally spaced |q|: plain scan kernel
    else:
        # the reference's scan generator: float-rounded fractions (parameters.cpp:1151), equally spaced to ~1e-8 only;
        # in scan-rounded mode these take the corrected scan kernel
        qls = synth.qlengths(*cfg["q"])
    NQ = len(qls) if scan else 1          # |q| values per step
    b = synth.factors(NA)
    u = synth.unit_vectors(NM, cfg["vseed"])
    m_off, m_cnt = div_assignment(world, rank, NM)
    f_off, f_cnt = div_assignment(world, rank, NF)
    by_frames = world > 1 and args.shard == "frames"

    ctx = sassena_b200.ScatterContext(local_rank)
    fp64_peak = ctx.measure_fp64_peak()

    # synthetic trajectory generated on the device (CPU twin: sassena_b200/synth.py), resident in HBM
    xyz = torch.empty(NF * NA * 3, dtype=torch.float32, device=dev)
    ctx.synth_trajectory(xyz.data_ptr(), NF, NA, cfg["box"], cfg["sigma"], cfg["seed"])

    def stage_resident():
        if by_frames:  # this rank's block of the timeline
            ctx.stage_frames_device(xyz.data_ptr() + f_off * NA * 12, f_cnt, NA)
            ctx.set_frame_window(NF, f_off)
        else:
            ctx.stage_frames_device(xyz.data_ptr(), NF, NA)
        ctx.set_factors(b)

    stage_resident()
    amp = torch.zeros(NQ * NM * NF * 2, dtype=torch.float64, device=dev) if by_frames else None
    plen = ctx.partial_len("autocorrelate")
    partial = torch.zeros(NQ * plen, dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    def compute_step(i):
        """one step on `world` GPUs, coordinates already staged: every |q| of the scan (scan mode) or one |q|"""
        q0 = float(qls[i % len(qls)])
        if world == 1:
            if scan:
                return ctx.compute_all_vectors_scan(u, qls)
            return ctx.compute_all_vectors(q0 * u)
        if by_frames:
            if scan:
                ctx.all_vectors_scan_amplitudes(u, qls, amp.data_ptr())
            else:
                ctx.all_vectors_amplitudes(q0 * u, amp.data_ptr())
            ctx.synchronize()
            dist.all_reduce(amp)  # exchange over NVSwitch: every rank ends up with the complete timelines
            torch.cuda.synchronize()
            for n in range(NQ):
                ctx.all_vectors_dsp_partial(amp.data_ptr() + n * NM * NF * 16, m_off, m_cnt, partial.data_ptr() + n * plen * 8)
        elif scan:
            ctx.compute_all_vectors_scan_partial(u[m_off:m_off + m_cnt], qls, partial.data_ptr())
        else:
            ctx.compute_all_vectors_partial(q0 * u[m_off:m_off + m_cnt], partial.data_ptr())
        ctx.synchronize()
        dist.all_reduce(partial)
        torch.cuda.synchronize()
        return [ctx.finalize(partial.data_ptr() + n * plen * 8, 1.0 / NM) for n in range(NQ)]

    # ---- device-resident measurement ----
    for i in range(args.warmup):
        compute_step(i)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    n0 = ctx.launch_count
    amp_ms = 0.0
    dsp_ms = 0.0
    ctx.timer_start()
    t0 = time.perf_counter()
    for i in range(args.steps):
        compute_step(args.warmup + i)
        amp_ms += ctx.last_amplitude_ms()
        dsp_ms += ctx.last_dsp_ms()
    ms = ctx.timer_stop()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = ctx.launch_count - n0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms, amp_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, amp_ms_max = float(t[0]), float(t[1])
    mine = {"rank": rank, "step_ms": ms / args.steps, "amp_ms": amp_ms / args.steps, "fp64_peak_tflops": fp64_peak}
    per_rank = [mine]
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    evals_step = float(NA) * NF * NM * NQ
    value = evals_step * args.steps / (ms_max * 1e-3)

    # ---- end to end: host buffers in, host results out, every step ----
    e2e = None
    if not args.no_e2e:
        host = ctx.pinned((f_cnt, NA, 3), np.float32)  # this rank's block of the trajectory, pinned
        ctx.memcpy_d2h(host.array, xyz.data_ptr() + f_off * NA * 12)
        slice_dev = torch.empty(f_cnt * NA * 3, dtype=torch.float32, device=dev) if (world > 1 and not by_frames) else None
        equal = all(div_assignment(world, r, NF)[1] == f_cnt for r in range(world))

        def e2e_step(i):
            if world == 1 or by_frames:
                # stager: chunked async H2D of this rank's frames on the copy stream; the amplitude launches wait per
                # chunk, so the copy overlaps the kernel
                ctx.stage_frames(host.array)
                if by_frames:
                    ctx.set_frame_window(NF, f_off)
                ctx.set_factors(b)
            else:
                # replicated coordinates: every rank loads its DivAssignment slice from the host, the slices are
                # all-gathered over NVLink (the reference's stage_fillpartitions broadcast, data_stager.cpp:102-118)
                ctx.memcpy_h2d(slice_dev.data_ptr(), host.array)
                if equal:
                    dist.all_gather_into_tensor(xyz, slice_dev)
                else:
                    parts = [xyz[div_assignment(world, r, NF)[0] * NA * 3:(sum(div_assignment(world, r, NF))) * NA * 3]
                             for r in range(world)]
                    dist.all_gather(parts, slice_dev)
                torch.cuda.synchronize()
                ctx.stage_frames_device(xyz.data_ptr(), NF, NA)
                ctx.set_factors(b)
            return compute_step(i)

        for i in range(min(args.warmup, 2)):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_step(args.warmup + i)
        barrier()
        e2e_s = time.perf_counter() - t0
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te[0])
        nvec_up = NM if (world == 1 or by_frames) else m_cnt
        e2e = {"value": evals_step * args.steps / e2e_s, "unit": "evals/s",
               "h2d_bytes_per_step": int(f_cnt * NA * 12 + NA * 8 + nvec_up * 24),
               "d2h_bytes_per_step": int(NQ * (NF * 16 + 32)),
               "ms_per_step": 1e3 * e2e_s / args.steps,
               "note": ("coordinates re-staged from pinned host memory every step (chunked async H2D overlapped "
                        "with the amplitude kernel); fqt/fq/fq2 of every |q| of the step read back to the host"
                        if world == 1 else
                        "every rank re-stages its own frame block from pinned host memory every step (chunked async H2D "
                        "overlapped with the amplitude kernel); amplitudes exchanged over NVLink" if by_frames else
                        "every rank H2D's its frame slice each step, slices all-gathered over NVLink, then compute")}
        stage_resident()  # restore the resident staging for anything that follows
        host.free()

    # ---- CPU baseline + parity on a bounded sample (rank 0, N=1) ----
    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as o
        cores = o.max_threads()
        NF_s, NM_s = cpu_sample_shape(cfg, cores, args.cpu_seconds)
        coords = np.empty((NF_s, NA, 3), dtype=np.float32)
        ctx.memcpy_d2h(coords, xyz.data_ptr())
        nmid = len(qls) // 2
        ql = qls[nmid]
        dt, (rfqt, rfq, rfq2), _ = run_cpu_oracle(cfg, NF_s, NM_s, ql, cores, coords)
        sample = (f"{NA} atoms x first {NF_s} frames x {NM_s} of {NM} subvectors of one |q| "
                  f"(amplitudes + FFT autocorrelation + store); " + cpu_sample_note(cores))
        cpu_baseline = {"value": float(NA) * NF_s * NM_s / dt, "unit": "evals/s", "cores": cores, "kind": cpu_reference_kind(),
                        "sample": sample, "seconds": dt}
        ctx.stage_frames_device(xyz.data_ptr(), NF_s, NA)
        ctx.set_factors(b)
        if scan:  # the same kernel as the timed path: the whole scan on the sample, compared at the sampled |q|
            fqts, fqs, fq2s = ctx.compute_all_vectors_scan(u[:NM_s], qls)
            fqt, fq = fqts[nmid], fqs[nmid]
        else:
            fqt, fq, _ = ctx.compute_all_vectors(ql * u[:NM_s])
        parity = {"fqt_rel_err": float(np.max(np.abs(fqt - rfqt)) / np.max(np.abs(rfqt))),
                  "fq_rel_err": float(abs(fq - rfq) / abs(rfqt[0])), "tolerance": 1e-9,
                  "vs": ("the reference's own AllVectorsScatterDevice (oracle/_ref build)" if cpu_reference_kind() == "reference"
                         else "oracle") + f" on the CPU sample, |q| index {nmid} of the scan"}
        stage_resident()

    if rank == 0:
        amp_s = amp_ms_max * 1e-3
        nf0 = div_assignment(world, 0, NF)[1] if by_frames else NF
        nm0 = NM if (by_frames or world == 1) else div_assignment(world, 0, NM)[1]
        evals_rank = float(NA) * nf0 * nm0 * NQ * args.steps
        scan_plan = ctx.last_scan_plan() if scan else None
        if scan:
            corrected = args.mode == "scan-rounded"
            passes = scan_passes(NQ, SCAN_MAX_PASS_CORR if corrected else SCAN_MAX_PASS)
            step_flop = SCAN_STEP_FLOP + (4.0 if corrected else 0.0)    # + D accumulation (2 FMA); E runs on the FP32 pipe
            step_instr = SCAN_STEP_INSTR + (2.0 if corrected else 0.0)
            flop_eval = (len(passes) * SCAN_SETUP_FLOP + sum(passes) * step_flop) / NQ
            instr_eval = (len(passes) * SCAN_SETUP_INSTR + sum(passes) * step_instr) / NQ
            kernel = "amplitude_scan_kernel" + (" (corrected)" if corrected else "")
        else:
            flop_eval, instr_eval, kernel = FLOP_PER_EVAL, FP64_INSTR_PER_EVAL, "amplitude_all_tiled_kernel"
        achieved = evals_rank * flop_eval / amp_s / 1e12
        traffic = None
        prof = os.path.join(ROOT, "profiles", "k1_traffic.json")
        if os.path.exists(prof):
            try:
                tj = json.load(open(prof))
                key = "scan_dram_bytes_per_frame" if scan else "dram_bytes_per_frame"
                traffic = tj[key] * nf0 if key in tj else None  # one launch covers all frames of the rank
            except Exception:
                traffic = None
        line = {
            "metric": "amplitude evals/s (atom*frame*q-vector)", "value": value, "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload], "NA": NA, "NF": NF, "NM_per_q": NM, "NQ": len(qls),
                       "step": (f"the whole |q| scan ({NQ} equally spaced |q| x {NM} orientation vectors) over the full "
                                "trajectory: amplitudes + FFT autocorrelation + average for every |q|" if scan else
                                "one |q| (compute() of the runner loop): amplitudes + FFT autocorrelation + average"),
                       "mode": args.mode, "scan_plan_plain_corrected_single": scan_plan,
                       "parallelism": (f"frame shard x{world} + amplitude all-reduce" if by_frames else
                                       f"q-vector shard x{world}" if world > 1 else "single GPU"),
                       "cache": f"inputs ({NF * NA * 12 / 1e9:.1f} GB coordinates) larger than L2"},
            "fqt_wall_time_s_all_q": ms_max / args.steps * 1e-3 * (len(qls) / NQ),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak, "traffic": traffic,
                         "kernel": kernel, "kernel_share_of_step": amp_ms_max / ms_max,
                         "algorithmic_flop_per_eval": flop_eval,
                         "fp64_pipe_util_executed": evals_rank * instr_eval * 2 / amp_s / 1e12 / fp64_peak,
                         "per_eval_formulation_equiv_tflops": evals_rank * FLOP_PER_EVAL / amp_s / 1e12,
                         "note": ("scan kernel: 2 sincos per (atom, direction, pass) + a 3-term recurrence per |q|; its own "
                                  "flop count is used for `achieved`.  per_eval_formulation_equiv_tflops is what the "
                                  "reference's one-sincos-per-evaluation formulation (45 flop/eval, SURVEY 8d) would need "
                                  "for the same evals/s" if scan else "45 flop per evaluation (SURVEY 8d)"),
                         "peak_source": "measured live: dependency-free DFMA chains on all SMs (sgpu_measure_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 entry"},
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity": parity,
            "host_wall_ms_per_step": wall_ms / args.steps,
            "dsp_ms_per_step": dsp_ms / args.steps,
            "per_rank": per_rank,
        }
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def mod_assignment_count(NN, rank, N):
    """ModAssignment (reference src/decomposition/assignment.cpp:82-132): indices rank, rank+NN, ..."""
    return (N - rank + NN - 1) // NN if rank < N else 0


def self_flop_per_timeline(NF):
    """45 flop per amplitude + one forward 2NF-point FFT per timeline (5 L log2 L); the inverse runs once per
    |q| on the summed power spectrum and is not counted"""
    L = 2.0 * NF
    return FLOP_PER_EVAL * NF + 5.0 * L * np.log2(L)


def self_cpu_sample(cfg, cores, target_core_seconds, timelines_per_core_s=70.0):
    """bounded CPU sample of the self workload: NA_s atoms x NM_s vectors of one |q|, all frames"""
    scale = 1e4 / cfg["NF"]
    n = max(cores, int(timelines_per_core_s * scale * cores * target_core_seconds))
    NM_s = min(cfg["NM"], 16)
    NA_s = max(cores, (n // NM_s // cores) * cores)
    return NA_s, NM_s


def run_reference_self(args):
    """--impl reference --workload C2: the oracle's self path (atoms over threads = ModAssignment over ranks)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from oracle import oracle as o
    from sassena_b200 import synth
    o.build()
    cfg = dict(synth.CONFIGS[args.workload])
    if args.frames:
        cfg["NF"] = args.frames
    if args.atoms:
        cfg["NA"] = args.atoms
    cores = o.max_threads()
    NA_s, NM_s = self_cpu_sample(cfg, cores, args.cpu_seconds)
    NF = cfg["NF"]
    qls = synth.qlengths(*cfg["q"])
    xa = np.ascontiguousarray(synth.trajectory(NF, NA_s, cfg["box"], cfg["sigma"], cfg["seed"]).transpose(1, 0, 2))
    b = synth.factors(cfg["NA"])[:NA_s]
    u = synth.unit_vectors(cfg["NM"], cfg["vseed"])[:NM_s]
    times = []
    for i in range(args.warmup + args.steps):
        dt, _ = run_cpu_self(xa, b, u, qls[i % len(qls)], cores)
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = float(NA_s) * NF * NM_s * len(times) / total
    sample = (f"{NA_s} of {cfg['NA']} atoms x all {NF} frames x {NM_s} of {cfg['NM']} vectors of one |q| per step (amplitude "
              f"timelines + FFT autocorrelation + store); " + cpu_sample_note(cores, "port") + " over atoms")
    line = {
        "impl": "reference", "metric": "amplitude evals/s (atom*frame*q-vector)", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "NA": cfg["NA"], "NF": NF, "NM": cfg["NM"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def _run_self(args, json_fd):
    """--workload C2: incoherent self scattering at full size.  A step is one compute() of the runner loop: one |q| with
    its 200 vectors over ALL atoms (6e6 timelines of 10k frames: amplitudes, FFT autocorrelation, store).  N GPUs: atoms
    sharded by ModAssignment (self_vectors_scatter_device.cpp:50), one NCCL all-reduce of the packed partial per |q|."""
    import torch
    import torch.distributed as dist
    import sassena_b200
    from sassena_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch multi-GPU runs with torchrun (one process per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(synth.CONFIGS[args.workload])
    if args.frames:
        cfg["NF"] = args.frames
    if args.atoms:
        cfg["NA"] = args.atoms
    NA, NF, NM = cfg["NA"], cfg["NF"], cfg["NM"]
    qls = synth.qlengths(*cfg["q"])
    u = synth.unit_vectors(NM, cfg["vseed"])
    b_all = synth.factors(NA)
    na_loc = mod_assignment_count(world, rank, NA)
    b_loc = np.ascontiguousarray(b_all[rank::world])

    ctx = sassena_b200.ScatterContext(local_rank)
    fp64_peak = ctx.measure_fp64_peak()
    # this rank's atoms (rank, rank+world, ...), atom-major [na_loc][NF][3], generated on the device (CPU twin in synth.py),
    # then kept in pinned host memory: the "trajectory on the host" that the end-to-end step stages from
    gen = torch.empty(na_loc * NF * 3, dtype=torch.float32, device=dev)
    ctx.synth_trajectory(gen.data_ptr(), NF, NA, cfg["box"], cfg["sigma"], cfg["seed"], layout=1, atom0=rank,
                         atom_stride=world, NA_out=na_loc)
    host = ctx.pinned((na_loc, NF, 3), np.float32)
    ctx.memcpy_d2h(host.array, gen.data_ptr())
    del gen
    torch.cuda.empty_cache()
    ctx.stage_atoms(host.array)  # library-owned buffer: the split path keeps it in its decimated frame order
    ctx.set_factors(b_loc)
    plen = ctx.partial_len("autocorrelate")
    partial = torch.zeros(plen, dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    def compute_step(i):
        q = float(qls[i % len(qls)]) * u
        if world == 1:
            return ctx.compute_self_vectors(q)
        ctx.compute_self_vectors_partial(q, partial.data_ptr())
        ctx.synchronize()
        dist.all_reduce(partial)  # the three boost::mpi::reduce calls of self_vectors_scatter_device.cpp:213-221
        torch.cuda.synchronize()
        return ctx.finalize(partial.data_ptr(), 1.0 / NM)

    for i in range(args.warmup):
        compute_step(i)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    n0 = ctx.launch_count
    amp_ms = 0.0
    ctx.timer_start()
    t0 = time.perf_counter()
    for i in range(args.steps):
        compute_step(args.warmup + i)
        amp_ms += ctx.last_amplitude_ms()
    ms = ctx.timer_stop()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = ctx.launch_count - n0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms, amp_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, amp_ms_max = float(t[0]), float(t[1])
    mine = {"rank": rank, "step_ms": ms / args.steps, "kernel_ms": amp_ms / args.steps, "atoms": na_loc,
            "fp64_peak_tflops": fp64_peak}
    per_rank = [mine]
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    evals_step = float(NA) * NF * NM
    value = evals_step * args.steps / (ms_max * 1e-3)

    e2e = None
    if not args.no_e2e:
        def e2e_step(i):
            ctx.stage_atoms(host.array)  # H2D of this rank's atoms from pinned host memory
            ctx.set_factors(b_loc)
            return compute_step(i)

        for i in range(min(args.warmup, 2)):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_step(args.warmup + i)
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te[0])
        e2e = {"value": evals_step * args.steps / e2e_s, "unit": "evals/s",
               "h2d_bytes_per_step": int(na_loc * NF * 12 + na_loc * 8 + NM * 24), "d2h_bytes_per_step": int(NF * 16 + 32),
               "ms_per_step": 1e3 * e2e_s / args.steps,
               "note": "every rank re-stages its atoms from pinned host memory every step (H2D, then the decimated frame "
                       "order of the split path is rebuilt on the device); fqt/fq/fq2 of the step's |q| read back"}

    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as o
        o.build()
        cores = o.max_threads()
        NA_s, NM_s = self_cpu_sample(cfg, cores, args.cpu_seconds)
        ql = qls[len(qls) // 2]
        xa = np.ascontiguousarray(host.array[:NA_s])
        dt, (rfqt, rfq, rfq2) = run_cpu_self(xa, b_all[:NA_s], u[:NM_s], ql, cores)
        sample = (f"{NA_s} of {NA} atoms x all {NF} frames x {NM_s} of {NM} vectors of one |q| (amplitude timelines + FFT "
                  f"autocorrelation + store); " + cpu_sample_note(cores, "port") + " over atoms")
        cpu_baseline = {"value": float(NA_s) * NF * NM_s / dt, "unit": "evals/s", "cores": cores, "kind": "port",
                        "sample": sample, "seconds": dt}
        ctx.stage_atoms(xa)
        ctx.set_factors(b_all[:NA_s])
        fqt, fq, _ = ctx.compute_self_vectors(ql * u[:NM_s])
        parity = {"fqt_rel_err": float(np.max(np.abs(fqt - rfqt)) / np.max(np.abs(rfqt))),
                  "fq_rel_err": float(abs(fq - rfq) / abs(rfqt[0])), "tolerance": 1e-9,
                  "vs": "oracle on the CPU sample (the oracle reproduces the reference's own SelfVectorsScatterDevice bit for bit, "
                        "tests/test_reference_devices.py)"}

    if rank == 0:
        kern_s = amp_ms_max * 1e-3
        tl_rank = float(mod_assignment_count(world, 0, NA)) * NM * args.steps
        achieved = tl_rank * self_flop_per_timeline(NF) / kern_s / 1e12
        traffic = None  # DRAM bytes per step of rank 0, from the ncu capture of the two split kernels (per timeline x timelines)
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
            if NF == 10000:  # the capture's timeline length
                traffic = tj["self_split_dram_bytes_per_timeline"] * tl_rank / args.steps
        except Exception:
            traffic = None
        line = {
            "metric": "amplitude evals/s (atom*frame*q-vector)", "value": value, "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload], "NA": NA, "NF": NF, "NM_per_q": NM, "NQ": len(qls),
                       "step": "one |q| (compute() of the runner loop) over all atoms: amplitude timelines, FFT "
                               "autocorrelation per (atom, vector), store/average",
                       "parallelism": f"atom shard (ModAssignment) x{world} + all-reduce of the packed partial" if world > 1
                       else "single GPU",
                       "cache": f"inputs ({NF * NA * 12 / 1e9:.1f} GB coordinates) larger than L2"},
            "timelines_per_s": float(NA) * NM * args.steps / (ms_max * 1e-3),
            "fqt_wall_time_s_all_q": ms_max / args.steps * 1e-3 * len(qls),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak, "traffic": traffic,
                         "kernel": "self_split_fft_kernel + self_split_combine_ring_kernel",
                         "kernel_share_of_step": amp_ms_max / ms_max,
                         "algorithmic_flop_per_timeline": self_flop_per_timeline(NF),
                         "note": "45 flop per amplitude + one forward 2NF-point FFT (5 L log2 L) per timeline (SURVEY 8d); "
                                 "the kernels transform L = R x 2^k >= 2NF-1 points",
                         "peak_source": "measured live: dependency-free DFMA chains on all SMs (sgpu_measure_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 entry"},
            "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "parity": parity,
            "host_wall_ms_per_step": wall_ms / args.steps, "per_rank": per_rank,
        }
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    host.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


MP_BATCH = 8  # |q| values per pass of the batched multipole kernel (MPSphereScatterDevice::runner batches them)


def mp_flop_per_atom_frame_q(L, nmom, Q=MP_BATCH):
    """SURVEY 8(d), multipole sphere: per (atom, frame) one Y_lm table (3 flop per (l, m >= 0) pair by recurrence + one
    sincos, 36 flop, per m) shared by the |q| of a pass; per |q| one j_l ladder (4 flop per l) and one complex MAC (8 flop)
    per moment"""
    pairs = (L + 1) * (L + 2) // 2
    return (3.0 * pairs + 36.0 * (L + 1)) / Q + 4.0 * (L + 1) + 8.0 * nmom


def run_reference_mpsphere(args):
    """--impl reference --workload C4: the oracle's MPSphere path on a bounded sample.  It reproduces the reference's own
    MPSphereScatterDevice bit for bit (tests/test_reference_devices.py); that device's oracle/_ref build evaluates sph_bessel /
    spherical_harmonic through the oracle's restatements (Boost.Math is absent), so the port is what is timed: kind "port"."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from oracle import oracle as o
    from sassena_b200 import synth
    o.build()
    cfg = dict(synth.CONFIGS[args.workload])
    if args.frames:
        cfg["NF"] = args.frames
    if args.atoms:
        cfg["NA"] = args.atoms
    cores = o.max_threads()
    mom = o.moments_sphere(cfg["L"])
    NF_s = 2
    NA_s = int(max(64, min(cfg["NA"], 7.8e7 / 16 * cores * args.cpu_seconds / (NF_s * len(mom)))))
    xyz = synth.trajectory(NF_s, NA_s, cfg["box"], cfg["sigma"], cfg["seed"], offset=cfg["offset"])
    sph = o.cart_to_spherical(xyz)
    b = synth.factors(cfg["NA"])[:NA_s]
    qls = synth.qlengths(*cfg["q"])
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        o.compute_mpsphere(sph, b, qls[(i * 7) % len(qls)], mom, dsp="square", nthreads=cores)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = float(NA_s) * NF_s * len(mom) * len(times) / total
    sample = (f"{NA_s} of {cfg['NA']} atoms x {NF_s} of {cfg['NF']} frames x {len(mom)} moments of one |q| per step (one "
              f"sph_bessel + spherical_harmonic evaluation per (moment, atom, frame) as the reference does), {cores} OpenMP threads")
    line = {
        "impl": "reference", "metric": "amplitude evals/s (atom*frame*q-vector)", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "NA": cfg["NA"], "NF": cfg["NF"], "moments": len(mom),
                   "unit_of_work": "one (atom, frame, |q|, moment) amplitude term", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def _run_mpsphere(args, json_fd):
    """--workload C4: multipole sphere averaging at full size.  A step is one pass of the batched multipole kernel: 8 |q|
    x 441 moments over all atoms and frames (amplitudes A_lm(q, t), dsp, store).  N GPUs: every rank holds the frames of
    its DivAssignment block of the ATOMS, the amplitudes (sums over atoms) are all-reduced before the DSP."""
    import torch
    import torch.distributed as dist
    import sassena_b200
    from sassena_b200 import synth
    from oracle import oracle as o  # moments list only here; the checker runs further down on rank 0

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch multi-GPU runs with torchrun (one process per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(synth.CONFIGS[args.workload])
    if args.frames:
        cfg["NF"] = args.frames
    if args.atoms:
        cfg["NA"] = args.atoms
    NA, NF, L = cfg["NA"], cfg["NF"], cfg["L"]
    mom = np.array([(0, 0)] + [(l, m) for l in range(1, L + 1) for m in range(-l, l + 1)], dtype=np.int64)  # parameters.cpp:1037-1075
    NM = len(mom)
    qls = synth.qlengths(*cfg["q"])
    batches = [qls[i:i + MP_BATCH] for i in range(0, len(qls) - MP_BATCH + 1, MP_BATCH)]
    a_off, a_cnt = div_assignment(world, rank, NA)
    b_loc = np.ascontiguousarray(synth.factors(NA)[a_off:a_off + a_cnt])

    ctx = sassena_b200.ScatterContext(local_rank)
    fp64_peak = ctx.measure_fp64_peak()
    gen = torch.empty(NF * a_cnt * 3, dtype=torch.float32, device=dev)
    ctx.synth_trajectory(gen.data_ptr(), NF, NA, cfg["box"], cfg["sigma"], cfg["seed"], layout=0, atom0=a_off, atom_stride=1,
                         NA_out=a_cnt, offset=cfg["offset"])
    host = ctx.pinned((NF, a_cnt, 3), np.float32)  # this rank's atoms of every frame, cartesian, pinned
    ctx.memcpy_d2h(host.array, gen.data_ptr())
    del gen
    torch.cuda.empty_cache()

    def stage():
        ctx.stage_frames(host.array)   # chunked async H2D into the library's buffer
        ctx.frames_to_spherical()      # SphericalCoordinateSet on the device (coordinate_set.cpp:303-315)

    stage()
    amp = torch.zeros(MP_BATCH * NM * NF * 2, dtype=torch.float64, device=dev)
    plen = ctx.partial_len("square")
    partial = torch.zeros(MP_BATCH * plen, dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    def compute_step(i):
        ql = batches[i % len(batches)]
        ctx.set_factors_batch(np.tile(b_loc, (len(ql), 1)))
        if world == 1:
            return ctx.compute_mpsphere_batch(ql, mom, dsp="square")
        ctx.mpsphere_amplitudes(ql, mom, 0, a_cnt, amp.data_ptr())
        ctx.synchronize()
        dist.all_reduce(amp)  # A_lm(q, t) are sums over atoms: complete them before the DSP
        torch.cuda.synchronize()
        ctx.mpsphere_dsp_partial(amp.data_ptr(), len(ql), NM, partial.data_ptr(), dsp="square")
        return [ctx.finalize(partial.data_ptr() + n * plen * 8, 1.0 / (4 * np.pi), dsp="square") for n in range(len(ql))]

    for i in range(args.warmup):
        compute_step(i)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    n0 = ctx.launch_count
    amp_ms = 0.0
    ctx.timer_start()
    t0 = time.perf_counter()
    for i in range(args.steps):
        compute_step(args.warmup + i)
        amp_ms += ctx.last_amplitude_ms()
    ms = ctx.timer_stop()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = ctx.launch_count - n0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms, amp_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, amp_ms_max = float(t[0]), float(t[1])
    mine = {"rank": rank, "step_ms": ms / args.steps, "kernel_ms": amp_ms / args.steps, "atoms": a_cnt, "fp64_peak_tflops": fp64_peak}
    per_rank = [mine]
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    evals_step = float(NA) * NF * MP_BATCH * NM
    value = evals_step * args.steps / (ms_max * 1e-3)

    e2e = None
    if not args.no_e2e:
        def e2e_step(i):
            stage()
            return compute_step(i)

        for i in range(min(args.warmup, 2)):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_step(args.warmup + i)
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te[0])
        e2e = {"value": evals_step * args.steps / e2e_s, "unit": "evals/s",
               "h2d_bytes_per_step": int(NF * a_cnt * 12 + MP_BATCH * a_cnt * 8 + NM * 16),
               "d2h_bytes_per_step": int(MP_BATCH * (NF * 16 + 32)), "ms_per_step": 1e3 * e2e_s / args.steps,
               "note": "every rank re-stages its atoms of all frames from pinned host memory every step (chunked async H2D), "
                       "converts them to (r, phi, theta) on the device, uploads the factors of the 8 |q| and reads fqt/fq/fq2 back"}

    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        o.build()
        cores = o.max_threads()
        NF_s = 2
        NA_s = int(max(64, min(NA, 7.8e7 / 16 * cores * args.cpu_seconds / (NF_s * NM))))
        sph = o.cart_to_spherical(np.ascontiguousarray(host.array[:NF_s, :NA_s]))
        ql = float(qls[len(qls) // 2])
        t0 = time.perf_counter()
        rfqt, rfq, rfq2 = o.compute_mpsphere(sph, b_loc[:NA_s], ql, mom, dsp="square", nthreads=cores)
        dt = time.perf_counter() - t0
        sample = (f"{NA_s} of {NA} atoms x {NF_s} of {NF} frames x {NM} moments of one |q| (one sph_bessel + spherical_harmonic "
                  f"evaluation per (moment, atom, frame) as the reference does), {cores} OpenMP threads")
        cpu_baseline = {"value": float(NA_s) * NF_s * NM / dt, "unit": "evals/s", "cores": cores, "kind": "port",
                        "sample": sample, "seconds": dt}
        ctx.stage_frames(sph, repr=sassena_b200.REPR_SPHERICAL)
        ctx.set_factors(b_loc[:NA_s])
        fqt, fq, _ = ctx.compute_mpsphere(ql, mom, dsp="square")
        parity = {"fqt_rel_err": float(np.max(np.abs(fqt - rfqt)) / np.max(np.abs(rfqt))),
                  "fq_rel_err": float(abs(fq - rfq) / abs(rfqt[0])), "tolerance": 1e-9, "vs": "oracle on the CPU sample"}

    if rank == 0:
        kern_s = amp_ms_max * 1e-3
        evals_rank = float(div_assignment(world, 0, NA)[1]) * NF * MP_BATCH * NM * args.steps
        flop_eval = mp_flop_per_atom_frame_q(L, NM) / NM
        achieved = evals_rank * flop_eval / kern_s / 1e12
        line = {
            "metric": "amplitude evals/s (atom*frame*q-vector)", "value": value, "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload], "NA": NA, "NF": NF, "NQ": len(qls), "moments": NM,
                       "unit_of_work": "one (atom, frame, |q|, moment) amplitude term",
                       "step": f"one pass of {MP_BATCH} |q| x {NM} moments over all atoms and frames: amplitudes, dsp (square), store",
                       "parallelism": f"atom shard (DivAssignment) x{world} + all-reduce of the amplitudes" if world > 1 else "single GPU",
                       "cache": f"inputs ({NF * NA * 12 / 1e9:.1f} GB coordinates) larger than L2"},
            "fqt_wall_time_s_all_q": ms_max / args.steps * 1e-3 * (len(qls) / MP_BATCH),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak, "traffic": None, "kernel": "multipole_gemm_kernel<8>",
                         "kernel_share_of_step": amp_ms_max / ms_max, "algorithmic_flop_per_eval": flop_eval,
                         "executed_fp64_flop_per_eval": (3.0 * 231 + 36.0 * 21 + MP_BATCH * (4.0 * 21 + 4.0 * 231)) / (MP_BATCH * NM)
                         if L == 20 else None,
                         "note": "algorithmic count of SURVEY 8(d): a Y_lm table per (atom, frame) shared by the |q| of a pass, a "
                                 "j_l ladder per |q| and one complex MAC (8 flop) per moment; the kernel evaluates only m >= 0 with "
                                 "real-times-complex products (4 flop per (l, m >= 0) pair), executed_fp64_flop_per_eval states that count",
                         "peak_source": "measured live: dependency-free DFMA chains on all SMs (sgpu_measure_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 entry"},
            "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "parity": parity,
            "host_wall_ms_per_step": wall_ms / args.steps, "per_rank": per_rank,
        }
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    host.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="override NF (debug; changes the workload)")
    ap.add_argument("--atoms", type=int, default=0, help="override NA (debug; changes the workload)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work per oracle sample, seconds per core")
    ap.add_argument("--mode", default="scan", choices=["scan", "scan-rounded", "per-q"],
                    help="scan: a step is the whole scan of equally spaced |q| through the scan kernel (default); "
                         "scan-rounded: the same with the reference's float-rounded |q| (corrected scan kernel); "
                         "per-q: a step is one |q| through the general kernel")
    ap.add_argument("--shard", default="frames", choices=["frames", "vectors"],
                    help="N>1: shard the frames (reference decomposition, default) or the subvectors (replicated coordinates)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 3) if not args.frames else args.warmup
    if args.impl == "reference":
        from sassena_b200 import synth
        if synth.CONFIGS[args.workload]["kind"] == "self":
            return run_reference_self(args)
        if synth.CONFIGS[args.workload]["kind"] == "mpsphere":
            return run_reference_mpsphere(args)
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
