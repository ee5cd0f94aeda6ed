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
FP64_INSTR_PER_EVAL = 21.0  # what the general kernel executes (sincos_qt.cuh)
# symmetric |q|-scan kernel (DESIGN.md "K1s", scan_sym.cu): per (atom, direction, pass) a set-up of one dot product, two
# phase scalings, two sincos and the b scaling (+ sigma z0 and the FP32 seeds for float-rounded scans); per PAIR of |q| two
# real Chebyshev recurrence steps and four real accumulations (+ four for the first-order correction sums)
SCAN_SETUP_INSTR, SCAN_SETUP_INSTR_CORR = 38.0, 40.0   # executed FP64-pipe instructions (SASS count of the hot loop)
SCAN_PAIR_INSTR, SCAN_PAIR_INSTR_CORR = 6.0, 10.0
SCAN_SETUP_FLOP = 65.0                                 # algorithmic FP64 flop of the set-up (as round 1)
SCAN_MAX_PASS, SCAN_MAX_PASS_CORR = 29, 17             # scan_sym.cu amplitude_scan_sym_max_pass
SCAN_MAX_PASS_CORR32 = 25                              # ... with the first-order correction sums in FP32 (CORR = 2)


def scan_passes(nq, max_b=SCAN_MAX_PASS):
    """pass lengths sgpu_capi.cu plan_scan uses for nq |q| values: even shares, at most max_b"""
    npass = (nq + max_b - 1) // max_b
    out, n0 = [], 0
    for p in range(npass):
        want = (nq - n0 + (npass - p) - 1) // (npass - p)
        L = min(want, max_b, nq - n0)
        out.append(L)
        n0 += L
    return out


def scan_work_per_eval(nq, corrected, fp32d=False):
    """(algorithmic flop, executed FP64-pipe instructions) per evaluation of a scan of nq |q| values: a pass of L values runs
    K = L // 2 pairs (an even L leaves one slot of the 2K+1 masked) plus the centre term.  fp32d: the corrected kernel whose
    first-order sums run as packed FP32 instructions (the FP64 pipe sees the plain kernel's pair work)"""
    passes = scan_passes(nq, (SCAN_MAX_PASS_CORR32 if fp32d else SCAN_MAX_PASS_CORR) if corrected else SCAN_MAX_PASS)
    pair = (SCAN_PAIR_INSTR if fp32d else SCAN_PAIR_INSTR_CORR) if corrected else SCAN_PAIR_INSTR
    setup = SCAN_SETUP_INSTR_CORR if corrected else SCAN_SETUP_INSTR
    instr = sum(setup + pair * max(1, L // 2) for L in passes)
    # flop: a DFMA is two flop, the set-up as counted in round 1 (65) plus the correction products
    flop = sum(SCAN_SETUP_FLOP + (6.0 if corrected else 0.0) + 2.0 * pair * max(1, L // 2) for L in passes)
    return flop / nq, instr / nq


def host_cores():
    """threads of the CPU legs: every core this process may run on -- never OMP_NUM_THREADS, which torchrun sets to 1"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


WORKLOADS = {
    # name: description (config key in sassena_b200.synth.CONFIGS)
    "C3": "coherent F(q,t): 100k atoms x 10k frames, 50 |q| x 500 sphere vectors",
    "C1": "synthetic 1k-atom box, 100 frames, 10 |q| x 100 sphere vectors, coherent",
    "C2": "incoherent self F_s(q,t): 30k atoms x 10k frames, 20 |q| x 200 vectors, per-atom FFT correlation (one |q| per step)",
    "C4": "static SAXS via multipole sphere averaging (MPSphereScatterDevice): 1M atoms, 1k frames, 200 |q|, moments l <= 20 "
          "(one batch of 8 |q| per step)",
    "C5": "stager-streamed 500k atoms x 50k frames self scattering (trajectory exceeds one GPU's HBM), atom-sharded with NCCL "
          "allreduce",
    "C5s": "bounded sample of: stager-streamed 500k atoms x 50k frames self scattering (waves of atoms streamed from pinned "
           "host memory, double-buffered), atom-sharded with NCCL allreduce",
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
    from oracle import oracle as o
    from sassena_b200 import synth
    cfg = dict(synth.CONFIGS[args.workload])
    if args.frames:
        cfg["NF"] = args.frames
    if args.atoms:
        cfg["NA"] = args.atoms
    cores = host_cores()
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
    return line


def ncu_figures(key):
    """figures of the committed ncu captures (profiles/ncu_figures.json: per kernel the FP64 pipe utilisation and DRAM bytes
    the summaries under profiles/ state), so that the line quotes the profiler's number next to the modelled one"""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_figures.json")))
        return d.get(key, {})
    except Exception:
        return {}


class Env:
    """one process per GPU: rank / world from torchrun's environment, NCCL process group for N > 1"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        if args.gpus > 1 and self.world == 1:
            raise SystemExit("launch multi-GPU runs with torchrun (one process per GPU)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()

    def attach(self, ctx):
        """give the context the job's NCCL communicator (inside the library): rank 0's unique id is broadcast through
        torch.distributed -- plumbing only, every data-path collective then runs on the library's streams"""
        if self.world > 1:
            import sassena_b200
            box = [sassena_b200.comm_unique_id() if self.rank == 0 else None]
            self.dist.broadcast_object_list(box, src=0)
            ctx.comm_init(box[0], self.world, self.rank)

    def free_cached(self):
        import gc
        gc.collect()
        self.torch.cuda.empty_cache()


def sub_args(args, workload, **over):
    """arguments of a secondary workload of the default invocation: few steps, bounded CPU sample"""
    a = argparse.Namespace(**vars(args))
    a.workload = workload
    a.steps = min(args.steps, 3)
    a.warmup = 3 if not (args.frames or args.atoms) else min(args.warmup, 3)
    a.cpu_seconds = min(args.cpu_seconds, 4.0)
    for k, v in over.items():
        setattr(a, k, v)
    return a


def run_workload(env, a):
    from sassena_b200 import synth
    if a.workload in ("C5", "C5s"):
        return bench_streamed_self(env, a)
    kind = synth.CONFIGS[a.workload]["kind"]
    if kind == "self":
        return bench_self(env, a)
    if kind == "mpsphere":
        return bench_mpsphere(env, a)
    return bench_coherent(env, a)


def run_ours(args):
    # exactly ONE line may reach stdout (the JSON line): NCCL / torchrun banners go to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        env = Env(args)
        if args.workload == "all":
            # headline: config 3 on the reference's own |q| generator; then the other north-star workloads, short
            line = run_workload(env, sub_args(args, "C3", steps=args.steps, warmup=args.warmup, cpu_seconds=args.cpu_seconds))
            subs = {}
            plan = [("C3_equally_spaced", sub_args(args, "C3", mode="scan", no_cpu=True)),
                    ("C2", sub_args(args, "C2")), ("C4", sub_args(args, "C4")), ("C5s", sub_args(args, "C5s"))]
            if env.world > 1 and args.shard == "frames":
                # the north star's partition of the coherent path -- q-vectors (here: the directions of the scan) over the
                # GPUs, every rank holding all frames, one all-reduce of the packed partials -- next to the headline's frame
                # partition (scatter_device_factory.cpp:104-138 splits q-vectors between partitions, frames inside one)
                plan.insert(1, ("C3_vector_sharded", sub_args(args, "C3", shard="vectors", no_cpu=True)))
            for name, a in plan:
                if name in args.skip:
                    continue
                env.free_cached()
                t0 = time.perf_counter()
                try:
                    sub = run_workload(env, a)
                except Exception as e:  # a secondary workload must not take the headline line down with it
                    sub = {"error": repr(e)[:400]} if env.rank == 0 else None
                    print(f"bench: workload {name} failed: {e!r}", file=sys.stderr)
                if sub is not None:
                    sub["bench_wall_s"] = time.perf_counter() - t0
                    subs[name] = sub
            if line is not None:
                line["workloads"] = subs
        else:
            line = run_workload(env, args)
        if line is not None:
            os.write(saved_stdout, (json.dumps(line) + "\n").encode())
        env.close()
        return 0
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)


def bench_coherent(env, args):
    """coherent workloads (C3, C1): returns the JSON line as a dict on rank 0, None elsewhere"""
    import sassena_b200
    from sassena_b200 import synth
    torch, dist = env.torch, env.dist
    world, rank, local_rank, dev = env.world, env.rank, env.local_rank, env.dev

    cfg = dict(synth.CONFIGS[args.workload])
    if args.frames:
        cfg["NF"] = args.frames
    if args.atoms:
        cfg["NA"] = args.atoms
    NA, NF, NM = cfg["NA"], cfg["NF"], cfg["NM"]
    scan = args.mode in ("scan", "scan-rounded")
    if args.mode == "scan":
        qls = np.linspace(cfg["q"][0], cfg["q"][1], cfg["q"][2])  # exactly equally spaced |q|: plain scan kernel
    else:
        # the reference's scan generator: float-rounded fractions (parameters.cpp:1151), equally spaced to ~1e-8 only;
        # the library plans the corrected scan kernel for them (default mode, "scan-rounded")
        qls = synth.qlengths(*cfg["q"])
    NQ = len(qls) if scan else 1          # |q| values per step
    b = synth.factors(NA)
    u = synth.unit_vectors(NM, cfg["vseed"])
    m_off, m_cnt = div_assignment(world, rank, NM)
    f_off, f_cnt = div_assignment(world, rank, NF)
    by_frames = world > 1 and args.shard == "frames"

    ctx = sassena_b200.ScatterContext(local_rank)
    env.attach(ctx)
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
            # local amplitudes -> exchange over NVLink (grouped ncclSend/ncclRecv, overlapped with the next pass) -> DSP of this
            # rank's timelines -> all-reduce of the packed partials: all inside the library, no host synchronisation
            if scan:
                ctx.compute_all_vectors_scan_sharded(u, qls, partial.data_ptr())
            else:
                ctx.compute_all_vectors_sharded(q0 * u, partial.data_ptr())
        else:
            if scan:
                ctx.compute_all_vectors_scan_partial(u[m_off:m_off + m_cnt], qls, partial.data_ptr())
            else:
                ctx.compute_all_vectors_partial(q0 * u[m_off:m_off + m_cnt], partial.data_ptr())
            ctx.comm_allreduce(partial.data_ptr(), NQ * plen)
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

        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        for i in range(min(args.warmup, 1)):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            e2e_step(args.warmup + i)
        barrier()
        e2e_s = time.perf_counter() - t0
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te[0])
        nvec_up = NM if (world == 1 or by_frames) else m_cnt
        e2e = {"value": evals_step * e2e_steps / e2e_s, "unit": "evals/s",
               "h2d_bytes_per_step": int(f_cnt * NA * 12 + NA * 8 + nvec_up * 24),
               "d2h_bytes_per_step": int(NQ * (NF * 16 + 32)),
               "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps,
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
        cores = host_cores()
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
        if scan:
            # every |q| of the scan, not only the sampled one: the scan kernel's amplitudes (recurrences, and for float-rounded
            # scans the first / second order corrections, part of them in FP32) against the general kernel K1, which evaluates
            # one FP64 sincos per (atom, frame, q-vector) as the reference does
            a_scan = torch.empty(NQ * NM_s * NF_s * 2, dtype=torch.float64, device=dev)
            a_one = torch.empty(NM_s * NF_s * 2, dtype=torch.float64, device=dev)
            ctx.all_vectors_scan_amplitudes(u[:NM_s], qls, a_scan.data_ptr())
            worst, where = 0.0, 0
            for n in range(NQ):
                ctx.all_vectors_amplitudes(float(qls[n]) * u[:NM_s], a_one.data_ptr())
                ctx.synchronize()
                torch.cuda.synchronize()
                d = float((a_scan.view(NQ, -1)[n] - a_one).abs().max() / a_one.abs().max())
                if d > worst:
                    worst, where = d, n
            parity["scan_vs_per_q_kernel_max_rel_diff"] = worst
            parity["scan_vs_per_q_kernel_worst_q_index"] = where
            del a_scan, a_one
        stage_resident()

    if rank == 0:
        amp_s = amp_ms_max * 1e-3
        nf0 = div_assignment(world, 0, NF)[1] if by_frames else NF
        nm0 = NM if (by_frames or world == 1) else div_assignment(world, 0, NM)[1]
        evals_rank = float(NA) * nf0 * nm0 * NQ * args.steps
        scan_plan = ctx.last_scan_plan() if scan else None
        corrected = fp32d = False
        if scan:
            corrected = bool(scan_plan) and scan_plan[1] > 0  # what the library planned for this |q| list
            # fewer corrected passes than the FP64-D kernel needs: the FP32-D kernel (25-|q| passes) was within its error bound
            fp32d = corrected and scan_plan[1] < len(scan_passes(NQ, SCAN_MAX_PASS_CORR))
            flop_eval, instr_eval = scan_work_per_eval(NQ, corrected, fp32d)
            kernel = "amplitude_scan_sym_kernel" + ((" (corrected: float-rounded scan, first-order sums in " +
                                                    ("packed FP32)" if fp32d else "FP64)")) if corrected else "")
        else:
            flop_eval, instr_eval, kernel = FLOP_PER_EVAL, FP64_INSTR_PER_EVAL, "amplitude_all_tiled_kernel"
        achieved = evals_rank * flop_eval / amp_s / 1e12
        ncu = ncu_figures(("scan_corrected_fp32d" if fp32d else "scan_corrected") if (scan and corrected) else
                          "scan_plain" if scan else "k1_tiled")
        traffic = ncu.get("dram_bytes_per_frame", None)
        traffic = traffic * nf0 if traffic is not None else None  # one launch covers all frames of the rank
        line = {
            "metric": "amplitude evals/s (atom*frame*q-vector)", "value": value, "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload], "NA": NA, "NF": NF, "NM_per_q": NM, "NQ": len(qls),
                       "step": (f"the whole |q| scan ({NQ} |q| x {NM} orientation vectors) over the full "
                                "trajectory: amplitudes + FFT autocorrelation + average for every |q|" if scan else
                                "one |q| (compute() of the runner loop): amplitudes + FFT autocorrelation + average"),
                       "mode": args.mode,
                       "q_list": ("the reference's scan generator (float-rounded fractions, parameters.cpp:1151)"
                                  if args.mode != "scan" else "np.linspace (exactly equally spaced)"),
                       "scan_plan_plain_corrected_single": scan_plan,
                       "parallelism": (f"frame shard x{world} + amplitude exchange over NVLink" if by_frames else
                                       f"q-vector shard x{world}" if world > 1 else "single GPU"),
                       "cache": f"inputs ({NF * NA * 12 / 1e9:.1f} GB coordinates) larger than L2"},
            "fqt_wall_time_s_all_q": ms_max / args.steps * 1e-3 * (len(qls) / NQ),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak, "traffic": traffic,
                         "kernel": kernel, "kernel_share_of_step": amp_ms_max / ms_max,
                         "algorithmic_flop_per_eval": flop_eval,
                         "fp64_instr_per_eval_executed": instr_eval,
                         "fp64_pipe_util_executed": evals_rank * instr_eval * 2 / amp_s / 1e12 / fp64_peak,
                         "fp64_pipe_util_ncu": ncu.get("fp64_pipe_pct"), "ncu_source": ncu.get("source"),
                         "survey_8d_flop_per_eval": FLOP_PER_EVAL,
                         "frac_by_survey_8d_count": evals_rank * FLOP_PER_EVAL / amp_s / 1e12 / fp64_peak,
                         "note": ("symmetric scan kernel: 2 sincos per (atom, direction, pass) + two real Chebyshev recurrence "
                                  "steps and four (corrected with FP64 first-order sums: eight; with FP32 sums: four + packed "
                                  "FP32) real accumulations per PAIR of |q|; `achieved` counts the FP64 flop this formulation "
                                  "needs.  frac_by_survey_8d_count prices the same evaluations at SURVEY "
                                  "8(d)'s one-sincos-per-evaluation figure (45 flop) and therefore exceeds 1: the kernel does "
                                  "not execute that work.  fp64_pipe_util_ncu is sm__inst_executed_pipe_fp64 of the committed "
                                  "capture" if scan else "45 flop per evaluation (SURVEY 8d)"),
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
    else:
        line = None
    ctx.close()
    del xyz, partial
    return line


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
    from oracle import oracle as o
    from sassena_b200 import synth
    o.build()
    cfg = dict(synth.CONFIGS[args.workload])
    if args.frames:
        cfg["NF"] = args.frames
    if args.atoms:
        cfg["NA"] = args.atoms
    cores = host_cores()
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
    return line


def bench_self(env, args):
    """--workload C2: incoherent self scattering at full size.  A step is one compute() of the runner loop: one |q| with
    its 200 vectors over ALL atoms (6e6 timelines of 10k frames: amplitudes, FFT autocorrelation, store).  N GPUs: atoms
    sharded by ModAssignment (self_vectors_scatter_device.cpp:50), one NCCL all-reduce of the packed partial per |q|."""
    import sassena_b200
    from sassena_b200 import synth
    torch, dist = env.torch, env.dist
    world, rank, local_rank, dev = env.world, env.rank, env.local_rank, env.dev
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
    env.attach(ctx)
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
        ctx.comm_allreduce(partial.data_ptr(), plen)  # the three boost::mpi::reduce calls of self_vectors_scatter_device.cpp:213-221
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

        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        for i in range(min(args.warmup, 1)):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            e2e_step(args.warmup + i)
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te[0])
        e2e = {"value": evals_step * e2e_steps / e2e_s, "unit": "evals/s", "steps": e2e_steps,
               "h2d_bytes_per_step": int(na_loc * NF * 12 + na_loc * 8 + NM * 24), "d2h_bytes_per_step": int(NF * 16 + 32),
               "ms_per_step": 1e3 * e2e_s / e2e_steps,
               "note": "every rank re-stages its atoms from pinned host memory every step (chunks of atoms on the copy stream; "
                       "each chunk is brought into the decimated frame order of the split path and evaluated as it lands, the "
                       "next copies run underneath); fqt/fq/fq2 of the step's |q| read back"}

    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as o
        o.build()
        cores = host_cores()
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
                         "kernel": "self_split_fft12_kernel + self_split_combine_ring_kernel",
                         "kernel_share_of_step": amp_ms_max / ms_max,
                         "algorithmic_flop_per_timeline": self_flop_per_timeline(NF),
                         "note": "45 flop per amplitude + one forward 2NF-point FFT (5 L log2 L) per timeline (SURVEY 8d); "
                                 "the kernels transform L = R x 2^k >= 2NF-1 points",
                         "peak_source": "measured live: dependency-free DFMA chains on all SMs (sgpu_measure_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 entry"},
            "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "parity": parity,
            "host_wall_ms_per_step": wall_ms / args.steps, "per_rank": per_rank,
        }
    else:
        line = None
    host.free()
    ctx.close()
    return line


def host_mem_available():
    """MemAvailable of /proc/meminfo in bytes (None if unreadable)"""
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) * 1024
    except OSError:
        pass
    return None


def bench_streamed_self(env, args):
    """--workload C5s / C5: BASELINE configs[4], stager-streamed self scattering (50k frames).  The rank's atoms (ModAssignment)
    sit atom-major in PINNED HOST memory and pass through the GPU in waves: sgpu_stage_atoms_prefetch queues wave w+1 on the
    copy stream while wave w is evaluated for every |q| of the step (sgpu_stage_atoms_swap, no host synchronisation); the packed
    partials -- sums over atoms -- accumulate per |q| on the device, one all-reduce + finalize per |q| after the last wave
    (self_vectors_scatter_device.cpp:213-238; DataStagerByAtom's buffering, data_stager.cpp:249-338, as a pipeline).
    C5s: a bounded sample (args.c5s_atoms atoms, one |q| per step) for the default invocation.  C5: all 500k atoms, args.nq |q|
    per step -- 300 GB of pinned host memory over the ranks, meant for 8 GPUs."""
    import sassena_b200
    from sassena_b200 import synth
    torch, dist = env.torch, env.dist
    world, rank, local_rank, dev = env.world, env.rank, env.local_rank, env.dev
    full = args.workload == "C5"
    cfg = dict(synth.CONFIGS["C5"])
    if args.frames:
        cfg["NF"] = args.frames
    NA = cfg["NA"] if full else min(cfg["NA"], args.c5s_atoms)
    if args.atoms:
        NA = args.atoms
    NF, NM = cfg["NF"], cfg["NM"]
    qls_all = synth.qlengths(*cfg["q"])
    NQ = min(len(qls_all), max(1, args.nq)) if full else 1
    u = synth.unit_vectors(NM, cfg["vseed"])
    b_all = synth.factors(cfg["NA"])
    atom_bytes = NF * 12
    # the ranks' atoms sit in pinned host memory: never ask for more than 70 % of what the box has free (a box driven out of
    # memory dies); if the configuration does not fit, run as many atoms as do and say so (NA < NA_config in the line)
    avail = host_mem_available()
    if avail:
        t_av = torch.tensor([float(avail)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_av, op=dist.ReduceOp.MIN)
        fit = int(0.7 * float(t_av[0]) // atom_bytes)
        if NA > fit:
            NA = max(world, fit // world * world)
    na_loc = mod_assignment_count(world, rank, NA)
    b_loc = np.ascontiguousarray(b_all[rank:NA:world])
    W = args.wave_atoms or max(64, int(2.0e9 // atom_bytes) // 64 * 64)  # ~2 GB per wave buffer
    W = max(1, min(W, na_loc))
    nwaves = (na_loc + W - 1) // W

    ctx = sassena_b200.ScatterContext(local_rank)
    env.attach(ctx)
    fp64_peak = ctx.measure_fp64_peak()
    free0, total_mem = torch.cuda.mem_get_info()
    # this rank's atoms, atom-major, generated wave by wave on the device (CPU twin: synth.py) into the pinned host buffer
    # the stager streams from
    t0 = time.perf_counter()
    host = ctx.pinned((na_loc, NF, 3), np.float32)
    gen = torch.empty(W * NF * 3, dtype=torch.float32, device=dev)
    for w in range(nwaves):
        first, cnt = w * W, min(W, na_loc - w * W)
        ctx.synth_trajectory(gen.data_ptr(), NF, cfg["NA"], cfg["box"], cfg["sigma"], cfg["seed"], layout=1,
                             atom0=rank + first * world, atom_stride=world, NA_out=cnt)
        ctx.memcpy_d2h(host.array[first:first + cnt], gen.data_ptr())
    del gen
    torch.cuda.empty_cache()
    gen_s = time.perf_counter() - t0

    plen = None
    acc = partial = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    # what the copy engine delivers for one wave with nothing else running (all ranks at once: they share the host's memory)
    w0 = min(W, na_loc)
    ctx.stage_atoms_prefetch(host.array[0:w0])
    ctx.stage_atoms_swap()
    barrier()
    t0 = time.perf_counter()
    ctx.stage_atoms_prefetch(host.array[0:w0])
    ctx.stage_atoms_swap()
    ctx.synchronize()
    torch.cuda.synchronize()
    h2d_idle_gbps = w0 * atom_bytes / (time.perf_counter() - t0) / 1e9

    kern_ms = [0.0]

    def stream_step(i):
        """all waves of this rank for the NQ |q| of step i; returns the finalized (fqt, fq, fq2) per |q|"""
        nonlocal plen, acc, partial
        qs = [float(qls_all[(i * NQ + n) % len(qls_all)]) for n in range(NQ)]
        ctx.stage_atoms_prefetch(host.array[0:min(W, na_loc)])
        for w in range(nwaves):
            first, cnt = w * W, min(W, na_loc - w * W)
            ctx.stage_atoms_swap()
            if w + 1 < nwaves:
                ctx.stage_atoms_prefetch(host.array[first + cnt:min(first + cnt + W, na_loc)])
            ctx.set_factors(b_loc[first:first + cnt])
            if plen is None:
                plen = ctx.partial_len("autocorrelate")
                acc = torch.zeros(NQ * plen, dtype=torch.float64, device=dev)
                partial = torch.zeros(plen, dtype=torch.float64, device=dev)
            for n, ql in enumerate(qs):
                dst = acc.data_ptr() + n * plen * 8
                if w == 0:
                    ctx.compute_self_vectors_partial(ql * u, dst)
                else:
                    ctx.compute_self_vectors_partial(ql * u, partial.data_ptr())
                    ctx.accumulate(dst, partial.data_ptr(), plen)
                kern_ms[0] += ctx.last_amplitude_ms()
        if world > 1:
            ctx.comm_allreduce(acc.data_ptr(), NQ * plen)  # the three boost::mpi::reduce calls of self_vectors_scatter_device.cpp:213-221
        return [ctx.finalize(acc.data_ptr() + n * plen * 8, 1.0 / NM) for n in range(NQ)]

    for i in range(args.warmup):
        stream_step(i)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    n0 = ctx.launch_count
    kern_ms[0] = 0.0
    ctx.timer_start()
    t0 = time.perf_counter()
    res = None
    for i in range(args.steps):
        res = stream_step(args.warmup + i)
    ms = ctx.timer_stop()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = ctx.launch_count - n0
    clocks = sampler.stop() if rank == 0 else None
    free1, _ = torch.cuda.mem_get_info()
    hbm_used = max(0, free0 - free1)
    # the wall clock (host, barrier on both sides) is the step time here: copies and kernels run on two streams
    t = torch.tensor([wall_ms, ms, kern_ms[0], float(hbm_used)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall_max, ms_max, kern_max, hbm_max = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    mine = {"rank": rank, "step_ms": wall_ms / args.steps, "compute_stream_ms": ms / args.steps, "kernel_ms": kern_ms[0] / args.steps,
            "atoms": na_loc, "waves": nwaves, "hbm_bytes": int(hbm_used), "library_buffers_bytes": ctx.device_bytes(),
            "pinned_host_bytes": int(na_loc * atom_bytes), "generate_s": gen_s}
    per_rank = [mine]
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    evals_step = float(NA) * NF * NM * NQ
    value = evals_step * args.steps / (wall_max * 1e-3)

    # the same atoms resident in HBM (when they fit beside the wave buffers): what the streaming costs
    resident = None
    if not full and na_loc * atom_bytes < 0.5 * free1:
        ctx.stage_atoms(host.array)
        ctx.set_factors(b_loc)
        pr = torch.zeros(plen, dtype=torch.float64, device=dev)
        ql = float(qls_all[(args.warmup + args.steps - 1) % len(qls_all)])
        ctx.compute_self_vectors_partial(ql * u, pr.data_ptr())
        barrier()
        ctx.timer_start()
        ctx.compute_self_vectors_partial(ql * u, pr.data_ptr())
        rms = ctx.timer_stop()
        barrier()
        if world > 1:
            ctx.comm_allreduce(pr.data_ptr(), plen)
        rres = ctx.finalize(pr.data_ptr(), 1.0 / NM)
        tr = torch.tensor([rms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        resident = {"value": float(NA) * NF * NM / (float(tr[0]) * 1e-3), "ms": float(tr[0]),
                    "streamed_vs_resident_rel_err": float(np.max(np.abs(res[-1][0] - rres[0])) / np.max(np.abs(rres[0]))),
                    "note": "the sample staged once and evaluated for the step's last |q| without streaming"}

    cpu_baseline = parity = None
    if rank == 0 and not args.no_cpu and (world == 1 or full):
        from oracle import oracle as o
        o.build()
        cores = host_cores()
        NA_s, NM_s = self_cpu_sample(cfg, cores, args.cpu_seconds)
        NA_s = min(NA_s, na_loc)
        ql = float(qls_all[len(qls_all) // 2])
        xa = np.ascontiguousarray(host.array[:NA_s])
        dt, (rfqt, rfq, rfq2) = run_cpu_self(xa, b_loc[:NA_s], u[:NM_s], ql, cores)
        sample = (f"{NA_s} atoms x all {NF} frames x {NM_s} of {NM} vectors of one |q| (amplitude timelines + FFT "
                  f"autocorrelation + store); " + cpu_sample_note(cores, "port") + " over atoms")
        cpu_baseline = {"value": float(NA_s) * NF * NM_s / dt, "unit": "evals/s", "cores": cores, "kind": "port",
                        "sample": sample, "seconds": dt}
        ctx.stage_atoms(xa)
        ctx.set_factors(b_loc[:NA_s])
        fqt, fq, _ = ctx.compute_self_vectors(ql * u[:NM_s])
        parity = {"fqt_rel_err": float(np.max(np.abs(fqt - rfqt)) / np.max(np.abs(rfqt))),
                  "fq_rel_err": float(abs(fq - rfq) / abs(rfqt[0])), "tolerance": 1e-9,
                  "vs": "oracle on the CPU sample: this rank's first atoms of the streamed trajectory (the oracle reproduces "
                        "the reference's own SelfVectorsScatterDevice bit for bit, tests/test_reference_devices.py)"}

    if rank == 0:
        tl_rank = float(mod_assignment_count(world, 0, NA)) * NM * NQ * args.steps
        achieved = tl_rank * self_flop_per_timeline(NF) / (kern_max * 1e-3) / 1e12
        h2d = int(na_loc * atom_bytes + na_loc * 8 + NQ * NM * 24)
        line = {
            "metric": "amplitude evals/s (atom*frame*q-vector)", "value": value, "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload], "NA": NA, "NA_config": cfg["NA"], "NF": NF, "NM_per_q": NM,
                       "NQ_per_step": NQ, "NQ_config": len(qls_all), "wave_atoms": W, "waves_per_rank": nwaves,
                       "step": f"every atom of the {'configuration' if full else 'sample'} streamed once from pinned host memory "
                               f"(waves of {W} atoms, double-buffered) and evaluated for {NQ} |q| x {NM} vectors: amplitude "
                               "timelines, FFT autocorrelation per (atom, vector), accumulation over waves, all-reduce, finalize",
                       "parallelism": f"atom shard (ModAssignment) x{world} + all-reduce of the packed partials" if world > 1
                       else "single GPU",
                       "cache": f"inputs ({NA * atom_bytes / 1e9:.1f} GB of coordinates) stream from the host every step"},
            "timelines_per_s": float(NA) * NM * NQ * args.steps / (wall_max * 1e-3),
            "fqt_wall_time_s_all_q": wall_max / args.steps * 1e-3 * (len(qls_all) / NQ) * (cfg["NA"] / NA),
            "staging": {"h2d_bytes_per_step_per_rank": int(na_loc * atom_bytes),
                        "h2d_gbps_needed_per_rank": na_loc * atom_bytes / (wall_max / args.steps * 1e-3) / 1e9,
                        "h2d_gbps_one_wave_idle_rank0": h2d_idle_gbps,
                        "host_mem_available_bytes": avail,
                        "kernel_share_of_step": kern_max / wall_max,
                        "overlap": "wave w+1 is copied on the copy stream while wave w is evaluated; a step's wall time minus its "
                                   "kernel time is what staging, launches and the final reduce add",
                        "hbm_high_water_bytes": int(hbm_max), "resident": resident},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak, "traffic": None,
                         "kernel": "self_split_fft12_kernel + self_split_combine_*",
                         "kernel_share_of_step": kern_max / wall_max,
                         "algorithmic_flop_per_timeline": self_flop_per_timeline(NF),
                         "note": "45 flop per amplitude + one forward 2NF-point FFT (5 L log2 L) per timeline (SURVEY 8d)",
                         "peak_source": "measured live: dependency-free DFMA chains on all SMs (sgpu_measure_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 entry"},
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(NQ * (NF * 16 + 32)),
                    "ms_per_step": wall_max / args.steps,
                    "note": "the streamed step IS end to end: every step copies all of the rank's atoms from pinned host memory "
                            "(inside the timed region, overlapped with the kernels) and reads fqt/fq/fq2 of its |q| back"},
            "gpu_launches": int(launches), "clocks": clocks, "parity": parity, "per_rank": per_rank,
        }
    else:
        line = None
    host.free()
    ctx.close()
    return line


MP_BATCH = 8  # |q| values per pass of the batched multipole kernel (MPSphereScatterDevice::runner batches as many)


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
    from oracle import oracle as o
    from sassena_b200 import synth
    o.build()
    cfg = dict(synth.CONFIGS[args.workload])
    if args.frames:
        cfg["NF"] = args.frames
    if args.atoms:
        cfg["NA"] = args.atoms
    cores = host_cores()
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
    return line


def bench_mpsphere(env, args):
    """--workload C4: multipole sphere averaging at full size.  A step is one pass of the batched multipole kernel: 8 |q|
    x 441 moments over all atoms and frames (amplitudes A_lm(q, t), dsp, store).  N GPUs: every rank holds the frames of
    its DivAssignment block of the ATOMS, the amplitudes (sums over atoms) are all-reduced before the DSP."""
    import sassena_b200
    from sassena_b200 import synth
    torch, dist = env.torch, env.dist
    world, rank, local_rank, dev = env.world, env.rank, env.local_rank, env.dev
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
    env.attach(ctx)
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
        ctx.comm_allreduce(amp.data_ptr(), amp.numel())  # A_lm(q, t) are sums over atoms: complete them before the DSP
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

        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        for i in range(min(args.warmup, 1)):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            e2e_step(args.warmup + i)
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te[0])
        e2e = {"value": evals_step * e2e_steps / e2e_s, "unit": "evals/s", "steps": e2e_steps,
               "h2d_bytes_per_step": int(NF * a_cnt * 12 + MP_BATCH * a_cnt * 8 + NM * 16),
               "d2h_bytes_per_step": int(MP_BATCH * (NF * 16 + 32)), "ms_per_step": 1e3 * e2e_s / e2e_steps,
               "note": "every rank re-stages its atoms of all frames from pinned host memory every step (chunked async H2D), "
                       "converts them to (r, phi, theta) on the device, uploads the factors of the pass's |q| and reads fqt/fq/fq2 back"}

    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as o  # the checker: rank 0, N = 1 only
        o.build()
        cores = host_cores()
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
                         "frac": achieved / fp64_peak, "traffic": None, "kernel": f"multipole_gemm_kernel<{MP_BATCH}>",
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
    else:
        line = None
    host.free()
    ctx.close()
    return line


def reference_line(args):
    """--impl reference: rank 0 alone runs the CPU arm and prints; the other ranks exit without work"""
    from sassena_b200 import synth

    def one(a):
        cfg_name = "C5" if a.workload == "C5s" else a.workload
        a = argparse.Namespace(**vars(a))
        a.workload = cfg_name
        kind = synth.CONFIGS[cfg_name]["kind"]
        if kind == "self":
            return run_reference_self(a)
        if kind == "mpsphere":
            return run_reference_mpsphere(a)
        return run_reference(a)

    if args.workload != "all":
        return one(args)
    line = one(sub_args(args, "C3", steps=args.steps, warmup=args.warmup, cpu_seconds=args.cpu_seconds))
    subs = {}
    for name in ("C2", "C4", "C5s"):
        if name in args.skip:
            continue
        a = sub_args(args, name, warmup=1)
        a.steps = min(a.steps, 2)
        subs[name] = one(a)
    line["workloads"] = subs
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all"] + sorted(WORKLOADS),
                    help="all (default): config 3 as the headline line plus C3 on equally spaced |q|, C2, C4 and the streamed C5 "
                         "sample under `workloads`; or one workload alone")
    ap.add_argument("--skip", default="", help="comma-separated secondary workloads to leave out of --workload all")
    ap.add_argument("--frames", type=int, default=0, help="override NF (debug; changes the workload)")
    ap.add_argument("--atoms", type=int, default=0, help="override NA (debug; changes the workload)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work per oracle sample, seconds per core")
    ap.add_argument("--mode", default="scan-rounded", choices=["scan", "scan-rounded", "per-q"],
                    help="scan-rounded (default): a step is the whole scan with the |q| list of the reference's generator "
                         "(float-rounded fractions -> corrected scan kernel); scan: exactly equally spaced |q| (plain scan "
                         "kernel); per-q: a step is one |q| through the general kernel")
    ap.add_argument("--shard", default="frames", choices=["frames", "vectors"],
                    help="N>1: shard the frames (reference decomposition, default) or the subvectors (replicated coordinates)")
    ap.add_argument("--e2e-steps", type=int, default=5, help="steps of the end-to-end (host buffers) measurement, at most --steps")
    ap.add_argument("--nq", type=int, default=2, help="--workload C5: |q| values evaluated (of the configuration's 20)")
    ap.add_argument("--c5s-atoms", type=int, default=8192, help="atoms of the streamed sample workload C5s (all ranks together)")
    ap.add_argument("--wave-atoms", type=int, default=0, help="C5 / C5s: atoms per streamed wave (0: chosen from the frame count)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.skip = [x for x in args.skip.split(",") if x]
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 3) if not (args.frames or args.atoms) else args.warmup
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        print(json.dumps(reference_line(args)))
        return 0
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
