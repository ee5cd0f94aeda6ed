"""Throughput of the |q|-scan amplitude kernels vs the per-|q| kernel K1 on a C3-shaped sample, and the deviation of every |q|
of the scan from K1 (run on the GPU box).  ROUNDED=1: the reference's float-rounded scan (corrected kernels;
SASSENA_SCAN_FP64_CORR=1 keeps the first-order sums in FP64), VARB=1: |q|-dependent factors; NA / NF / NM / NQ override the shape."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sassena_b200  # noqa: E402
from sassena_b200 import synth  # noqa: E402


def main():
    NA = int(os.environ.get("NA", 100000))
    NF = int(os.environ.get("NF", 600))
    NM = int(os.environ.get("NM", 500))
    NQ = int(os.environ.get("NQ", 32))
    dev = torch.device("cuda", 0)
    ctx = sassena_b200.ScatterContext(0)
    xyz = torch.empty(NF * NA * 3, dtype=torch.float32, device=dev)
    ctx.synth_trajectory(xyz.data_ptr(), NF, NA, 100.0, 0.05, 5)
    ctx.stage_frames_device(xyz.data_ptr(), NF, NA)
    b = synth.factors(NA)
    ctx.set_factors(b)
    u = synth.unit_vectors(NM, 6)
    if os.environ.get('VARB'):  # |q|-dependent factors: one row per |q|
        ctx.set_factors_batch(np.stack([b * (1.0 + 0.01 * n) for n in range(NQ)]))
    s0, ds = 0.1, 0.1
    sv = s0 + ds * np.arange(NQ)
    if os.environ.get('ROUNDED'):  # the reference's float-rounded scan fractions
        sv = s0 + np.float32(np.arange(NQ) / (NQ - 1)).astype(np.float64) * (ds * (NQ - 1))
    amp = torch.empty(NQ * NM * NF * 2, dtype=torch.float64, device=dev)
    for rep in range(2):
        ctx.synchronize()
        t0 = time.perf_counter()
        ctx.all_vectors_scan_amplitudes(u, sv, amp.data_ptr())
        ctx.synchronize()
        dt = time.perf_counter() - t0
        ms = ctx.last_amplitude_ms()
    evals = float(NA) * NF * NM * NQ
    print(f"scan plan {ctx.last_scan_plan()}: NQ={NQ} {evals / (ms * 1e-3):.3e} evals/s (kernel {ms:.1f} ms, wall {dt * 1e3:.1f} ms)")
    # per-|q| kernel for comparison, and every |q| of the scan against it
    a1 = torch.empty(NM * NF * 2, dtype=torch.float64, device=dev)
    worst, where = 0.0, -1
    for n in range(NQ):
        for rep in range(2 if n == NQ - 1 else 1):
            ctx.all_vectors_amplitudes(sv[n] * u, a1.data_ptr())
            ctx.synchronize()
            ms1 = ctx.last_amplitude_ms()
        torch.cuda.synchronize()
        ref = a1.view(NM, NF, 2)
        got = amp.view(NQ, NM, NF, 2)[n]
        err = float((got - ref).abs().max() / ref.abs().max())
        if err > worst:
            worst, where = err, n
    print(f"per-|q| kernel: {float(NA) * NF * NM / (ms1 * 1e-3):.3e} evals/s")
    print(f"scan vs per-|q| kernel over all {NQ} |q|: max rel diff {worst:.2e} (at |q| index {where})")


if __name__ == "__main__":
    main()
