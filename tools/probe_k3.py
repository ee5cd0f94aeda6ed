"""K3 standalone (corr_colfft + corr_rowfft + reductions) on a batch of complete timelines — the DSP of the coherent and
multipole devices.  Shapes from the environment: NM timelines of NF frames (default BASELINE config 4's 8 x 441 = 3528 x 1000).
Prints the time per batch and the achieved HBM rate against the algorithmic bytes (16 NF read per timeline)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import sassena_b200  # noqa: E402
from sassena_b200 import synth  # noqa: E402


def main():
    NM = int(os.environ.get("NM", 3528))
    NF = int(os.environ.get("NF", 1000))
    reps = int(os.environ.get("REPS", 20))
    dev = torch.device("cuda", 0)
    ctx = sassena_b200.ScatterContext(0)
    ctx.stage_frames(synth.trajectory(NF, 4, 30.0, 0.1, 1))
    ctx.set_factors(synth.factors(4))
    amp = torch.randn(NM * NF * 2, dtype=torch.float64, device=dev)
    plen = ctx.partial_len("autocorrelate")
    partial = torch.zeros(plen, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    for _ in range(3):
        ctx.all_vectors_dsp_partial(amp.data_ptr(), 0, NM, partial.data_ptr())
    ctx.synchronize()
    ctx.timer_start()
    for _ in range(reps):
        ctx.all_vectors_dsp_partial(amp.data_ptr(), 0, NM, partial.data_ptr())
    ms = ctx.timer_stop() / reps
    alg = 16.0 * NF * NM
    print(f"K3 NM={NM} NF={NF}: {ms:.4f} ms per batch, algorithmic {alg / 1e6:.1f} MB -> {alg / (ms * 1e-3) / 1e9:.1f} GB/s")


if __name__ == "__main__":
    main()
