"""K1 throughput vs number of q-vectors per launch (sharding granularity study)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sassena_b200
from sassena_b200 import synth
ctx = sassena_b200.ScatterContext(0)
NA, NF = 100000, 2000
d = ctx.device_alloc(NF * NA * 12)
ctx.synth_trajectory(d, NF, NA, 100.0, 0.05, 5)
ctx.stage_frames_device(d, NF, NA)
ctx.set_factors(synth.factors(NA))
u = synth.unit_vectors(500, 6)
for NM in [int(x) for x in sys.argv[1:]] or [48, 62, 63, 96, 125, 144, 250, 480, 500]:
    q = 2.0 * u[:NM]
    best = 1e9
    for it in range(3):
        ctx.compute_all_vectors(q)
        best = min(best, ctx.last_amplitude_ms())
    print(f"NM={NM:4d}: amp {best:8.2f} ms  {NA*NF*NM/(best*1e-3):.3e} evals/s")
