"""Quick device sanity/perf probe (not the bench): FP64 peak and K1 throughput on a cfg3-shaped slice."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sassena_b200
from sassena_b200 import synth

ctx = sassena_b200.ScatterContext(0)
print("fp64 peak TFLOP/s:", ctx.measure_fp64_peak())
NA, NF, NM = 100000, int(sys.argv[1]) if len(sys.argv) > 1 else 2000, 500
d = ctx.device_alloc(NF * NA * 12)
ctx.synth_trajectory(d, NF, NA, 100.0, 0.05, 5)
ctx.stage_frames_device(d, NF, NA)
ctx.set_factors(synth.factors(NA))
q = 2.0 * synth.unit_vectors(NM, 6)
for it in range(3):
    t0 = time.time()
    fqt, fq, fq2 = ctx.compute_all_vectors(q)
    dt = time.time() - t0
    amp = ctx.last_amplitude_ms()
    print(f"iter {it}: wall {dt*1e3:.1f} ms  amplitude {amp:.2f} ms  dsp {ctx.last_dsp_ms():.2f} ms  "
          f"evals/s (kernel) {NA*NF*NM/(amp*1e-3):.3e}  fq0={fqt[0]:.6e}")
