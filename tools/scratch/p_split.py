import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import sassena_b200
from sassena_b200 import synth
NF, NA, NM = int(sys.argv[1]), int(sys.argv[2]), 200
ctx = sassena_b200.ScatterContext(0)
d = ctx.device_alloc(NA * NF * 12)
ctx.synth_trajectory(d, NF, 30000, 70.0, 0.05, 3, layout=1, NA_out=NA)
h = np.empty((NA, NF, 3), dtype=np.float32); ctx.memcpy_d2h(h, d); ctx.device_free(d)
ctx.stage_atoms(h)
ctx.set_factors(synth.factors(NA))
q = 1.0 * synth.unit_vectors(NM, 4)
ctx.compute_self_vectors(q)
