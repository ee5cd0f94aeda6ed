import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import sassena_b200
from sassena_b200 import synth
NF, NA, NM = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
xyz = synth.trajectory(NF, NA, 30.0, 0.1, 17, layout=1)
ctx = sassena_b200.ScatterContext(0)
ctx.stage_atoms(xyz); ctx.set_factors(synth.factors(NA))
r = ctx.compute_self_vectors(1.3 * synth.unit_vectors(NM, 18))
print("ok", r[1])
