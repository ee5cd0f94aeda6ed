"""Error of the three-term twiddle recurrences of self_split_fft12_kernel (selffused.cu, twiddle16): powers w^k, k < 16, of
W_4096^tid / W_256^n3 / W_L^(256 r) generated as w_{k+2} = 2 cos(2 theta) w_k - w_{k-2} from w^0..w^3."""
import numpy as np


def powers(z0, z1, cg2):
    z = [z0, z1, cg2 * z1 - z0]
    z.append(cg2 * z[2] - z1)
    C2 = cg2 * cg2 - 2.0
    for k in range(4, 16):
        z.append(C2 * z[k - 2] - z[k - 4])
    return z


worst = 0.0
for N, ids in ((4096, range(256)), (256, range(16))):
    for t in ids:
        w = np.exp(-2j * np.pi * t / N)
        z = powers(1.0 + 0j, w, 2 * w.real)
        worst = max(worst, max(abs(z[k] - np.exp(-2j * np.pi * t * k / N)) for k in range(16)))
for R in (2, 5, 25, 64):
    L = R * 4096
    for r in range(R):
        g = np.exp(-2j * np.pi * 256 * r / L)
        for tid in (1, 77, 255):
            h = np.exp(-2j * np.pi * r * tid / L)
            z = powers(h, h * g, 2 * g.real)
            worst = max(worst, max(abs(z[k] - h * g ** k) for k in range(16)))
print("worst absolute twiddle error:", worst)
assert worst < 5e-14
