import numpy as np


def rel_err(got, ref):
    """norm-wise relative error max|got-ref| / max|ref| (the form the 1e-9 tolerance on fq/fqt is checked in:
    element-wise relative error is meaningless where F(q,t) passes through zero)."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    scale = np.max(np.abs(ref))
    if scale == 0:
        return float(np.max(np.abs(got)))
    return float(np.max(np.abs(got - ref)) / scale)


# north_star: "relative tolerance of 1e-9 on fq/fqt"
TOL = 1e-9
