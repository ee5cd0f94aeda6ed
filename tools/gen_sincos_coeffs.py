"""Generate the polynomial coefficients used by the amplitude kernels (sassena_b200/csrc/kernels/sincos_qt.cuh).

The kernels work in quarter-turn units: u = phase*(2/pi), k = rint(u), f = u-k in [-1/2,1/2],
  sin(pi/2 f) = f * S(f^2),   cos(pi/2 f) = 1 + f^2 * C(f^2).
S and C are near-minimax polynomials in s=f^2 on [0,1/4]: Chebyshev-node interpolation evaluated with mpmath
(50 digits), converted to the monomial basis and rounded to double.  Prints C arrays as hex floats and the
max abs error of the rounded polynomials.
"""
import sys
import mpmath as mp

mp.mp.dps = 60


def cheb_fit(func, deg, a, b):
    n = deg + 1
    nodes = [mp.cos(mp.pi * (2 * i + 1) / (2 * n)) for i in range(n)]
    xs = [(a + b) / 2 + (b - a) / 2 * t for t in nodes]
    ys = [func(x) for x in xs]
    # solve Vandermonde in high precision (small n)
    V = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            V[i, j] = xs[i] ** j
    c = mp.lu_solve(V, mp.matrix(ys))
    return [c[i] for i in range(n)]


def S_exact(s):
    if s == 0:
        return mp.pi / 2
    f = mp.sqrt(s)
    return mp.sin(mp.pi / 2 * f) / f


def C_exact(s):
    if s == 0:
        return -(mp.pi / 2) ** 2 / 2
    f = mp.sqrt(s)
    return (mp.cos(mp.pi / 2 * f) - 1) / s


def maxerr(coefs, kind):
    worst = mp.mpf(0)
    for i in range(0, 2001):
        f = mp.mpf(i) / 4000
        s = f * f
        p = mp.mpf(0)
        for c in reversed(coefs):
            p = p * s + mp.mpf(c)
        val = f * p if kind == "sin" else 1 + s * p
        ex = mp.sin(mp.pi / 2 * f) if kind == "sin" else mp.cos(mp.pi / 2 * f)
        worst = max(worst, abs(val - ex))
    return worst


def main():
    for name, fn, kind, degs in (("S", S_exact, "sin", (4, 5, 6)), ("C", C_exact, "cos", (4, 5, 6))):
        for deg in degs:
            c = cheb_fit(fn, deg, mp.mpf(0), mp.mpf(1) / 4)
            cd = [float(x) for x in c]
            err = maxerr(cd, kind)
            print(f"// {kind}: degree {deg} in f^2, max abs err {mp.nstr(err, 3)}")
            print(f"static constexpr double {name}{deg}[{deg + 1}] = {{" + ", ".join(x.hex() for x in cd) + "};")
            print("//   = " + ", ".join(repr(x) for x in cd))


if __name__ == "__main__":
    main()
