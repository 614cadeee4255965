"""Feasibility study (numpy, CPU) of the certified fixed-point route for the integer passes:
accumulate sum(v_k * round(w_k * 2^S)) exactly in int64, bound |acc_scipy - acc_int / 2^S| from the
weights, take trunc() from the integer sum wherever its fraction is farther than the bound from an integer.
Reports the bound, the share of ambiguous outputs, and checks that every unambiguous output equals scipy's."""
import sys
from fractions import Fraction

import numpy as np
import scipy.ndimage as ndi

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from oracle import oracle  # noqa: E402


def study(dtype, sigma, order, data, axis=0):
    radius = int(4.0 * sigma + 0.5)
    w = oracle.gaussian_kernel1d(sigma, order, radius)[::-1]
    L = len(w)
    M = int(np.iinfo(dtype).max)
    wmax = float(np.abs(w).max())
    S = 0
    while wmax * 2.0 ** (S + 1) < 2 ** 31 - 1 and S < 52:
        S += 1
    wq = np.array([int(round(Fraction(float(x)) * 2 ** S)) for x in w], dtype=np.int64)
    assert np.abs(wq).max() < 2 ** 31
    # rigorous bound, exact rationals: quantisation + scipy's own rounding (gamma_n * sum |v w|)
    quant = sum(abs(Fraction(float(x)) - Fraction(int(q), 2 ** S)) for x, q in zip(w, wq)) * M
    n = L + 1
    gamma = Fraction(n, 2 ** 53) / (1 - Fraction(n, 2 ** 53))
    own = gamma * sum(abs(Fraction(float(x))) for x in w) * M
    eps = quant + own
    delta = int(eps * 2 ** S) + 2                      # in units of 2^-S
    ref = ndi.correlate1d(data, w, axis=axis, mode="reflect")            # scipy: the truth
    ext = np.pad(data.astype(np.int64), [(radius, radius) if a == axis else (0, 0) for a in range(data.ndim)],
                 mode="symmetric")
    acc = np.zeros(data.shape, dtype=np.int64)
    for k in range(L):
        sl = [slice(None)] * data.ndim
        sl[axis] = slice(k, k + data.shape[axis])
        acc += ext[tuple(sl)] * wq[k]
    a = np.abs(acc)
    q, f = a >> S, a & ((1 << S) - 1)
    amb = ((f < delta) | (f > (1 << S) - delta)) & ~((q == 0) & (f < delta))
    val = np.where(acc < 0, -q, q)
    got = val.astype(np.int32).astype(dtype)                              # the cast's wrap for u16/u8
    ok = np.array_equal(got[~amb], ref[~amb])
    return dict(dtype=np.dtype(dtype).name, sigma=sigma, order=order, taps=L, S=S, eps=float(eps),
                ambiguous=float(amb.mean()), unambiguous_equal=bool(ok),
                wrong_if_ignored=int((got != ref).sum()))


def main():
    rng = np.random.default_rng(0)
    for dtype in (np.uint8, np.uint16):
        M = np.iinfo(dtype).max
        noisy = rng.integers(0, M + 1, (256, 512)).astype(dtype)
        cam = np.clip(rng.normal(0.1 * M, 0.02 * M, (256, 512)), 0, M).astype(dtype)   # offset + noise
        sat = np.full((256, 512), M, dtype)                                            # saturated region
        for name, d in (("uniform noise", noisy), ("camera-like", cam), ("saturated", sat)):
            for sigma, order in ((1.0, 0), (2.0, 0), (4.0, 0), (2.0, 1)):
                r = study(dtype, sigma, order, d)
                print("%-14s %s" % (name, r), flush=True)


if __name__ == "__main__":
    main()
