"""CPU check of the epilogue's GELU evaluation (csrc/common.cuh: gelu_erf) against float64.

Emulates the fp32 instruction sequence (every fma rounded once to fp32) for the current form and candidate
cheaper forms, and prints max absolute / relative-to-max(1,|x|) errors.  No GPU needed."""
import numpy as np
from scipy.special import erfc

f32 = np.float32
C = [-3.6413832276593894e-05, 0.000372989394236356, -0.0012582261115312576, -0.0011454012710601091,
     0.02857113443315029, -0.1486237645149231, -0.9183861017227173, -1.62791109085083, -0.9999999403953552]


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def poly(t, coef):
    p = np.full_like(t, f32(coef[0]))
    for c in coef[1:]:
        p = fma(p, t, np.full_like(t, f32(c)))
    return p


def ex2(p):
    return np.exp2(p.astype(np.float64)).astype(f32)


def gelu_old(x):
    t = np.minimum((np.abs(x) * f32(0.70710678118654752440)).astype(f32), f32(4.2))
    h = ex2(poly(t, C))
    phi = np.where(x < 0, h, (f32(1.0) - h).astype(f32))
    return (x * phi).astype(f32)


def gelu_new(x, scaled):
    if scaled:
        s = 0.70710678118654752440
        coef = [c * s ** (8 - i) for i, c in enumerate(C)]
        t = np.minimum(np.abs(x), f32(4.2 / s))
    else:
        coef = C
        t = np.minimum((np.abs(x) * f32(0.70710678118654752440)).astype(f32), f32(4.2))
    h = ex2(poly(t, coef))
    return fma(-np.abs(x), h, np.maximum(x, f32(0)))     # relu(x) - |x| h


if __name__ == "__main__":
    x = np.concatenate([np.linspace(-8, 8, 4_000_001), np.random.default_rng(0).standard_normal(2_000_000) * 2]).astype(f32)
    ref = x.astype(np.float64) * 0.5 * erfc(-x.astype(np.float64) / np.sqrt(2.0))
    erf_form = (x * (f32(0.5) * (f32(1) + np.vectorize(lambda v: v)(__import__("scipy.special").special.erf((x * f32(0.7071067811865476)).astype(f32)).astype(f32))).astype(f32)).astype(f32)).astype(f32)
    for name, y in (("erf form fp32", erf_form), ("current", gelu_old(x)), ("relu-|x|h", gelu_new(x, False)),
                    ("relu-|x|h, scaled coef", gelu_new(x, True))):
        e = np.abs(y.astype(np.float64) - ref)
        print(f"{name:26s} max abs {e.max():.3e}  max rel-to-max(1,|x|) {(e / np.maximum(1, np.abs(x))).max():.3e}")
