#!/usr/bin/env python
"""Coefficients and error report for a polynomial 2^x on the FMA pipe (taking part of the softmax's
`ex2` off the XU pipe, 16 lanes/clk/SM on B200).  Runs on the CPU; emulates the fp32 instruction sequence

    t = max(x, -126)                     FMNMX
    r = t + 12582912.0f                  FADD   (1.5 * 2^23: round-to-nearest-even integer n in the low mantissa bits)
    f = t - (r - 12582912.0f)            FADD, FADD     f in [-0.5, 0.5]
    p = c0 + f*(c1 + f*(c2 + ... ))      DEG x FFMA
    y = as_float(as_int(p) + (as_int(r) << 23))     SHL, IADD   (p in [0.70, 1.42]: the exponent add cannot carry wrongly)

and compares with float64 2^x over the range a softmax produces (x <= 0).

    python tools/exp2_poly.py [degree]
"""
import sys

import numpy as np


def fit(deg: int) -> np.ndarray:
    # Chebyshev-node least squares on [-0.5, 0.5] of 2^f, then a few Remez-like reweighting rounds on the relative error
    k = np.arange(4096)
    f = 0.5 * np.cos(np.pi * (k + 0.5) / 4096)
    w = np.ones_like(f)
    for _ in range(30):
        A = np.vander(f, deg + 1, increasing=True) * (w / 2.0 ** f)[:, None]
        c, *_ = np.linalg.lstsq(A, w, rcond=None)
        err = np.abs(np.vander(f, deg + 1, increasing=True) @ c / 2.0 ** f - 1.0)
        w = w * (1.0 + 0.5 * err / err.max())
    return c


def emulate(x: np.ndarray, c: np.ndarray) -> np.ndarray:
    f32 = np.float32
    t = np.maximum(x.astype(f32), f32(-126.0))
    magic = f32(12582912.0)
    r = (t + magic).astype(f32)
    n = (r.view(np.int32) & 0x7FFFFF) - 0x400000             # low mantissa bits of 1.5*2^23 + n
    f = (t - (r - magic).astype(f32)).astype(f32)
    p = np.full_like(f, f32(c[-1]))
    for ck in c[-2::-1]:
        p = (p.astype(np.float64) * f.astype(np.float64) + np.float64(f32(ck))).astype(f32)   # one rounding per FFMA
    y = (p.view(np.int32) + (n.astype(np.int32) << 23)).view(f32)
    return y


def main() -> int:
    deg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    rng = np.random.default_rng(0)
    x = np.concatenate([-rng.random(2_000_000) * 30.0, -rng.random(200_000) * 126.0, np.linspace(-1, 0, 100001), [0.0, -0.5, -1.5]])
    ref = np.exp2(x.astype(np.float32).astype(np.float64))
    for d in sorted({3, 4, 5, deg}):
        c = fit(d)
        y = emulate(x, c).astype(np.float64)
        rel = np.abs(y / ref - 1.0)
        print(f"degree {d}: max rel err {rel.max():.3e}  (tf32 rounding of P is 2.4e-4, fp32 ulp 6e-8)")
        print("   coefficients c0..c%d: " % d + ", ".join(f"{np.float32(v)!r}".replace("np.float32(", "").rstrip(")") + "f" for v in c))
    return 0


if __name__ == "__main__":
    sys.exit(main())
