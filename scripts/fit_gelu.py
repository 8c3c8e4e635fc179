"""Fit of the single-exp2 erf-GELU used in the GEMM epilogue (csrc/common.cuh: gelu_erf_fast).
gelu(x) = max(x,0) - |x| * 2^Q(min(|x|,6)),  Q ~ log2(0.5 erfc(t / sqrt 2)), weighted so the ABSOLUTE gelu error is minimised."""
import numpy as np
from scipy.special import erf, erfc

T, DEG = 6.0, 6
t = np.linspace(0, T, 20001)
Q = np.log2(0.5 * erfc(t / np.sqrt(2)))
w = np.maximum(t, 1e-3) * np.exp2(Q)
ww = w.copy()
V = np.vander(t, DEG + 1, increasing=True)
for _ in range(60):                      # iteratively re-weighted least squares -> approximately minimax
    c, *_ = np.linalg.lstsq(V * ww[:, None], Q * ww, rcond=None)
    err = (V @ c - Q) * w
    ww = ww * (1 + 4 * np.abs(err) / np.abs(err).max())
    ww /= ww.max()
c32 = c.astype(np.float32)
x = np.linspace(-8, 8, 400001).astype(np.float32)
a = np.minimum(np.abs(x), np.float32(T))
q = np.full_like(a, c32[-1])
for k in range(DEG - 1, -1, -1):
    q = (q * a + c32[k]).astype(np.float32)
g = np.maximum(x, 0) - np.abs(x) * np.exp2(q.astype(np.float64))
ref = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
print("coefficients (c0..c6):", [float(v) for v in c32])
print("max abs error: %.3e" % np.abs(g - ref).max())
