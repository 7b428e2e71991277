import numpy as np


def relerr(a, ref, scale=1.0):
    """|a - ref| / max(|ref|, scale): the tolerance metric of every floating-point parity test
    (scale = 1.0 unless a test says otherwise)."""
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return np.abs(a - ref) / np.maximum(np.abs(ref), scale)
