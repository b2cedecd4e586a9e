import glob
import os

import numpy as np

G = os.path.join(os.path.dirname(__file__), "golden")


def golden(pattern):
    files = sorted(glob.glob(os.path.join(G, pattern)))
    assert files, pattern
    return files


def have_gpu():
    try:
        import flashattn_b200 as fb
        lib = fb._lib.load("flashattention_kernel")
        return lib.fa_device_count() > 0
    except Exception:
        return False


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))
