"""Host side of the legacy ABI's bf16 wire format (csrc/host_convert.cpp, linked into flashattention_kernel.so):
fp32 -> bf16 narrowing must be bit-identical to round-to-nearest-even (what cvt.rn.bf16.f32 does on the device and what
the oracle's round_bf16 restates), NaN must stay NaN, +-inf / max-finite / denormals must behave, and widening is the
exact inverse on bf16 values.  Runs without a GPU (the library loads; no CUDA call is made)."""
import ctypes

import numpy as np

from flashattn_b200 import _lib
from flashattn_b200 import device as dev


def _lib_fns():
    lib = ctypes.CDLL(_lib.KERNEL_DIR + "/flashattention_kernel.so")
    lib.fa_host_narrow_f32_bf16.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    lib.fa_host_widen_bf16_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    lib.fa_host_convert_isa.restype = ctypes.c_int
    return lib


def test_narrow_matches_round_to_nearest_even_and_widen_inverts():
    lib = _lib_fns()
    rng = np.random.default_rng(0)
    n = (1 << 18) + 13                      # not a multiple of the 16-element vector body
    bits = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    special = np.array([0x7F800000, 0xFF800000, 0x7F7FFFFF, 0x00000001, 0x80000001, 0x3F808000, 0x3F818000, 0x3F807FFF,
                        0x7FC00001, 0xFFC00000, 0x7F800001, 0x00000000, 0x80000000], dtype=np.uint32)
    bits[:special.size] = special
    x = bits.view(np.float32)
    got = np.empty(n, dtype=np.uint16)
    lib.fa_host_narrow_f32_bf16(got.ctypes.data, x.ctypes.data, n)
    finite = ~np.isnan(x)
    want = dev.to_bf16_bits(x)
    np.testing.assert_array_equal(got[finite], want[finite])
    back = dev.from_bf16_bits(got)
    assert np.all(np.isnan(back[~finite]))            # NaN stays NaN (never rounds up into infinity or -0)
    assert np.isposinf(back[0]) and np.isneginf(back[1]) and np.isposinf(back[2])   # max finite rounds to +inf (RNE)
    wide = np.empty(n, dtype=np.float32)
    lib.fa_host_widen_bf16_f32(wide.ctypes.data, got.ctypes.data, n)
    np.testing.assert_array_equal(wide.view(np.uint32), got.astype(np.uint32) << 16)
    assert lib.fa_host_convert_isa() in (0, 2)
