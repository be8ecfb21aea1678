"""Exhaustive GPU checks of the branch-free float32 helpers the integrate kernel uses in place of the
compiler's IEEE sequences: they must be bit-identical to the IEEE round-to-nearest results."""
import ctypes

import pytest

pytestmark = pytest.mark.gpu


def test_fast_reciprocal_is_ieee_exact_over_the_normal_range(cuda_device):
    from mq3d_b200 import _lib
    bad = ctypes.c_ulonglong(123)
    # every float in [2^-126, 2^126], both signs (~4.2e9 values)
    _lib.check(_lib.lib().mq3d_selftest_rcp(0x00800000, 0x7E800000, ctypes.byref(bad)))
    assert bad.value == 0
