"""Host logic of the legacy ABI's transfer split (csrc/legacy_pipeline.cuh, plan_hybrid): which quarters of each
page-locked tensor go through the staging threads (bf16 on the wire) instead of direct fp32 DMA.  No GPU needed: the
library loads and the function makes no CUDA call.  Model: per direction, a direct tensor costs 1, a staged one 0.5;
the host costs `cost` per staged tensor; quarters move to the host while that shortens max(H2D, D2H, host)."""
import ctypes

import pytest

import flashattn_b200 as fb


def plan(in_pageable, out_pageable, cost):
    lib = fb._lib.load("flashattention_kernel")
    I, O = (ctypes.c_int * max(1, len(in_pageable))), (ctypes.c_int * max(1, len(out_pageable)))
    ip, op, iq, oq = I(*in_pageable), O(*out_pageable), I(), O()
    rc = lib.fa_plan_transfer_preview(len(in_pageable), ip, len(out_pageable), op, cost, iq, oq)
    assert rc == 0
    return list(iq)[:len(in_pageable)], list(oq)[:len(out_pageable)]


def modelled_time(iq, oq, cost):
    t_in = sum(1 - q / 8 for q in iq)
    t_out = sum(1 - q / 8 for q in oq)
    return max(t_in, t_out, cost * (sum(iq) + sum(oq)) / 4)


def test_forward_with_page_locked_inputs_balances_link_and_host():
    iq, oq = plan([0, 0, 0], [0], 0.8)        # Q K V up, O down: measured host cost on the 16-core box
    assert oq == [0]                          # the short direction stays direct
    assert iq == [4, 4, 1]                    # one tensor at a time, 9 quarters: 3 - 9/8 = 1.875 vs host 1.8
    assert modelled_time(iq, oq, 0.8) < modelled_time([0, 0, 0], [0], 0.8)
    assert modelled_time(iq, oq, 0.8) < modelled_time([4, 4, 4], [0], 0.8)


def test_backward_counts_the_pageable_dO_against_the_host_budget():
    iq, oq = plan([1], [0, 0, 0], 0.8)        # dO from numpy (staged throughout), dQ dK dV into page-locked outputs
    assert iq == [4]
    assert sum(oq) < 9                        # less than with a page-locked dO: the host already narrows dO
    iq2, oq2 = plan([0], [0, 0, 0], 0.8)
    assert sum(oq2) == 9 and iq2 == [0]


def test_pageable_tensors_are_never_taken_off_the_staging_route():
    iq, oq = plan([1, 1, 1], [1], 0.8)
    assert iq == [4, 4, 4] and oq == [4]


@pytest.mark.parametrize("cost", [0.1, 0.4, 0.8, 1.6, 6.4, 50.0])
def test_more_expensive_host_means_fewer_staged_quarters(cost):
    iq, oq = plan([0, 0, 0], [0, 0, 0], cost)
    iq2, oq2 = plan([0, 0, 0], [0, 0, 0], cost * 2)
    assert sum(iq2) + sum(oq2) <= sum(iq) + sum(oq)
    assert all(0 <= q <= 4 for q in iq + oq)
    # never worse than all-direct under the model it optimises
    assert modelled_time(iq, oq, cost) <= modelled_time([0, 0, 0], [0, 0, 0], cost) + 1e-12


def test_few_threads_keep_everything_direct():
    # 8 ranks on a 16-core host: 2 threads per rank -> cost 0.8 * 16 / 2 = 6.4 per tensor
    iq, oq = plan([0, 0, 0], [0], 6.4)
    assert sum(iq) <= 1 and oq == [0]


def test_disabled_and_bad_arguments():
    assert plan([0, 0, 0], [0], -1.0) == ([0, 0, 0], [0])
    lib = fb._lib.load("flashattention_kernel")
    assert lib.fa_plan_transfer_preview(17, None, 0, None, 0.8, None, None) != 0
