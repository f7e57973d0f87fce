"""Model of the MMA-issuer schedule of ``gemm_mxfp8_2cta_split_kernel`` (csrc/kernels/gemm_mxfp8.cu): three rotating
128-column accumulators, half 0 of a 256-wide tile LAG K blocks ahead at the start of a tile and LAG early at its end.
The constants are read from the source; the checks are the invariants the kernel's barriers rely on."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "csrc", "kernels", "gemm_mxfp8.cu")).read()


def _const(name):
    m = re.search(r"struct Mx8SplitCfg \{.*?\b%s = (\d+)" % name, SRC, re.S)
    assert m, name
    return int(m.group(1))


STAGES, LAG, NACC = _const("STAGES"), _const("LAG"), _const("NACC")


def schedule(num_k):
    """Issue order of one tile: list of (half, kb), plus the position at which the second buffer is waited for."""
    ev, wait_b1_at = [], None
    if num_k >= 2 * LAG:
        ev += [(0, kb) for kb in range(LAG)]
        wait_b1_at = len(ev)
        ev += [(1, kb) for kb in range(LAG)]
        for kb in range(LAG, num_k - LAG):
            ev += [(0, kb), (1, kb)]
        ev += [(0, kb) for kb in range(num_k - LAG, num_k)]
        ev += [(1, kb) for kb in range(num_k - LAG, num_k)]
    else:
        wait_b1_at = 0
        for kb in range(num_k):
            ev += [(0, kb), (1, kb)]
    return ev, wait_b1_at


def test_constants_fit_tensor_memory_and_the_stage_ring():
    assert NACC * 128 + 12 * STAGES <= 512           # accumulators + per-stage scale-factor columns
    assert LAG + 1 < STAGES                          # stages held by the lagging half leave room for the TMA producer
    assert NACC == 3


@pytest.mark.parametrize("num_k", [1, 2, 5, 6, 7, 24, 120])
def test_every_k_block_once_per_half_in_order_and_bounded_stage_hold(num_k):
    ev, wait_at = schedule(num_k)
    for h in (0, 1):
        assert [kb for hh, kb in ev if hh == h] == list(range(num_k))        # stage cursors advance by one per event
    issued0 = issued1 = 0
    for i, (h, kb) in enumerate(ev):
        if h == 0:
            issued0 += 1
        else:
            assert kb < issued0, "half 1 touches a stage half 0 has not waited for"
            assert i >= wait_at, "half 1 issued before its accumulator buffer was waited for"
            issued1 += 1
        assert issued0 - issued1 <= (LAG if num_k >= 2 * LAG else 1) <= STAGES - 2
    last0 = max(i for i, (h, _) in enumerate(ev) if h == 0)
    last1 = max(i for i, (h, _) in enumerate(ev) if h == 1)
    if num_k >= 2 * LAG:
        assert last1 - last0 == LAG                  # half 0 is handed to the epilogue LAG half-blocks early ...
        assert wait_at == LAG                        # ... and the next tile needs that buffer LAG half-blocks after its start


def test_buffer_rotation_gives_every_tile_a_spare_and_reuses_the_oldest():
    for it in range(12):
        b0, b1 = (2 * it) % NACC, (2 * it + 1) % NACC
        n0, n1 = (2 * (it + 1)) % NACC, (2 * (it + 1) + 1) % NACC
        assert n0 not in (b0, b1)                    # next tile's half 0 starts in the spare buffer at once
        assert n1 == b0                              # its half 1 takes the buffer that finished first
    # use counts -> barrier parities seen by MMA issuer and epilogue agree
    uses = [0] * NACC
    for u in range(40):
        buf = u % NACC
        assert uses[buf] == u // NACC
        uses[buf] += 1
