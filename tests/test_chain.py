"""Weight normalisation + split arithmetic ([PROBE] vectors)."""
import pytest
from hypothesis import given, settings, strategies as st

from comfyui_parallelanything_b200 import chain


def test_normalize_need_not_sum_to_100():
    assert chain.normalize_weights([50, 50]) == [0.5, 0.5]
    assert chain.normalize_weights([10, 30]) == [0.25, 0.75]
    assert chain.normalize_weights([0, 0]) == [0.5, 0.5]
    assert chain.normalize_weights([]) == []


@pytest.mark.parametrize("pcts,batch,expect", [
    ([12.5] * 8, 16, [2] * 8),
    ([40, 40, 15, 5], 32, [12, 12, 4, 4]),
    ([40, 40, 15, 5], 21, [8, 8, 3, 2]),
    ([50, 50], 21, [10, 11]),
])
def test_compat_vectors(pcts, batch, expect):
    w = chain.normalize_weights(pcts)
    assert chain.split_sizes(batch, w, "compat") == expect


def test_compat_repairs_negative_remainder():
    # reference arithmetic gives [3,1,1,-1] here and then crashes in torch.split
    w = [0.9, 0.03, 0.03, 0.04]
    assert chain.split_compat(4, w) == [3, 1, 1, -1]
    s = chain.split_sizes(4, w, "compat")
    assert sum(s) == 4 and min(s) >= 0


def test_exact_is_largest_remainder():
    assert chain.split_exact(32, chain.normalize_weights([40, 40, 15, 5])) == [13, 13, 5, 1]
    assert chain.split_exact(3, [0.5, 0.5]) == [2, 1]


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 257), st.lists(st.floats(0.01, 100.0), min_size=1, max_size=8),
       st.sampled_from(["compat", "exact"]))
def test_split_conserves_batch(batch, pcts, mode):
    s = chain.split_sizes(batch, chain.normalize_weights(pcts), mode)
    assert sum(s) == batch and min(s) >= 0 and len(s) == len(pcts)


def test_vram_adjust_conserves_batch():
    free = {"cuda:0": 1000.0, "cuda:1": 3000.0}.get
    w = chain.vram_adjusted_weights(["cuda:0", "cuda:1"], [0.5, 0.5], lambda d: free(d, 0.0))
    assert abs(sum(w) - 1) < 1e-9 and w[1] > w[0]
    assert w == pytest.approx([0.7 * 0.5 + 0.3 * 0.25, 0.7 * 0.5 + 0.3 * 0.75])
    s = chain.split_sizes_vram(21, ["cpu", "cpu", "cpu", "cpu"], chain.normalize_weights([40, 40, 15, 5]))
    assert sum(s) == 21            # reference early-out returns [8,8,3,1] (sum 20)


def test_assign_blocks():
    assert chain.assign_blocks(19, [0.5, 0.5]) == [0] * 10 + [1] * 9
    assert chain.assign_blocks(4, [0.25, 0.25, 0.25, 0.25]) == [0, 1, 2, 3]
    owners = chain.assign_blocks(38, chain.normalize_weights([40, 40, 15, 5]))
    assert len(owners) == 38 and owners == sorted(owners)


def test_parse_chain_roundtrip():
    c = [chain.make_entry("cpu", 30), {"device": "cuda:1", "weight": 0.7}, ("cpu", 5)]
    e = chain.parse_chain(c)
    assert [x.device for x in e] == ["cpu", "cuda:1", "cpu"]
    assert e[1].percentage == pytest.approx(70.0)
    assert chain.validate_devices(["cpu", "not-a-device"]) == "not-a-device"
