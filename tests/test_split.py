"""kwargs splitting / gather rules (ADP:1210-1285) incl. the deliberate fixes."""
import torch

from comfyui_parallelanything_b200.parallel import split as sp


def test_batch_size():
    assert sp.get_batch_size(torch.zeros(5, 2)) == 5
    assert sp.get_batch_size([torch.zeros(3, 2), torch.zeros(3)]) == 3
    assert sp.get_batch_size(["a", "b"]) == 2
    assert sp.get_batch_size(7) == 1


def test_split_value_and_concat_roundtrip():
    x = torch.arange(12.).view(6, 2)
    parts = sp.split_value(x, [4, 2])
    assert [p.shape[0] for p in parts] == [4, 2]
    tup = sp.split_value((x, "k", x + 1), [1, 5])
    assert tup[1][1] == "k" and tup[1][2].shape[0] == 5
    assert torch.equal(sp.concatenate_results(parts), x)
    merged = sp.concatenate_results([(p, "m") for p in parts])
    assert torch.equal(merged[0], x) and merged[1] == "m"


def test_kwargs_rules():
    B = 4
    kw = dict(y=torch.randn(B, 3), scalar=2.5, shared=torch.randn(7),
              lst=[torch.randn(B, 1), torch.randn(B, 2)],
              ragged=[torch.randn(B, 1), torch.randn(3, 2)],
              control={"input": [torch.randn(B, 2)], "scale": 0.5},
              transformer_options={"cond_or_uncond": [0, 1]})
    out = sp.split_kwargs(kw, [3, 1], B)
    assert out[0]["y"].shape[0] == 3 and out[1]["y"].shape[0] == 1
    assert out[0]["scalar"] == 2.5 and out[1]["shared"] is kw["shared"]
    assert out[1]["lst"][1].shape == (1, 2)
    # fix: ragged lists are replicated per element, not dropped
    assert out[0]["ragged"][0].shape[0] == 3 and out[0]["ragged"][1].shape[0] == 3
    # fix: dicts are recursed into
    assert out[1]["control"]["input"][0].shape[0] == 1 and out[1]["control"]["scale"] == 0.5
    assert out[0]["transformer_options"] == {"cond_or_uncond": [0, 1]}
    # strict compat reproduces the reference: ragged key dropped, dict by reference
    c = sp.split_kwargs(kw, [3, 1], B, strict_compat=True)
    assert "ragged" not in c[0] and c[0]["control"] is kw["control"]


def test_output_like_write_rows():
    first = torch.ones(2, 3)
    buf = sp.output_like(first, 5, torch.device("cpu"))
    sp.write_rows(buf, first, 0)
    sp.write_rows(buf, torch.full((3, 3), 2.0), 2)
    assert buf[:2].eq(1).all() and buf[2:].eq(2).all()


def test_move_recurses_dicts():
    d = {"a": [torch.zeros(1)], "b": {"c": torch.zeros(2)}}
    m = sp.move_to_device(d, "cpu")
    assert m["b"]["c"].device.type == "cpu"
