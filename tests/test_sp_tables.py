"""Sequence-parallel exchange descriptor tables (exec/flux_sp.py, exec/wan_sp.py), executed on the CPU: the tables are pure
functions of the geometry and of the ranks' buffer addresses; a byte-level interpreter of the copy descriptors stands in for
csrc/comm/sp_a2a.cu and the result is compared with the plain index arithmetic the exchange is meant to implement
([my tokens, all heads] <-> [all tokens, my heads])."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    """Import exec/<name>.py's ``exchange_tables`` without importing the CUDA-dependent rest of the module."""
    src = open(os.path.join(ROOT, "comfyui_parallelanything_b200", "exec", name + ".py")).read()
    a = src.index("def exchange_tables(")
    b = src.index("\nclass ", a)
    ns = {}
    exec(compile(src[a:b], name + ".exchange_tables", "exec"), ns)
    return ns["exchange_tables"]


class Heap:
    """One flat byte array for all ranks' buffers; 'pointers' are offsets into it."""

    def __init__(self):
        self.chunks, self.size = [], 0

    def alloc(self, n_elems):           # bf16 = 2 bytes, stored as uint16
        off = self.size
        self.size += n_elems * 2 + 64   # gaps catch out-of-bounds descriptors
        return off

    def build(self):
        self.mem = np.zeros(self.size, dtype=np.uint8)

    def view(self, off, shape):
        n = int(np.prod(shape))
        return self.mem[off:off + 2 * n].view(np.uint16).reshape(shape)

    def run(self, descs):
        for src, dst, sp, dp, rows, rb in descs:
            for r in range(rows):
                self.mem[dst + r * dp:dst + r * dp + rb] = self.mem[src + r * sp:src + r * sp + rb]


@pytest.mark.parametrize("n", [2, 4])
def test_wan_exchange_tables_transpose_tokens_and_heads(n):
    exchange_tables = _load("wan_sp")
    heads, L = 8, 64
    dim, hpg, Ll = heads * 128, heads // n, L // n
    heap = Heap()
    ptrs = [{"QKV": heap.alloc(Ll * 3 * dim), "QF": heap.alloc(L * hpg * 128), "KF": heap.alloc(L * hpg * 128),
             "VF": heap.alloc(L * hpg * 128), "ATTF": heap.alloc(L * hpg * 128), "ATT": heap.alloc(Ll * dim)} for _ in range(n)]
    heap.build()
    rng = np.random.default_rng(0)
    full_qkv = rng.integers(0, 65535, size=(L, 3, heads, 128), dtype=np.uint16)      # global [token, q/k/v, head, d]
    full_att = rng.integers(0, 65535, size=(L, heads, 128), dtype=np.uint16)         # attention output, global
    for r in range(n):
        heap.view(ptrs[r]["QKV"], (Ll, 3, heads, 128))[:] = full_qkv[r * Ll:(r + 1) * Ll]
        heap.view(ptrs[r]["ATTF"], (L, hpg, 128))[:] = full_att[:, r * hpg:(r + 1) * hpg]
    for g in range(n):
        qkv, att = exchange_tables(g, n, Ll, dim, hpg, ptrs)
        heap.run(qkv)
        heap.run(att)
    for g in range(n):
        for sec, name in enumerate(("QF", "KF", "VF")):
            got = heap.view(ptrs[g][name], (L, hpg, 128))
            assert np.array_equal(got, full_qkv[:, sec, g * hpg:(g + 1) * hpg]), (g, name)
        got = heap.view(ptrs[g]["ATT"], (Ll, heads, 128))
        assert np.array_equal(got, full_att[g * Ll:(g + 1) * Ll]), g


@pytest.mark.parametrize("n", [2, 4])
def test_flux_exchange_tables_keep_txt_img_order(n):
    exchange_tables = _load("flux_sp")
    heads, Lt, Li, hid, mlp = 8, 16, 48, 8 * 128, 256
    hpg, Ltl, Lil = heads // n, Lt // n, Li // n
    Ll, L = Ltl + Lil, Lt + Li
    heap = Heap()
    ptrs = [{"Q": heap.alloc(heads * Ll * 128), "K": heap.alloc(heads * Ll * 128), "V": heap.alloc(heads * Ll * 128),
             "QF": heap.alloc(hpg * L * 128), "KF": heap.alloc(hpg * L * 128), "VF": heap.alloc(hpg * L * 128),
             "ATTF": heap.alloc(L * hpg * 128), "CAT": heap.alloc(Ll * (hid + mlp))} for _ in range(n)]
    heap.build()
    rng = np.random.default_rng(1)
    full = {k: rng.integers(0, 65535, size=(heads, L, 128), dtype=np.uint16) for k in "QKV"}      # global [head, token, d]
    full_att = rng.integers(0, 65535, size=(L, heads, 128), dtype=np.uint16)
    rows_of = lambda r: list(range(r * Ltl, (r + 1) * Ltl)) + list(range(Lt + r * Lil, Lt + (r + 1) * Lil))  # noqa: E731
    for r in range(n):
        for k in "QKV":
            heap.view(ptrs[r][k], (heads, Ll, 128))[:] = full[k][:, rows_of(r)]
        heap.view(ptrs[r]["ATTF"], (L, hpg, 128))[:] = full_att[:, r * hpg:(r + 1) * hpg]
    for g in range(n):
        qkv, att = exchange_tables(g, n, Lt, Li, hid, mlp, hpg, ptrs)
        heap.run(qkv)
        heap.run(att)
    for g in range(n):
        for k in "QKV":
            got = heap.view(ptrs[g][k + "F"], (hpg, L, 128))
            assert np.array_equal(got, full[k][g * hpg:(g + 1) * hpg]), (g, k)
        cat = heap.view(ptrs[g]["CAT"], (Ll, hid + mlp))
        got = cat[:, :hid].reshape(Ll, heads, 128)
        assert np.array_equal(got, full_att[rows_of(g)]), g
        assert not cat[:, hid:].any()                      # the MLP half of the concat buffer is not touched


def test_image_only_tables_of_the_nextdit_refiner():
    """Z-Image's noise refiner attends over the image tokens only: the FLUX table builder with an empty caption segment
    (zero-sized descriptors dropped, as exec/zimage_sp.py does)."""
    exchange_tables = _load("flux_sp")
    n, heads, Li, dim = 2, 4, 32, 4 * 128
    hpg, Lil = heads // n, Li // n
    heap = Heap()
    ptrs = [{"Q": heap.alloc(heads * Lil * 128), "K": heap.alloc(heads * Lil * 128), "V": heap.alloc(heads * Lil * 128),
             "QF": heap.alloc(hpg * Li * 128), "KF": heap.alloc(hpg * Li * 128), "VF": heap.alloc(hpg * Li * 128),
             "ATTF": heap.alloc(Li * hpg * 128), "CAT": heap.alloc(Lil * dim)} for _ in range(n)]
    heap.build()
    rng = np.random.default_rng(2)
    full = {k: rng.integers(0, 65535, size=(heads, Li, 128), dtype=np.uint16) for k in "QKV"}
    full_att = rng.integers(0, 65535, size=(Li, heads, 128), dtype=np.uint16)
    for r in range(n):
        for k in "QKV":
            heap.view(ptrs[r][k], (heads, Lil, 128))[:] = full[k][:, r * Lil:(r + 1) * Lil]
        heap.view(ptrs[r]["ATTF"], (Li, hpg, 128))[:] = full_att[:, r * hpg:(r + 1) * hpg]
    for g in range(n):
        qkv, att = exchange_tables(g, n, 0, Li, dim, 0, hpg, ptrs)
        qkv = [d for d in qkv if d[4] > 0 and d[5] > 0]
        att = [d for d in att if d[4] > 0 and d[5] > 0]
        assert len(qkv) == 3 * n and len(att) == n
        heap.run(qkv)
        heap.run(att)
    for g in range(n):
        for k in "QKV":
            assert np.array_equal(heap.view(ptrs[g][k + "F"], (hpg, Li, 128)), full[k][g * hpg:(g + 1) * hpg]), (g, k)
        assert np.array_equal(heap.view(ptrs[g]["CAT"], (Lil, heads, 128)), full_att[g * Lil:(g + 1) * Lil]), g


def test_staging_copies_conditioning_only_when_it_changed():
    """UlyssesBase.stage: the timestep is copied every step, the conditioning only when the source tensor object / storage /
    version changed (the reference re-sends constant conditioning every step)."""
    import torch
    spec = importlib.util.spec_from_file_location("_sp_common_src", os.path.join(
        ROOT, "comfyui_parallelanything_b200", "exec", "sp_common.py"))
    src = open(spec.origin).read().replace("from .. import ops", "ops = None")
    ns = {"__name__": "_sp_common_src"}
    exec(compile(src, spec.origin, "exec"), ns)
    base = object.__new__(ns["UlyssesBase"])                 # no GPUs: skip __init__, stage() needs no state
    ctx = torch.randn(1, 8, 16)
    st = {"t": torch.zeros(1, dtype=torch.bfloat16), "ctx": torch.zeros(1, 8, 16, dtype=torch.bfloat16), "ctx_src": None}
    base.stage(st, torch.tensor([0.75]), ctx, {}, True)
    assert float(st["t"]) == 0.75 and torch.allclose(st["ctx"].float(), ctx, atol=0.02)
    v0 = st["ctx"]._version
    base.stage(st, torch.tensor([0.5]), ctx, {}, True)
    assert float(st["t"]) == 0.5 and st["ctx"]._version == v0, "unchanged conditioning was copied again"
    ctx.mul_(2.0)                                            # in-place edit bumps the version -> copied again
    base.stage(st, torch.tensor([0.5]), ctx, {}, True)
    assert st["ctx"]._version > v0 and torch.allclose(st["ctx"].float(), ctx, atol=0.04)
    v1 = st["ctx"]._version
    base.stage(st, torch.tensor([0.5]), ctx, {}, False)      # caching off: always copied
    assert st["ctx"]._version > v1
