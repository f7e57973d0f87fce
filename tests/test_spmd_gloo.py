"""Multi-process host-side protocol on CPU (gloo, world_size 2/3): uneven shards, kwargs rules,
small-batch fallback, worker error propagation."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Toy(nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(7)
        self.a = nn.Linear(4, 8)
        self.b = nn.Linear(8, 4)
        self.fail_on_rank = -1

    def forward(self, x, timesteps, context=None, y=None, scale=1.0, control=None, **kw):
        if dist.get_rank() == self.fail_on_rank:
            raise ValueError("boom")
        h = torch.tanh(self.a(x)) * scale + timesteps[:, None]
        if context is not None:
            h = h + context.mean(1)
        if y is not None:
            h = h + y
        if control is not None:
            h = h + control["input"][0]
        return self.b(h), h.sum(-1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from comfyui_parallelanything_b200.parallel.spmd_generic import SpmdModuleEngine
    m = Toy().eval()
    weights = [40, 40, 20][:world] if case == "weighted" else None
    eng = SpmdModuleEngine(m, weights=weights)
    try:
        if rank != 0:
            if case == "fail":
                m.fail_on_rank = 1
            n = eng.serve()
            q.put((rank, "served", n))
        else:
            res = {}
            for B in (7, 2, 1):
                g = torch.Generator().manual_seed(B)
                x, t = torch.randn(B, 4, generator=g), torch.rand(B, generator=g)
                ctx, y = torch.randn(B, 3, 8, generator=g), torch.randn(B, 8, generator=g)
                ctl = {"input": [torch.randn(B, 8, generator=g)], "w": 0.5}
                want = m(x, t, context=ctx, y=y, scale=2.0, control=ctl)
                if case == "fail" and B >= world:
                    with pytest.raises(RuntimeError, match="rank 1 failed"):
                        eng.forward(x, t, context=ctx, y=y, scale=2.0, control=ctl)
                    res[B] = True
                    continue
                got = eng.forward(x, t, context=ctx, y=y, scale=2.0, control=ctl)
                res[B] = bool(torch.allclose(got[0], want[0], atol=1e-6) and torch.allclose(got[1], want[1], atol=1e-6))
            eng.stop()
            q.put((0, "results", res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,case", [(2, "even"), (3, "weighted"), (2, "fail")])
def test_spmd_generic_gloo(world, case):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = [o for o in out if o[1] == "results"][0][2]
    assert all(res.values()), res
    served = {o[0]: o[2] for o in out if o[1] == "served"}
    assert all(v == 3 for v in served.values()), served   # B=7, B=2 (skip for world 3), B=1 (skip) all seen
