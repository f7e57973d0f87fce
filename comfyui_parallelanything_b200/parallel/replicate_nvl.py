"""Weight replication over NVLink / NVSwitch (SURVEY K9, §2.4 "one-to-all payloads ... NVLS multicast").

The reference clones a model by moving the source to the CPU **in place** and re-uploading it key by key for every
device (/root/reference/any_device_parallel.py:600-663: one D2H + N H2D passes over PCIe plus a Python loop per
key).  Here the PACKED executor weights (concatenated modulation tables, fused QKV matrices, MXFP8 bytes + scale
chunks ...) are produced once on the lead GPU and replicated device-to-device:

  * ``nvls``  - one ``multimem.st`` kernel on the lead GPU per chunk (csrc/comm/multicast.cu): the lead reads its HBM
    once, the NVSwitch fans every 16-byte store out into the slab each GPU bound to a multicast object
    (csrc/runtime/multicast.cpp), receivers drain their slab into the final tensors with a local copy.  Lead egress is
    paid once instead of once per peer.
  * ``p2p``   - per-tensor peer copies (cudaMemcpyPeerAsync), the fallback when the fabric has no multicast.
  * ``nccl``  - (multi-process only) bucketed ``dist.broadcast`` - the library baseline.

``shell_like`` builds an executor of the same class / geometry on another device WITHOUT a model: same packed-table
keys and shapes, uninitialised tensors - the receive side of a replication.
"""
from __future__ import annotations

import os
import socket
import time
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .. import ops
from ..exec.pack_cache import packed_table
from ..utils import log

_SKIP = ("_ws", "_io", "_side", "_ev_fork", "_ev_join", "_tblocks")


def shell_like(executor, device) -> Any:
    """Structural clone of a native executor on ``device``: every packed tensor becomes ``empty_like`` there, small
    python state is copied, per-device runtime state (workspaces, graph caches, streams, events) is re-created."""
    from ..exec.graphs import GraphCache
    device = torch.device(device)
    memo: Dict[int, Any] = {}

    def conv(v):
        if id(v) in memo:
            return memo[id(v)]
        if isinstance(v, torch.Tensor):
            out = torch.empty_like(v, device=device) if v.is_cuda else v.clone()
        elif isinstance(v, GraphCache):
            out = GraphCache(device, enabled=v.enabled, limit=v.limit)
        elif isinstance(v, torch.device):
            out = device
        elif isinstance(v, (torch.cuda.Stream, torch.cuda.Event)):
            out = None
        elif isinstance(v, dict):
            out = {}
            memo[id(v)] = out
            for k, x in v.items():
                out[k] = conv(x)
            return out
        elif isinstance(v, list):
            out = []
            memo[id(v)] = out
            out.extend(conv(x) for x in v)
            return out
        elif isinstance(v, tuple):
            out = tuple(conv(x) for x in v)
        elif isinstance(v, (int, float, str, bool, bytes, type(None), torch.dtype)):
            return v
        elif isinstance(v, nn.Module) and not getattr(v, "pa_native", False):
            raise TypeError("executor holds a torch module; cannot make a shell of it")
        elif hasattr(v, "__dict__") and not isinstance(v, type) and not callable(v):
            out = type(v).__new__(type(v))
            memo[id(v)] = out
            for k, x in vars(v).items():
                out.__dict__[k] = None if k in ("_kv", "_kv_sig") else conv(x)
            return out
        else:
            out = v                                   # SimpleNamespace handled above; functions / classes by reference
        memo[id(v)] = out
        return out

    new = object.__new__(type(executor))
    nn.Module.__init__(new)
    module_internals = set(vars(new))
    for k, v in vars(executor).items():
        if k in module_internals:
            continue
        if k in ("_ws", "_io"):
            new.__dict__[k] = {}
        elif k in ("_side", "_ev_fork", "_ev_join"):
            new.__dict__[k] = None
        elif k == "_tblocks":
            continue
        else:
            new.__dict__[k] = conv(v)
    if "_tblocks" in vars(executor):                  # UNet: the flat list of transformer blocks aliases objects in inp/mid/outb
        new.__dict__["_tblocks"] = [memo[id(tb)] for tb in executor._tblocks]
    return new


def _pairs(src, dsts: Sequence[Any]) -> List[Tuple[torch.Tensor, List[torch.Tensor]]]:
    ts = packed_table(src)
    tds = [packed_table(d) for d in dsts]
    out = []
    for k, v in ts.items():
        if not isinstance(v, torch.Tensor) or not v.is_cuda:
            continue
        row = []
        for td in tds:
            w = td.get(k)
            if not isinstance(w, torch.Tensor) or w.shape != v.shape or w.dtype != v.dtype:
                raise KeyError(f"replica is missing packed tensor {k!r} (or it has another shape)")
            row.append(w)
        out.append((v, row))
    return out


def _nbytes(t: torch.Tensor) -> int:
    return t.numel() * t.element_size()


def replicate_inprocess(src_executor, dst_executors: Sequence[Any], method: str = "auto",
                        slot_bytes: int = 256 << 20) -> dict:
    """Fill ``dst_executors`` (shells on other GPUs of THIS process) from ``src_executor``.  Returns a report
    ``{"method", "bytes", "seconds", "gbps", "why"}``."""
    if not dst_executors:
        return {"method": "none", "bytes": 0, "seconds": 0.0}
    C = ops.require()
    pairs = _pairs(src_executor, dst_executors)
    src_dev = src_executor.device
    devs = [src_dev] + [d.device for d in dst_executors]
    total = sum(_nbytes(s) for s, _ in pairs)
    why = ""
    timing: Dict[str, float] = {}
    t0 = time.perf_counter()
    use = method
    if method in ("auto", "nvls"):
        try:
            if not all(C.multicast_supported(d.index) for d in devs):
                raise RuntimeError("CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED is 0 on a device of the chain")
            timing = _nvls_inprocess(C, pairs, devs, slot_bytes)
            use = "nvls"
        except Exception as e:                        # fabric without multicast, or the driver refused the object
            if method == "nvls":
                log.warn("NVLS multicast replication unavailable (%s); using peer copies", e)
            why = str(e)[:200]
            use = "p2p"
    if use == "p2p":
        for s, ds in pairs:
            for d in ds:
                d.copy_(s, non_blocking=True)
    for d in devs:
        torch.cuda.synchronize(d)
    dt = time.perf_counter() - t0
    rep = {"method": use, "bytes": total, "seconds": round(dt, 4), "gbps": round(total / max(dt, 1e-9) / 1e9, 1),
           "receivers": len(dst_executors)}
    if timing:                                        # multicast object creation / binding vs the stores themselves
        rep["team_setup_s"] = round(timing["setup"], 4)
        rep["transfer_s"] = round(timing["transfer"], 4)
        rep["transfer_gbps_per_receiver"] = round(total / max(timing["transfer"], 1e-9) / 1e9, 1)
    if why:
        rep["why_not_nvls"] = why
    return rep


def _chunks(pairs, slot_bytes: int):
    """Greedy packing of (tensor, byte range) pieces into slots; offsets 256-byte aligned, sizes multiples of 16."""
    cur, used = [], 0
    for idx, (s, _ds) in enumerate(pairs):
        n = _nbytes(s)
        pos = 0
        while pos < n:
            room = slot_bytes - used
            if room < 4096:
                yield cur
                cur, used, room = [], 0, slot_bytes
            take = min(n - pos, room)
            if take < n - pos:
                take -= take % 256
            cur.append((idx, pos, take, used))
            used += (take + 255) // 256 * 256
            pos += take
    if cur:
        yield cur


def _nvls_inprocess(C, pairs, devs, slot_bytes: int) -> Dict[str, float]:
    bad = [s for s, _ in pairs if not s.is_contiguous() or s.data_ptr() % 16]
    if bad:
        raise RuntimeError("packed tensors must be contiguous and 16-byte aligned")
    t0 = time.perf_counter()
    team = C.MulticastTeam([d.index for d in devs], 2 * slot_bytes, 0, False)
    try:
        team.bind_all()
        for d in devs:
            torch.cuda.synchronize(d)
        t1 = time.perf_counter()
        slot = team.size() // 2
        slot -= slot % 256
        lead = devs[0]
        streams = [torch.cuda.Stream(device=d) for d in devs]
        free_evs: List[List[torch.cuda.Event]] = [[], []]
        k = 0
        for chunk in _chunks(pairs, min(slot, slot_bytes)):
            s_ = k & 1
            base = s_ * slot
            with torch.cuda.device(lead):
                for ev in free_evs[s_]:
                    streams[0].wait_event(ev)             # receivers drained this slot two chunks ago
                for idx, pos, take, off in chunk:
                    src = pairs[idx][0]
                    nb = (take + 15) // 16 * 16           # the tail of an odd-sized tensor stays inside its 512-B block
                    team.bcast(0, src.data_ptr() + pos, base + off, nb, streams[0].cuda_stream)
                ready = torch.cuda.Event()
                ready.record(streams[0])
            free_evs[s_] = []
            for j in range(1, len(devs)):
                with torch.cuda.device(devs[j]):
                    streams[j].wait_event(ready)
                    for idx, pos, take, off in chunk:
                        dst = pairs[idx][1][j - 1]
                        C.memcpy_async(dst.data_ptr() + pos, team.uc_ptr(j) + base + off, take, 3,
                                       streams[j].cuda_stream)
                    ev = torch.cuda.Event()
                    ev.record(streams[j])
                    free_evs[s_].append(ev)
            k += 1
        for d in devs:
            torch.cuda.synchronize(d)
        return {"setup": t1 - t0, "transfer": time.perf_counter() - t1}
    finally:
        team.close()


# ----------------------------------------------------------------------------- one process per GPU
def _send_fd_to_peers(fd: int, world: int, path: str, meta: bytes) -> None:
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    try:
        if os.path.exists(path):
            os.unlink(path)
        srv.bind(path)
        srv.listen(world)
        import torch.distributed as dist
        dist.barrier()                                     # peers connect only after the socket exists
        for _ in range(world - 1):
            conn, _a = srv.accept()
            with conn:
                socket.send_fds(conn, [meta], [fd])
    finally:
        srv.close()
        try:
            os.unlink(path)
        except OSError:
            pass


def _recv_fd(path: str) -> Tuple[int, bytes]:
    import torch.distributed as dist
    dist.barrier()
    c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    for _ in range(200):
        try:
            c.connect(path)
            break
        except (FileNotFoundError, ConnectionRefusedError):
            time.sleep(0.05)
    with c:
        msg, fds, _f, _a = socket.recv_fds(c, 256, 1)
    return fds[0], msg


def broadcast_executor(executor, src: int = 0, method: str = "nvls", slot_bytes: int = 256 << 20) -> dict:
    """SPMD (torchrun) replication: rank ``src`` holds the packed weights, every other rank's executor (same class /
    geometry, any content) receives them.  ``nvls``: multicast object shared through a POSIX fd, ``multimem.st`` kernel
    on ``src``, NCCL barriers only for slot hand-off; falls back to ``nccl`` (bucketed broadcast) when the fabric has
    no multicast.  Collective: every rank calls it."""
    import torch.distributed as dist
    from .spmd import broadcast_executor_weights
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = executor.device
    C = ops.require()
    t0 = time.perf_counter()
    rep: Dict[str, Any] = {"requested": method}
    done = False
    if method == "nvls":
        ok = torch.tensor([1 if C.multicast_supported(dev.index) else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            try:
                rep.update(_nvls_spmd(C, executor, src, rank, world, dev, slot_bytes))
                done = True
                rep["method"] = "nvls"
            except Exception as e:                       # noqa: BLE001 - reported, then the NCCL path runs on all ranks
                rep["why_not_nvls"] = str(e)[:200]
                log.warn("NVLS multicast broadcast failed on rank %d (%s)", rank, e)
            flag = torch.tensor([1 if done else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            done = int(flag.item()) == 1
        else:
            rep["why_not_nvls"] = "CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED is 0 on a rank"
    if not done:
        broadcast_executor_weights(executor, src=src)
        rep["method"] = "nccl"
    torch.cuda.synchronize(dev)
    dist.barrier()
    dt = time.perf_counter() - t0
    total = sum(_nbytes(v) for v in packed_table(executor).values() if isinstance(v, torch.Tensor) and v.is_cuda)
    rep.update(bytes=total, seconds=round(dt, 4), gbps=round(total / max(dt, 1e-9) / 1e9, 1))
    return rep


def _nvls_spmd(C, executor, src: int, rank: int, world: int, dev, slot_bytes: int) -> Dict[str, float]:
    import torch.distributed as dist
    t0 = time.perf_counter()
    path = f"/tmp/pa_mc_{os.environ.get('MASTER_PORT', '0')}_{os.getuid()}.sock"
    if rank == src:
        team = C.MulticastTeam([dev.index], 2 * slot_bytes, world, True)
        meta = f"{team.size()} {team.granularity()}".encode()
        fd = team.export_fd()
        try:
            _send_fd_to_peers(fd, world, path, meta)
        finally:
            os.close(fd)
    else:
        fd, meta = _recv_fd(path)
        size, gran = (int(v) for v in meta.decode().split())
        try:
            team = C.MulticastTeam.from_fd(fd, dev.index, size, gran)
        finally:
            os.close(fd)
    try:
        dist.barrier()                                    # every device added ...
        team.bind_all()                                   # ... before anyone binds
        dist.barrier()
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        pairs = [(v, [v]) for v in packed_table(executor).values() if isinstance(v, torch.Tensor) and v.is_cuda]
        slot = team.size() // 2
        slot -= slot % 256
        st = torch.cuda.current_stream(dev)
        tick = torch.zeros(1, device=dev)
        k = 0
        for chunk in _chunks(pairs, min(slot, slot_bytes)):
            base = (k & 1) * slot
            if rank == src:
                for idx, pos, take, off in chunk:
                    t = pairs[idx][0]
                    team.bcast(0, t.data_ptr() + pos, base + off, (take + 15) // 16 * 16, st.cuda_stream)
            # stream-ordered hand-off: after this all-reduce the lead's stores of chunk k are complete on every rank,
            # and it cannot start chunk k+2 (same slot) before every receiver has entered the NEXT all-reduce, i.e.
            # has drained chunk k
            dist.all_reduce(tick)
            if rank != src:
                for idx, pos, take, off in chunk:
                    t = pairs[idx][0]
                    C.memcpy_async(t.data_ptr() + pos, team.uc_ptr(0) + base + off, take, 3, st.cuda_stream)
            k += 1
        dist.all_reduce(tick)
        torch.cuda.synchronize(dev)
        t2 = time.perf_counter()
        nb = sum(_nbytes(t) for t, _ in pairs)
        return {"team_setup_s": round(t1 - t0, 4), "transfer_s": round(t2 - t1, 4),
                "transfer_gbps_per_receiver": round(nb / max(t2 - t1, 1e-9) / 1e9, 1), "chunks": k}
    finally:
        team.close()
