#!/usr/bin/env python
"""Randomised-delay stress of the cross-GPU flag protocol (SURVEY §4.5 / §5 "race detection"): N ranks, E epochs.

Every epoch follows the engine's step protocol (parallel/spmd.py) with the replica forward replaced by a random device
delay, so producer / consumer skew is varied on purpose:

  rank 0:  random delay -> write the epoch's input pattern into its staging area -> signal_flags(all ranks, slot 0, e)
  rank r:  wait_flags(slot 0 >= e) -> random delay -> READ rank 0's input pattern over NVLink (peer mapping), check it
           is epoch e's (a stale or torn payload is counted on the device), write f(e, r) into rank 0's output rows
           (peer stores) -> signal_flags(rank 0, slot 8 + r, e)
  rank 0:  wait_flags(slots 8.. >= e) -> check every rank's rows carry f(e, r); the done-flags also fence input reuse.

No host synchronisation inside the loop: mismatches accumulate in device counters, read once at the end together with
the flag watchdog's error word.     torchrun --nproc-per-node N tools/flag_stress.py --epochs 100000
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=100000)
    ap.add_argument("--max-delay-cycles", type=int, default=60000)
    ap.add_argument("--words", type=int, default=4096, help="payload words per rank")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from comfyui_parallelanything_b200 import ops
    from comfyui_parallelanything_b200.parallel.spmd import SymmetricHeap
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    C = ops.require()
    W = a.words
    heap = SymmetricHeap(256 + 4 * W + 4 * W * world + 4096)
    off_flags = heap.carve(256)
    off_in = heap.carve(4 * W)
    off_out = heap.carve(4 * W * world)
    flags = heap.view(off_flags, (64,), torch.int32)
    inp0 = heap.view(off_in, (W,), torch.int32)                       # rank 0's staging area (local view)
    out0 = heap.view(off_out, (world, W), torch.int32)                # rank 0's output rows (local view)
    peer_in = C.tensor_from_ptr(heap.peer_ptr(0, off_in), 4 * W, dev.index).view(torch.int32)
    peer_out = C.tensor_from_ptr(heap.peer_ptr(0, off_out) + 4 * W * rank, 4 * W, dev.index).view(torch.int32)
    peer_tab = torch.tensor([heap.peer_ptr(r, off_flags) for r in range(world)], dtype=torch.int64, device=dev)
    lead_tab = torch.tensor([heap.peer_ptr(0, off_flags)], dtype=torch.int64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    bad_in = torch.zeros(1, dtype=torch.int64, device=dev)
    bad_out = torch.zeros(1, dtype=torch.int64, device=dev)
    timeout = int(20000 * 1.9e6)
    g = torch.Generator().manual_seed(1234 + rank)
    delays = torch.randint(0, a.max_delay_cycles, (a.epochs, 2), generator=g).tolist()
    rmul = torch.arange(world, device=dev, dtype=torch.int32).view(world, 1) * 7919
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for e in range(1, a.epochs + 1):
        d0, d1 = delays[e - 1]
        if rank == 0:
            torch.cuda._sleep(d0)
            inp0.fill_(e * 31 + 5)
            C.signal_flags(peer_tab, world, 0, e)
        C.wait_flags(flags, 0, 1, e, timeout, err)
        torch.cuda._sleep(d1)
        bad_in += (peer_in != e * 31 + 5).sum()                      # NVLink peer read of rank 0's inputs
        peer_out.fill_(e * 13 + rank * 7919)                         # NVLink peer stores into rank 0's rows
        C.signal_flags(lead_tab, 1, 8 + rank, e)
        if rank == 0:
            C.wait_flags(flags, 8, world, e, timeout, err)
            bad_out += (out0 != (e * 13 + rmul)).sum()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res = torch.stack([bad_in[0], bad_out[0], (err[0] & 0xFFFFFFFF).long()])
    allr = [torch.empty_like(res) for _ in range(world)]
    dist.all_gather(allr, res)
    if rank == 0:
        tot = torch.stack(allr).sum(0).tolist()
        ok = tot[0] == 0 and tot[1] == 0 and tot[2] == 0
        print("PA_FLAGS " + json.dumps(dict(world=world, epochs=a.epochs, ok=ok, stale_or_torn_inputs=tot[0],
                                            wrong_output_words=tot[1], watchdog_error_words=tot[2],
                                            max_delay_cycles=a.max_delay_cycles, payload_words=W,
                                            seconds=round(dt, 2), us_per_epoch=round(dt / a.epochs * 1e6, 2))), flush=True)
    heap.close()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
