#!/usr/bin/env python
"""Run every native-kernel self-check on the current GPU and write ``gpurun_out/selfcheck.json``.

Checks run sequentially in a child process; if a kernel traps (poisoned CUDA context) or hangs past
the timeout, the child is killed and a fresh one continues with the next check — one bad kernel
cannot hide the state of the rest.

    python tools/gpu_check.py                 # all checks
    python tools/gpu_check.py --only gemm,attention     # substring filter(s)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(names) -> int:
    import torch
    from comfyui_parallelanything_b200.utils import selfcheck
    for name in names:
        print("PA_START " + name, flush=True)
        t0 = time.time()
        fatal = False
        try:
            r = selfcheck.CHECKS[name]()
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            msg = f"{type(e).__name__}: {e}"
            r = dict(name=name, ok=False, error=msg[:600])
            fatal = "CUDA" in msg or "cuda" in msg or "launch failure" in msg
        r["seconds"] = round(time.time() - t0, 2)
        print("PA_CHECK " + json.dumps(r), flush=True)
        if fatal:
            return 3          # context is gone: let the parent restart us
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", nargs="*")
    ap.add_argument("--only", default="")
    ap.add_argument("--timeout", type=int, default=240, help="per-child wall clock limit (s)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "selfcheck.json"))
    a = ap.parse_args()
    if a.child is not None:
        return child(a.child)
    from comfyui_parallelanything_b200.utils import selfcheck
    todo = [n for n in selfcheck.CHECKS if any(s in n for s in a.only.split(","))]
    results = {}
    while todo:
        p = subprocess.Popen([sys.executable, __file__, "--child"] + todo, stdout=subprocess.PIPE,
                             stderr=subprocess.PIPE, text=True)
        try:
            out, err = p.communicate(timeout=a.timeout)
        except subprocess.TimeoutExpired:
            p.kill()
            out, err = p.communicate()
            err += "\n[gpu_check] child timed out"
        started = [l.split(" ", 1)[1] for l in out.splitlines() if l.startswith("PA_START ")]
        for l in out.splitlines():
            if l.startswith("PA_CHECK "):
                r = json.loads(l[len("PA_CHECK "):])
                results[r["name"] if r["name"] in todo else started[len(results) % max(1, len(started))]] = r
        done = [n for n in started if any(r is results.get(n) for r in [results.get(n)]) and n in results]
        if started and started[-1] not in results:       # died inside this check
            results[started[-1]] = dict(name=started[-1], ok=False, error="child died/timed out",
                                        stderr=err[-1500:])
            done.append(started[-1])
        if not started:
            results[todo[0]] = dict(name=todo[0], ok=False, error="child produced no output", stderr=err[-1500:])
            done.append(todo[0])
        todo = [n for n in todo if n not in done and n not in results]
    ordered = [results[n] for n in selfcheck.CHECKS if n in results]
    for r in ordered:
        print(("PASS " if r.get("ok") else "FAIL ") + json.dumps(r), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(ordered, f, indent=1)
    bad = [r["name"] for r in ordered if not r.get("ok")]
    print(f"{len(ordered) - len(bad)}/{len(ordered)} checks passed; failed: {bad}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
