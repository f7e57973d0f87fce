#!/usr/bin/env python
"""Run every native-kernel self-check on the current GPU, one process per check (a
trapped/hung kernel cannot poison the rest), and write ``gpurun_out/selfcheck.json``.

    python tools/gpu_check.py                 # all checks
    python tools/gpu_check.py --only gemm     # substring filter
    python tools/gpu_check.py --check NAME    # (internal) run one check in-process
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


def run_one(name: str) -> int:
    import torch
    from comfyui_parallelanything_b200.utils import selfcheck
    t0 = time.time()
    try:
        r = selfcheck.CHECKS[name]()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        r = dict(name=name, ok=False, error=f"{type(e).__name__}: {e}"[:500])
    r["seconds"] = round(time.time() - t0, 2)
    print("PA_CHECK " + json.dumps(r), flush=True)
    return 0 if r.get("ok") else 1


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--check")
    ap.add_argument("--only", default="")
    ap.add_argument("--timeout", type=int, default=180)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "selfcheck.json"))
    a = ap.parse_args()
    if a.check:
        return run_one(a.check)
    from comfyui_parallelanything_b200.utils import selfcheck
    names = [n for n in selfcheck.CHECKS if a.only in n]
    results = []
    for n in names:
        try:
            p = subprocess.run([sys.executable, __file__, "--check", n], capture_output=True, text=True,
                               timeout=a.timeout)
            line = [l for l in p.stdout.splitlines() if l.startswith("PA_CHECK ")]
            if line:
                r = json.loads(line[-1][len("PA_CHECK "):])
            else:
                r = dict(name=n, ok=False, error="no result", stdout=p.stdout[-800:], stderr=p.stderr[-1500:])
        except subprocess.TimeoutExpired:
            r = dict(name=n, ok=False, error=f"timeout after {a.timeout}s")
        results.append(r)
        print(("PASS " if r.get("ok") else "FAIL ") + json.dumps(r), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(results, f, indent=1)
    bad = [r["name"] for r in results if not r.get("ok")]
    print(f"{len(results) - len(bad)}/{len(results)} checks passed; failed: {bad}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
