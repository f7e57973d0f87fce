"""One-process-per-GPU engine on >= 2 real GPUs: fused in-kernel NVLink scatter/gather, TMA over peer memory
and the NCCL baseline must all reproduce the single-GPU executor result (tools/spmd_check.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


@pytest.mark.parametrize("family", ["flux", "unet", "wan", "zimage"])
def test_spmd_fused_matches_single_gpu(family):
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "spmd_check.py"), family]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("PA_SPMD ")]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(lines[-1][len("PA_SPMD "):])
    assert res["ok"] and res["world"] == n, res
    for name, v in res["results"].items():
        assert v["mean_rel"] < 5e-3, (name, v)


def test_spmd_weight_broadcast_nvls_or_nccl():
    """K9: packed executor weights from rank 0 to every rank (multimem.st through an NVSwitch multicast object shared via
    POSIX fds; NCCL broadcast when the fabric has no multicast) - all ranks must end up with rank 0's bytes."""
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29537", os.path.join(ROOT, "tools", "spmd_check.py"), "bcast"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("PA_SPMD ")]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(lines[-1][len("PA_SPMD "):])
    assert res["ok"] and res["world"] == n, res
    print("weight broadcast:", {k: (v["method"], v.get("gbps"), v.get("why_not_nvls")) for k, v in res["results"].items()})


def test_flag_protocol_randomised_delay_stress():
    """the release/acquire flag protocol under randomised producer/consumer skew (20 000 epochs here;
    the 100 000-epoch run is committed under profiles/)."""
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29539", os.path.join(ROOT, "tools", "flag_stress.py"),
           "--epochs", "20000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("PA_FLAGS ")]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(lines[-1][len("PA_FLAGS "):])
    assert res["ok"] and res["world"] == n and res["epochs"] == 20000, res
