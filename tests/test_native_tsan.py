"""ThreadSanitizer build of the C++ host runtime's executor (SURVEY §5 race detection: "TSAN build of the C++ host
runtime"): csrc/runtime/host_executor.h is CUDA- and Python-free, so the exact code runtime.cpp wraps is compiled here
with -fsanitize=thread and stressed from several producer threads."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_host_executor_is_race_free_under_tsan(tmp_path):
    exe = str(tmp_path / "host_executor_tsan")
    src = os.path.join(ROOT, "tests", "native", "host_executor_tsan.cpp")
    b = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-pthread", src, "-o", exe],
                       capture_output=True, text=True)
    if b.returncode != 0 and "tsan" in (b.stderr or "").lower():
        pytest.skip("libtsan not available: " + b.stderr[-200:])
    assert b.returncode == 0, b.stderr[-2000:]
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 second_deadlock_stack=1")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-3000:]
    assert "ThreadSanitizer" not in r.stderr and "host_executor_tsan: ok" in r.stdout, r.stderr[-3000:]
