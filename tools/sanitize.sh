#!/bin/bash
# compute-sanitizer passes over the small self-checks (SURVEY §5 "Race detection / sanitizers").
# Usage (on a GPU box): tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck] [check-name-substring | =exact-name]
# tcgen05/TMA kernels are async-proxy heavy: racecheck only models generic-proxy shared-memory accesses, so
# it is meaningful for the elementwise / layout / flag kernels; memcheck + synccheck cover everything.
TOOL=${1:-memcheck}
ONLY=${2:-layout}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
compute-sanitizer --tool "$TOOL" --error-exitcode 9 --log-file "gpurun_out/sanitizer_${TOOL}_${ONLY}.log" \
  python tools/gpu_check.py --child $(python - <<PY
import sys
sys.path.insert(0, ".")
from comfyui_parallelanything_b200.utils import selfcheck
only = "$ONLY"
print(" ".join(n for n in selfcheck.CHECKS if (n == only[1:] if only.startswith("=") else only in n)))
PY
)
rc=$?
tail -5 "gpurun_out/sanitizer_${TOOL}_${ONLY}.log"
exit $rc
