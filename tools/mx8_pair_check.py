"""CTA-pair MXFP8 GEMM: numerics vs the exact product of the dequantised operands, then pair vs one-CTA timing on the
FLUX step's shapes.  Run under `timeout` (a protocol error in a cluster kernel hangs instead of failing)."""
import json
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
import comfyui_parallelanything_b200 as pa  # noqa: E402
from comfyui_parallelanything_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
res = {}


def rel(a, b):
    return ((a.float() - b.float()).abs().mean() / b.float().abs().mean().clamp_min(1e-9)).item()


for (B, M, K, N) in ((1, 256, 256, 256), (1, 256, 1024, 512), (2, 320, 1024, 768), (3, 512, 3072, 1344), (2, 1024, 3072, 2304)):
    a = torch.randn(B, M, K, device=dev, dtype=torch.bfloat16)
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.03
    aq, sfa = ops.quantize_mxfp8(a)
    ad = ops.dequantize_mxfp8(aq, sfa)
    for tile in (224, 256):
        wq, sfb = ops.quantize_mxfp8(w, tile)
        exact = ad @ ops.dequantize_mxfp8(wq, sfb, tile)[0].t()
        outs = []
        for pair in ((0, 1, 2) if tile == 256 else (0, 1)):
            out = torch.zeros(B, M, N, dtype=torch.bfloat16, device=dev)
            ops.gemm_fp8(aq, sfa, wq, sfb, "bias", tile, out=out, pair=pair)
            torch.cuda.synchronize()
            outs.append(out)
        res[f"{B}x{M}x{K}x{N}_t{tile}"] = {"pair_vs_exact": rel(outs[1], exact), "one_vs_exact": rel(outs[0], exact),
                                           "pair_eq_one": bool(torch.equal(outs[0], outs[1])),
                                           "split_eq_one": bool(torch.equal(outs[0], outs[-1]))}
        print(f"{B}x{M}x{K}x{N} t{tile}", res[f"{B}x{M}x{K}x{N}_t{tile}"], flush=True)

e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
SHAPES = {"linear1_b8": (8, 4608, 3072, 21504, (256,)), "linear2_b8": (8, 4608, 15360, 3072, (224, 256)),
          "img_mlp_b8": (8, 4096, 3072, 12288, (224, 256)), "img_proj_b8": (8, 4096, 3072, 3072, (224, 256)),
          "txt_mlp_b8": (8, 512, 3072, 12288, (224,)), "linear1_b1": (1, 4608, 3072, 21504, (256,)),
          "linear2_b1": (1, 4608, 15360, 3072, (224, 256)), "img_proj_b1": (1, 4096, 3072, 3072, (224, 256)),
          "wan_o_b2": (2, 14400, 5120, 5120, (224, 256)), "wan_ffn2_b2": (2, 14400, 13824, 5120, (224, 256))}
only = [a for a in sys.argv[1:] if not a.startswith("-")]
for name, (B, M, K, N, tiles) in SHAPES.items():
    if only and name not in only:
        continue
    a = torch.randn(B, M, K, device=dev, dtype=torch.bfloat16)
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.02
    aq, sfa = ops.quantize_mxfp8(a)
    out = torch.empty(B, M, N, dtype=torch.bfloat16, device=dev)
    del a
    row = {}
    for tile in tiles:
        wq, sfb = ops.quantize_mxfp8(w, tile)
        for pair in ((0, 1, 2) if tile == 256 else (0, 1)):
            for _ in range(3):
                ops.gemm_fp8(aq, sfa, wq, sfb, "bias", tile, out=out, pair=pair)
            ts = []
            for _ in range(8):
                flush.fill_(1)
                e0.record()
                ops.gemm_fp8(aq, sfa, wq, sfb, "bias", tile, out=out, pair=pair)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            row[f"t{tile}_" + ("one", "pair", "pair_split")[pair]] = {"ms": round(ms, 4), "tflops": round(2.0 * B * M * N * K / ms / 1e9, 1)}
    del w
    res[name] = row
    print(name, row, flush=True)
json.dump(res, open("gpurun_out/mx8_pair_check.json", "w"), indent=1)
