#!/usr/bin/env python
"""Per-kernel SASS listings of the named kernels:  cuobjdump -sass ops/_C.so, split per function into profiles/r2/sass/
(listings over 800 instructions gzipped, their first 160 lines kept as *.head.txt) + a README with the tensor / TMA / TMEM
mnemonic counts.   python tools/sass_split.py"""
import collections
import gzip
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "comfyui_parallelanything_b200", "ops", "_C.so")
OUT = os.path.join(ROOT, "profiles", "r2", "sass")
WANT = {"scatter_patch_embed_kernel": "scatter_patch_embed", "scatter_conv_in_kernel": "scatter_conv_in",
        "gemm_mxfp8_kernelILi256ELi1": "gemm_mxfp8_256x1_drainfirst", "gemm_mxfp8_kernelILi224ELi2": "gemm_mxfp8_224x2",
        "gemm_mxfp8_2cta_kernelILi224ELi2": "gemm_mxfp8_ctapair_224x2", "gemm_mxfp8_2cta_split_kernel": "gemm_mxfp8_ctapair_splitn",
        "xattn_cluster_kernel": "xattn_cluster",
        "gemm_bf16_2cta_kernel": "gemm_bf16_2cta", "attention2_kernelILi128ELi0ELi1": "attention2_d128",
        "gn_cluster_kernel": "groupnorm_cluster", "multimem_bcast_kernel": "multimem_bcast", "sp_pull_kernel": "sp_pull",
        "sp_signal_kernel": "sp_signal", "ln_mod_fast_kernelILi12ELi12ELb1": "ln_mod_fp8out",
        "gemm_bf16_tcgen05_kernelILi64": "gemm_bf16_bn64_euler_gather", "unet_out_gather": "unet_out_gather"}
KEY = ("UTC", "UTMA", "LDTM", "STTM", "UBLKCP", "USETMAXREG", "MUFU", "HMMA", "LDG", "STG", "RED", "ATOM", "MEMBAR", "SYNCS",
       "UCGABAR")


def main():
    os.makedirs(OUT, exist_ok=True)
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    parts = re.split(r"(?m)^\s*Function : ", txt)
    rows = []
    for p in parts[1:]:
        name = p.split("\n", 1)[0].strip()
        for k, short in WANT.items():
            if k not in name:
                continue
            body = "        Function : " + p
            ops = collections.Counter()
            for line in body.splitlines():
                m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s", line)
                if m:
                    op = m.group(1)
                    ops[op if op.startswith(("UTC", "UTMA", "LDTM", "STTM", "UBLKCP", "USETMAXREG", "UCGABAR")) else op.split(".")[0]] += 1
            n = sum(ops.values())
            fn = os.path.join(OUT, short + ".sass")
            if n > 800:
                with gzip.open(fn + ".gz", "wt") as f:
                    f.write(body)
                with open(fn + ".head.txt", "w") as f:
                    f.write("\n".join(body.splitlines()[:160]) + f"\n... full listing: {short}.sass.gz\n")
            else:
                with open(fn, "w") as f:
                    f.write(body)
            key = sorted(((o, c) for o, c in ops.items() if o.startswith(KEY)), key=lambda x: -x[1])
            rows.append((short, name[:90], n, key))
            break
    with open(os.path.join(OUT, "README.md"), "w") as f:
        f.write("# SASS listings of the named kernels (round 2)\n\n`python tools/sass_split.py` = `cuobjdump -sass "
                "comfyui_parallelanything_b200/ops/_C.so` split per kernel.  Listings over 800 instructions are gzipped "
                "(`*.sass.gz`), their first 160 lines are in `*.head.txt`.  `UTCHMMA` / `UTCQMMA` = tcgen05.mma (bf16 / "
                "block-scaled fp8), `UTMALDG` / `UTMASTG` = TMA load / store, `LDTM` / `STTM` = tcgen05.ld / st, `UTCCP` = "
                "tcgen05.cp, `UBLKCP` = cp.async.bulk, `UCGABAR_*` = cluster barrier, `USETMAXREG` = setmaxnreg.  "
                "`multimem.st` assembles to `STG.E.128.STRONG.SYS` on a multicast address (see the PTX in "
                "`csrc/comm/multicast.cu`).\n\n| file | kernel | instructions | tensor / TMA / TMEM / memory mnemonics |\n|---|---|---|---|\n")
        for short, name, n, key in sorted(rows):
            f.write(f"| `{short}.sass*` | `{name}` | {n} | " + ", ".join(f"{o} x{c}" for o, c in key[:14]) + " |\n")
    print(f"{len(rows)} kernels -> {OUT}")


if __name__ == "__main__":
    main()
