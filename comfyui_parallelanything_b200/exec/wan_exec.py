"""WAN2.x video DiT on hand-written sm_100a kernels.

Same building blocks as the FLUX executor (tcgen05 GEMM with fused epilogues, TMEM attention, AdaLN
LayerNorm kernel) plus what WAN needs on top:

  * the video latent ``[B, 16, T, H, W]`` is patchified by the fused scatter kernel as a ``T*H`` tall image
    (Conv3d(1,2,2) == the same 2x2 patch GEMM applied per frame), timestep sinusoid included;
  * q/k use a *full-width* RMSNorm (over all heads) followed by 3-D RoPE: a dedicated in-place kernel on
    the fused QKV GEMM output; attention then reads q/k/v through strided 4-D TMA views;
  * per-block modulation = learned table + time projection: one broadcast-add kernel builds all 40x6
    vectors, which the LayerNorm kernel / gated-residual epilogues index directly;
  * text cross-attention K/V depend only on the (step-invariant) context -> computed once per sampling run
    and cached (SURVEY K3 "constant conditioning is re-sent every step" in the reference);
  * head: AdaLN + Linear + unpatchify (+ Euler update, + NVLink peer store) in one GEMM epilogue.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from .. import ops
from . import block_linear, cache_workspace, quantize_block_weights, recognize, touch_workspace
from .graphs import GraphCache
from ..models import flux as flux_model
from ..models import wan as wan_model


def _bf(t: torch.Tensor, d) -> torch.Tensor:
    return t.detach().to(device=d, dtype=torch.bfloat16).contiguous()


class WanExecutor(nn.Module):
    pa_family = "wan"
    pa_native = True

    def __init__(self, model: "wan_model.WanModel", device, cuda_graphs: bool = False, fp8: bool = False):
        super().__init__()
        ops.require()
        d = self.device = torch.device(device)
        p = self.params = recognize.params_of(model, "wan")      # from weight shapes / ComfyUI attribute names
        self.dim, self.heads = p.dim, p.num_heads
        if p.dim // p.num_heads != 128 or tuple(p.patch_size) != (1, 2, 2) or p.in_dim != 16:
            raise ValueError("WanExecutor is specialised for head_dim 128, patch (1,2,2), 16 latent channels")
        W: Dict[str, Optional[torch.Tensor]] = {}

        def lin(name, m):
            W[name + ".w"] = _bf(m.weight, d)
            W[name + ".b"] = _bf(m.bias, d) if m.bias is not None else None

        W["patch.w"] = _bf(model.patch_embedding.weight.reshape(p.dim, -1), d)          # [dim, 16*1*2*2]
        W["patch.b"] = _bf(model.patch_embedding.bias, d)
        lin("text0", model.text_embedding[0])
        lin("text2", model.text_embedding[2])
        lin("time0", model.time_embedding[0])
        lin("time2", model.time_embedding[2])
        lin("tproj", model.time_projection[1])
        mods = []
        for i, blk in enumerate(model.blocks):
            sa, ca = blk.self_attn, blk.cross_attn
            W[f"b{i}.qkv.w"] = torch.cat([_bf(sa.q.weight, d), _bf(sa.k.weight, d), _bf(sa.v.weight, d)], 0).contiguous()
            W[f"b{i}.qkv.b"] = torch.cat([_bf(sa.q.bias, d), _bf(sa.k.bias, d), _bf(sa.v.bias, d)], 0).contiguous()
            W[f"b{i}.nq"], W[f"b{i}.nk"] = _bf(recognize.norm_scale(sa.norm_q), d), _bf(recognize.norm_scale(sa.norm_k), d)
            lin(f"b{i}.o", sa.o)
            lin(f"b{i}.cq", ca.q)
            W[f"b{i}.ckv.w"] = torch.cat([_bf(ca.k.weight, d), _bf(ca.v.weight, d)], 0).contiguous()
            W[f"b{i}.ckv.b"] = torch.cat([_bf(ca.k.bias, d), _bf(ca.v.bias, d)], 0).contiguous()
            W[f"b{i}.cnq"], W[f"b{i}.cnk"] = _bf(recognize.norm_scale(ca.norm_q), d), _bf(recognize.norm_scale(ca.norm_k), d)
            lin(f"b{i}.co", ca.o)
            W[f"b{i}.n3.g"], W[f"b{i}.n3.b"] = _bf(blk.norm3.weight, d), _bf(blk.norm3.bias, d)
            lin(f"b{i}.f0", blk.ffn[0])
            lin(f"b{i}.f2", blk.ffn[2])
            mods.append(_bf(blk.modulation.reshape(6 * p.dim), d))
        W["mod_table"] = torch.stack(mods, 0).contiguous()                               # [n_blocks, 6*dim]
        hm = _bf(model.head.modulation.reshape(2, p.dim), d)
        W["head_shift"], W["head_scale"] = hm[0:1].contiguous(), hm[1:2].contiguous()
        # WAN's head emits features ordered (pt, ph, pw, c); the fused unpatchify epilogue expects (c, ph, pw):
        # permute the rows once at build time instead of shuffling activations every step.
        hw_, hb_ = model.head.head.weight.detach(), model.head.head.bias.detach()
        oc = p.out_dim
        W["head.w"] = _bf(hw_.view(4, oc, p.dim).permute(1, 0, 2).reshape(4 * oc, p.dim), d)
        W["head.b"] = _bf(hb_.view(4, oc).permute(1, 0).reshape(4 * oc), d)
        self.W = W
        self.n_blocks = len(model.blocks)
        # fp8=True: the block GEMMs (qkv / o / cross q / cross o / ffn = ~all of the FLOPs) run as MXFP8 block-scaled
        # tcgen05 GEMMs (weights quantised once here, activations per GEMM); embedders, text K/V and the head stay bf16
        self.fp8 = bool(fp8)
        if self.fp8:
            names = [f"b{i}.{n_}" for i in range(self.n_blocks) for n_ in ("qkv", "o", "cq", "co", "f0", "f2")]
            quantize_block_weights(W, names, lambda _n: 224)
            torch.cuda.empty_cache()
        self.eps = p.eps
        self._ws: Dict[Tuple, dict] = {}
        self._graphs = GraphCache(self.device, enabled=cuda_graphs)
        self.launches_per_step = 0

    def parameters(self, recurse: bool = True):  # type: ignore[override]
        return iter(())

    def invalidate_conditioning(self) -> None:
        """Forget the cached text embedding / cross-attention K/V (call when the prompt changes in place)."""
        for ws in self._ws.values():
            ws["ctx_sig"] = None

    def release(self) -> None:
        self.W.clear()
        self._ws.clear()
        self._graphs.clear()

    def workspace(self, B: int, T: int, H: int, Wd: int, Lc: int) -> dict:
        key = (B, T, H, Wd, Lc)
        ws = touch_workspace(self._ws, key)
        if ws is not None:
            return ws
        d, dim = self.device, self.dim
        L = T * (H // 2) * (Wd // 2)
        e = lambda *s: torch.empty(*s, dtype=torch.bfloat16, device=d)  # noqa: E731
        ws = dict(B=B, T=T, H=H, Wd=Wd, L=L, Lc=Lc)
        ws["X"], ws["XM"] = e(B, L, dim), e(B, L, dim)
        ws["QKV"] = e(B, L, 3 * dim)
        ws["ATT"] = e(B, L, dim)
        ws["FF"] = e(B, L, self.params.ffn_dim)
        ws["T1"], ws["E_H"], ws["E"], ws["SE"] = e(B, 256), e(B, dim), e(B, dim), e(B, dim)
        ws["E0"] = e(B, 6 * dim)
        ws["MOD"] = e(B, self.n_blocks, 6 * dim)
        ws["HSHIFT"], ws["HSCALE"] = e(B, 1, dim), e(B, 1, dim)
        ws["CTX_H"], ws["CTX"] = e(B, Lc, dim), e(B, Lc, dim)
        ws["CQ"] = e(B, L, dim)
        ws["CKV"] = [e(B, Lc, 2 * dim) for _ in range(self.n_blocks)]
        ws["OUT"] = e(B, self.params.out_dim, T, H, Wd)
        m = wan_model.WanModel.__new__(wan_model.WanModel)
        ids = wan_model.WanModel.make_ids(m, 1, T, H // 2, Wd // 2, d)
        dd = 128
        pe = flux_model.EmbedND(dd, 10000, [dd - 4 * (dd // 6), 2 * (dd // 6), 2 * (dd // 6)])(ids)
        ws["ROPE"] = torch.stack([pe[0, 0, :, :, 0, 0], pe[0, 0, :, :, 1, 0]], -1).float().contiguous()
        ws["ctx_sig"] = None
        cache_workspace(self._ws, key, ws, device=self.device, on_evict=lambda _k: self._graphs.clear())
        return ws

    def _heads(self, t: torch.Tensor, which: int, n: int) -> torch.Tensor:
        """[B, L, n*dim] -> the ``which``-th [B, H, L, 128] strided view (no copy)."""
        b, l, _ = t.shape
        return t.view(b, l, n, self.heads, 128)[:, :, which].permute(0, 2, 1, 3)

    def _prepare_ctx(self, ws, ctx) -> int:
        """Text embedding + every block's cross-attention K/V (+ k RMSNorm): functions of the conditioning only, so
        they run EAGERLY and only when the conditioning changed (tensor identity / version, or after
        ``invalidate_conditioning``) - never inside the per-step CUDA graph, which reads the results from the
        workspace's fixed ``CKV`` buffers (the reference re-sends constant conditioning every step)."""
        sig = (ctx.data_ptr(), tuple(ctx.shape), ctx._version)
        if ws["ctx_sig"] == sig:
            return 0
        W, dim = self.W, self.dim
        C = ops.require()
        ops.gemm(ctx, W["text0.w"], "gelu", out=ws["CTX_H"], bias=W["text0.b"])
        ops.gemm(ws["CTX_H"], W["text2.w"], "bias", out=ws["CTX"], bias=W["text2.b"])
        for i in range(self.n_blocks):
            ops.gemm(ws["CTX"], W[f"b{i}.ckv.w"], "bias", out=ws["CKV"][i], bias=W[f"b{i}.ckv.b"])
            C.rms_rope(ws["CKV"][i][:, :, :dim], W[f"b{i}.cnk"], None, self.eps)
        ws["ctx_sig"] = sig
        return 2 + 2 * self.n_blocks

    def _run(self, ws, x_ptr, t, ctx, out, x_in=None, sigmas=None, out_ptr=None, out_sample_off=0, t_ptr=None,
             x_copy=None):
        W, dim, B, L = self.W, self.dim, ws["B"], ws["L"]
        C = ops.require()
        n = 0
        X, XM, QKV, ATT, FF, MOD, ROPE = ws["X"], ws["XM"], ws["QKV"], ws["ATT"], ws["FF"], ws["MOD"], ws["ROPE"]
        TH = ws["T"] * ws["H"]
        # fused scatter: video latent (as a T*H tall image) -> patch GEMM ; timestep sinusoid
        C.scatter_patch_embed(W["patch.w"], W["patch.b"], x_ptr, t_ptr if t_ptr is not None else t.data_ptr(), 0,
                              ws["T1"], None, x_copy, X, 16, TH, ws["Wd"], 1.0)
        ops.gemm(ws["T1"], W["time0.w"], "silu", out=ws["E_H"], bias=W["time0.b"])
        ops.gemm(ws["E_H"], W["time2.w"], "bias", out=ws["E"], bias=W["time2.b"])
        C.silu(ws["E"], ws["SE"])
        ops.gemm(ws["SE"], W["tproj.w"], "bias", out=ws["E0"], bias=W["tproj.b"])
        C.bcast_add(ws["E0"], W["mod_table"], MOD)
        C.bcast_add(ws["E"], W["head_shift"], ws["HSHIFT"])        # head uses e (not the 6-way projection)
        C.bcast_add(ws["E"], W["head_scale"], ws["HSCALE"])
        n += 8

        def mod(i, j):
            return MOD[:, i, j * dim:(j + 1) * dim]

        for i in range(self.n_blocks):
            # ---- self attention
            ops.layernorm_modulate(X, XM, scale=mod(i, 1), shift=mod(i, 0), eps=self.eps)
            block_linear(W, XM, f"b{i}.qkv", "bias", out=QKV)
            C.rms_rope(QKV[:, :, :dim], W[f"b{i}.nq"], ROPE, self.eps)
            C.rms_rope(QKV[:, :, dim:2 * dim], W[f"b{i}.nk"], ROPE, self.eps)
            ops.attention(self._heads(QKV, 0, 3), self._heads(QKV, 1, 3), self._heads(QKV, 2, 3), out=ATT)
            block_linear(W, ATT, f"b{i}.o", "gate_res", out=X, residual=X, gate=mod(i, 2))
            # ---- text cross attention
            ops.layernorm_modulate(X, XM, gamma=W[f"b{i}.n3.g"], beta=W[f"b{i}.n3.b"], eps=self.eps)
            block_linear(W, XM, f"b{i}.cq", "bias", out=ws["CQ"])
            C.rms_rope(ws["CQ"], W[f"b{i}.cnq"], None, self.eps)
            ops.attention(self._heads(ws["CQ"], 0, 1), self._heads(ws["CKV"][i], 0, 2), self._heads(ws["CKV"][i], 1, 2),
                          out=ATT)
            block_linear(W, ATT, f"b{i}.co", "res", out=X, residual=X)
            # ---- FFN
            ops.layernorm_modulate(X, XM, scale=mod(i, 4), shift=mod(i, 3), eps=self.eps)
            block_linear(W, XM, f"b{i}.f0", "gelu", out=FF)
            block_linear(W, FF, f"b{i}.f2", "gate_res", out=X, residual=X, gate=mod(i, 5))
            n += 14
        # ---- head: AdaLN + Linear + unpatchify (+ Euler, + peer store)
        ops.layernorm_modulate(X, XM, scale=ws["HSCALE"][:, 0], shift=ws["HSHIFT"][:, 0], eps=self.eps)
        kw = dict(bias=W["head.b"], C=self.params.out_dim, Hl=TH, Wl=ws["Wd"], xout_sample_off=out_sample_off)
        if out_ptr is not None:
            kw["x_out_ptr"] = out_ptr
        else:
            kw["x_out"] = out
        if sigmas is not None:
            kw["sigmas"], kw["x_in"] = sigmas, x_in
        ops.gemm(XM, W["head.w"], "euler_unpatch", **kw)
        n += 2
        self.launches_per_step = n
        return out

    def _prep(self, x, timesteps, context):
        d = self.device
        bf = lambda t: t.to(device=d, dtype=torch.bfloat16).contiguous()  # noqa: E731
        return bf(x), bf(timesteps), bf(context)

    @torch.no_grad()
    def forward(self, x, timesteps, context=None, clip_fea=None, transformer_options=None, **kwargs):
        with torch.cuda.device(self.device):
            x, timesteps, context = self._prep(x, timesteps, context)
            B, _, T, H, Wd = x.shape
            ws = self.workspace(B, T, H, Wd, context.shape[1])
            out = torch.empty_like(x)
            self._prepare_ctx(ws, context)
            self._run(ws, x.data_ptr(), timesteps, context, out)
            return out

    def _shard_args(self, x_src_ptr, shape, timesteps, context, out_ptr, out_sample_off):
        d = self.device
        timesteps = timesteps.to(device=d, dtype=torch.bfloat16).contiguous()
        context = context.to(device=d, dtype=torch.bfloat16).contiguous()
        key = ("shard", tuple(shape), x_src_ptr, timesteps.data_ptr(), context.data_ptr(), tuple(context.shape), out_ptr,
               out_sample_off)
        return key, timesteps, context

    @torch.no_grad()
    def forward_shard(self, x_src_ptr: int, shape, timesteps, context, out_ptr: int, out_sample_off: int, **_ignored):
        with torch.cuda.device(self.device):
            key, timesteps, context = self._shard_args(x_src_ptr, shape, timesteps, context, out_ptr, out_sample_off)
            ws = self.workspace(shape[0], shape[2], shape[3], shape[4], context.shape[1])
            self._prepare_ctx(ws, context)
            self._graphs.run(key, lambda: self._run(ws, x_src_ptr, timesteps, context, None, out_ptr=out_ptr,
                                                    out_sample_off=out_sample_off))

    def shard_graph_handle(self, x_src_ptr: int, shape, timesteps, context, out_ptr: int, out_sample_off: int,
                           **_ignored) -> int:
        with torch.cuda.device(self.device):
            key, _t, context = self._shard_args(x_src_ptr, shape, timesteps, context, out_ptr, out_sample_off)
            ws = self.workspace(shape[0], shape[2], shape[3], shape[4], context.shape[1])
            if ws["ctx_sig"] != (context.data_ptr(), tuple(context.shape), context._version):
                return 0                    # conditioning changed: take the Python path once (eager K/V precompute)
        return self._graphs.exec_handle(key)

    @torch.no_grad()
    def denoise_step(self, x, timesteps, context, sigmas, out=None, out_ptr=None, out_sample_off=0,
                     x_src_ptr: Optional[int] = None, t_src_ptr: Optional[int] = None):
        with torch.cuda.device(self.device):
            B, _, T, H, Wd = x.shape
            ws = self.workspace(B, T, H, Wd, context.shape[1])
            if out is None and out_ptr is None:
                out = ws["OUT"]
            self._prepare_ctx(ws, context)

            def body():
                self._run(ws, x_src_ptr if x_src_ptr is not None else x.data_ptr(), timesteps, context, out, x_in=x,
                          sigmas=sigmas, out_ptr=out_ptr, out_sample_off=out_sample_off, t_ptr=t_src_ptr,
                          x_copy=x if x_src_ptr is not None else None)

            key = (tuple(x.shape), x.data_ptr(), timesteps.data_ptr(), context.data_ptr(), tuple(context.shape),
                   sigmas.data_ptr(), out.data_ptr() if out is not None else 0, out_ptr or 0, out_sample_off,
                   x_src_ptr or 0, t_src_ptr or 0)
            self._graphs.run(key, body)
            return out


def build_wan_executor(model: nn.Module, device, **kw) -> WanExecutor:
    return WanExecutor(model, device, **kw)
