"""Drop-in mirror of the reference's LDM attention modules (surface 2).

    CrossAttention          ldm/modules/attention.py:146-222
    GEGLU / FeedForward     ldm/modules/attention.py:31-58
    BasicTransformerBlock   ldm/modules/attention.py:225-252

Module / parameter names are those of the reference, so SD-1.5 LDM checkpoints load with ``load_state_dict``.
Inside a block, LayerNorm -> projection GEMMs -> attention -> out-projection(+bias +residual) and
LayerNorm -> GEGLU GEMM -> out GEMM(+bias +residual) are kernels of libadaface_b200.so; the residual adds are
folded into the GEMM epilogues.  CUDA only, no fallback.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from . import autograd as ag


def _ver(*ts):
    return tuple(None if t is None else (t.data_ptr(), t._version) for t in ts)


def _bf16(t):
    return t.detach().to(torch.bfloat16).contiguous()


def _f32(t):
    return None if t is None else t.detach().float().contiguous()


def _processor_key_mask(mask, B):
    """The diffusers-processor rule for the same mask (dalc:254-273, ``CrossAttention.mask_mode = 'processor'``): if ANY instance's
    mask is all zero the mask is dropped for the whole batch; no uniform-attention instances."""
    m = mask.reshape(B, -1) != 0
    drop = (m.sum(dim=1) == 0).any()
    return (m | drop).to(torch.uint8).contiguous(), None


def _ldm_key_mask(mask, B):
    """attention.py:185-194 fills masked scores with ``-finfo.max`` (not -inf): every key of an instance whose mask is ALL zero
    gets the same score, so that instance attends uniformly to all keys (out = mean of V) -- realistic when the nearest resize of
    ``SpatialTransformer`` (:298) makes a small face mask vanish at 8 x 8 / 16 x 16.  Returns the uint8 key mask with such
    instances unmasked (so the kernels never see an empty row) and the per-instance empty flags; evaluated on the device."""
    m = mask.reshape(B, -1) != 0
    empty = m.sum(dim=1) == 0
    return (m | empty[:, None]).to(torch.uint8).contiguous(), empty


def _uniform_where_empty(o, v, empty):
    """Replace the output rows of empty-mask instances by the uniform-attention result mean_j V[j] (per channel = per head)."""
    vmean = v.float().mean(dim=1, keepdim=True).to(o.dtype)
    return torch.where(empty[:, None, None], vmean.expand_as(o), o)


class CrossAttention(nn.Module):
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.):
        super().__init__()
        inner_dim = dim_head * heads
        context_dim = context_dim if context_dim is not None else query_dim
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner_dim, bias=False)
        self.to_k = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_v = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, query_dim), nn.Dropout(dropout))
        if dropout != 0.:
            raise NotImplementedError("SD-1.5 uses dropout 0 in every attention block")
        self.save_cross_attn_vars = False
        self.cached_activations = None
        self._pack_key, self._pack = None, None
        # Optional AttnProcessor_LoRA_Capture (the diffusers surface of the same operator) installed by
        # unet_wrapper.set_up_attn_processors on the captured layers: the block then routes this module through it
        # (LoRA / normalize / capture with the processor's keys and C^-1/4 factor).  ``processor_kwargs`` = the per-call
        # cross_attention_kwargs ({'img_mask', 'subj_indices'}, ddpm.py:4227-4229).
        self.processor = None
        self.processor_kwargs = None
        self.mask_mode = "ldm"            # 'ldm': attention.py:185-194 (-finfo.max fill); 'processor': dalc:254-273 (drop if any empty)

    def run_processor(self, x16, context):
        kw = {k: v for k, v in (self.processor_kwargs or {}).items() if k in ("img_mask", "subj_indices")}
        if context is not None:
            kw.pop("img_mask", None)                 # the mask only reaches self-attention (dalc:254: `if img_mask is not None and not is_cross`)
        return self.processor(self, x16, encoder_hidden_states=context, **kw)

    def _weights(self):
        ws = (self.to_q.weight, self.to_k.weight, self.to_v.weight, self.to_out[0].weight, self.to_out[0].bias)
        key = _ver(*ws)
        if key != self._pack_key:
            with torch.no_grad():
                pk = {"wq": _bf16(ws[0]), "wo": _bf16(ws[3]), "bo": _f32(ws[4]),
                      "wkv": torch.cat([_bf16(ws[1]), _bf16(ws[2])], dim=0)}
                inner = pk["wq"].shape[0]
                pk["wk"], pk["wv"] = pk["wkv"][:inner], pk["wkv"][inner:]      # row-slice views (contiguous)
                if ws[0].shape[1] == ws[1].shape[1]:
                    pk["wqkv"] = torch.cat([pk["wq"], pk["wkv"]], dim=0)
            self._pack, self._pack_key = pk, key
        return self._pack

    def _attend_train(self, x16, context, mask, residual=None, out_dtype=torch.bfloat16):
        """``_attend`` with every kernel paired with its backward (autograd.py): gradients reach x and the context;
        the frozen U-Net weights get none (ddpm.py:637-638)."""
        B, N, Cq = x16.shape
        H = self.heads
        pk = self._weights()
        C = pk["wq"].shape[0]
        x2d = x16.reshape(B * N, Cq)
        q = prob = score = None
        key_mask, empty = (None, None) if mask is None else (_processor_key_mask if self.mask_mode == "processor" else _ldm_key_mask)(mask, B)
        if context is None:
            if self.save_cross_attn_vars:
                raise NotImplementedError("save_cross_attn_vars is only set on cross-attention layers (attn2)")
            qkv = ag.linear(x2d, pk, "wqkv").view(B, N, 3 * C)
            o = ag.attention(qkv, qkv, qkv, (0, C, 2 * C), C, C, H, self.scale, key_mask)
            if empty is not None:
                o = _uniform_where_empty(o, qkv[:, :, 2 * C:], empty)
        else:
            ctx = context.to(torch.bfloat16).contiguous()
            S = ctx.shape[1]
            c2d = ctx.view(B * S, ctx.shape[2])
            if self.save_cross_attn_vars:
                if key_mask is not None or S > 128:
                    raise NotImplementedError("capture is only defined for cross-attention contexts (<= 128 keys, no mask)")
                q = ag.linear(x2d, pk, "wq", out_dtype=torch.float32).view(B, N, C)
                k = ag.linear(c2d, pk, "wk", out_dtype=torch.float32).view(B, S, C)
                v = ag.linear(c2d, pk, "wv", out_dtype=torch.float32).view(B, S, C)
                one = torch.ones((), device=x16.device)
                o, prob, score, _ = ag.CrossCaptureFn.apply(q, k, v, one, H, self.scale, True, True, None, None, False, 1.0)
            else:
                q = ag.linear(x2d, pk, "wq").view(B, N, C)
                kv = ag.linear(c2d, pk, "wkv").view(B, S, 2 * C)
                o = ag.attention(q, kv, kv, (0, 0, C), C, C, H, self.scale, key_mask)
                if empty is not None:
                    o = _uniform_where_empty(o, kv[:, :, C:], empty)
        res2d = None if residual is None else residual.reshape(B * N, -1)
        out = ag.linear(o.view(B * N, C), pk, "wo", "bo", residual=res2d, out_dtype=out_dtype).view(B, N, -1)
        if self.save_cross_attn_vars:
            if residual is not None:
                raise RuntimeError("capture with a fused residual would corrupt cached 'attn_out'")
            self.cached_activations = {"q": ag.ChanMajorFn.apply(q, math.sqrt(self.scale)), "attn": prob, "attnscore": score,
                                       "attn_out": ag.ChanMajorFn.apply(out, 1.0)}
        return out

    def _attend(self, x16, context, mask, residual=None, out_dtype=torch.bfloat16):
        """x16 [B,N,C] bf16 contiguous -> [B,N,query_dim]; ``residual`` is added in the out-projection epilogue."""
        B, N, Cq = x16.shape
        H = self.heads
        pk = self._weights()
        C = pk["wq"].shape[0]
        x2d = x16.view(B * N, Cq)
        prob = score = None
        key_mask, empty = (None, None) if mask is None else (_processor_key_mask if self.mask_mode == "processor" else _ldm_key_mask)(mask, B)       # attention.py:185-194
        if context is None:
            if self.save_cross_attn_vars:
                raise NotImplementedError("save_cross_attn_vars is only set on cross-attention layers (attn2)")
            if key_mask is None:
                o = ops.self_attention_fused_qkv(x2d, pk["wqkv"], None, B, N, H, self.scale)
            else:
                qkv = ops.proj(x2d, pk["wqkv"]).view(B, N, 3 * C)
                o = ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], H, self.scale, key_mask=key_mask)
                o = _uniform_where_empty(o, qkv[:, :, 2 * C:], empty)
        else:
            ctx = context.to(torch.bfloat16).contiguous()
            S = ctx.shape[1]
            c2d = ctx.view(B * S, ctx.shape[2])
            if self.save_cross_attn_vars:
                if key_mask is not None or S > 128:
                    raise NotImplementedError("capture is only defined for cross-attention contexts (<= 128 keys, no mask)")
                q = ops.proj(x2d, pk["wq"], out_dtype=torch.float32).view(B, N, C)      # fp32 q/k/v feed the capture kernel
                kv = ops.proj(c2d, pk["wkv"], out_dtype=torch.float32).view(B, S, 2 * C)
                o, prob, score, _ = ops.attention_cross_capture(q, kv[:, :, :C], kv[:, :, C:], H, self.scale)
            elif key_mask is None:
                o = ops.cross_attention_fused(x2d, pk["wq"], None, c2d, pk["wkv"], None, B, N, S, H, self.scale)
            else:
                q = ops.proj(x2d, pk["wq"]).view(B, N, C)
                kv = ops.proj(c2d, pk["wkv"]).view(B, S, 2 * C)
                o = ops.attention(q, kv[:, :, :C], kv[:, :, C:], H, self.scale, key_mask=key_mask)
                o = _uniform_where_empty(o, kv[:, :, C:], empty)
        res2d = None if residual is None else residual.view(B * N, -1)
        out = ops.proj(o.view(B * N, C), pk["wo"], bias=pk["bo"], residual=res2d, out_dtype=out_dtype).view(B, N, -1)
        if self.save_cross_attn_vars:                                             # attention.py:207-220
            if residual is not None:
                raise RuntimeError("capture with a fused residual would corrupt cached 'attn_out'")
            self.cached_activations = {"q": ops.chan_major(q, math.sqrt(self.scale)), "attn": prob, "attnscore": score,
                                       "attn_out": ops.chan_major(out, 1.0)}
        return out

    def forward(self, x, context=None, mask=None):
        if not x.is_cuda:
            raise RuntimeError("adaface_b200 CrossAttention runs on CUDA only (no CPU fallback)")
        if torch.is_grad_enabled() and (x.requires_grad or (context is not None and context.requires_grad)):
            return self._attend_train(x.to(torch.bfloat16).contiguous(), context, mask).to(x.dtype)
        return self._attend(x.to(torch.bfloat16).contiguous(), context, mask).to(x.dtype)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(nn.Module):
    """FeedForward(dim, glu=True): net = [GEGLU(dim, 4 dim), Dropout, Linear(4 dim, dim)] (attention.py:41-58)."""

    def __init__(self, dim, dim_out=None, mult=4, glu=True, dropout=0.):
        super().__init__()
        if not glu:
            raise NotImplementedError("SD-1.5 transformer blocks use the gated feed-forward (gated_ff=True)")
        inner_dim = int(dim * mult)
        if inner_dim % 64:
            raise ValueError("GEGLU inner dim must be a multiple of 64 (packed [a|gate] tiles)")
        self.net = nn.Sequential(GEGLU(dim, inner_dim), nn.Dropout(dropout), nn.Linear(inner_dim, dim_out or dim))
        self._pack_key, self._pack = None, None

    def _weights(self):
        p, o = self.net[0].proj, self.net[2]
        key = _ver(p.weight, p.bias, o.weight, o.bias)
        if key != self._pack_key:
            with torch.no_grad():
                inner = p.weight.shape[0] // 2
                # pack rows as [a(64) | gate(64)] per 128-column tile so that the epilogue sees both halves
                idx = torch.arange(inner, device=p.weight.device).view(-1, 64)
                perm = torch.cat([idx, idx + inner], dim=1).reshape(-1)
                self._pack = {"w1": _bf16(p.weight[perm]), "b1": _f32(p.bias[perm]), "w2": _bf16(o.weight), "b2": _f32(o.bias)}
            self._pack_key = key
        return self._pack

    def _ff(self, h16, residual=None, out_dtype=torch.bfloat16):
        pk = self._weights()
        g = ops.proj(h16, pk["w1"], bias=pk["b1"], act=ops.ACT_GEGLU)
        return ops.proj(g, pk["w2"], bias=pk["b2"], residual=residual, out_dtype=out_dtype)

    def _ff_train(self, h16, residual=None, out_dtype=torch.bfloat16):
        """Training form: the packed pre-activation [a | gate] is kept for the GEGLU backward kernel."""
        pk = self._weights()
        u = ag.linear(h16, pk, "w1", "b1")
        return ag.linear(ag.ActFn.apply(u, ops.ACT_GEGLU), pk, "w2", "b2", residual=residual, out_dtype=out_dtype)

    def forward(self, x):
        shp = x.shape
        ff = self._ff_train if (torch.is_grad_enabled() and x.requires_grad) else self._ff
        y = ff(x.to(torch.bfloat16).contiguous().view(-1, shp[-1]))
        return y.view(*shp[:-1], -1).to(x.dtype)


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, n_heads, d_head, dropout=0., context_dim=None, gated_ff=True, checkpoint=True):
        super().__init__()
        self.attn1 = CrossAttention(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = CrossAttention(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.checkpoint = checkpoint     # activation checkpointing is a no-op for the forward-only kernels

    def _ln(self, x2d, norm):
        return ops.layernorm(x2d, norm.weight.detach().float(), norm.bias.detach().float(), norm.eps)

    def forward(self, x, context=None, mask=None):
        """attention.py:242-252: x1 = attn1(LN1 x, mask) + x; x2 = x1 + attn2(LN2 x1, ctx); x3 = FF(LN3 x2) + x2."""
        if not x.is_cuda:
            raise RuntimeError("adaface_b200 BasicTransformerBlock runs on CUDA only (no CPU fallback)")
        B, N, C = x.shape
        proc = self.attn2.processor
        if torch.is_grad_enabled() and (x.requires_grad or (context is not None and context.requires_grad)
                                        or (proc is not None and proc._has_trainable())):
            return self._forward_train(x, context, mask)
        x0 = x.to(torch.bfloat16).contiguous()
        capture = self.attn2.save_cross_attn_vars
        h = self._ln(x0.view(B * N, C), self.norm1).view(B, N, C)
        x1 = self.attn1._attend(h, None, mask, residual=x0)
        h = self._ln(x1.view(B * N, C), self.norm2).view(B, N, C)
        if self.attn2.processor is not None:
            x2 = self.attn2.run_processor(h, context) + x1
        elif capture:      # cached attn_out must be the bare attention output (attention.py:220)
            x2 = self.attn2._attend(h, context, None) + x1
        else:
            x2 = self.attn2._attend(h, context, None, residual=x1)
        h = self._ln(x2.view(B * N, C), self.norm3)
        x3 = self.ff._ff(h, residual=x2.view(B * N, C)).view(B, N, C)
        return x3.to(x.dtype)

    def _forward_train(self, x, context, mask):
        """The same block with every kernel paired with its backward; LayerNorm affine parameters are frozen U-Net
        weights, so only dx is produced.  With activation checkpointing on in the reference (attention.py:239) the
        block is recomputed there; here the backward is recompute-form at the kernel level (attention) and keeps only
        the bf16 activations between kernels."""
        B, N, C = x.shape
        ln = lambda t, norm: ag.LayerNormFn.apply(t, norm.weight.detach(), norm.bias.detach(), norm.eps, torch.bfloat16)
        x0 = x.to(torch.bfloat16).contiguous().view(B * N, C)
        capture = self.attn2.save_cross_attn_vars
        x1 = self.attn1._attend_train(ln(x0, self.norm1).view(B, N, C), None, mask, residual=x0).view(B * N, C)
        h = ln(x1, self.norm2).view(B, N, C)
        if self.attn2.processor is not None:
            x2 = (self.attn2.run_processor(h, context) + x1.view(B, N, C)).view(B * N, C)
        elif capture:
            x2 = (self.attn2._attend_train(h, context, None) + x1.view(B, N, C)).view(B * N, C)
        else:
            x2 = self.attn2._attend_train(h, context, None, residual=x1).view(B * N, C)
        x3 = self.ff._ff_train(ln(x2, self.norm3), residual=x2).view(B, N, C)
        return x3.to(x.dtype)


class SpatialTransformer(nn.Module):
    """ldm/modules/attention.py:254-304: GroupNorm(32, eps 1e-6) -> 1x1 proj_in -> tokens ->
    BasicTransformerBlock x depth -> 1x1 proj_out (zero-initialised) -> + input.  Same module / parameter names as the
    reference (norm, proj_in, transformer_blocks.N, proj_out), so SD-1.5 LDM checkpoints load as they are.
    The norm is fused with the NCHW -> tokens re-layout, the 1x1 convolutions are projection GEMMs, and the way back to
    NCHW is fused with the residual.  ``forward_tokens`` is the NHWC-resident, differentiable entry (frozen weights)."""

    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0., context_dim=None):
        super().__init__()
        self.in_channels = in_channels
        inner_dim = n_heads * d_head
        self.norm = nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner_dim, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner_dim, n_heads, d_head, dropout=dropout, context_dim=context_dim) for _ in range(depth)])
        self.proj_out = nn.Conv2d(inner_dim, in_channels, kernel_size=1, stride=1, padding=0)
        for p in self.proj_out.parameters():          # zero_module (attention.py:280)
            nn.init.zeros_(p)
        self._pack_key, self._pack = None, None

    def _weights(self):
        ps = (self.norm.weight, self.norm.bias, self.proj_in.weight, self.proj_in.bias, self.proj_out.weight, self.proj_out.bias)
        key = _ver(*ps)
        if key != self._pack_key:
            with torch.no_grad():
                self._pack = {"gn_w": _f32(ps[0]), "gn_b": _f32(ps[1]), "w_in": _bf16(ps[2].flatten(1)), "b_in": _f32(ps[3]),
                              "w_out": _bf16(ps[4].flatten(1)), "b_out": _f32(ps[5])}
            self._pack_key = key
        return self._pack

    def forward_tokens(self, t_in, hw, context=None, mask=None):
        """NHWC-resident entry (the U-Net mirror keeps activations as tokens): t_in bf16 [B, h*w, C] -> bf16 [B, h*w, C].
        Same arithmetic as forward(); the norm runs tokens -> tokens and the residual rides in proj_out's epilogue.
        Differentiable w.r.t. the tokens and the context (frozen weights: GroupNormActFn, FrozenLinearFn, the block's
        training path)."""
        b, n, c = t_in.shape
        h, w = hw
        pk = self._weights()
        train = ag.needs_grad(t_in, context)
        t = ag.groupnorm_act(t_in, pk["gn_w"], pk["gn_b"], self.norm.num_groups, self.norm.eps, False)                # :291
        if train:
            t = ag.linear(t.view(b * n, c), pk, "w_in", "b_in").view(b, n, -1)
        else:
            t = ops.proj(t.view(b * n, c), pk["w_in"], bias=pk["b_in"]).view(b, n, -1)                               # :292
        for block in self.transformer_blocks:
            block.attn2.infeat_size = (h, w)
            mask2 = F.interpolate(mask, size=(h, w), mode="nearest") if mask is not None else None
            t = block(t, context=context, mask=mask2)
        if train or ag.needs_grad(t):                  # (a trainable adapter inside a block also starts the autograd graph)
            return ag.linear(t.reshape(b * n, -1), pk, "w_out", "b_out", residual=t_in.view(b * n, c)).view(b, n, c)
        return ops.proj(t.reshape(b * n, -1), pk["w_out"], bias=pk["b_out"], residual=t_in.view(b * n, c)).view(b, n, c)   # :303-304

    def forward(self, x, context=None, mask=None):
        if not x.is_cuda:
            raise RuntimeError("adaface_b200 SpatialTransformer runs on CUDA only (no CPU fallback)")
        if torch.is_grad_enabled() and x.requires_grad:
            raise NotImplementedError("SpatialTransformer.forward: gradients w.r.t. an NCHW input are not built; use forward_tokens "
                                      "(differentiable w.r.t. tokens and context) or the UNetModel mirror")
        b, c, h, w = x.shape
        if ag.needs_grad(context):                     # training through the context: the tokens path carries the backward
            from .ldm_unet_blocks import _to_nchw, _to_tokens
            return _to_nchw(self.forward_tokens(_to_tokens(x), (h, w), context=context, mask=mask), (h, w), x.dtype)
        pk = self._weights()
        x_in = x.contiguous()
        t = ops.groupnorm_tokens(x_in, pk["gn_w"], pk["gn_b"], self.norm.num_groups, self.norm.eps)      # :291, 293
        t = ops.proj(t.view(b * h * w, c), pk["w_in"], bias=pk["b_in"]).view(b, h * w, -1)                # :292
        for block in self.transformer_blocks:
            block.attn2.infeat_size = (h, w)                                                             # :296
            mask2 = F.interpolate(mask, size=(h, w), mode="nearest") if mask is not None else None       # :298
            t = block(t, context=context, mask=mask2)
        t = ops.proj(t.reshape(b * h * w, -1), pk["w_out"], bias=pk["b_out"]).view(b, h * w, c)           # :303
        return ops.tokens_to_nchw_add(t, x_in)                                                            # :301, 304
