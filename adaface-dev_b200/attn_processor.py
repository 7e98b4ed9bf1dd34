"""Drop-in mirror of the reference's attention processor surface (surface 1).

    AttnProcessor_LoRA_Capture          adaface/diffusers_attn_lora_capture.py:142-364
    ScaleGrad / GradientScaler / gen_gradient_scaler                          :23-67
    LoraDoraLinear                      stands in for peft ``lora.Linear(..., use_dora=True)`` (:171-181)
    Attention                           minimal holder with the attributes the processor reads from a
                                        diffusers ``Attention`` (:212-342); diffusers itself is not required.

Same constructor, ``reset_attn_cache_and_flags``, ``cached_activations`` and ``__call__`` contract; the
arithmetic runs in libadaface_b200.so (bf16 in, fp32 accumulate).  There is no PyTorch fallback.
"""
import math
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


# ------------------------------------------------------------------------------------------------ A5
class ScaleGrad(torch.autograd.Function):
    """Identity forward, grad * alpha backward (dalc:23-42)."""

    @staticmethod
    def forward(ctx, input_, alpha_, debug=False):
        ctx.save_for_backward(alpha_)
        return input_.view_as(input_)

    @staticmethod
    def backward(ctx, grad_output):
        (alpha_,) = ctx.saved_tensors
        return (grad_output * alpha_ if ctx.needs_input_grad[0] else None), None, None


class GradientScaler(nn.Module):
    def __init__(self, alpha=1.0, debug=False):
        super().__init__()
        self._alpha = torch.tensor(alpha, requires_grad=False)
        self._debug = torch.tensor(debug, requires_grad=False)

    def forward(self, input_):
        if not (torch.is_grad_enabled() and input_.requires_grad):
            return input_                      # identity forward: nothing to scale without a backward pass
        if self._alpha.device != input_.device:
            self._alpha = self._alpha.to(input_.device)      # moved once, not per call (CUDA-graph friendly)
        return ScaleGrad.apply(input_, self._alpha, False)


def gen_gradient_scaler(alpha, debug=False):
    """dalc:59-67: alpha == 1 -> Identity, alpha == 0 -> detach, otherwise GradientScaler."""
    if alpha == 1:
        return nn.Identity()
    if alpha > 0:
        return GradientScaler(alpha, debug=debug)
    if alpha != 0:
        raise ValueError("gradient scale must be >= 0")
    return torch.detach


# ------------------------------------------------------------------------------------------------ A4
class _Magnitude(nn.Module):
    def __init__(self, w):
        super().__init__()
        self.weight = nn.Parameter(w)


class LoraDoraLinear(nn.Module):
    """Parameter container with peft's ``lora.Linear`` layout: ``base_layer``, ``lora_A['default']``,
    ``lora_B['default']`` (nn.Linear, no bias) and ``lora_magnitude_vector['default'].weight``.
    Init as peft: A Kaiming-uniform(a=sqrt 5), B zero, magnitude = ||W||_row (identity adapter).
    The arithmetic is fused into the projection kernel by the processor."""

    def __init__(self, base_layer, adapter_name="default", r=192, lora_alpha=16, use_dora=True, lora_dropout=0.1):
        super().__init__()
        if not use_dora:
            raise NotImplementedError("the reference always uses DoRA (lora_uses_dora=True, dalc:499)")
        self.base_layer = base_layer
        self.r, self.lora_alpha, self.scaling = r, lora_alpha, lora_alpha / r
        self.adapter = adapter_name
        dev = base_layer.weight.device
        self.lora_A = nn.ModuleDict({adapter_name: nn.Linear(base_layer.in_features, r, bias=False, device=dev)})
        self.lora_B = nn.ModuleDict({adapter_name: nn.Linear(r, base_layer.out_features, bias=False, device=dev)})
        nn.init.kaiming_uniform_(self.lora_A[adapter_name].weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B[adapter_name].weight)
        mag = torch.linalg.norm(base_layer.weight.detach().float(), dim=1)
        self.lora_magnitude_vector = nn.ModuleDict({adapter_name: _Magnitude(mag)})
        self._pack_key, self._pack = None, None
        # adapter management is neutered on these modules by the reference (dalc:531-532)
        self.enable_adapters = lambda *a, **k: None
        self.set_adapter = lambda *a, **k: None

    def pack(self):
        """bf16 operands of the fused kernel: A [r,in], Bs = s*B [out,r], colscale = m / ||W + s B A||_row (fp32,
        detached).  Rebuilt only when a parameter changed (training: every step; inference: once)."""
        A, B = self.lora_A[self.adapter].weight, self.lora_B[self.adapter].weight
        m, W = self.lora_magnitude_vector[self.adapter].weight, self.base_layer.weight
        key = tuple((t.data_ptr(), t._version) for t in (A, B, m, W))
        if key != self._pack_key:
            with torch.no_grad():
                A16 = A.detach().to(torch.bfloat16).contiguous()
                B16 = B.detach().to(torch.bfloat16).contiguous()
                Wd = W.detach()
                cs = ops.dora_colscale(Wd if Wd.dtype in (torch.float32, torch.bfloat16) else Wd.float(), A16, B16, self.scaling, m)
                self._pack = (A16, (B.detach().float() * self.scaling).to(torch.bfloat16).contiguous(), cs)
            self._pack_key = key
        return self._pack

    def invalidate(self):
        """Force the next pack() to rebuild the operands from the live parameters (call before capturing a training step into a
        CUDA graph, so that the rebuild is part of the graph and every replay sees the optimiser's latest update)."""
        self._pack_key = None


class Attention(nn.Module):
    """The attributes AttnProcessor_LoRA_Capture reads from a diffusers ``Attention``: for SD-1.5
    all norms are None, to_q/k/v have no bias, to_out = [Linear(bias), Dropout(0)], no residual, rescale 1."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, processor=None, device=None):
        super().__init__()
        inner = heads * dim_head
        ctx = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.spatial_norm = self.group_norm = self.norm_cross = self.norm_q = self.norm_k = None
        self.residual_connection, self.rescale_output_factor = False, 1.0
        self.to_q = nn.Linear(query_dim, inner, bias=False, device=device)
        self.to_k = nn.Linear(ctx, inner, bias=False, device=device)
        self.to_v = nn.Linear(ctx, inner, bias=False, device=device)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, device=device), nn.Dropout(0.0)])
        self.processor = processor if processor is not None else AttnProcessor_LoRA_Capture()

    def set_processor(self, processor):
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **cross_attention_kwargs)


# ------------------------------------------------------------------------------------------------ weights
def _bf16_pack(attn):
    """bf16 copies of the frozen projection weights, fused where one GEMM can serve several projections.
    Cached on the ``attn`` module and refreshed when any weight changes (``_version`` / storage)."""
    ws = (attn.to_q.weight, attn.to_k.weight, attn.to_v.weight, attn.to_out[0].weight)
    bs = (attn.to_q.bias, attn.to_k.bias, attn.to_v.bias, attn.to_out[0].bias)
    key = tuple((t.data_ptr(), t._version) for t in ws) + tuple(None if b is None else (b.data_ptr(), b._version) for b in bs)
    pk = getattr(attn, "_adaface_b200_pack", None)
    if pk is not None and pk["key"] == key:
        return pk
    with torch.no_grad():
        to16 = lambda t: t.detach().to(torch.bfloat16).contiguous()
        f32 = lambda t: None if t is None else t.detach().float().contiguous()
        pk = {"key": key, "wq": to16(ws[0]), "wk": to16(ws[1]), "wv": to16(ws[2]), "wo": to16(ws[3]),
              "bq": f32(bs[0]), "bk": f32(bs[1]), "bv": f32(bs[2]), "bo": f32(bs[3])}
        pk["wkv"] = torch.cat([pk["wk"], pk["wv"]], dim=0)
        pk["bkv"] = None if bs[1] is None and bs[2] is None else torch.cat(
            [f32(b) if b is not None else torch.zeros(w.shape[0], device=w.device) for b, w in ((bs[1], ws[1]), (bs[2], ws[2]))])
        if ws[0].shape[1] == ws[1].shape[1]:
            pk["wqkv"] = torch.cat([pk["wq"], pk["wk"], pk["wv"]], dim=0)
            pk["bqkv"] = None if all(b is None for b in bs[:3]) else torch.cat(
                [f32(b) if b is not None else torch.zeros(w.shape[0], device=w.device) for b, w in zip(bs[:3], ws[:3])])
    attn._adaface_b200_pack = pk
    return pk


def _linear(x2d, w16, bias, lora: Optional[LoraDoraLinear], **kw):
    """x W^T (+ DoRA-scaled LoRA update) + bias, one fused GEMM launch (+ one skinny launch for T = x A^T)."""
    if lora is None:
        return ops.proj(x2d, w16, bias=bias, **kw)
    A16, Bs16, colscale = lora.pack()
    t = ops.proj(x2d, A16)
    return ops.proj(x2d, w16, t=t, bs=Bs16, colscale=colscale, bias=bias, **kw)


def _token_flags(B, S, ib, in_, device):
    """uint8 [B, S] with ones at (ib[k], in_[k]).  The value operand is a DEVICE tensor: ``flag[ib, in_] = 1`` would stage the
    Python scalar through a host tensor, which CUDA-graph capture forbids."""
    flag = torch.zeros((B, S), device=device, dtype=torch.uint8)
    ib, in_ = ib.to(device).long(), in_.to(device).long()
    return flag.index_put_((ib, in_), torch.ones(ib.numel(), device=device, dtype=torch.uint8))


def img_mask_to_key_mask(img_mask, n_tokens):
    """dalc:254-273: nearest-resize the [B,1,H,W] mask to sqrt(N) x sqrt(N), use it as a KEY mask, and drop it for
    the whole batch if any instance's resized mask is all zero -- evaluated on the device (no host sync)."""
    ms = int(math.sqrt(n_tokens))
    m = F.interpolate(img_mask.float(), size=(ms, ms), mode="nearest").reshape(img_mask.shape[0], -1) != 0
    drop = (m.sum(dim=1) == 0).any()
    return (m | drop).to(torch.uint8).contiguous()


# ------------------------------------------------------------------------------------------------ A1
class AttnProcessor_LoRA_Capture(nn.Module):
    r"""B200-native ``AttnProcessor_LoRA_Capture`` (dalc:142-364); same constructor and call contract."""

    def __init__(self, capture_ca_activations: bool = False, enable_lora: bool = False, lora_uses_dora=True,
                 lora_proj_layers=None, lora_rank: int = 192, lora_alpha: float = 16, q_lora_updates_query=False,
                 attn_proc_idx=-1):
        super().__init__()
        self.global_enable_lora = enable_lora
        self.attn_proc_idx = attn_proc_idx
        self.reset_attn_cache_and_flags(capture_ca_activations, False, False, enable_lora)
        self.lora_rank, self.lora_alpha = lora_rank, lora_alpha
        self.lora_scale = self.lora_alpha / self.lora_rank
        self.q_lora_updates_query = q_lora_updates_query
        # subject-columns-only capture (optional mode): also emit cached_activations['attn_subj']
        self.capture_subj_cols_only = False
        self.capture_consumers = None          # set_capture_consumers(): fused reductions of the captured map
        self.to_q_lora = self.to_k_lora = self.to_v_lora = self.to_out_lora = None
        # Reference quirk 2 (fixed): always defined, so capture works with LoRA globally off (dalc:164-168, 314).
        self.cross_attn_scale_factor = nn.Parameter(torch.tensor(0.8), requires_grad=True)
        if self.global_enable_lora:
            for name, layer in (lora_proj_layers or {}).items():
                if name not in ("q", "k", "v", "out"):
                    raise ValueError(f"unknown LoRA projection '{name}'")
                setattr(self, f"to_{name}_lora", LoraDoraLinear(layer, "default", r=lora_rank, lora_alpha=lora_alpha,
                                                                use_dora=lora_uses_dora, lora_dropout=0.1))

    def set_capture_consumers(self, subj_sum=False, ref_attn=None, keep_attn=False):
        """B200 extension: while capturing, REDUCE the probability map inside the attention kernel instead of
        writing [B,8,N,S] fp32 -- ``cached_activations['attn_subj_sum']`` [B,H,N] = mass on each instance's subject columns (what
        calc_subj_masked_bg_suppress_loss reads, ldm/util.py:1862-1868) when ``subj_sum``; ``['attn_sqdiff']`` [B] = sum over
        (h, i, j) of (attn - ref_attn)^2 against the detached sc_rep map ``ref_attn`` [B,H,N,S] (calc_sc_rep_attn_distill_loss,
        :2084-2089).  Both are differentiable.  'attn' / 'attnscore' are then only produced when ``keep_attn``.  Call with no
        arguments to switch the consumers off."""
        self.capture_consumers = {"subj_sum": bool(subj_sum), "ref_attn": ref_attn, "keep_attn": bool(keep_attn)} \
            if (subj_sum or ref_attn is not None) else None

    def _consume(self, q, k, v, H, sm_scale, B, S, subj_indices, train):
        """The capture call with fused consumers (both the autograd and the no-grad form)."""
        from . import autograd as ag
        cons = self.capture_consumers
        if self.mix_attn_mats_in_batch:
            raise NotImplementedError("capture consumers are not defined together with mix_attn_mats_in_batch")
        col_flag, _ = self._subj_aux(B, S, subj_indices, q.device)
        sum_flag = None
        if cons["subj_sum"]:
            if subj_indices is None:
                raise ValueError("capture consumer 'subj_sum' requires subj_indices")
            ib, in_ = subj_indices
            sum_flag = _token_flags(B, S, ib, in_, q.device)
        ref = cons["ref_attn"]
        if ref is not None:
            ref = ref.detach().float().contiguous()
        if self.cross_attn_scale_factor.device != q.device:
            self.cross_attn_scale_factor.data = self.cross_attn_scale_factor.data.to(q.device)
        prob = None
        if train:
            o, subj_sum, sqdiff = ag.CrossConsumeFn.apply(q, k, v, self.cross_attn_scale_factor, H, sm_scale, col_flag, sum_flag, ref, 10.0)
        else:
            qm = ops.qmean(q) if col_flag is not None else None
            o, subj_sum, sqdiff, prob = ops.attention_cross_consume(
                q, k, v, H, sm_scale, sum_flag=sum_flag, ref_prob=ref, want_prob=cons["keep_attn"], col_flag=col_flag, qmean=qm,
                ca_scale=self.cross_attn_scale_factor.detach().float().reshape(1))
        extra = {}
        if sum_flag is not None:
            extra["attn_subj_sum"] = subj_sum
        if ref is not None:
            extra["attn_sqdiff"] = sqdiff
        return o, prob, extra

    def reset_attn_cache_and_flags(self, capture_ca_activations, normalize_cross_attn, mix_attn_mats_in_batch, enable_lora):
        """dalc:184-190."""
        self.capture_ca_activations = capture_ca_activations
        self.normalize_cross_attn = normalize_cross_attn
        self.mix_attn_mats_in_batch = mix_attn_mats_in_batch
        self.cached_activations = {}
        self.enable_lora = enable_lora and self.global_enable_lora

    def __call__(self, attn, hidden_states: torch.Tensor, encoder_hidden_states: Optional[torch.Tensor] = None,
                 attention_mask: Optional[torch.Tensor] = None, temb: Optional[torch.Tensor] = None,
                 img_mask: Optional[torch.Tensor] = None,
                 subj_indices: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, debug: bool = False, *args,
                 **kwargs) -> torch.Tensor:
        for name in ("spatial_norm", "group_norm", "norm_cross", "norm_q", "norm_k"):
            if getattr(attn, name, None) is not None:
                raise NotImplementedError(f"attn.{name} is None for every SD-1.5 attention (dalc:211-233); got a module")
        if attention_mask is not None:
            raise NotImplementedError("attention_mask is None at every reference call site (SURVEY 8c); use img_mask")
        if not hidden_states.is_cuda:
            raise RuntimeError("adaface_b200 AttnProcessor_LoRA_Capture runs on CUDA only (no CPU fallback)")

        if torch.is_grad_enabled() and (hidden_states.requires_grad or (
                encoder_hidden_states is not None and encoder_hidden_states.requires_grad) or self._has_trainable()):
            return self._call_train(attn, hidden_states, encoder_hidden_states, img_mask, subj_indices)

        residual = hidden_states
        in_dtype = hidden_states.dtype
        input_ndim = hidden_states.ndim
        if input_ndim == 4:                                                          # dalc:217-220
            bsz, channel, height, width = hidden_states.shape
            hidden_states = hidden_states.view(bsz, channel, height * width).transpose(1, 2)
        B, N, C_in = hidden_states.shape
        H = attn.heads
        x = hidden_states.to(torch.bfloat16).contiguous()
        x2d = x.view(B * N, C_in)
        pk = _bf16_pack(attn)
        lora = (lambda n: getattr(self, f"to_{n}_lora")) if self.enable_lora else (lambda n: None)
        is_cross = encoder_hidden_states is not None
        C = pk["wq"].shape[0]
        d = C // H
        sm_scale = 1.0 / math.sqrt(d)

        if not is_cross:
            # ---- self-attention: one fused QKV GEMM unless a LoRA adapter splits it
            key_mask = img_mask_to_key_mask(img_mask, N) if img_mask is not None else None
            if lora("q") is None and lora("k") is None and lora("v") is None and "wqkv" in pk:
                o = ops.self_attention_fused_qkv(x2d, pk["wqkv"], pk["bqkv"], B, N, H, sm_scale, key_mask)
            else:
                q = ops.proj(x2d, pk["wq"], bias=pk["bq"]).view(B, N, C)
                if lora("q") is not None and self.q_lora_updates_query:
                    q = _linear(x2d, pk["wq"], pk["bq"], lora("q")).view(B, N, C)
                k = _linear(x2d, pk["wk"], pk["bk"], lora("k")).view(B, N, C)
                v = _linear(x2d, pk["wv"], pk["bv"], lora("v")).view(B, N, C)
                o = ops.attention(q, k, v, H, sm_scale, key_mask=key_mask)
        else:
            ctx = encoder_hidden_states.to(torch.bfloat16).contiguous()
            S = ctx.shape[1]
            c2d = ctx.view(B * S, ctx.shape[2])
            # The slow-SDPA path (capture / normalize) keeps q, k, v in fp32 out of the projection GEMMs: the capture
            # kernel then evaluates q.k with a bf16 hi/lo split, which is what holds probabilities to 1e-3.
            hp = bool(self.capture_ca_activations or self.normalize_cross_attn)
            pdt = torch.float32 if hp else torch.bfloat16
            fast = (not hp) and all(lora(n) is None for n in ("q", "k", "v"))
            q = q2 = k = v = prob = score = prob_subj = consumed = None
            if fast:                                                                 # dalc:320-322: 3 launches in all
                o = ops.cross_attention_fused(x2d, pk["wq"], pk["bq"], c2d, pk["wkv"], pk["bkv"], B, N, S, H, sm_scale)
            else:
                q = q2 = ops.proj(x2d, pk["wq"], bias=pk["bq"], out_dtype=pdt).view(B, N, C)   # dalc:235
                if lora("q") is not None:                                            # dalc:239-249
                    q2 = _linear(x2d, pk["wq"], pk["bq"], lora("q"), out_dtype=pdt).view(B, N, C)
                    if self.q_lora_updates_query:
                        q = q2
                if lora("k") is None and lora("v") is None:
                    kv = ops.proj(c2d, pk["wkv"], bias=pk["bkv"], out_dtype=pdt).view(B, S, 2 * C)
                    k, v = kv[:, :, :C], kv[:, :, C:]
                else:
                    k = _linear(c2d, pk["wk"], pk["bk"], lora("k"), out_dtype=pdt).view(B, S, C)   # dalc:280-283
                    v = _linear(c2d, pk["wv"], pk["bv"], lora("v"), out_dtype=pdt).view(B, S, C)   # dalc:285-288
            if fast:
                pass
            elif hp:                                                                 # dalc:309-315
                mix = bool(self.mix_attn_mats_in_batch)
                if mix and B % 2:
                    raise ValueError("mix_attn_mats_in_batch needs an even batch ordered [sc.., mc..] (dalc:113)")
                col_flag, subj_cols = self._subj_aux(B, S, subj_indices, x.device)
                qm = ops.qmean(q) if col_flag is not None else None
                cap = bool(self.capture_ca_activations)
                if self.cross_attn_scale_factor.device != x.device:     # processor left on the CPU by the caller
                    self.cross_attn_scale_factor.data = self.cross_attn_scale_factor.data.to(x.device)
                if cap and self.capture_consumers is not None:
                    o, prob, consumed = self._consume(q, k, v, H, sm_scale, B, S, subj_indices, train=False)
                    score = prob_subj = None
                else:
                    o, prob, score, prob_subj = ops.attention_cross_capture(
                        q, k, v, H, sm_scale, want_prob=cap, want_score=cap, col_flag=col_flag, qmean=qm,
                        ca_scale=self.cross_attn_scale_factor.detach().float().reshape(1), mix=mix, subj_cols=subj_cols)
            else:                                                                    # dalc:320-322
                o = ops.attention(q, k, v, H, sm_scale)
                prob = score = prob_subj = None

        out = _linear(o.view(B * N, C), pk["wo"], pk["bo"], lora("out")).view(B, N, -1)   # dalc:328-334
        hidden_out = out.to(in_dtype)
        if input_ndim == 4:
            hidden_out = hidden_out.transpose(-1, -2).reshape(bsz, channel, height, width)
        if getattr(attn, "residual_connection", False):
            hidden_out = hidden_out + residual
        rescale = getattr(attn, "rescale_output_factor", 1.0)
        if rescale != 1.0:
            hidden_out = hidden_out / rescale

        if is_cross and self.capture_ca_activations:                                 # dalc:344-362
            f = math.sqrt(1.0 / math.sqrt(C))     # quirk 1: the scale is taken before the head split => C^-1/4
            ca = self.cached_activations
            ca["q"] = ops.chan_major(q, f)
            ca["q2"] = ca["q"] if q2 is q else ops.chan_major(q2, f)
            ca["k"] = ops.chan_major(k, f)
            ca["v"] = ops.chan_major(v, f)
            ca["attn"], ca["attnscore"] = prob, score
            ca["attn_out"] = ops.chan_major(out, 1.0) if (input_ndim == 3 and rescale == 1.0 and not getattr(
                attn, "residual_connection", False)) else hidden_out.float().permute(0, 2, 1).contiguous()
            if prob_subj is not None:
                ca["attn_subj"] = prob_subj
            if consumed:
                ca.update(consumed)
        return hidden_out

    forward = __call__

    # -------------------------------------------------------------------------------------------- training path
    def _has_trainable(self):
        if self.enable_lora and any(p.requires_grad for n in ("q", "k", "v", "out") if getattr(self, f"to_{n}_lora") is not None
                                    for p in getattr(self, f"to_{n}_lora").parameters()):
            return True
        return bool(self.normalize_cross_attn and self.cross_attn_scale_factor.requires_grad)

    def _subj_aux(self, B, S, subj_indices, device):
        """col_flag [B,S] uint8 for normalize (dalc:119-133) and subj_cols [B,n] int32 for the subject-columns capture."""
        col_flag = subj_cols = None
        if self.normalize_cross_attn and not self.mix_attn_mats_in_batch:
            if subj_indices is None:
                raise ValueError("normalize_cross_attn=True requires subj_indices (dalc:120)")
            ib, in_ = subj_indices
            col_flag = _token_flags(B, S, ib, in_, device)
        if self.capture_subj_cols_only and subj_indices is not None:
            ib, in_ = subj_indices
            slot, n_sub = self._subj_slots(ib.to(device), B)
            subj_cols = torch.full((B, n_sub), -1, device=device, dtype=torch.int32)
            subj_cols[ib.long(), slot] = in_.to(device=device, dtype=torch.int32)
        return col_flag, subj_cols

    def _subj_slots(self, ib, B):
        """Slot of every subject token inside its own instance's row of ``subj_cols`` = how many tokens of the same instance
        precede it, and n_sub = the largest per-instance count.  The reference's ``subj_indices`` routinely cover only part of the
        batch (the conditional half of a CFG batch, ddim.py:239-243; sliced batches), so the counts are NOT assumed equal.
        n_sub sizes an allocation, hence one host read per distinct index tensor (cached on its storage + version); inside a CUDA
        graph capture an uncached index tensor is an error rather than a silent wrong shape."""
        key = (ib.data_ptr(), ib._version, ib.numel(), B, ib.device)
        hit = getattr(self, "_slot_cache", None)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        if ib.numel() == 0:
            return ib.long(), 1
        if ib.is_cuda and torch.cuda.is_current_stream_capturing():
            raise RuntimeError("capture_subj_cols_only: call once outside CUDA-graph capture with these subj_indices first")
        onehot = ib.long()[:, None] == torch.arange(B, device=ib.device)[None, :]
        if not bool(onehot.any(dim=1).all()):
            raise IndexError(f"subj_indices: batch index outside [0, {B})")
        slot = (onehot.cumsum(dim=0) - 1)[torch.arange(ib.numel(), device=ib.device), ib.long()]
        n_sub = int(onehot.sum(dim=0).max().item())
        self._slot_cache = (key, slot, n_sub)
        return slot, n_sub

    def _call_train(self, attn, hidden_states, encoder_hidden_states, img_mask, subj_indices):
        """Same arithmetic as ``__call__`` with every kernel paired with its backward (autograd.py): gradients reach
        hidden_states, encoder_hidden_states (=> SubjBasisGenerator), the LoRA A / B / DoRA magnitudes and
        cross_attn_scale_factor (x10 GradientScaler, dalc:129), including through every cached activation.  The
        frozen base weights get none.  Layout differences from the inference path: q/k/v stay in plain [B, L, C]
        buffers (they are saved for the recompute-form backward), K and V are projected separately on the capture path."""
        from . import autograd as ag
        residual = hidden_states
        in_dtype = hidden_states.dtype
        input_ndim = hidden_states.ndim
        if input_ndim == 4:
            bsz, channel, height, width = hidden_states.shape
            hidden_states = hidden_states.view(bsz, channel, height * width).transpose(1, 2)
        B, N, C_in = hidden_states.shape
        H = attn.heads
        x2d = hidden_states.to(torch.bfloat16).contiguous().view(B * N, C_in)
        pk = _bf16_pack(attn)
        lora = (lambda n: getattr(self, f"to_{n}_lora")) if self.enable_lora else (lambda n: None)
        is_cross = encoder_hidden_states is not None
        C = pk["wq"].shape[0]
        sm_scale = 1.0 / math.sqrt(C // H)
        q = q2 = k = v = prob = score = prob_subj = consumed = None

        if not is_cross:
            key_mask = img_mask_to_key_mask(img_mask, N) if img_mask is not None else None
            if all(lora(n) is None for n in ("q", "k", "v")) and "wqkv" in pk:
                qkv = ag.linear(x2d, pk, "wqkv", "bqkv").view(B, N, 3 * C)
                o = ag.attention(qkv, qkv, qkv, (0, C, 2 * C), C, C, H, sm_scale, key_mask)
            else:
                q = ag.linear(x2d, pk, "wq", "bq", lora=lora("q") if self.q_lora_updates_query else None).view(B, N, C)
                k = ag.linear(x2d, pk, "wk", "bk", lora=lora("k")).view(B, N, C)
                v = ag.linear(x2d, pk, "wv", "bv", lora=lora("v")).view(B, N, C)
                o = ag.attention(q, k, v, (0, 0, 0), C, C, H, sm_scale, key_mask)
        else:
            ctx = encoder_hidden_states.to(torch.bfloat16).contiguous()
            S = ctx.shape[1]
            c2d = ctx.view(B * S, ctx.shape[2])
            hp = bool(self.capture_ca_activations or self.normalize_cross_attn)
            pdt = torch.float32 if hp else torch.bfloat16
            if lora("q") is not None:                                                # dalc:239-249
                q2 = ag.linear(x2d, pk, "wq", "bq", lora=lora("q"), out_dtype=pdt).view(B, N, C)
                q = q2 if self.q_lora_updates_query else ag.linear(x2d, pk, "wq", "bq", out_dtype=pdt).view(B, N, C)
            else:
                q = q2 = ag.linear(x2d, pk, "wq", "bq", out_dtype=pdt).view(B, N, C)
            if hp or lora("k") is not None or lora("v") is not None:
                k = ag.linear(c2d, pk, "wk", "bk", lora=lora("k"), out_dtype=pdt).view(B, S, C)   # dalc:280-283
                v = ag.linear(c2d, pk, "wv", "bv", lora=lora("v"), out_dtype=pdt).view(B, S, C)   # dalc:285-288
            if hp:                                                                   # dalc:309-315
                mix = bool(self.mix_attn_mats_in_batch)
                if mix and B % 2:
                    raise ValueError("mix_attn_mats_in_batch needs an even batch ordered [sc.., mc..] (dalc:113)")
                col_flag, subj_cols = self._subj_aux(B, S, subj_indices, x2d.device)
                if self.cross_attn_scale_factor.device != x2d.device:
                    self.cross_attn_scale_factor.data = self.cross_attn_scale_factor.data.to(x2d.device)
                cap = bool(self.capture_ca_activations)
                if cap and self.capture_consumers is not None:
                    o, prob, consumed = self._consume(q, k, v, H, sm_scale, B, S, subj_indices, train=True)
                else:
                    o, prob, score, prob_subj = ag.CrossCaptureFn.apply(q, k, v, self.cross_attn_scale_factor, H, sm_scale, cap,
                                                                        cap, col_flag, subj_cols, mix, 10.0)
            elif k is None:
                kv = ag.linear(c2d, pk, "wkv", "bkv").view(B, S, 2 * C)
                o = ag.attention(q, kv, kv, (0, 0, C), C, C, H, sm_scale)
            else:
                o = ag.attention(q, k, v, (0, 0, 0), C, C, H, sm_scale)

        out = ag.linear(o.view(B * N, C), pk, "wo", "bo", lora=lora("out")).view(B, N, -1)   # dalc:328-334
        hidden_out = out.to(in_dtype)
        if input_ndim == 4:
            hidden_out = hidden_out.transpose(-1, -2).reshape(bsz, channel, height, width)
        if getattr(attn, "residual_connection", False):
            hidden_out = hidden_out + residual
        rescale = getattr(attn, "rescale_output_factor", 1.0)
        if rescale != 1.0:
            hidden_out = hidden_out / rescale
        if is_cross and self.capture_ca_activations:                                 # dalc:344-362
            f = math.sqrt(1.0 / math.sqrt(C))
            ca = self.cached_activations
            ca["q"] = ag.ChanMajorFn.apply(q, f)
            ca["q2"] = ca["q"] if q2 is q else ag.ChanMajorFn.apply(q2, f)
            ca["k"] = ag.ChanMajorFn.apply(k, f)
            ca["v"] = ag.ChanMajorFn.apply(v, f)
            ca["attn"], ca["attnscore"] = prob, score
            ca["attn_out"] = ag.ChanMajorFn.apply(out, 1.0) if (input_ndim == 3 and rescale == 1.0 and not getattr(
                attn, "residual_connection", False)) else hidden_out.float().permute(0, 2, 1).contiguous()
            if prob_subj is not None:
                ca["attn_subj"] = prob_subj
            if consumed:
                ca.update(consumed)
        return hidden_out
