"""Tensor-level wrappers around the C-ABI (include/adaface_b200.h).

PyTorch is only plumbing here: it owns device memory and the current stream; every arithmetic step of the
hot path is a kernel of libadaface_b200.so.  All functions raise (never fall back) on bad input.
"""
import ctypes
import math

import torch

from . import _lib

BF16, F32 = 0, 1
ACT_NONE, ACT_QUICK_GELU, ACT_GEGLU = 0, 1, 2


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _dt(t):
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    raise TypeError(f"adaface_b200: unsupported dtype {t.dtype} (bf16 / fp32 only)")


def _need(t, name, dtype=None, last_contig=True):
    if not t.is_cuda:
        raise RuntimeError(f"adaface_b200: `{name}` must be a CUDA tensor (no CPU fallback exists)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"adaface_b200: `{name}` must be {dtype}, got {t.dtype}")
    if last_contig and t.stride(-1) != 1:
        raise ValueError(f"adaface_b200: `{name}` must have unit stride in its last dimension")
    return t


def proj(x, w, *, t=None, bs=None, colscale=None, bias=None, residual=None, out=None, out_dtype=torch.bfloat16,
         act=ACT_NONE):
    """Y = act(colscale * (X W^T + T Bs^T) + bias) + residual  (adaface_proj_lora_fwd).

    x [M,K] bf16 (row stride free), w [N,K] bf16 contiguous, t [M,R] / bs [N,R] bf16, colscale / bias [N] fp32,
    residual [M,N'] bf16|fp32, out [M,N'] bf16|fp32 with N' = N (N/2 for GEGLU)."""
    _need(x, "x", torch.bfloat16)
    _need(w, "w", torch.bfloat16)
    if x.dim() != 2 or w.dim() != 2 or not w.is_contiguous() or x.shape[1] != w.shape[1]:
        raise ValueError(f"proj: bad shapes x{tuple(x.shape)} w{tuple(w.shape)}")
    M, K = x.shape
    N = w.shape[0]
    R, ldt = 0, 0
    if t is not None or bs is not None:
        _need(t, "t", torch.bfloat16)
        _need(bs, "bs", torch.bfloat16)
        if t.shape[0] != M or bs.shape[0] != N or t.shape[1] != bs.shape[1] or not bs.is_contiguous():
            raise ValueError(f"proj: bad LoRA shapes t{tuple(t.shape)} bs{tuple(bs.shape)}")
        R, ldt = t.shape[1], t.stride(0)
    for v, nm in ((colscale, "colscale"), (bias, "bias")):
        if v is not None:
            _need(v, nm, torch.float32)
            if v.numel() != N or not v.is_contiguous():
                raise ValueError(f"proj: `{nm}` must be a contiguous [N] tensor")
    n_out = N // 2 if act == ACT_GEGLU else N
    if out is None:
        out = torch.empty((M, n_out), device=x.device, dtype=out_dtype)
    _need(out, "out")
    if tuple(out.shape) != (M, n_out):
        raise ValueError(f"proj: out has shape {tuple(out.shape)}, expected {(M, n_out)}")
    ldr, rdt = 0, BF16
    if residual is not None:
        _need(residual, "residual")
        if tuple(residual.shape) != (M, n_out):
            raise ValueError("proj: residual shape mismatch")
        ldr, rdt = residual.stride(0), _dt(residual)
    _lib.call("adaface_proj_lora_fwd", _ptr(x), x.stride(0), _ptr(w), _ptr(t), ldt, _ptr(bs), _ptr(colscale),
              _ptr(bias), _ptr(residual), ldr, rdt, _ptr(out), out.stride(0), _dt(out), M, N, K, R, act, _stream())
    return out


def _view3(t, name, dtype=torch.bfloat16):
    _need(t, name, dtype)
    if t.dim() != 3:
        raise ValueError(f"attention: `{name}` must be [B, L, H*d]")
    return t


def attention(q, k, v, heads, scale, *, key_mask=None, causal_mult=0, out=None, lse=None):
    """softmax(scale * q k^T + masks) v for [B, L, H*d] views (adaface_attn_fwd).
    With causal_mult = M > 1 (CLIPAttentionMKV) k / v are [B, T, M*H*d] views: token t carries its M keys back to
    back, Lk = T*M, and key j is visible to query i iff j // M <= i."""
    _view3(q, "q"), _view3(k, "k"), _view3(v, "v")
    B, Lq, C = q.shape
    kv_mult = max(1, int(causal_mult))
    Lk = k.shape[1] * kv_mult
    if C % heads or k.shape != v.shape or k.shape[0] != B or k.shape[2] != C * kv_mult:
        raise ValueError(f"attention: inconsistent shapes q{tuple(q.shape)} k{tuple(k.shape)} v{tuple(v.shape)}")
    d = C // heads
    if out is None:
        out = torch.empty((B, Lq, C), device=q.device, dtype=torch.bfloat16)
    _view3(out, "out")
    if key_mask is not None:
        _need(key_mask, "key_mask", torch.uint8)
        if tuple(key_mask.shape) != (B, Lk) or not key_mask.is_contiguous():
            raise ValueError("attention: key_mask must be a contiguous uint8 [B, Lk] tensor")
    _lib.call("adaface_attn_fwd", _ptr(q), q.stride(0), q.stride(1), _ptr(k), k.stride(0), k.stride(1), _ptr(v),
              v.stride(0), v.stride(1), _ptr(out), out.stride(0), out.stride(1), B, heads, Lq, Lk, d, _ptr(key_mask),
              int(causal_mult), float(scale), _ptr(_lse_buf(lse, B, heads, Lq)), _stream())
    return out


def _lse_buf(lse, B, H, Lq):
    if lse is not None:
        _need(lse, "lse", torch.float32)
        if tuple(lse.shape) != (B, H, Lq) or not lse.is_contiguous():
            raise ValueError(f"attention: lse must be a contiguous fp32 [B, H, Lq] tensor, got {tuple(lse.shape)}")
    return lse


def attention_headmajor(q, k, v, scale, *, d=None, out=None, lse=None):
    """Unmasked attention for head-major views q [B, H, Lq, drow_q], k / v [B, H, Lk, drow_kv] (any batch / head /
    token strides, unit stride in the last dim).  `d` = true head dim when rows are zero-padded (drow > d; default:
    the k/v row width).  Output in the reference layout [B, Lq, H*d] (adaface_attn_headmajor_fwd)."""
    for t, nm in ((q, "q"), (k, "k"), (v, "v")):
        _need(t, nm, torch.bfloat16)
        if t.dim() != 4:
            raise ValueError(f"attention_headmajor: `{nm}` must be [B, H, L, d]")
    B, H, Lq, drow_q = q.shape
    Lk, drow_kv = k.shape[2], k.shape[3]
    d = drow_kv if d is None else d
    if v.shape != k.shape or k.shape[:2] != (B, H) or drow_q < d or drow_kv < d:
        raise ValueError(f"attention_headmajor: inconsistent shapes q{tuple(q.shape)} k{tuple(k.shape)} v{tuple(v.shape)} d={d}")
    if out is None:
        out = torch.empty((B, Lq, H * d), device=q.device, dtype=torch.bfloat16)
    _lib.call("adaface_attn_headmajor_fwd", _ptr(q), q.stride(0), q.stride(1), q.stride(2), _ptr(k), k.stride(0), k.stride(1),
              k.stride(2), _ptr(v), v.stride(0), v.stride(1), v.stride(2), _ptr(out), out.stride(0), out.stride(1), B, H, Lq,
              Lk, d, drow_q, drow_kv, float(scale), _ptr(_lse_buf(lse, B, H, Lq)), _stream())
    return out


_HEADS_WS = {}
# A/B switch.  The four-tile attention kernel fetches every K/V tile once per 512 queries, so the slow in-row-OOB TMA
# path of the interleaved layout no longer matters and the padded head-major detour (60 % more QKV store bytes) is
# off by default: measured 470 vs 459 TFLOP/s on the whole stack.
_SELF_HEADMAJOR = __import__("os").environ.get("ADAFACE_SELF_HEADMAJOR", "0") == "1"
_CROSS_HEADMAJOR = __import__("os").environ.get("ADAFACE_CROSS_HEADMAJOR", "0") == "1"    # measured in-graph, level A: 49.8 us per block head-major vs 40.0 us plain


def heads_workspace(n_which, B, H, L, dpad, device):
    """Zero-initialised [n_which, B, H, L, dpad] bf16 workspace, cached per shape: proj_heads never writes the pad
    columns, so they stay zero across calls."""
    key = (n_which, B, H, L, dpad, str(device))
    ws = _HEADS_WS.get(key)
    if ws is None:
        ws = torch.zeros((n_which, B, H, L, dpad), device=device, dtype=torch.bfloat16)
        _HEADS_WS[key] = ws
    return ws


def proj_heads(x, w, heads, d, rows_per_batch, *, dpad=None, t=None, bs=None, colscale=None, bias=None, out=None):
    """Projection whose bf16 output is scattered head-major: returns y[which, b, h, n, :dpad] with
    y[which, b, h, n, dd] = (x W^T ...)[b*rows_per_batch + n, which*heads*d + h*d + dd]  (adaface_proj_lora_heads_fwd)."""
    _need(x, "x", torch.bfloat16)
    _need(w, "w", torch.bfloat16)
    M, K = x.shape
    N = w.shape[0]
    if dpad is None:
        dpad = (d + 63) // 64 * 64
    n_which = N // (heads * d)
    B = M // rows_per_batch
    if out is None:
        out = heads_workspace(n_which, B, heads, rows_per_batch, dpad, x.device)
    if tuple(out.shape) != (n_which, B, heads, rows_per_batch, dpad) or not out.is_contiguous():
        raise ValueError("proj_heads: out must be a contiguous [N/(H*d), B, H, rows_per_batch, dpad] tensor")
    R, ldt = 0, 0
    if t is not None:
        R, ldt = t.shape[1], t.stride(0)
    _lib.call("adaface_proj_lora_heads_fwd", _ptr(x), x.stride(0), _ptr(w), _ptr(t), ldt, _ptr(bs), _ptr(colscale), _ptr(bias),
              _ptr(out), M, N, K, R, heads, d, dpad, rows_per_batch, _stream())
    return out


def attention_cross_capture(q, k, v, heads, scale, *, want_prob=True, want_score=True, col_flag=None, qmean=None,
                            ca_scale=None, mix=False, subj_cols=None, out=None):
    """The slow SDPA of dalc:79-139 as one kernel (adaface_attn_cross_capture_fwd).
    q / k / v are all bf16 or all fp32 views (fp32 = high-precision scores for capture).
    Returns (out [B,Lq,C] bf16, prob [B,H,Lq,S] fp32 | None, score | None, prob_subj [B,H,Lq,n_subj] | None)."""
    _view3(q, "q", q.dtype), _view3(k, "k", q.dtype), _view3(v, "v", q.dtype)
    in_dt = _dt(q)
    B, Lq, C = q.shape
    S = k.shape[1]
    d = C // heads
    dev = q.device
    if out is None:
        out = torch.empty((B, Lq, C), device=dev, dtype=torch.bfloat16)
    prob = torch.empty((B, heads, Lq, S), device=dev, dtype=torch.float32) if want_prob else None
    score = torch.empty((B, heads, Lq, S), device=dev, dtype=torch.float32) if want_score else None
    prob_subj, n_subj = None, 0
    if subj_cols is not None:
        _need(subj_cols, "subj_cols", torch.int32)
        if subj_cols.dim() != 2 or subj_cols.shape[0] != B or not subj_cols.is_contiguous():
            raise ValueError("attention_cross_capture: subj_cols must be a contiguous int32 [B, n_subj] tensor")
        n_subj = subj_cols.shape[1]
        prob_subj = torch.empty((B, heads, Lq, n_subj), device=dev, dtype=torch.float32)
    if col_flag is not None:
        _need(col_flag, "col_flag", torch.uint8)
        _need(qmean, "qmean", torch.float32)
        if tuple(col_flag.shape) != (B, S) or tuple(qmean.shape) != (B, C):
            raise ValueError("attention_cross_capture: col_flag must be [B,S] and qmean [B,C]")
        if ca_scale is not None:
            _need(ca_scale, "ca_scale", torch.float32)   # device scalar: cross_attn_scale_factor
    _lib.call("adaface_attn_cross_capture_fwd", _ptr(q), q.stride(0), q.stride(1), _ptr(k), k.stride(0), k.stride(1),
              _ptr(v), v.stride(0), v.stride(1), _ptr(out), out.stride(0), out.stride(1), B, heads, Lq, S, d,
              float(scale), _ptr(prob), _ptr(score), _ptr(prob_subj), _ptr(subj_cols), n_subj, _ptr(col_flag),
              _ptr(qmean), _ptr(ca_scale), int(bool(mix)), in_dt, _stream())
    return out, prob, score, prob_subj


def _consume_checks(q, k, heads, col_flag, qmean_, sum_flag, ref_prob):
    B, Lq, C = q.shape
    S = k.shape[1]
    if col_flag is not None:
        _need(col_flag, "col_flag", torch.uint8)
        _need(qmean_, "qmean", torch.float32)
        if tuple(col_flag.shape) != (B, S) or tuple(qmean_.shape) != (B, C):
            raise ValueError("attention_cross_consume: col_flag must be [B,S] and qmean [B,C]")
    if sum_flag is not None:
        _need(sum_flag, "sum_flag", torch.uint8)
        if tuple(sum_flag.shape) != (B, S) or not sum_flag.is_contiguous():
            raise ValueError("attention_cross_consume: sum_flag must be a contiguous uint8 [B,S] tensor")
    if ref_prob is not None:
        _need(ref_prob, "ref_prob", torch.float32)
        if tuple(ref_prob.shape) != (B, heads, Lq, S) or not ref_prob.is_contiguous():
            raise ValueError("attention_cross_consume: ref_prob must be a contiguous fp32 [B,H,Lq,S] tensor")


def attention_cross_consume(q, k, v, heads, scale, *, sum_flag=None, ref_prob=None, want_prob=False, col_flag=None, qmean=None,
                            ca_scale=None, out=None):
    """Capture with fused consumers (adaface_attn_cross_consume_fwd): the slow SDPA of dalc:79-139 whose
    probability map is reduced in registers instead of being written.  Returns (out [B,Lq,C] bf16, subj_sum [B,H,Lq] fp32 | None
    = mass on the columns flagged in sum_flag [B,S], sqdiff [B] fp32 | None = sum_{h,i,j} (prob - ref_prob)^2, prob | None)."""
    _view3(q, "q", q.dtype), _view3(k, "k", q.dtype), _view3(v, "v", q.dtype)
    _consume_checks(q, k, heads, col_flag, qmean, sum_flag, ref_prob)
    B, Lq, C = q.shape
    S = k.shape[1]
    dev = q.device
    if out is None:
        out = torch.empty((B, Lq, C), device=dev, dtype=torch.bfloat16)
    prob = torch.empty((B, heads, Lq, S), device=dev, dtype=torch.float32) if want_prob else None
    subj_sum = torch.empty((B, heads, Lq), device=dev, dtype=torch.float32) if sum_flag is not None else None
    slots = (Lq + 63) // 64
    sq_part = torch.zeros((B, heads * slots * 4), device=dev, dtype=torch.float32) if ref_prob is not None else None
    _lib.call("adaface_attn_cross_consume_fwd", _ptr(q), q.stride(0), q.stride(1), _ptr(k), k.stride(0), k.stride(1), _ptr(v),
              v.stride(0), v.stride(1), _ptr(out), out.stride(0), out.stride(1), B, heads, Lq, S, C // heads, float(scale), _ptr(prob),
              _ptr(col_flag), _ptr(qmean), _ptr(ca_scale), _dt(q), _ptr(sum_flag), _ptr(subj_sum), _ptr(ref_prob), _ptr(sq_part),
              slots, _stream())
    sqdiff = sq_part.sum(dim=1) if sq_part is not None else None          # fixed-order reduction of the per-CTA partials
    return out, subj_sum, sqdiff, prob


def attention_cross_consume_bwd(q, k, v, dout, heads, scale, *, sum_flag=None, g_subj=None, ref_prob=None, mse_coef=None, dprob=None,
                                col_flag=None, qmean=None, ca_scale=None, dca_mul=1.0, dkv_dtype=torch.float32):
    """Backward of attention_cross_consume (adaface_attn_cross_consume_bwd): the map's gradient is implicit --
    g_subj [B,H,Lq] on the flagged columns + mse_coef (device scalar) * (P - ref_prob) (+ an optional dense dprob).
    Returns (dq bf16 [B,Lq,C], dk, dv [B,S,C] of dkv_dtype, dca fp32 [1])."""
    _view3(q, "q", q.dtype), _view3(k, "k", q.dtype), _view3(v, "v", q.dtype), _view3(dout, "dout")
    _consume_checks(q, k, heads, col_flag, qmean, sum_flag, ref_prob)
    B, Lq, C = q.shape
    S = k.shape[1]
    d = C // heads
    dev = q.device
    if g_subj is not None:
        _need(g_subj, "g_subj", torch.float32)
        if tuple(g_subj.shape) != (B, heads, Lq) or not g_subj.is_contiguous() or sum_flag is None:
            raise ValueError("attention_cross_consume_bwd: g_subj must be a contiguous fp32 [B,H,Lq] tensor and needs sum_flag")
    if (ref_prob is None) != (mse_coef is None):
        raise ValueError("attention_cross_consume_bwd: ref_prob and mse_coef go together")
    if mse_coef is not None:
        _need(mse_coef, "mse_coef", torch.float32)
        if mse_coef.numel() != B or not mse_coef.is_contiguous():
            raise ValueError("attention_cross_consume_bwd: mse_coef must be a contiguous fp32 [B] tensor")
    if dprob is not None:
        _need(dprob, "dprob", torch.float32)
        if tuple(dprob.shape) != (B, heads, Lq, S) or not dprob.is_contiguous():
            raise ValueError("attention_cross_consume_bwd: dprob must be a contiguous fp32 [B,H,Lq,S] tensor")
    chunks = _lib.cross_capture_bwd_chunks(B, heads, Lq)
    dq = torch.empty((B, Lq, C), device=dev, dtype=torch.bfloat16)
    dk = torch.empty((B, S, C), device=dev, dtype=dkv_dtype)
    dv = torch.empty((B, S, C), device=dev, dtype=dkv_dtype)
    dca = torch.zeros(1, device=dev, dtype=torch.float32)
    part = torch.empty((2, B * heads * chunks * S * d), device=dev, dtype=torch.float32)
    dca_part = torch.empty(B * heads * chunks, device=dev, dtype=torch.float32)
    _lib.call("adaface_attn_cross_consume_bwd", _ptr(q), q.stride(0), q.stride(1), _ptr(k), k.stride(0), k.stride(1), _ptr(v),
              v.stride(0), v.stride(1), _ptr(dout), dout.stride(0), dout.stride(1), _ptr(dprob), B, heads, Lq, S, d, float(scale),
              _ptr(col_flag), _ptr(qmean), _ptr(ca_scale), _dt(q), _ptr(dq), dq.stride(0), dq.stride(1), _ptr(dk), dk.stride(0),
              dk.stride(1), _ptr(dv), dv.stride(0), dv.stride(1), _dt(dk), _ptr(dca), float(dca_mul), _ptr(part[0]), _ptr(part[1]),
              _ptr(dca_part), _ptr(sum_flag), _ptr(g_subj), _ptr(ref_prob), _ptr(mse_coef), _stream())
    return dq, dk, dv, dca


def qmean(q):
    """Mean over the queries: q [B, L, C] bf16|fp32 view -> [B, C] fp32 (adaface_qmean)."""
    _view3(q, "q", q.dtype)
    B, L, C = q.shape
    out = torch.empty((B, C), device=q.device, dtype=torch.float32)
    _lib.call("adaface_qmean", _ptr(q), _dt(q), q.stride(0), q.stride(1), B, L, C, _ptr(out), _stream())
    return out


def chan_major(src, factor):
    """dst[b, c, n] = factor * src[b, n, c] -> fp32 [B, C, L] (adaface_capture_chan_major, dalc:349-362)."""
    _need(src, "src")
    B, L, C = src.shape
    dst = torch.empty((B, C, L), device=src.device, dtype=torch.float32)
    _lib.call("adaface_capture_chan_major", _ptr(src), _dt(src), src.stride(0), src.stride(1), B, L, C, float(factor),
              _ptr(dst), _stream())
    return dst


def layernorm(x, w, b, eps=1e-5, out_dtype=torch.bfloat16):
    """LayerNorm over the last dim of a 2-D view; fp32 statistics (adaface_layernorm_fwd)."""
    _need(x, "x")
    _need(w, "w", torch.float32), _need(b, "b", torch.float32)
    M, C = x.shape
    y = torch.empty((M, C), device=x.device, dtype=out_dtype)
    _lib.call("adaface_layernorm_fwd", _ptr(x), _dt(x), x.stride(0), _ptr(w), _ptr(b), _ptr(y), _dt(y), y.stride(0), M, C,
              float(eps), _stream())
    return y


def sbg_head(hs, layer_weights, w, b, eps=1e-5):
    """LayerNorm(sum_l wl[l] * h_l) for up to 4 fp32 [M, C] hidden states (adaface_sbg_head_fwd)."""
    if not 1 <= len(hs) <= 4 or (layer_weights.numel() if torch.is_tensor(layer_weights) else len(layer_weights)) != len(hs):
        raise ValueError("sbg_head: 1..4 hidden states with one weight each")
    M, C = hs[0].shape
    for h in hs:
        _need(h, "h", torch.float32)
        if tuple(h.shape) != (M, C) or h.stride(0) != hs[0].stride(0):
            raise ValueError("sbg_head: hidden states must share shape and row stride")
    out = torch.empty((M, C), device=hs[0].device, dtype=torch.float32)
    ptrs = [_ptr(h) for h in hs] + [ctypes.c_void_p(0)] * (4 - len(hs))
    if torch.is_tensor(layer_weights):        # device weights (fp32 [n], already normalised): no host read, graph-capturable
        _need(layer_weights, "layer_weights", torch.float32)
        _lib.call("adaface_sbg_head_fwd_dev", *ptrs, _ptr(layer_weights), len(hs), hs[0].stride(0), _ptr(w), _ptr(b), _ptr(out),
                  out.stride(0), M, C, float(eps), _stream())
        return out
    wl = (ctypes.c_float * len(hs))(*[float(x) for x in layer_weights])
    _lib.call("adaface_sbg_head_fwd", *ptrs, wl, len(hs), hs[0].stride(0), _ptr(w), _ptr(b), _ptr(out), out.stride(0), M,
              C, float(eps), _stream())
    return out


def groupnorm_tokens(x, gamma, beta, groups=32, eps=1e-6):
    """GroupNorm over [B, C, h, w] fused with 'b c h w -> b (h w) c': returns bf16 [B, h*w, C] (adaface_groupnorm_tokens_fwd)."""
    _need(x, "x")
    if x.dim() != 4 or not x.is_contiguous():
        raise ValueError("groupnorm_tokens: x must be a contiguous [B, C, h, w] tensor")
    B, C, h, w = x.shape
    _need(gamma, "gamma", torch.float32), _need(beta, "beta", torch.float32)
    ws = torch.empty((2, B, C), device=x.device, dtype=torch.float32)
    y = torch.empty((B, h * w, C), device=x.device, dtype=torch.bfloat16)
    _lib.call("adaface_groupnorm_tokens_fwd", _ptr(x), _dt(x), _ptr(gamma), _ptr(beta), B, C, h * w, int(groups), float(eps),
              _ptr(ws[0]), _ptr(ws[1]), _ptr(y), _stream())
    return y


def tokens_to_nchw_add(t, x_in):
    """out[b, c, h, w] = t[b, (h w), c] + x_in[b, c, h, w] (adaface_tokens_to_nchw_add); t bf16, out like x_in."""
    _need(t, "t", torch.bfloat16), _need(x_in, "x_in")
    B, C, h, w = x_in.shape
    if tuple(t.shape) != (B, h * w, C) or not t.is_contiguous() or not x_in.is_contiguous():
        raise ValueError("tokens_to_nchw_add: t must be a contiguous [B, h*w, C] tensor matching x_in [B, C, h, w]")
    out = torch.empty_like(x_in)
    _lib.call("adaface_tokens_to_nchw_add", _ptr(t), _ptr(x_in), _dt(x_in), _ptr(out), B, C, h * w, _stream())
    return out


def pack_conv3x3_weight(weight):
    """[Cout, Cin, 3, 3] (nn.Conv2d) -> bf16 [Cout, 9 * Kc], Kc = Cin rounded up to 64, K index = (ky*3 + kx) * Kc + ci:
    the K-major operand adaface_conv3x3_fwd reads (one 64-channel slab per tap and chunk).  Load-time re-layout."""
    Cout, Cin, kh, kw = weight.shape
    if (kh, kw) != (3, 3):
        raise ValueError(f"pack_conv3x3_weight: expected a 3x3 kernel, got {kh}x{kw}")
    kc = (Cin + 63) // 64 * 64
    w = torch.zeros((Cout, 9, kc), device=weight.device, dtype=torch.bfloat16)
    w[:, :, :Cin] = weight.detach().permute(0, 2, 3, 1).reshape(Cout, 9, Cin).to(torch.bfloat16)
    return w.reshape(Cout, 9 * kc).contiguous()


def conv3x3(x, w_packed, hw, *, stride=1, bias=None, rowbias=None, residual=None, t=None, bs=None, colscale=None, out=None,
            out_dtype=torch.bfloat16):
    """3x3 convolution, padding 1, over NHWC tokens (adaface_conv3x3_fwd): x bf16 [B, h*w, Cin] contiguous with hw = (h, w),
    w_packed from pack_conv3x3_weight, bias fp32 [Cout], rowbias fp32 [B, Cout] (added per image), residual / out
    [B, ho*wo, Cout]; t [B*ho*wo, R] / bs [Cout, R] / colscale [Cout] = conv-LoRA / DoRA tail.  Returns [B, ho*wo, Cout]."""
    _need(x, "x", torch.bfloat16), _need(w_packed, "w_packed", torch.bfloat16)
    h, w = hw
    if x.dim() != 3 or not x.is_contiguous() or x.shape[1] != h * w or not w_packed.is_contiguous():
        raise ValueError(f"conv3x3: x must be a contiguous [B, h*w, Cin] tensor (got {tuple(x.shape)}, hw={hw})")
    B, _, Cin = x.shape
    Cout = w_packed.shape[0]
    if w_packed.shape[1] != 9 * ((Cin + 63) // 64 * 64):
        raise ValueError(f"conv3x3: w_packed{tuple(w_packed.shape)} does not match Cin={Cin}")
    if stride not in (1, 2):
        raise ValueError("conv3x3: stride must be 1 or 2")
    ho, wo = h // stride, w // stride
    M = B * ho * wo
    for v, nm, shape in ((bias, "bias", (Cout,)), (colscale, "colscale", (Cout,)), (rowbias, "rowbias", (B, Cout))):
        if v is not None:
            _need(v, nm, torch.float32)
            if tuple(v.shape) != shape or not v.is_contiguous():
                raise ValueError(f"conv3x3: `{nm}` must be a contiguous fp32 {shape} tensor")
    R, ldt = 0, 0
    if t is not None or bs is not None:
        _need(t, "t", torch.bfloat16), _need(bs, "bs", torch.bfloat16)
        if t.shape[0] != M or bs.shape[0] != Cout or t.shape[1] != bs.shape[1] or not bs.is_contiguous():
            raise ValueError(f"conv3x3: bad LoRA shapes t{tuple(t.shape)} bs{tuple(bs.shape)}")
        R, ldt = t.shape[1], t.stride(0)
    if out is None:
        out = torch.empty((B, ho * wo, Cout), device=x.device, dtype=out_dtype)
    _need(out, "out")
    if tuple(out.shape) != (B, ho * wo, Cout) or not out.is_contiguous():
        raise ValueError(f"conv3x3: out has shape {tuple(out.shape)}, expected contiguous {(B, ho * wo, Cout)}")
    ldr, rdt = 0, BF16
    if residual is not None:
        _need(residual, "residual")
        if tuple(residual.shape) != (B, ho * wo, Cout) or not residual.is_contiguous():
            raise ValueError("conv3x3: residual shape mismatch")
        ldr, rdt = Cout, _dt(residual)
    _lib.call("adaface_conv3x3_fwd", _ptr(x), B, h, w, Cin, _ptr(w_packed), _ptr(t), ldt, _ptr(bs), R, _ptr(colscale), _ptr(bias),
              _ptr(rowbias), _ptr(residual), ldr, rdt, _ptr(out), Cout, _dt(out), Cout, int(stride), ACT_NONE, _stream())
    return out


def groupnorm_act_tokens(x, gamma, beta, groups=32, eps=1e-5, silu=True):
    """act(GroupNorm(x)) over tokens: x bf16 [B, HW, C] -> bf16 [B, HW, C] (adaface_groupnorm_act_tokens_fwd)."""
    _need(x, "x", torch.bfloat16), _need(gamma, "gamma", torch.float32), _need(beta, "beta", torch.float32)
    if x.dim() != 3 or not x.is_contiguous():
        raise ValueError("groupnorm_act_tokens: x must be a contiguous [B, HW, C] tensor")
    B, HW, C = x.shape
    n_ws = int(_lib.load().adaface_groupnorm_act_tokens_ws_floats(B, HW, C, int(groups)))
    ws = torch.empty(2 * B * C + n_ws, device=x.device, dtype=torch.float32)
    y = torch.empty_like(x)
    _lib.call("adaface_groupnorm_act_tokens_fwd", _ptr(x), _ptr(gamma), _ptr(beta), B, HW, C, int(groups), float(eps), 1 if silu else 0,
              _ptr(ws[2 * B * C:]) if n_ws else ctypes.c_void_p(0), _ptr(ws), _ptr(ws[B * C:]), _ptr(y), _stream())
    return y


def groupnorm_act_tokens_bwd(x, dy, gamma, beta, groups=32, eps=1e-5, silu=True):
    """dX of groupnorm_act_tokens (frozen gamma / beta): x, dy bf16 [B, HW, C] -> bf16 (adaface_groupnorm_act_tokens_bwd)."""
    _need(x, "x", torch.bfloat16), _need(dy, "dy", torch.bfloat16), _need(gamma, "gamma", torch.float32), _need(beta, "beta", torch.float32)
    if x.dim() != 3 or not x.is_contiguous() or dy.shape != x.shape or not dy.is_contiguous():
        raise ValueError("groupnorm_act_tokens_bwd: x and dy must be contiguous [B, HW, C] tensors of one shape")
    B, HW, C = x.shape
    ws = torch.empty((4, B, C), device=x.device, dtype=torch.float32)
    dx = torch.empty_like(x)
    _lib.call("adaface_groupnorm_act_tokens_bwd", _ptr(x), _ptr(dy), _ptr(gamma), _ptr(beta), B, HW, C, int(groups), float(eps),
              1 if silu else 0, _ptr(ws), _ptr(dx), _stream())
    return dx


def resample2x_bwd(x, hw_low, mode):
    """mode 0: 2x2 sum-pool [B, 4*h*w, C] -> [B, h*w, C] (backward of upsample2x_tokens); mode 1: zero-insert [B, h*w, C] ->
    [B, 4*h*w, C] (first step of the backward of a stride-2 conv3x3).  hw_low = (h, w) of the LOW resolution."""
    _need(x, "x", torch.bfloat16)
    h, w = hw_low
    n_in = 4 * h * w if mode == 0 else h * w
    if x.dim() != 3 or not x.is_contiguous() or x.shape[1] != n_in:
        raise ValueError(f"resample2x_bwd: x must be a contiguous [B, {n_in}, C] tensor")
    B, _, C = x.shape
    y = torch.empty((B, h * w if mode == 0 else 4 * h * w, C), device=x.device, dtype=torch.bfloat16)
    _lib.call("adaface_resample2x_bwd", _ptr(x), _ptr(y), B, h, w, C, int(mode), _stream())
    return y


def dora_colscale(W, A16, B16, scaling, m):
    """colscale = m / ||W + s B A||_row (adaface_dora_colscale) with the product B A on the projection GEMM: W [N, K] fp32 | bf16
    contiguous, A16 bf16 [r, K], B16 bf16 [N, r] (UNscaled), m fp32 [N].  Returns fp32 [N].  No library arithmetic."""
    _need(W, "W")
    N, K = W.shape
    if not W.is_contiguous():
        raise ValueError("dora_colscale: W must be contiguous")
    rp = (A16.shape[0] + 7) // 8 * 8
    At = transpose(A16, pad_to=1)                                  # [K, r]
    if rp != A16.shape[0]:                                         # reduction length must be a multiple of 8 for the GEMM
        At = torch.nn.functional.pad(At, (0, rp - A16.shape[0]))
        B16 = torch.nn.functional.pad(B16, (0, rp - B16.shape[1]))
    BA = proj(B16.contiguous(), At.contiguous(), out_dtype=torch.float32)      # [N, K] = B A
    out = torch.empty(N, device=W.device, dtype=torch.float32)
    mf = m.detach().float().contiguous()
    _lib.call("adaface_dora_colscale", _ptr(W), _dt(W), _ptr(BA), BA.stride(0), float(scaling), _ptr(mf), _ptr(out), N, K, _stream())
    return out


def im2col3x3_tokens(x, hw):
    """bf16 NHWC tokens [B, h*w, C] -> bf16 [B*h*w, 9*Kc] in the K order of pack_conv3x3_weight (adaface_im2col3x3_tokens): the
    X operand of a 3x3 adapter's weight gradient dA = dT^T im2col(X)."""
    _need(x, "x", torch.bfloat16)
    h, w = hw
    if x.dim() != 3 or not x.is_contiguous() or x.shape[1] != h * w or x.shape[2] % 8:
        raise ValueError("im2col3x3_tokens: x must be a contiguous [B, h*w, C] tensor with C a multiple of 8")
    B, _, C = x.shape
    kc = (C + 63) // 64 * 64
    col = torch.empty((B * h * w, 9 * kc), device=x.device, dtype=torch.bfloat16)
    _lib.call("adaface_im2col3x3_tokens", _ptr(x), _ptr(col), B, h, w, C, _stream())
    return col


def unpack_conv3x3_weight(w_packed, cin):
    """Inverse of pack_conv3x3_weight for a gradient: [Cout, 9*Kc] -> [Cout, cin, 3, 3] (same dtype)."""
    cout = w_packed.shape[0]
    kc = w_packed.shape[1] // 9
    return w_packed.view(cout, 3, 3, kc)[..., :cin].permute(0, 3, 1, 2).contiguous()


def pack_conv3x3_weight_dx(weight):
    """Packed operand of the convolution's input gradient: dX = conv3x3(dY, W') with W'[ci, co, ky, kx] = W[co, ci, 2-ky, 2-kx]
    (flipped taps, transposed channels) -- the same implicit-GEMM kernel computes it."""
    return pack_conv3x3_weight(weight.detach().flip(2, 3).transpose(0, 1).contiguous())


def silu(x):
    """SiLU(x) -> bf16 (adaface_silu_fwd); x bf16 | fp32 contiguous."""
    _need(x, "x")
    if not x.is_contiguous():
        raise ValueError("silu: x must be contiguous")
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    _lib.call("adaface_silu_fwd", _ptr(x), _dt(x), _ptr(y), x.numel(), _stream())
    return y


def timestep_embedding(timesteps, dim, max_period=10000):
    """Sinusoidal embedding of a 1-D batch of timesteps -> bf16 [B, dim] (adaface_timestep_embedding)."""
    if not timesteps.is_cuda or timesteps.dim() != 1:
        raise RuntimeError("timestep_embedding: `timesteps` must be a 1-D CUDA tensor (no CPU fallback exists)")
    t = timesteps.float().contiguous()
    out = torch.empty((t.shape[0], dim), device=t.device, dtype=torch.bfloat16)
    _lib.call("adaface_timestep_embedding", _ptr(t), t.shape[0], int(dim), float(max_period), _ptr(out), _stream())
    return out


def upsample2x_tokens(x, hw):
    """Nearest 2x of NHWC tokens: bf16 [B, h*w, C] -> [B, 4*h*w, C] (adaface_upsample2x_tokens)."""
    _need(x, "x", torch.bfloat16)
    h, w = hw
    if x.dim() != 3 or not x.is_contiguous() or x.shape[1] != h * w:
        raise ValueError("upsample2x_tokens: x must be a contiguous [B, h*w, C] tensor")
    B, _, C = x.shape
    y = torch.empty((B, 4 * h * w, C), device=x.device, dtype=torch.bfloat16)
    _lib.call("adaface_upsample2x_tokens", _ptr(x), _ptr(y), B, h, w, C, _stream())
    return y


def softmax_scale(d):
    return 1.0 / math.sqrt(d)


def self_attention_fused_qkv(x2d, wqkv, bqkv, B, N, heads, scale, key_mask=None):
    """Fused QKV projection + unmasked/masked self-attention; returns o [B, N, C] bf16 from the interleaved [B, N, 3C]
    buffer.  (ADAFACE_SELF_HEADMAJOR=1 restores the d = 40 detour through the zero-padded head-major workspace, which
    the single-tile kernels' TMA loads 3x faster than 80-byte head slices.)"""
    C = wqkv.shape[0] // 3
    d = C // heads
    if key_mask is None and d == 40 and _SELF_HEADMAJOR:
        ws = proj_heads(x2d, wqkv, heads, d, N, bias=bqkv)
        return attention_headmajor(ws[0], ws[1], ws[2], scale, d=d)
    qkv = proj(x2d, wqkv, bias=bqkv).view(B, N, 3 * C)
    return attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], heads, scale, key_mask=key_mask)


def cross_attention_fused(x2d, wq, bq, ctx2d, wkv, bkv, B, N, S, heads, scale):
    """q and fused k|v projections + unmasked cross-attention (fast path, dalc:321); returns o [B, N, C] bf16."""
    C = wq.shape[0]
    d = C // heads
    kv = proj(ctx2d, wkv, bias=bkv).view(B, S, 2 * C)
    if d == 40 and _CROSS_HEADMAJOR:
        q = proj_heads(x2d, wq, heads, d, N, bias=bq)[0]                     # [B, H, N, 64]
        k = kv[:, :, :C].unflatten(2, (heads, d)).transpose(1, 2)             # views [B, H, S, d] of the interleaved buffer
        v = kv[:, :, C:].unflatten(2, (heads, d)).transpose(1, 2)
        return attention_headmajor(q, k, v, scale, d=d)
    q = proj(x2d, wq, bias=bq).view(B, N, C)
    return attention(q, kv[:, :, :C], kv[:, :, C:], heads, scale)


# ================================================================================================ backward (K5)
def attention_bwd(q, k, v, o, dout, lse, heads, scale, dq, dk, dv, *, key_mask=None, causal_mult=0):
    """Flash-attention backward (adaface_attn_bwd): fills the bf16 views dq / dk / dv (shaped like q / k / v)."""
    for t, nm in ((q, "q"), (k, "k"), (v, "v"), (o, "o"), (dout, "dout"), (dq, "dq"), (dk, "dk"), (dv, "dv")):
        _view3(t, nm)
    B, Lq, C = q.shape
    kv_mult = max(1, int(causal_mult))
    Lk = k.shape[1] * kv_mult
    d = C // heads
    _lse_buf(lse, B, heads, Lq)
    delta = torch.empty_like(lse)
    _lib.call("adaface_attn_bwd", _ptr(q), q.stride(0), q.stride(1), _ptr(k), k.stride(0), k.stride(1), _ptr(v), v.stride(0),
              v.stride(1), _ptr(o), o.stride(0), o.stride(1), _ptr(dout), dout.stride(0), dout.stride(1), _ptr(lse),
              _ptr(delta), _ptr(dq), dq.stride(0), dq.stride(1), _ptr(dk), dk.stride(0), dk.stride(1), _ptr(dv), dv.stride(0),
              dv.stride(1), B, heads, Lq, Lk, d, _ptr(key_mask), int(causal_mult), float(scale), _stream())


def attention_cross_capture_bwd(q, k, v, dout, heads, scale, *, dprob=None, dscore=None, col_flag=None, qmean=None,
                                ca_scale=None, mix=False, dca_mul=1.0, dkv_dtype=torch.float32):
    """Backward of attention_cross_capture (adaface_attn_cross_capture_bwd).
    Returns (dq bf16 [B,Lq,C], dk, dv [B,S,C] of dkv_dtype, dca fp32 [1])."""
    _view3(q, "q", q.dtype), _view3(k, "k", q.dtype), _view3(v, "v", q.dtype), _view3(dout, "dout")
    B, Lq, C = q.shape
    S = k.shape[1]
    d = C // heads
    dev = q.device
    for g, nm in ((dprob, "dprob"), (dscore, "dscore")):
        if g is not None:
            _need(g, nm, torch.float32)
            if tuple(g.shape) != (B, heads, Lq, S) or not g.is_contiguous():
                raise ValueError(f"attention_cross_capture_bwd: `{nm}` must be a contiguous fp32 [B,H,Lq,S] tensor")
    chunks = _lib.cross_capture_bwd_chunks(B, heads, Lq)
    dq = torch.empty((B, Lq, C), device=dev, dtype=torch.bfloat16)
    dk = torch.empty((B, S, C), device=dev, dtype=dkv_dtype)
    dv = torch.empty((B, S, C), device=dev, dtype=dkv_dtype)
    dca = torch.zeros(1, device=dev, dtype=torch.float32)
    part = torch.empty((2, B * heads * chunks * S * d), device=dev, dtype=torch.float32)
    dca_part = torch.empty(B * heads * chunks, device=dev, dtype=torch.float32)
    _lib.call("adaface_attn_cross_capture_bwd", _ptr(q), q.stride(0), q.stride(1), _ptr(k), k.stride(0), k.stride(1), _ptr(v),
              v.stride(0), v.stride(1), _ptr(dout), dout.stride(0), dout.stride(1), _ptr(dprob), _ptr(dscore), B, heads, Lq, S,
              d, float(scale), _ptr(col_flag), _ptr(qmean), _ptr(ca_scale), int(bool(mix)), _dt(q), _ptr(dq), dq.stride(0), dq.stride(1),
              _ptr(dk), dk.stride(0), dk.stride(1), _ptr(dv), dv.stride(0), dv.stride(1), _dt(dk), _ptr(dca), float(dca_mul),
              _ptr(part[0]), _ptr(part[1]), _ptr(dca_part), _stream())
    return dq, dk, dv, dca


def transpose(src, *, out_dtype=None, alpha=1.0, colscale=None, rowscale=None, pad_to=1):
    """dst[b, j, i] = alpha * colscale[j] * rowscale[i] * src[b, i, j] (adaface_transpose).  src [I, J] or [B, I, J]
    with unit stride in J.  `pad_to`: the destination row pitch is rounded up to this multiple and the padding is zero
    (the GEMM needs reduction lengths / row pitches that are multiples of 8); the returned tensor includes the padding."""
    _need(src, "src")
    squeeze = src.dim() == 2
    s3 = src.unsqueeze(0) if squeeze else src
    B, I, J = s3.shape
    out_dtype = out_dtype or src.dtype
    Ip = (I + pad_to - 1) // pad_to * pad_to
    dst = (torch.zeros if Ip != I else torch.empty)((B, J, Ip), device=src.device, dtype=out_dtype)
    _lib.call("adaface_transpose", _ptr(s3), _dt(s3), s3.stride(0), s3.stride(1), _ptr(dst), _dt(dst), dst.stride(0),
              dst.stride(1), B, I, J, float(alpha), _ptr(colscale), _ptr(rowscale), _stream())
    return dst[0] if squeeze else dst


def colsum(a, *, b=None, bias=None, colmul=None, out=None):
    """out[j] += colmul[j] * sum_i a[i, j] * (b[i, j] - bias[j]) (adaface_colsum); returns fp32 [N]."""
    _need(a, "a")
    M, N = a.shape
    if out is None:
        out = torch.zeros(N, device=a.device, dtype=torch.float32)
    ldb, bdt = 0, BF16
    if b is not None:
        _need(b, "b")
        ldb, bdt = b.stride(0), _dt(b)
    _lib.call("adaface_colsum", _ptr(a), _dt(a), a.stride(0), _ptr(b), bdt, ldb, _ptr(bias), _ptr(colmul), _ptr(out), M, N,
              _stream())
    return out


def layernorm_bwd(x, dy, w, eps=1e-5, *, want_wgrad=False):
    """LayerNorm backward (adaface_layernorm_bwd): returns (dx like x, dw, db) -- dw / db None unless want_wgrad."""
    _need(x, "x"), _need(dy, "dy"), _need(w, "w", torch.float32)
    M, C = x.shape
    dx = torch.empty((M, C), device=x.device, dtype=x.dtype)
    dw = db = None
    if want_wgrad:
        dw = torch.zeros(C, device=x.device, dtype=torch.float32)
        db = torch.zeros(C, device=x.device, dtype=torch.float32)
    _lib.call("adaface_layernorm_bwd", _ptr(x), _dt(x), x.stride(0), _ptr(dy), _dt(dy), dy.stride(0), _ptr(w), _ptr(dx),
              dx.stride(0), _ptr(dw), _ptr(db), M, C, float(eps), _stream())
    return dx, dw, db


def act_fwd(u, act):
    """h = act(u): quick-GELU [M,N] -> [M,N]; GEGLU packed [M,2N] -> [M,N] (adaface_act_fwd)."""
    _need(u, "u", torch.bfloat16)
    M, W = u.shape
    n_out = W // 2 if act == ACT_GEGLU else W
    h = torch.empty((M, n_out), device=u.device, dtype=torch.bfloat16)
    _lib.call("adaface_act_fwd", _ptr(u), u.stride(0), _ptr(h), h.stride(0), M, n_out, int(act), _stream())
    return h


def act_bwd(u, dh, act):
    """du = dh * act'(u) (adaface_act_bwd)."""
    _need(u, "u", torch.bfloat16), _need(dh, "dh", torch.bfloat16)
    M, n_out = dh.shape
    du = torch.empty_like(u)
    _lib.call("adaface_act_bwd", _ptr(u), u.stride(0), _ptr(dh), dh.stride(0), _ptr(du), du.stride(0), M, n_out, int(act),
              _stream())
    return du


def sbg_head_bwd(hs, layer_weights, w, dout, eps=1e-5):
    """Backward of sbg_head: returns ([dh_l], dwl fp32 [n], dw, db) (adaface_sbg_head_bwd)."""
    n = len(hs)
    M, C = hs[0].shape
    dev = hs[0].device
    _need(dout, "dout", torch.float32)
    dhs = [torch.empty((M, C), device=dev, dtype=torch.float32) for _ in hs]
    for h in hs:
        if h.stride(0) != C or h.stride(1) != 1:
            raise ValueError("sbg_head_bwd: hidden states must be contiguous")
    dwl = torch.zeros(4, device=dev, dtype=torch.float32)
    dw = torch.zeros(C, device=dev, dtype=torch.float32)
    db = torch.zeros(C, device=dev, dtype=torch.float32)
    pad = [ctypes.c_void_p(0)] * (4 - n)
    if torch.is_tensor(layer_weights):
        _need(layer_weights, "layer_weights", torch.float32)
        _lib.call("adaface_sbg_head_bwd_dev", *([_ptr(h) for h in hs] + pad), _ptr(layer_weights), n, C, _ptr(w), _ptr(dout), dout.stride(0),
                  *([_ptr(d) for d in dhs] + pad), _ptr(dwl), _ptr(dw), _ptr(db), M, C, float(eps), _stream())
        return dhs, dwl[:n], dw, db
    wl = (ctypes.c_float * n)(*[float(x) for x in layer_weights])
    _lib.call("adaface_sbg_head_bwd", *([_ptr(h) for h in hs] + pad), wl, n, C, _ptr(w), _ptr(dout), dout.stride(0),
              *([_ptr(d) for d in dhs] + pad), _ptr(dwl), _ptr(dw), _ptr(db), M, C, float(eps), _stream())
    return dhs, dwl[:n], dw, db
