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


def attention(q, k, v, heads, scale, *, key_mask=None, causal_mult=0, out=None):
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
              int(causal_mult), float(scale), _stream())
    return out


def attention_headmajor(q, k, v, scale, *, d=None, out=None):
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
              Lk, d, drow_q, drow_kv, float(scale), _stream())
    return out


_HEADS_WS = {}


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
    if not 1 <= len(hs) <= 4 or len(layer_weights) != len(hs):
        raise ValueError("sbg_head: 1..4 hidden states with one weight each")
    M, C = hs[0].shape
    for h in hs:
        _need(h, "h", torch.float32)
        if tuple(h.shape) != (M, C) or h.stride(0) != hs[0].stride(0):
            raise ValueError("sbg_head: hidden states must share shape and row stride")
    out = torch.empty((M, C), device=hs[0].device, dtype=torch.float32)
    wl = (ctypes.c_float * len(hs))(*[float(x) for x in layer_weights])
    ptrs = [_ptr(h) for h in hs] + [ctypes.c_void_p(0)] * (4 - len(hs))
    _lib.call("adaface_sbg_head_fwd", *ptrs, wl, len(hs), hs[0].stride(0), _ptr(w), _ptr(b), _ptr(out), out.stride(0), M,
              C, float(eps), _stream())
    return out


def softmax_scale(d):
    return 1.0 / math.sqrt(d)


def self_attention_fused_qkv(x2d, wqkv, bqkv, B, N, heads, scale, key_mask=None):
    """Fused QKV projection + unmasked/masked self-attention; returns o [B, N, C] bf16.
    d = 40 without a mask: the projection scatters q/k/v into the zero-padded head-major workspace (128-byte rows per
    (batch, head)) that the tcgen05 attention kernel's TMA loads 3x faster than 80-byte head slices; every other
    case keeps the interleaved [B, N, 3C] buffer."""
    C = wqkv.shape[0] // 3
    d = C // heads
    if key_mask is None and d == 40:
        ws = proj_heads(x2d, wqkv, heads, d, N, bias=bqkv)
        return attention_headmajor(ws[0], ws[1], ws[2], scale, d=d)
    qkv = proj(x2d, wqkv, bias=bqkv).view(B, N, 3 * C)
    return attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], heads, scale, key_mask=key_mask)


def cross_attention_fused(x2d, wq, bq, ctx2d, wkv, bkv, B, N, S, heads, scale):
    """q and fused k|v projections + unmasked cross-attention (fast path, dalc:321); returns o [B, N, C] bf16."""
    C = wq.shape[0]
    d = C // heads
    kv = proj(ctx2d, wkv, bias=bkv).view(B, S, 2 * C)
    if d == 40:
        q = proj_heads(x2d, wq, heads, d, N, bias=bq)[0]                     # [B, H, N, 64]
        k = kv[:, :, :C].unflatten(2, (heads, d)).transpose(1, 2)             # views [B, H, S, d] of the interleaved buffer
        v = kv[:, :, C:].unflatten(2, (heads, d)).transpose(1, 2)
        return attention_headmajor(q, k, v, scale, d=d)
    q = proj(x2d, wq, bias=bq).view(B, N, C)
    return attention(q, kv[:, :, :C], kv[:, :, C:], heads, scale)
