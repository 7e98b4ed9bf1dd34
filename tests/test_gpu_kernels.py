"""GPU: each kernel of libadaface_b200.so, called through the C-ABI, against a plain fp32 restatement of the
same op on seeded inputs (edge cases: ragged tiles, masks, strided views, LoRA tails, every epilogue)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


def ops():
    import adaface_dev_b200 as a
    return a.ops


def a_lib():
    import adaface_dev_b200 as a
    return a._lib


def rnd(*shape, std=1.0, seed=0, dtype=BF):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * std).to(dtype).cuda()


def maxerr(a, b, rtol=0.0):
    """max over elements of |a - b| - rtol * |b|.  rtol = 2^-8 absorbs the final bf16 rounding of large outputs
    (bf16 half-ulp = 2^-9 relative), so `atol` bounds the error that is NOT just output quantisation."""
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs() - rtol * b.abs()).max().item()


R8 = 2.0 ** -8


# ------------------------------------------------------------------------------------------- K1 GEMM
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 320, 320), (4096, 320, 320), (154, 640, 768), (1280, 2304, 768),
                                   (1280, 768, 3072), (77, 24, 320), (1000, 960, 320), (64, 1280, 1280), (130, 200, 72)])
def test_proj_plain(M, N, K):
    x, w = rnd(M, K, seed=1), rnd(N, K, std=1 / math.sqrt(K), seed=2)
    y = ops().proj(x, w)
    ref = x.float() @ w.float().T
    assert y.shape == (M, N) and y.dtype == BF
    assert maxerr(y, ref, R8) < 1e-2


@pytest.mark.parametrize("R", [8, 16, 192])
@pytest.mark.parametrize("M,N,K", [(512, 320, 320), (154, 320, 768), (333, 640, 640)])
def test_proj_lora_dora_bias(M, N, K, R):
    """Y = colscale * (X W^T + T Bs^T) + bias with T = X A^T  (SURVEY 8a A4)."""
    x, w = rnd(M, K, seed=1), rnd(N, K, std=1 / math.sqrt(K), seed=2)
    A, Bs = rnd(R, K, std=1 / math.sqrt(K), seed=3), rnd(N, R, std=0.05, seed=4)
    cs, bias = rnd(N, seed=5, dtype=torch.float32).abs() + 0.5, rnd(N, seed=6, dtype=torch.float32)
    t = ops().proj(x, A)
    assert maxerr(t, x.float() @ A.float().T, R8) < 1e-2
    y = ops().proj(x, w, t=t, bs=Bs, colscale=cs, bias=bias)
    ref = cs * (x.float() @ w.float().T + t.float() @ Bs.float().T) + bias
    assert maxerr(y, ref, R8) < 1e-2


def test_proj_epilogues():
    M, N, K = 260, 768, 768
    x, w = rnd(M, K, seed=1), rnd(N, K, std=1 / math.sqrt(K), seed=2)
    bias = rnd(N, seed=3, dtype=torch.float32)
    res32, res16 = rnd(M, N, seed=4, dtype=torch.float32), rnd(M, N, seed=5)
    base = x.float() @ w.float().T + bias
    y = ops().proj(x, w, bias=bias, residual=res32, out_dtype=torch.float32)
    assert y.dtype == torch.float32 and maxerr(y, base + res32) < 5e-3
    y = ops().proj(x, w, bias=bias, residual=res16)
    assert maxerr(y, base + res16.float(), R8) < 1e-2
    y = ops().proj(x, w, bias=bias, act=1)
    assert maxerr(y, base * torch.sigmoid(1.702 * base), R8) < 1e-2
    # strided input view (a column slice of a wider buffer) and strided output view
    wide = rnd(M, 3 * K, seed=7)
    outbuf = torch.zeros(M, 2 * N, device="cuda", dtype=BF)
    ops().proj(wide[:, K:2 * K], w, out=outbuf[:, N:])
    assert maxerr(outbuf[:, N:], wide[:, K:2 * K].float() @ w.float().T, R8) < 1e-2
    assert outbuf[:, :N].abs().max().item() == 0


def test_proj_geglu():
    """Packed [a(64)|gate(64)] tiles: out = a * gelu(gate)  (ldm/modules/attention.py:31-38)."""
    M, C = 200, 320
    x = rnd(M, C, seed=1)
    W, b = rnd(8 * C, C, std=1 / math.sqrt(C), seed=2), rnd(8 * C, seed=3, dtype=torch.float32)
    inner = 4 * C
    idx = torch.arange(inner).view(-1, 64)
    perm = torch.cat([idx, idx + inner], dim=1).reshape(-1).cuda()
    y = ops().proj(x, W[perm].contiguous(), bias=b[perm].contiguous(), act=2)
    h = x.float() @ W.float().T + b
    ref = h[:, :inner] * F.gelu(h[:, inner:])
    assert y.shape == (M, inner)
    assert maxerr(y, ref, R8) < 1e-2


def test_proj_rejects_bad_input():
    x, w = rnd(16, 20, seed=1), rnd(8, 20, seed=2)       # K = 20 is not a multiple of 8
    with pytest.raises(RuntimeError):
        ops().proj(x, w)
    with pytest.raises(RuntimeError):
        ops().proj(x.cpu(), w.cpu())


# ------------------------------------------------------------------------------------------- K2 attention
def ref_attn(q, k, v, H, scale, key_mask=None, causal_mult=0):
    B, Lq, C = q.shape
    d = C // H
    qh = q.float().cpu().view(B, Lq, H, d).transpose(1, 2)
    kh = k.float().cpu().reshape(B, -1, H, d).transpose(1, 2)
    vh = v.float().cpu().reshape(B, -1, H, d).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) * scale
    Lk = kh.shape[2]
    if key_mask is not None:
        s = s.masked_fill(~key_mask.cpu().bool()[:, None, None, :], float("-inf"))
    if causal_mult:
        i = torch.arange(Lq)[:, None]
        j = torch.arange(Lk)[None, :]
        s = s.masked_fill((j // causal_mult) > i, float("-inf"))
    p = s.softmax(-1)
    return (p @ vh).transpose(1, 2).reshape(B, Lq, C), s, p


@pytest.mark.parametrize("d,Lq,Lk", [(40, 256, 256), (40, 1000, 1000), (40, 77, 300), (80, 192, 130), (160, 64, 64),
                                     (160, 100, 77), (64, 20, 20), (40, 4096, 77), (40, 1, 1),
                                     # persistent short-context tcgen05 kernel (Lk <= 128, Lq >= 512)
                                     (80, 1024, 77), (40, 1000, 97), (40, 4096, 128), (80, 600, 16), (40, 513, 1), (40, 2048, 120),
                                     # four-tile ("quad") kernel: d = 40, Lq >= 1024, Lk > 128; ragged key and query tails
                                     (40, 2048, 1111), (40, 1500, 4096), (40, 1024, 129)])
def test_attention(d, Lq, Lk):
    B, H = 2, 8 if d != 64 else 12
    C = H * d
    q, k, v = rnd(B, Lq, C, seed=1), rnd(B, Lk, C, seed=2), rnd(B, Lk, C, seed=3)
    o = ops().attention(q, k, v, H, d ** -0.5)
    ref, _, _ = ref_attn(q, k, v, H, d ** -0.5)
    assert maxerr(o, ref) < 2e-2


@pytest.mark.parametrize("B,Lq,Lk", [(5, 2048, 2048), (5, 2048, 1111), (3, 4096, 4096)])
def test_attention_quad_wave_tail(B, Lq, Lk):
    """More (batch, head, 512-query) units than SMs: the four-tile kernel runs whole waves and hands the remainder to its
    tail launch -- two tiles x two key halves merged in the epilogue (even number of full key tiles) or two-tile CTAs
    (ragged keys).  Also checks the log-sum-exp the training path keeps."""
    import math
    H, d = 8, 40
    C = H * d
    qkv = rnd(B, max(Lq, Lk), 3 * C, seed=7)
    q, k, v = qkv[:, :Lq, :C], qkv[:, :Lk, C:2 * C], qkv[:, :Lk, 2 * C:]
    lse = torch.empty(B, H, Lq, device="cuda")
    o = ops().attention(q, k, v, H, d ** -0.5, lse=lse)
    hd = lambda t: t.float().reshape(B, -1, H, d).transpose(1, 2)
    ref = torch.nn.functional.scaled_dot_product_attention(hd(q), hd(k), hd(v)).transpose(1, 2).reshape(B, Lq, C)
    assert (o.float() - ref).abs().max().item() < 2e-2
    rlse = torch.logsumexp(hd(q) @ hd(k).transpose(-1, -2) * d ** -0.5, dim=-1) * math.log2(math.e)
    assert (lse - rlse).abs().max().item() < 2e-2


def test_attention_fused_qkv_views_and_key_mask():
    B, N, H, d = 2, 320, 8, 40
    C = H * d
    qkv = rnd(B, N, 3 * C, seed=1)
    q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
    g = torch.Generator().manual_seed(5)
    mask = (torch.rand(B, N, generator=g) > 0.4).to(torch.uint8).cuda()
    o = ops().attention(q, k, v, H, d ** -0.5, key_mask=mask)
    ref, _, _ = ref_attn(q, k, v, H, d ** -0.5, key_mask=mask)
    assert maxerr(o, ref) < 2e-2


@pytest.mark.parametrize("B,Lq,Lk,pattern", [(2, 4096, 4096, "random"), (2, 2048, 1111, "random"), (5, 2048, 2048, "lead"), (3, 4096, 4096, "blocks"),
                                             (1, 1024, 129, "last_only")])
def test_attention_key_mask_on_tcgen05(B, Lq, Lk, pattern):
    """img_mask self-attention at level A (dalc:254-273) on the four-tile tcgen05 kernel: the mask rides in the spare K column of
    the zero-padded head dim.  Patterns: random keys, the first key tiles entirely masked (their partial results must be erased by
    the first rescale), whole 64-key tiles masked here and there, a single surviving key in the ragged last tile; bulk + tail
    launches (B = 5 / 3) and the log-sum-exp of the training path."""
    import math
    H, d = 8, 40
    C = H * d
    qkv = rnd(B, max(Lq, Lk), 3 * C, seed=11)
    q, k, v = qkv[:, :Lq, :C], qkv[:, :Lk, C:2 * C], qkv[:, :Lk, 2 * C:]
    g = torch.Generator().manual_seed(5)
    if pattern == "random":
        mask = torch.rand(B, Lk, generator=g) > 0.4
    elif pattern == "lead":
        mask = torch.ones(B, Lk, dtype=torch.bool)
        mask[:, :200] = False
        mask[0, :1500] = False
    elif pattern == "blocks":
        mask = (torch.rand(B, Lk // 64, generator=g) > 0.5).repeat_interleave(64, dim=1)
        mask[:, -1] = True
    else:
        mask = torch.zeros(B, Lk, dtype=torch.bool)
        mask[:, -1] = True
    mask = mask.to(torch.uint8).cuda()
    lse = torch.empty(B, H, Lq, device="cuda")
    n0 = a_lib().launch_count()
    o = ops().attention(q, k, v, H, d ** -0.5, key_mask=mask, lse=lse)
    ref, s, _ = ref_attn(q, k, v, H, d ** -0.5, key_mask=mask)
    assert maxerr(o, ref) < 2e-2
    rlse = torch.logsumexp(s, dim=-1) * math.log2(math.e)
    assert (lse.cpu() - rlse).abs().max().item() < 2e-2
    assert a_lib().launch_count() - n0 <= 2      # bulk (+ tail) launch of the four-tile kernel: the mask no longer falls to the warp-MMA kernel


@pytest.mark.parametrize("d,B,N,pattern", [(80, 3, 1024, "random"), (80, 2, 1024, "lead"), (40, 2, 512, "random"), (80, 1, 256, "last_only")])
def test_attention_key_mask_small_cta_kernel(d, B, N, pattern):
    """img_mask self-attention at level B (d = 80, 1024 tokens) and on short d = 40 maps: the small-CTA tcgen05 kernel masks the scores
    in registers (mask bytes of the key tile, broadcast loads).  Leading key tiles entirely masked, a single surviving key, lse."""
    import math
    H = 8
    C = H * d
    qkv = rnd(B, N, 3 * C, seed=13)
    q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
    g = torch.Generator().manual_seed(7)
    if pattern == "random":
        mask = torch.rand(B, N, generator=g) > 0.4
    elif pattern == "lead":
        mask = torch.ones(B, N, dtype=torch.bool)
        mask[:, :130] = False
        mask[0, :700] = False
    else:
        mask = torch.zeros(B, N, dtype=torch.bool)
        mask[:, -1] = True
    mask = mask.to(torch.uint8).cuda()
    lse = torch.empty(B, H, N, device="cuda")
    n0 = a_lib().launch_count()
    o = ops().attention(q, k, v, H, d ** -0.5, key_mask=mask, lse=lse)
    assert a_lib().launch_count() - n0 == 1
    ref, s, _ = ref_attn(q, k, v, H, d ** -0.5, key_mask=mask)
    assert maxerr(o, ref) < 2e-2
    rlse = torch.logsumexp(s, dim=-1) * math.log2(math.e)
    assert (lse.cpu() - rlse).abs().max().item() < 2e-2


@pytest.mark.parametrize("mult,T", [(1, 20), (2, 20), (4, 24), (1, 77), (2, 77), (8, 77)])
def test_attention_causal_multi_kv(mult, T):
    """CLIPAttentionMKV: each token carries `mult` keys back to back; key j visible iff j // mult <= i."""
    B, H, d = 3, 12, 64
    E = H * d
    buf = rnd(B, T, E * (1 + 2 * mult), seed=1)
    q, k, v = buf[:, :, :E], buf[:, :, E:E + E * mult], buf[:, :, E + E * mult:]
    o = ops().attention(q, k, v, H, d ** -0.5, causal_mult=mult)
    ref, _, _ = ref_attn(q, k.reshape(B, T * mult, E), v.reshape(B, T * mult, E), H, d ** -0.5, causal_mult=mult)
    assert maxerr(o, ref) < 2e-2


@pytest.mark.parametrize("d,N,B", [(40, 320, 2), (80, 200, 1), (160, 64, 2)])
def test_proj_heads_and_headmajor_attention(d, N, B):
    """Fused QKV projection scattered into the padded head-major workspace + attention on that layout == the
    interleaved-layout path (same arithmetic, different addresses)."""
    H = 8
    C = H * d
    x, w = rnd(B * N, C, seed=1), rnd(3 * C, C, std=1 / math.sqrt(C), seed=2)
    bias = rnd(3 * C, seed=3, dtype=torch.float32)
    ws = ops().proj_heads(x, w, H, d, N, bias=bias)
    dpad = ws.shape[-1]
    ref = (x.float() @ w.float().T + bias).view(B, N, 3, H, d).permute(2, 0, 3, 1, 4)
    assert maxerr(ws[..., :d], ref, R8) < 1e-2
    assert ws[..., d:].abs().max().item() == 0 if dpad > d else True
    qkv = ops().proj(x, w, bias=bias).view(B, N, 3 * C)
    o_ref = ops().attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], H, d ** -0.5)
    o = ops().attention_headmajor(ws[0], ws[1], ws[2], d ** -0.5, d=d)
    assert maxerr(o, o_ref) < 1e-2


# ------------------------------------------------------------------------------------------- K3 capture
@pytest.mark.parametrize("d,Lq,S", [(40, 256, 77), (40, 100, 97), (80, 64, 77), (160, 64, 77), (40, 4096, 77), (40, 30, 128)])
def test_cross_capture_plain(d, Lq, S):
    B, H = 2, 8
    C = H * d
    q, k, v = rnd(B, Lq, C, seed=1), rnd(B, S, C, seed=2), rnd(B, S, C, seed=3)
    o, prob, score, _ = ops().attention_cross_capture(q, k, v, H, d ** -0.5)
    ref, s, p = ref_attn(q, k, v, H, d ** -0.5)
    assert maxerr(o, ref) < 2e-2
    assert maxerr(score, s) < 2e-3            # bf16 products, fp32 accumulate: exact up to summation order
    assert maxerr(prob, p) < 1e-3
    assert abs(prob.sum(-1).mean().item() - 1) < 1e-5


def test_cross_capture_fp32_inputs_split_precision():
    """fp32 q/k/v (projection GEMM fp32 output): scores via bf16 hi/lo split are accurate to ~1e-5."""
    B, H, d, Lq, S = 2, 8, 40, 200, 77
    C = H * d
    q, k, v = (rnd(B, L, C, seed=s, dtype=torch.float32) for L, s in ((Lq, 1), (S, 2), (S, 3)))
    o, prob, score, _ = ops().attention_cross_capture(q, k, v, H, d ** -0.5)
    ref, s_, p_ = ref_attn(q, k, v, H, d ** -0.5)
    assert maxerr(score, s_) < 2e-4 and maxerr(prob, p_) < 5e-5 and maxerr(o, ref) < 2e-2


def test_cross_capture_normalize_and_subj_cols():
    B, H, d, Lq, S = 2, 8, 40, 300, 77
    C = H * d
    q, k, v = rnd(B, Lq, C, seed=1), rnd(B, S, C, seed=2), rnd(B, S, C, seed=3)
    ib = torch.arange(B).repeat_interleave(16)
    in_ = torch.arange(4, 20).repeat(B)
    flag = torch.zeros(B, S, dtype=torch.uint8)
    flag[ib, in_] = 1
    ca = torch.tensor([0.8], device="cuda")
    cols = in_.view(B, 16).to(torch.int32).cuda()
    qm = ops().qmean(q)
    assert maxerr(qm, q.float().mean(dim=1)) < 1e-4
    o, prob, score, psub = ops().attention_cross_capture(q, k, v, H, d ** -0.5, col_flag=flag.cuda(), qmean=qm,
                                                        ca_scale=ca, subj_cols=cols)
    _, s, _ = ref_attn(q, k, v, H, d ** -0.5)
    sub = s[ib, :, :, in_]
    sub = (sub - sub.mean(dim=2, keepdim=True)) * 0.8
    s2 = s.clone()
    s2[ib, :, :, in_] = sub
    p2 = s2.softmax(-1)
    vh = v.float().cpu().view(B, S, H, d).transpose(1, 2)
    ref = (p2 @ vh).transpose(1, 2).reshape(B, Lq, C)
    assert maxerr(score, s2) < 3e-3 and maxerr(prob, p2) < 1e-3 and maxerr(o, ref) < 2e-2
    assert maxerr(psub, p2[:, :, :, 4:20]) < 1e-3


def test_cross_capture_mix():
    B, H, d, Lq, S = 4, 8, 40, 130, 77
    C = H * d
    q, k, v = rnd(B, Lq, C, seed=1), rnd(B, S, C, seed=2), rnd(B, S, C, seed=3)
    o, prob, score, _ = ops().attention_cross_capture(q, k, v, H, d ** -0.5, mix=True)
    _, s, _ = ref_attn(q, k, v, H, d ** -0.5)
    sm = ((s[:2] + s[2:]) / 2).repeat(2, 1, 1, 1)
    p = sm.softmax(-1)
    vh = v.float().cpu().view(B, S, H, d).transpose(1, 2)
    ref = (p @ vh).transpose(1, 2).reshape(B, Lq, C)
    assert maxerr(score, sm) < 3e-3 and maxerr(prob, p) < 1e-3 and maxerr(o, ref) < 2e-2


def test_cross_capture_rejects():
    q, k = rnd(3, 64, 320, seed=1), rnd(3, 77, 320, seed=2)
    with pytest.raises(RuntimeError):
        ops().attention_cross_capture(q, k, k, 8, 0.1, mix=True)          # odd batch
    k2 = rnd(2, 200, 320, seed=3)
    with pytest.raises(RuntimeError):
        ops().attention_cross_capture(rnd(2, 64, 320), k2, k2, 8, 0.1)    # > 128 keys


# ------------------------------------------------------------------------------------------- K4 & helpers
@pytest.mark.parametrize("C", [320, 640, 768, 1280])
@pytest.mark.parametrize("dtype", [BF, torch.float32])
def test_layernorm(C, dtype):
    M = 1000
    x = rnd(M, C, std=2.0, seed=1, dtype=dtype) + 0.5
    w, b = rnd(C, seed=2, dtype=torch.float32), rnd(C, seed=3, dtype=torch.float32)
    y = ops().layernorm(x, w, b, 1e-5)
    ref = F.layer_norm(x.float(), (C,), w, b, 1e-5)
    assert maxerr(y, ref, R8) < 1e-2
    y32 = ops().layernorm(x.float(), w, b, 1e-5, out_dtype=torch.float32)
    assert maxerr(y32, ref) < 1e-4


def test_chan_major_and_sbg_head():
    x = rnd(2, 300, 320, seed=1)
    y = ops().chan_major(x, 0.25)
    assert maxerr(y, x.float().permute(0, 2, 1) * 0.25) < 1e-6
    x32 = rnd(2, 77, 320, seed=2, dtype=torch.float32)
    assert maxerr(ops().chan_major(x32, 1.0), x32.permute(0, 2, 1)) == 0
    hs = [rnd(130, 768, seed=s, dtype=torch.float32) for s in (1, 2, 3)]
    w, b = rnd(768, seed=4, dtype=torch.float32), rnd(768, seed=5, dtype=torch.float32)
    wl = [1 / 7, 2 / 7, 4 / 7]
    out = ops().sbg_head(hs, wl, w, b)
    ref = F.layer_norm(sum(a * h for a, h in zip(wl, hs)), (768,), w, b, 1e-5)
    assert maxerr(out, ref) < 1e-4


def test_dora_pack_matches_peft_formula():
    """LoraDoraLinear.pack(): A, s*B and colscale = m / ||W + s B A||_row reproduce the oracle's
    lora_dora_linear when applied as y = colscale o (x W^T + (x A^T)(sB)^T) + b."""
    torch.manual_seed(0)
    import oracle
    import adaface_dev_b200 as a
    base = torch.nn.Linear(48, 32).cuda()
    lora = a.LoraDoraLinear(base, r=8, lora_alpha=2).cuda()
    with torch.no_grad():
        lora.lora_B["default"].weight.normal_(std=0.05)
        lora.lora_magnitude_vector["default"].weight.mul_(1.1)
    A16, Bs16, cs = lora.pack()
    assert A16.dtype == torch.bfloat16 and Bs16.dtype == torch.bfloat16 and cs.dtype == torch.float32
    x = torch.randn(5, 48).cuda()
    y = cs * (x @ base.weight.T + (x @ A16.float().T) @ Bs16.float().T) + base.bias
    c_ = lambda t_: t_.detach().cpu()
    ref = oracle.lora_dora_linear(c_(x), c_(base.weight), c_(base.bias), c_(lora.lora_A["default"].weight), c_(lora.lora_B["default"].weight),
                                  c_(lora.lora_magnitude_vector["default"].weight), lora.scaling)
    assert (y.detach().cpu() - ref).abs().max().item() < 2e-2          # bf16 rounding of A and s*B only
    # the DoRA column scale itself (adaface_dora_colscale + B.A on the projection GEMM) against fp32 torch
    wn = torch.linalg.norm(base.weight.float() + lora.scaling * lora.lora_B["default"].weight.float() @ lora.lora_A["default"].weight.float(), dim=1)
    assert ((cs - lora.lora_magnitude_vector["default"].weight.float() / wn).abs().max() / cs.abs().max()).item() < 2e-3
    # identity at init (peft: B = 0, m = ||W||_row)
    fresh = a.LoraDoraLinear(torch.nn.Linear(48, 32).cuda(), r=8, lora_alpha=2).cuda()
    _, Bs0, cs0 = fresh.pack()
    assert Bs0.abs().max().item() == 0 and (cs0 - 1).abs().max().item() < 1e-6
    # the pack is cached until a parameter changes
    assert lora.pack()[0] is A16
    with torch.no_grad():
        lora.lora_A["default"].weight.add_(1.0)
    assert lora.pack()[0] is not A16
