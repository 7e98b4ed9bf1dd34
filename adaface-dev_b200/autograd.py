"""Autograd glue for the stage-2 training step (north-star item 5; reference: autograd over eager PyTorch ops,
ddpm.py:1645-1707 back-propagating through the `sc` U-Net instance into the LoRA/DoRA adapters,
``cross_attn_scale_factor`` and -- through the prompt context -- SubjBasisGenerator).

Every ``torch.autograd.Function`` here is a thin pairing of one forward kernel call with its backward kernel calls
(include/adaface_b200.h, K5).  PyTorch's engine only orders the calls and sums gradients that fan in; no gradient
arithmetic runs in PyTorch.  Conventions:

* activations bf16 (fp32 for the SubjBasisGenerator residual stream and the high-precision capture path);
  an upstream gradient arrives in the dtype of the output it belongs to and is cast to bf16 once for the GEMMs;
* base U-Net weights are FROZEN (the reference freezes the U-Net, ddpm.py:637-638): they receive no gradient even if
  ``requires_grad`` is left on; trainable are the LoRA A / B / DoRA magnitude, ``cross_attn_scale_factor`` and every
  SubjBasisGenerator parameter except the embeddings;
* weight gradients dW = dY^T X run on the same K-major tcgen05 GEMM over operands re-laid by ``adaface_transpose``.
"""
import torch
from torch.autograd import Function

from . import ops

BF16 = torch.bfloat16


def _b16(t):
    t = t if t.dtype == BF16 else t.to(BF16)
    return t if t.stride(-1) == 1 else t.contiguous()


def _transposed(pack, key):
    """[K, N] copy of the frozen bf16 weight pack[key] ([N, K]), cached next to it (dX = dY W)."""
    wt = pack.get(key + "_t")
    if wt is None:
        wt = pack[key + "_t"] = ops.transpose(pack[key])
    return wt


class FrozenLinearFn(Function):
    """y = x W^T + bias (+ residual) with a frozen weight: backward is one GEMM, dX = dY W."""

    @staticmethod
    def forward(ctx, x, pack, wkey, bkey, residual, out_dtype):
        ctx.pack, ctx.wkey = pack, wkey
        ctx.res_dtype = None if residual is None else residual.dtype
        return ops.proj(x, pack[wkey], bias=pack.get(bkey) if bkey else None, residual=residual, out_dtype=out_dtype)

    @staticmethod
    def backward(ctx, dy):
        dx = dres = None
        if ctx.needs_input_grad[0]:
            dx = ops.proj(_b16(dy), _transposed(ctx.pack, ctx.wkey))
        if ctx.res_dtype is not None and ctx.needs_input_grad[4]:
            dres = dy if dy.dtype == ctx.res_dtype else dy.to(ctx.res_dtype)
        return dx, None, None, None, dres, None


class TrainLinearFn(Function):
    """y = x W^T + bias (+ residual) with TRAINABLE fp32 parameters (SubjBasisGenerator).  ``params`` =
    (w0, b0, w1, b1, ...) are the nn.Linear parameters whose rows are concatenated in the bf16 pack (fused Q|K|V)."""

    @staticmethod
    def forward(ctx, x, pack, wkey, bkey, residual, out_dtype, *params):
        ctx.pack, ctx.wkey = pack, wkey
        ctx.res_dtype = None if residual is None else residual.dtype
        ctx.rows = [p.shape[0] for p in params[0::2]]
        ctx.has_bias = [b is not None for b in params[1::2]]
        ctx.save_for_backward(x)
        return ops.proj(x, pack[wkey], bias=pack.get(bkey) if bkey else None, residual=residual, out_dtype=out_dtype)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy16 = _b16(dy)
        dx = dres = None
        if ctx.needs_input_grad[0]:
            dx = ops.proj(dy16, _transposed(ctx.pack, ctx.wkey))
        if ctx.res_dtype is not None and ctx.needs_input_grad[4]:
            dres = dy if dy.dtype == ctx.res_dtype else dy.to(ctx.res_dtype)
        dyt = ops.transpose(dy16, pad_to=8)                       # [N, Mp]
        xt = ops.transpose(x, pad_to=8)                           # [K, Mp]
        dw = ops.proj(dyt, xt, out_dtype=torch.float32)           # [N, K] = dY^T X
        db = ops.colsum(dy16)
        grads, r0 = [], 0
        for n, hb in zip(ctx.rows, ctx.has_bias):
            grads += [dw[r0:r0 + n], db[r0:r0 + n] if hb else None]
            r0 += n
        return (dx, None, None, None, dres, None) + tuple(grads)


class LoraLinearFn(Function):
    """y = colscale o (x W^T + s (x A^T) B^T) + bias -- peft lora.Linear with DoRA in eval form, base
    weight frozen, A / B / magnitude trainable, colscale = m / ||W + s B A||_row detached as in peft."""

    @staticmethod
    def forward(ctx, x, pack, wkey, bkey, lora, A, B, m, out_dtype):
        A16, Bs16, cs = lora.pack()
        bias = pack.get(bkey) if bkey else None
        t = ops.proj(x, A16)
        y = ops.proj(x, pack[wkey], t=t, bs=Bs16, colscale=cs, bias=bias, out_dtype=out_dtype)
        ctx.pack, ctx.wkey, ctx.bias, ctx.scaling = pack, wkey, bias, float(lora.scaling)
        ctx.save_for_backward(x, t, y, A16, Bs16, cs, m)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, t, y, A16, Bs16, cs, m = ctx.saved_tensors
        dy16 = _b16(dy)
        w16 = ctx.pack[ctx.wkey]
        # dZ = dY o cs is never materialised: cs is folded into the B operands (rowscale) or the transpose (colscale)
        dt = ops.proj(dy16, ops.transpose(Bs16, rowscale=cs))                       # [M, R] = dZ Bs
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.proj(dy16, ops.transpose(w16, rowscale=cs), t=dt, bs=ops.transpose(A16))   # dZ W + dT A
        dzt = ops.transpose(dy16, colscale=cs, alpha=ctx.scaling, pad_to=8)         # [N, Mp] = s dZ^T
        dB = ops.proj(dzt, ops.transpose(t, pad_to=8), out_dtype=torch.float32)     # [N, R] = s dZ^T T
        dA = ops.proj(ops.transpose(dt, pad_to=8), ops.transpose(x, pad_to=8), out_dtype=torch.float32)   # [R, K] = dT^T X
        dm = ops.colsum(dy16, b=y, bias=ctx.bias, colmul=m.detach().float().reciprocal().contiguous())
        return dx, None, None, None, None, dA, dB, dm.to(m.dtype), None


class AttentionFn(Function):
    """Flash attention over column slices of (possibly one and the same) projection buffers: q = q_t[..., q_off:+Cq],
    k = k_t[..., k_off:+Ckv], v = v_t[..., v_off:+Ckv].  Passing the fused QKV buffer as all three tensors makes the
    backward kernels write straight into one fused dQKV buffer (no slice-gradient copies)."""

    @staticmethod
    def forward(ctx, q_t, k_t, v_t, offs, Cq, Ckv, heads, scale, key_mask, causal_mult):
        qo, ko, vo = offs
        q, k, v = q_t[:, :, qo:qo + Cq], k_t[:, :, ko:ko + Ckv], v_t[:, :, vo:vo + Ckv]
        B, Lq = q.shape[0], q.shape[1]
        o = torch.empty((B, Lq, Cq), device=q.device, dtype=BF16)
        lse = torch.empty((B, heads, Lq), device=q.device, dtype=torch.float32)
        ops.attention(q, k, v, heads, scale, key_mask=key_mask, causal_mult=causal_mult, out=o, lse=lse)
        ctx.cfg = (offs, Cq, Ckv, heads, scale, causal_mult)
        ctx.same_kq, ctx.same_vq, ctx.same_vk = k_t is q_t, v_t is q_t, v_t is k_t
        ctx.save_for_backward(q_t, k_t, v_t, o, lse, key_mask)
        return o

    @staticmethod
    def backward(ctx, do):
        q_t, k_t, v_t, o, lse, key_mask = ctx.saved_tensors
        (qo, ko, vo), Cq, Ckv, heads, scale, causal_mult = ctx.cfg
        if ctx.same_kq:
            k_t = q_t
        if ctx.same_vq:
            v_t = q_t
        elif ctx.same_vk:
            v_t = k_t
        used = {}
        for t, w in ((q_t, Cq), (k_t, Ckv), (v_t, Ckv)):
            used[id(t)] = used.get(id(t), 0) + w
        grads = {}
        for t in (q_t, k_t, v_t):
            if id(t) not in grads:
                alloc = torch.empty_like if used[id(t)] == t.shape[2] else torch.zeros_like
                grads[id(t)] = alloc(t, memory_format=torch.contiguous_format)
        dq = grads[id(q_t)][:, :, qo:qo + Cq]
        dk = grads[id(k_t)][:, :, ko:ko + Ckv]
        dv = grads[id(v_t)][:, :, vo:vo + Ckv]
        ops.attention_bwd(q_t[:, :, qo:qo + Cq], k_t[:, :, ko:ko + Ckv], v_t[:, :, vo:vo + Ckv], o, _b16(do), lse, heads, scale,
                          dq, dk, dv, key_mask=key_mask, causal_mult=causal_mult)
        gk = None if ctx.same_kq else grads[id(k_t)]
        gv = None if (ctx.same_vq or ctx.same_vk) else grads[id(v_t)]
        return grads[id(q_t)], gk, gv, None, None, None, None, None, None, None


class CrossCaptureFn(Function):
    """The slow SDPA of dalc:79-139 (capture / normalize) with gradients through out, prob and score at once."""

    @staticmethod
    def forward(ctx, q, k, v, ca_param, heads, scale, want_prob, want_score, col_flag, subj_cols, mix, dca_mul):
        qm = ops.qmean(q) if col_flag is not None else None
        ca = ca_param.detach().float().reshape(1)
        out, prob, score, prob_subj = ops.attention_cross_capture(q, k, v, heads, scale, want_prob=want_prob,
                                                                  want_score=want_score, col_flag=col_flag, qmean=qm,
                                                                  ca_scale=ca, mix=mix, subj_cols=subj_cols)
        ctx.cfg = (heads, scale, mix, dca_mul)
        ctx.save_for_backward(q, k, v, ca, col_flag, qm, ca_param)
        if prob_subj is not None:
            ctx.mark_non_differentiable(prob_subj)
        return out, prob, score, prob_subj

    @staticmethod
    def backward(ctx, dout, dprob, dscore, _dsubj):
        q, k, v, ca, col_flag, qm, ca_param = ctx.saved_tensors
        heads, scale, mix, dca_mul = ctx.cfg
        B, Lq, C = q.shape
        if dout is None:
            dout = torch.zeros((B, Lq, C), device=q.device, dtype=BF16)
        fix = lambda g: None if g is None else g.float().contiguous()
        dq, dk, dv, dca = ops.attention_cross_capture_bwd(q, k, v, _b16(dout), heads, scale, dprob=fix(dprob), dscore=fix(dscore),
                                                          col_flag=col_flag, qmean=qm, ca_scale=ca, mix=mix,
                                                          dca_mul=dca_mul, dkv_dtype=k.dtype)
        dcap = dca.reshape(ca_param.shape).to(ca_param.dtype) if ctx.needs_input_grad[3] else None
        return dq.to(q.dtype), dk, dv, dcap, None, None, None, None, None, None, None, None


class CrossConsumeFn(Function):
    """Capture with fused consumers: out, subj_sum [B,H,Lq] and sqdiff [B] = sum (prob - ref_prob)^2, all
    differentiable w.r.t. q, k, v (and cross_attn_scale_factor under normalize) -- no [B,H,Lq,S] map in either direction."""

    @staticmethod
    def forward(ctx, q, k, v, ca_param, heads, scale, col_flag, sum_flag, ref_prob, dca_mul):
        qm = ops.qmean(q) if col_flag is not None else None
        ca = ca_param.detach().float().reshape(1)
        out, subj_sum, sqdiff, _ = ops.attention_cross_consume(q, k, v, heads, scale, sum_flag=sum_flag, ref_prob=ref_prob,
                                                               col_flag=col_flag, qmean=qm, ca_scale=ca)
        ctx.cfg = (heads, scale, dca_mul)
        ctx.save_for_backward(q, k, v, ca, col_flag, qm, ca_param, sum_flag, ref_prob)
        dev = q.device
        if subj_sum is None:
            subj_sum = torch.zeros((), device=dev)
            ctx.mark_non_differentiable(subj_sum)
        if sqdiff is None:
            sqdiff = torch.zeros((), device=dev)
            ctx.mark_non_differentiable(sqdiff)
        return out, subj_sum, sqdiff

    @staticmethod
    def backward(ctx, dout, dsum, dsq):
        q, k, v, ca, col_flag, qm, ca_param, sum_flag, ref_prob = ctx.saved_tensors
        heads, scale, dca_mul = ctx.cfg
        B, Lq, C = q.shape
        if dout is None:
            dout = torch.zeros((B, Lq, C), device=q.device, dtype=BF16)
        g_subj = dsum.float().contiguous() if (dsum is not None and sum_flag is not None) else None
        coef = None
        if ref_prob is not None and dsq is not None:
            coef = (2.0 * dsq.float().reshape(-1).expand(B)).contiguous()    # d/dP sum (P - R)^2 = 2 (P - R), times upstream, per instance
        dq, dk, dv, dca = ops.attention_cross_consume_bwd(q, k, v, _b16(dout), heads, scale, sum_flag=sum_flag if g_subj is not None else None,
                                                          g_subj=g_subj, ref_prob=ref_prob if coef is not None else None, mse_coef=coef,
                                                          col_flag=col_flag, qmean=qm, ca_scale=ca, dca_mul=dca_mul, dkv_dtype=k.dtype)
        dcap = dca.reshape(ca_param.shape).to(ca_param.dtype) if ctx.needs_input_grad[3] else None
        return dq.to(q.dtype), dk, dv, dcap, None, None, None, None, None, None


class ChanMajorFn(Function):
    """cached q / q2 / k / v / attn_out: 'b n c -> b c n' times a factor, fp32 (dalc:349-362)."""

    @staticmethod
    def forward(ctx, src, factor):
        ctx.factor, ctx.dtype = factor, src.dtype
        return ops.chan_major(src, factor)

    @staticmethod
    def backward(ctx, dcap):
        return ops.transpose(dcap.float().contiguous(), out_dtype=ctx.dtype, alpha=ctx.factor), None


class LayerNormFn(Function):
    @staticmethod
    def forward(ctx, x, w, b, eps, out_dtype):
        wf = w.detach().float().contiguous()
        ctx.eps = eps
        ctx.save_for_backward(x, wf)
        return ops.layernorm(x, wf, b.detach().float().contiguous(), eps, out_dtype=out_dtype)

    @staticmethod
    def backward(ctx, dy):
        x, wf = ctx.saved_tensors
        wg = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dx, dw, db = ops.layernorm_bwd(x, dy if dy.stride(-1) == 1 else dy.contiguous(), wf, ctx.eps, want_wgrad=wg)
        return dx, dw, db, None, None


class ActFn(Function):
    """Stand-alone activation of the training path (quick-GELU / packed GEGLU); the pre-activation is what is saved."""

    @staticmethod
    def forward(ctx, u, act):
        ctx.act = act
        ctx.save_for_backward(u)
        return ops.act_fwd(u, act)

    @staticmethod
    def backward(ctx, dh):
        (u,) = ctx.saved_tensors
        return ops.act_bwd(u, _b16(dh), ctx.act), None


class SbgHeadFn(Function):
    """LayerNorm(sum_l wl[l] h_l) (arc2face_models.py:291-306) with gradients to the hidden states, the layer weights
    and the final LayerNorm."""

    @staticmethod
    def forward(ctx, wl_t, ln_w, ln_b, eps, *hs):
        # the (normalised) layer weights stay on the DEVICE: no host read per step, the whole step is graph-capturable
        wl = wl_t.detach().float().reshape(-1).contiguous()
        wf = ln_w.detach().float().contiguous()
        ctx.eps, ctx.wl_shape, ctx.wl_dtype = eps, wl_t.shape, wl_t.dtype
        ctx.save_for_backward(wf, wl, *hs)
        return ops.sbg_head(list(hs), wl, wf, ln_b.detach().float().contiguous(), eps)

    @staticmethod
    def backward(ctx, dout):
        wf, wl, *hs = ctx.saved_tensors
        dhs, dwl, dw, db = ops.sbg_head_bwd(hs, wl, wf, dout.float().contiguous(), ctx.eps)
        return (dwl.reshape(ctx.wl_shape).to(ctx.wl_dtype), dw, db, None) + tuple(dhs)


# ------------------------------------------------------------------------------------------------ helpers
# ------------------------------------------------------------------------------------------------ U-Net convolutional blocks
def _b16c(t):
    t = t if t.dtype == BF16 else t.to(BF16)
    return t if t.is_contiguous() else t.contiguous()


def needs_grad(*ts):
    return torch.is_grad_enabled() and any(t is not None and torch.is_tensor(t) and t.requires_grad for t in ts)


class GroupNormActFn(Function):
    """act(GroupNorm(x)) over NHWC tokens with frozen affine parameters: backward = adaface_groupnorm_act_tokens_bwd."""

    @staticmethod
    def forward(ctx, x, gamma, beta, groups, eps, silu):
        ctx.cfg = (groups, eps, silu)
        ctx.save_for_backward(x, gamma, beta)
        return ops.groupnorm_act_tokens(x, gamma, beta, groups, eps, silu=silu)

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta = ctx.saved_tensors
        groups, eps, silu = ctx.cfg
        return ops.groupnorm_act_tokens_bwd(x, _b16c(dy), gamma, beta, groups, eps, silu=silu), None, None, None, None, None


class Conv3x3Fn(Function):
    """3x3 convolution with a FROZEN weight (+ bias, per-image bias, residual).  dX is the same implicit-GEMM kernel over the
    flipped / transposed weight pack (stride 2: after zero-insertion of dY); the residual's gradient is dY itself."""

    @staticmethod
    def forward(ctx, x, pack, wkey, weight, hw, stride, bias, rowbias, residual, out_dtype):
        ctx.pack, ctx.wkey, ctx.weight, ctx.hw, ctx.stride = pack, wkey, weight, hw, stride
        ctx.res_dtype = None if residual is None else residual.dtype
        return ops.conv3x3(x, pack[wkey], hw, stride=stride, bias=bias, rowbias=rowbias, residual=residual, out_dtype=out_dtype)

    @staticmethod
    def backward(ctx, dy):
        dx = dres = None
        if ctx.needs_input_grad[0]:
            key = ctx.wkey + "_dx"
            wdx = ctx.pack.get(key)
            if wdx is None:
                wdx = ctx.pack[key] = ops.pack_conv3x3_weight_dx(ctx.weight)
            g = _b16c(dy)
            if g.shape[2] % 8:                       # the U-Net's 4-channel output convolution: TMA needs 16-byte pixel strides
                g = torch.nn.functional.pad(g, (0, 8 - g.shape[2] % 8))
            if ctx.stride == 2:
                g = ops.resample2x_bwd(g, (ctx.hw[0] // 2, ctx.hw[1] // 2), 1)
            dx = ops.conv3x3(g, wdx, ctx.hw)
        if ctx.res_dtype is not None and ctx.needs_input_grad[8]:
            dres = dy if dy.dtype == ctx.res_dtype else dy.to(ctx.res_dtype)
        return dx, None, None, None, None, None, None, None, dres, None


class ConvLoraFn(Function):
    """peft lora.Conv2d with DoRA in eval form on a frozen 3x3 / 1x1 base convolution (dalc:541-591; SURVEY 8a A4, conv form):
        y = colscale o (conv(x, W) + s conv1x1(conv(x, A), B)) + bias,   colscale = m / ||W + s B.A|| (detached, as in peft)
    with A / B / magnitude trainable.  Forward = the implicit-GEMM kernel twice (T = conv(x, A), then the base convolution with
    the rank-r tail in the same TMEM tile).  Backward: dZ = dY o colscale; dT = dZ (sB); dX = conv(dZ, W') + conv(dT, A') on the
    same kernel over flipped / transposed weights; dB = s dZ^T T and dA = dT^T im2col(x) on the K-major projection GEMM;
    d magnitude as for the linear adapter."""

    @staticmethod
    def forward(ctx, x, lora, A, B, m, hw):
        wp, ap, bs, cs, bias = lora.pack()
        b, n, _ = x.shape
        if lora.is_3x3:
            t = ops.conv3x3(x, ap, hw).view(b * n, -1)
            y = ops.conv3x3(x, wp, hw, bias=bias, t=t, bs=bs, colscale=cs)
        else:
            t = ops.proj(x.view(b * n, -1), ap)
            y = ops.proj(x.view(b * n, -1), wp, t=t, bs=bs, colscale=cs, bias=bias).view(b, n, -1)
        ctx.lora, ctx.hw, ctx.bias = lora, hw, bias
        ctx.save_for_backward(x, t, y, bs, cs, m)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, t, y, bs, cs, m = ctx.saved_tensors
        lora, hw = ctx.lora, ctx.hw
        b, n, cin = x.shape
        dy16 = _b16c(dy)
        dy2d = dy16.view(b * n, -1)
        dt = ops.proj(dy2d, ops.transpose(bs, rowscale=cs))                                        # [P, r] = dZ (sB)
        W, A = lora.base_layer.weight, lora.lora_A[lora.adapter].weight
        dx = None
        if ctx.needs_input_grad[0]:
            if lora.is_3x3:
                # conv(dZ, W') with dZ = dY o colscale: the column scale is folded into W's output-channel axis before packing
                dx = ops.conv3x3(dy16, ops.pack_conv3x3_weight_dx(W.detach().float() * cs.view(-1, 1, 1, 1)), hw)
                dx = ops.conv3x3(dt.view(b, n, -1), ops.pack_conv3x3_weight_dx(A), hw, residual=dx)
            else:
                w16 = W.detach().flatten(1).to(BF16)
                dx = ops.proj(dy2d, ops.transpose(w16, rowscale=cs), t=dt, bs=ops.transpose(A.detach().flatten(1).to(BF16))).view(b, n, cin)
        dzt = ops.transpose(dy2d, colscale=cs, alpha=float(lora.scaling), pad_to=8)               # [cout, Pp] = s dZ^T
        dB = ops.proj(dzt, ops.transpose(t, pad_to=8), out_dtype=torch.float32)                   # [cout, r]
        dtt = ops.transpose(dt, pad_to=8)                                                          # [r, Pp]
        if lora.is_3x3:
            col = ops.im2col3x3_tokens(x, hw)                                                      # [P, 9 Kc]
            dA = ops.unpack_conv3x3_weight(ops.proj(dtt, ops.transpose(col, pad_to=8), out_dtype=torch.float32), cin)
        else:
            dA = ops.proj(dtt, ops.transpose(x.view(b * n, cin), pad_to=8), out_dtype=torch.float32).view_as(A)
        dm = ops.colsum(dy2d, b=y.view(b * n, -1), bias=ctx.bias, colmul=m.detach().float().reciprocal().contiguous())
        return dx, None, dA.view_as(A).to(A.dtype), dB.view(lora.lora_B[lora.adapter].weight.shape).to(A.dtype), dm.to(m.dtype), None


class Upsample2xFn(Function):
    @staticmethod
    def forward(ctx, x, hw):
        ctx.hw = hw
        return ops.upsample2x_tokens(x, hw)

    @staticmethod
    def backward(ctx, dy):
        return ops.resample2x_bwd(_b16c(dy), ctx.hw, 0), None


def groupnorm_act(x, gamma, beta, groups, eps, silu):
    if needs_grad(x):
        return GroupNormActFn.apply(x, gamma, beta, groups, eps, silu)
    return ops.groupnorm_act_tokens(x, gamma, beta, groups, eps, silu=silu)


def conv3x3(x, pack, wkey, weight, hw, *, stride=1, bias=None, rowbias=None, residual=None, out_dtype=BF16):
    """``weight`` = the module's [Cout, Cin, 3, 3] parameter (only read to build the dX pack on the first backward)."""
    if needs_grad(x, residual):
        return Conv3x3Fn.apply(x, pack, wkey, weight, hw, stride, bias, rowbias, residual, out_dtype)
    return ops.conv3x3(x, pack[wkey], hw, stride=stride, bias=bias, rowbias=rowbias, residual=residual, out_dtype=out_dtype)


def upsample2x(x, hw):
    if needs_grad(x):
        return Upsample2xFn.apply(x, hw)
    return ops.upsample2x_tokens(x, hw)


def linear(x, pack, wkey, bkey=None, *, lora=None, params=None, residual=None, out_dtype=BF16):
    """One projection of the training path: frozen weight (default), frozen weight + DoRA adapter (``lora``), or
    trainable nn.Linear parameters (``params`` = (w0, b0, w1, b1, ...))."""
    if lora is not None:
        if residual is not None:
            raise ValueError("linear: residual is not fused into the LoRA projection")
        ad = lora.adapter
        return LoraLinearFn.apply(x, pack, wkey, bkey, lora, lora.lora_A[ad].weight, lora.lora_B[ad].weight,
                                  lora.lora_magnitude_vector[ad].weight, out_dtype)
    if params is not None:
        return TrainLinearFn.apply(x, pack, wkey, bkey, residual, out_dtype, *params)
    return FrozenLinearFn.apply(x, pack, wkey, bkey, residual, out_dtype)


def attention(q_t, k_t, v_t, offs, Cq, Ckv, heads, scale, key_mask=None, causal_mult=0):
    return AttentionFn.apply(q_t, k_t, v_t, offs, Cq, Ckv, heads, scale, key_mask, causal_mult)
