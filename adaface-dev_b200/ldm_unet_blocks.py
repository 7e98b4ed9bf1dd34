"""Drop-in mirrors of the convolutional blocks of the reference's LDM U-Net.

    ResBlock     ldm/modules/diffusionmodules/openaimodel.py:164-260   (SD-1.5: use_scale_shift_norm=False, no up / down)
    Upsample     ldm/modules/diffusionmodules/openaimodel.py:92-119    (nearest 2x + 3x3 convolution)
    Downsample   ldm/modules/diffusionmodules/openaimodel.py:135-161   (3x3 convolution, stride 2)

Module / parameter names are the reference's (in_layers.0/2, emb_layers.1, out_layers.0/3, skip_connection, conv, op), so
SD-1.5 LDM checkpoints load with ``load_state_dict``.

B200 design: activations live as NHWC bf16 "tokens" [B, h*w, C] -- the layout of the attention path, so a
ResBlock -> SpatialTransformer chain needs no re-layout -- and a 3x3 convolution is an implicit GEMM on tcgen05 whose
activation tiles are shifted TMA boxes (zero padding = TMA out-of-bounds fill; adaface_conv3x3_fwd).  GroupNorm + SiLU is
one pass over the tokens, the time-embedding term rides in the first convolution's epilogue as a per-image bias and the
skip connection in the second one's as the residual.  ``forward`` keeps the reference's NCHW signature (one transpose
in, one out); ``forward_tokens`` is the NHWC-resident entry.  Training: ``forward_tokens`` is differentiable w.r.t. its token
input (autograd.py: GroupNormActFn / Conv3x3Fn / Upsample2xFn; the weights are the frozen U-Net's and get no gradient; the conv-LoRA
adapters -- ResBlock.conv_loras, LoraDoraConv2d -- train through ConvLoraFn).  CUDA only, no fallback.
"""
import torch
import torch.nn as nn

from . import ops
from . import autograd as ag
from .ldm_attention import _bf16, _f32, _ver


def _to_tokens(x):
    b, c, h, w = x.shape
    if x.dtype not in (torch.bfloat16, torch.float32):
        x = x.float()
    return ops.transpose(x.contiguous().view(b, c, h * w), out_dtype=torch.bfloat16)          # [B, h*w, C]


def _to_nchw(t, hw, dtype):
    b, _, c = t.shape
    if ag.needs_grad(t):                                        # training: the transposition has a backward (fp32 out)
        return ag.ChanMajorFn.apply(t, 1.0).view(b, c, *hw).to(dtype)
    if dtype not in (torch.bfloat16, torch.float32):            # fp16 callers (the reference's autocast): through fp32
        return ops.transpose(t, out_dtype=torch.float32).view(b, c, *hw).to(dtype)
    return ops.transpose(t, out_dtype=dtype).view(b, c, *hw)


def _no_grad_only(name, *ts):
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts):
        raise NotImplementedError(f"{name}: the training path of this module is not built yet")


class ResBlock(nn.Module):
    def __init__(self, channels, emb_channels, dropout=0., out_channels=None, use_conv=False, use_scale_shift_norm=False, dims=2,
                 use_checkpoint=False, up=False, down=False):
        super().__init__()
        if use_scale_shift_norm or up or down or dims != 2:
            raise NotImplementedError("ResBlock: only the SD-1.5 configuration (dims=2, no scale-shift norm, no up / down) is built")
        self.channels, self.emb_channels, self.dropout = channels, emb_channels, dropout
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.in_layers = nn.Sequential(nn.GroupNorm(32, channels), nn.SiLU(), nn.Conv2d(channels, self.out_channels, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(nn.GroupNorm(32, self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                        nn.Conv2d(self.out_channels, self.out_channels, 3, padding=1))
        for p in self.out_layers[3].parameters():                   # zero_module (openaimodel.py:233)
            nn.init.zeros_(p)
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        elif use_conv:
            self.skip_connection = nn.Conv2d(channels, self.out_channels, 3, padding=1)
        else:
            self.skip_connection = nn.Conv2d(channels, self.out_channels, 1)
        self._pack_key, self._pack = None, None
        # conv-LoRA adapters by diffusers name ('conv1', 'conv2', 'conv_shortcut'), installed by set_up_ffn_loras
        # (dalc:541-591: up_blocks.3.resnets.[12].conv*); consulted only while ``ffn_lora_on`` (set_lora_and_capture_flags)
        self.conv_loras = nn.ModuleDict()
        self.ffn_lora_on = False

    def _weights(self):
        gn1, conv1, lin, gn2, conv2 = self.in_layers[0], self.in_layers[2], self.emb_layers[1], self.out_layers[0], self.out_layers[3]
        ps = [gn1.weight, gn1.bias, conv1.weight, conv1.bias, lin.weight, lin.bias, gn2.weight, gn2.bias, conv2.weight, conv2.bias]
        skip = self.skip_connection if isinstance(self.skip_connection, nn.Conv2d) else None
        if skip is not None:
            ps += [skip.weight, skip.bias]
        key = _ver(*ps)
        if key != self._pack_key:
            with torch.no_grad():
                pk = {"gn1_w": _f32(ps[0]), "gn1_b": _f32(ps[1]), "w1": ops.pack_conv3x3_weight(ps[2]), "b1": _f32(ps[3]),
                      "w_emb": _bf16(ps[4]), "b_emb": _f32(ps[5]), "gn2_w": _f32(ps[6]), "gn2_b": _f32(ps[7]),
                      "w2": ops.pack_conv3x3_weight(ps[8]), "b2": _f32(ps[9])}
                if skip is not None:
                    pk["w_skip"] = ops.pack_conv3x3_weight(ps[10]) if self.use_conv else _bf16(ps[10].flatten(1))
                    pk["b_skip"] = _f32(ps[11])
            self._pack, self._pack_key = pk, key
        return self._pack

    def forward_tokens(self, t, emb, hw, emb_act=None):
        """t bf16 [B, h*w, C] (NHWC), emb [B, emb_channels] -> bf16 [B, h*w, out_channels].  ``emb_act`` = SiLU(emb) as
        bf16 when the caller has already evaluated it (the U-Net shares it between its ResBlocks)."""
        _no_grad_only("ResBlock (time embedding)", emb, emb_act)      # the time embedding is frozen: gradients flow through t only
        if self.training and self.dropout > 0:
            raise NotImplementedError("ResBlock: dropout > 0 in training mode is not built (the reference trains with dropout 0)")
        pk = self._weights()
        b = t.shape[0]
        gn1, gn2, conv1, conv2 = self.in_layers[0], self.out_layers[0], self.in_layers[2], self.out_layers[3]
        h = ag.groupnorm_act(t, pk["gn1_w"], pk["gn1_b"], gn1.num_groups, gn1.eps, True)                              # :240 in_layers[:-1]
        if emb_act is None:
            emb_act = ops.silu(emb.contiguous())
        emb_out = ops.proj(emb_act, pk["w_emb"], bias=pk["b_emb"], out_dtype=torch.float32)                           # :248
        lo = self.conv_loras if (self.ffn_lora_on and len(self.conv_loras)) else {}
        if "conv1" in lo:
            h = lo["conv1"].forward_tokens(h, hw, rowbias=emb_out)
        else:
            h = ag.conv3x3(h, pk, "w1", conv1.weight, hw, bias=pk["b1"], rowbias=emb_out)                             # :247 + :257
        h = ag.groupnorm_act(h, pk["gn2_w"], pk["gn2_b"], gn2.num_groups, gn2.eps, True)                              # :258
        if "w_skip" not in pk:
            skip = t
        elif "conv_shortcut" in lo:
            skip = lo["conv_shortcut"].forward_tokens(t, hw)
        elif self.use_conv:
            skip = ag.conv3x3(t, pk, "w_skip", self.skip_connection.weight, hw, bias=pk["b_skip"])
        elif ag.needs_grad(t):
            skip = ag.linear(t.view(b * hw[0] * hw[1], -1), pk, "w_skip", "b_skip").view(b, hw[0] * hw[1], -1)
        else:
            skip = ops.proj(t.view(b * hw[0] * hw[1], -1), pk["w_skip"], bias=pk["b_skip"]).view(b, hw[0] * hw[1], -1)
        if "conv2" in lo:
            return lo["conv2"].forward_tokens(h, hw, residual=skip)
        return ag.conv3x3(h, pk, "w2", conv2.weight, hw, bias=pk["b2"], residual=skip)                                # :258-260

    def forward(self, x, emb):
        if not x.is_cuda:
            raise RuntimeError("adaface_b200 ResBlock runs on CUDA only (no CPU fallback)")
        hw = tuple(x.shape[2:])
        return _to_nchw(self.forward_tokens(_to_tokens(x), emb, hw), hw, x.dtype)


class Upsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        if dims != 2 or padding != 1:
            raise NotImplementedError("Upsample: dims=2, padding=1 only")
        self.channels, self.out_channels, self.use_conv = channels, out_channels or channels, use_conv
        if use_conv:
            self.conv = nn.Conv2d(channels, self.out_channels, 3, padding=1)
        self._pack_key, self._pack = None, None

    def forward_tokens(self, t, hw):
        up = ag.upsample2x(t, hw)                                               # :116
        if not self.use_conv:
            return up
        key = _ver(self.conv.weight, self.conv.bias)
        if key != self._pack_key:
            self._pack, self._pack_key = {"w": ops.pack_conv3x3_weight(self.conv.weight), "b": _f32(self.conv.bias)}, key
        return ag.conv3x3(up, self._pack, "w", self.conv.weight, (2 * hw[0], 2 * hw[1]), bias=self._pack["b"])      # :118

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("adaface_b200 Upsample runs on CUDA only (no CPU fallback)")
        hw = tuple(x.shape[2:])
        return _to_nchw(self.forward_tokens(_to_tokens(x), hw), (2 * hw[0], 2 * hw[1]), x.dtype)


class Downsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        if dims != 2 or padding != 1 or not use_conv:
            raise NotImplementedError("Downsample: the SD-1.5 form (dims=2, use_conv, padding=1) only")
        self.channels, self.out_channels, self.use_conv = channels, out_channels or channels, use_conv
        self.op = nn.Conv2d(channels, self.out_channels, 3, stride=2, padding=1)
        self._pack_key, self._pack = None, None

    def forward_tokens(self, t, hw):
        key = _ver(self.op.weight, self.op.bias)
        if key != self._pack_key:
            self._pack, self._pack_key = {"w": ops.pack_conv3x3_weight(self.op.weight), "b": _f32(self.op.bias)}, key
        return ag.conv3x3(t, self._pack, "w", self.op.weight, hw, stride=2, bias=self._pack["b"])   # :160

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("adaface_b200 Downsample runs on CUDA only (no CPU fallback)")
        hw = tuple(x.shape[2:])
        return _to_nchw(self.forward_tokens(_to_tokens(x), hw), (hw[0] // 2, hw[1] // 2), x.dtype)


# ------------------------------------------------------------------------------------------------ conv-LoRA (dalc:541-591)
class _ConvMagnitude(nn.Module):
    def __init__(self, mag):
        super().__init__()
        self.weight = nn.Parameter(mag)


class LoraDoraConv2d(nn.Module):
    """Parameter container with peft's ``lora.Conv2d(use_dora=True)`` layout for a 3x3 (or 1x1) base convolution -- the
    ``up_blocks.3.resnets.[12].conv1 / conv2 / conv_shortcut`` adapters the reference installs with
    ``set_up_ffn_loras`` (adaface/diffusers_attn_lora_capture.py:541-591; r = 192, alpha = 16, DoRA):
    ``base_layer``, ``lora_A[adapter]`` = Conv2d(cin, r, k, padding, bias=False), ``lora_B[adapter]`` = Conv2d(r, cout, 1,
    bias=False), ``lora_magnitude_vector[adapter].weight`` [cout].  Init as peft: A Kaiming-uniform(a = sqrt 5), B zero,
    magnitude = ||W|| per output channel (identity adapter).  Eval-mode arithmetic (convolution form):
        y = bias + m / ||W + s B.A||_(cin,kh,kw) * (conv(x, W) + s conv1x1(conv(x, A), B))
    runs as two launches of the implicit-GEMM kernel: T = conv(x, A) (Cout = r), then the base convolution with the rank-r
    tail T (sB)^T accumulated into the same TMEM tile and the DoRA column scale + bias in its epilogue.  Training: ConvLoraFn
    (autograd.py) pairs it with dX, dA (through an im2col of x), dB and d magnitude."""

    def __init__(self, base_layer, adapter_name="default", r=192, lora_alpha=16, use_dora=True, lora_dropout=0.1):
        super().__init__()
        if not use_dora:
            raise NotImplementedError("the reference always uses DoRA (lora_uses_dora=True)")
        k = base_layer.kernel_size
        if k not in ((3, 3), (1, 1)) or base_layer.stride != (1, 1) or base_layer.padding != (k[0] // 2, k[0] // 2):
            raise NotImplementedError("LoraDoraConv2d: 3x3 (padding 1) or 1x1 base convolutions with stride 1 only")
        self.base_layer = base_layer
        self.r, self.lora_alpha, self.scaling, self.adapter = r, lora_alpha, lora_alpha / r, adapter_name
        dev = base_layer.weight.device
        cin, cout = base_layer.in_channels, base_layer.out_channels
        self.lora_A = nn.ModuleDict({adapter_name: nn.Conv2d(cin, r, k, padding=base_layer.padding, bias=False, device=dev)})
        self.lora_B = nn.ModuleDict({adapter_name: nn.Conv2d(r, cout, 1, bias=False, device=dev)})
        nn.init.kaiming_uniform_(self.lora_A[adapter_name].weight, a=5 ** 0.5)
        nn.init.zeros_(self.lora_B[adapter_name].weight)
        self.lora_magnitude_vector = nn.ModuleDict({adapter_name: _ConvMagnitude(torch.linalg.norm(base_layer.weight.detach().float().flatten(1), dim=1))})
        self._pack_key, self._pack = None, None
        self.enable_adapters = lambda *a, **k_: None
        self.set_adapter = lambda *a, **k_: None

    def add_adapter(self, adapter_name):
        """A further adapter on the same base convolution (the reference keeps 'recon_loss', 'unet_distill' and 'comp_distill'
        sets side by side, dalc:553-556); same init as the first."""
        if adapter_name in self.lora_A:
            return
        first = self.adapter
        a0, b0 = self.lora_A[first], self.lora_B[first]
        dev = a0.weight.device
        self.lora_A[adapter_name] = nn.Conv2d(a0.in_channels, a0.out_channels, a0.kernel_size, padding=a0.padding, bias=False, device=dev)
        self.lora_B[adapter_name] = nn.Conv2d(b0.in_channels, b0.out_channels, 1, bias=False, device=dev)
        nn.init.kaiming_uniform_(self.lora_A[adapter_name].weight, a=5 ** 0.5)
        nn.init.zeros_(self.lora_B[adapter_name].weight)
        self.lora_magnitude_vector[adapter_name] = _ConvMagnitude(torch.linalg.norm(self.base_layer.weight.detach().float().flatten(1), dim=1))

    def invalidate(self):
        """Force the next pack() to rebuild from the live parameters (see LoraDoraLinear.invalidate)."""
        self._pack_key = None

    def set_active_adapter(self, adapter_name):
        if adapter_name not in self.lora_A:
            raise KeyError(f"LoraDoraConv2d: unknown adapter {adapter_name!r} (have {list(self.lora_A)})")
        if adapter_name != self.adapter:
            self.adapter, self._pack_key = adapter_name, None

    @property
    def is_3x3(self):
        return self.base_layer.kernel_size == (3, 3)

    def pack(self):
        """(W, A, s*B [cout, r], colscale = m / ||W + s B.A||, bias): W / A in the kernel's K-major tap layout for 3x3, plain
        [cout, cin] / [r, cin] for 1x1.  Rebuilt only when a parameter changed."""
        A, B = self.lora_A[self.adapter].weight, self.lora_B[self.adapter].weight
        m, W, b = self.lora_magnitude_vector[self.adapter].weight, self.base_layer.weight, self.base_layer.bias
        key = _ver(A, B, m, W, b)
        if key != self._pack_key:
            with torch.no_grad():
                Bf = B.detach().float().flatten(1)                                           # [cout, r]
                # ||W + s B.A|| over (cin, kh, kw) with B.A on the projection GEMM (no library arithmetic, graph-capturable)
                Wf = W.detach().flatten(1)
                cs = ops.dora_colscale((Wf if Wf.dtype in (torch.float32, torch.bfloat16) else Wf.float()).contiguous(),
                                       _bf16(A.flatten(1)), _bf16(Bf), self.scaling, m)
                if self.is_3x3:
                    wp, ap = ops.pack_conv3x3_weight(W), ops.pack_conv3x3_weight(A)
                else:
                    wp, ap = _bf16(W.flatten(1)), _bf16(A.flatten(1))
                self._pack = (wp, ap, (Bf * self.scaling).to(torch.bfloat16).contiguous(), cs, _f32(b))
            self._pack_key = key
        return self._pack

    def forward_tokens(self, t, hw, rowbias=None, residual=None):
        """t bf16 [B, h*w, cin] -> bf16 [B, h*w, cout]; ``rowbias`` / ``residual`` as in ops.conv3x3 (3x3 only / both)."""
        b, n, _ = t.shape
        ad = self.adapter
        A, B, m = self.lora_A[ad].weight, self.lora_B[ad].weight, self.lora_magnitude_vector[ad].weight
        if torch.is_grad_enabled() and (t.requires_grad or A.requires_grad or B.requires_grad or m.requires_grad):
            # training (eval-mode arithmetic: the U-Net and its adapters stay in .eval(), ddpm.py:637-638, so lora_dropout is
            # inactive): the adapter-fused convolution with its backward; the per-image bias / skip are added by autograd ops
            y = ag.ConvLoraFn.apply(t if t.dtype == torch.bfloat16 else t.to(torch.bfloat16), self, A, B, m, hw)
            if rowbias is not None:
                y = y + rowbias.to(y.dtype)[:, None, :]
            if residual is not None:
                y = y + residual
            return y
        wp, ap, bs, cs, bias = self.pack()
        if self.is_3x3:
            ta = ops.conv3x3(t, ap, hw)
            return ops.conv3x3(t, wp, hw, bias=bias, rowbias=rowbias, residual=residual, t=ta.view(b * n, -1), bs=bs, colscale=cs)
        if rowbias is not None:
            raise NotImplementedError("LoraDoraConv2d: rowbias is only defined for the 3x3 form")
        x2d = t.view(b * n, -1)
        ta = ops.proj(x2d, ap)
        res2d = None if residual is None else residual.view(b * n, -1)
        return ops.proj(x2d, wp, t=ta, bs=bs, colscale=cs, bias=bias, residual=res2d).view(b, n, -1)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("adaface_b200 LoraDoraConv2d runs on CUDA only (no CPU fallback)")
        hw = tuple(x.shape[2:])
        return _to_nchw(self.forward_tokens(_to_tokens(x), hw), hw, x.dtype)
