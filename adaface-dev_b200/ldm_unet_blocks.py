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
in, one out); ``forward_tokens`` is the NHWC-resident entry.  Forward only; CUDA only, no fallback.
"""
import torch
import torch.nn as nn

from . import ops
from .ldm_attention import _bf16, _f32, _ver


def _to_tokens(x):
    b, c, h, w = x.shape
    if x.dtype not in (torch.bfloat16, torch.float32):
        x = x.float()
    return ops.transpose(x.contiguous().view(b, c, h * w), out_dtype=torch.bfloat16)          # [B, h*w, C]


def _to_nchw(t, hw, dtype):
    b, _, c = t.shape
    if dtype not in (torch.bfloat16, torch.float32):            # fp16 callers (the reference's autocast): through fp32
        return ops.transpose(t, out_dtype=torch.float32).view(b, c, *hw).to(dtype)
    return ops.transpose(t, out_dtype=dtype).view(b, c, *hw)


def _no_grad_only(name, *ts):
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts):
        raise NotImplementedError(f"{name}: backward of the frozen U-Net's convolutional blocks is not built yet")


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
        _no_grad_only("ResBlock", t, emb)
        if self.training and self.dropout > 0:
            raise NotImplementedError("ResBlock: dropout > 0 in training mode is not built (the reference trains with dropout 0)")
        pk = self._weights()
        b = t.shape[0]
        gn1, gn2 = self.in_layers[0], self.out_layers[0]
        h = ops.groupnorm_act_tokens(t, pk["gn1_w"], pk["gn1_b"], gn1.num_groups, gn1.eps, silu=True)                 # :240 in_layers[:-1]
        if emb_act is None:
            emb_act = ops.silu(emb.contiguous())
        emb_out = ops.proj(emb_act, pk["w_emb"], bias=pk["b_emb"], out_dtype=torch.float32)                           # :248
        h = ops.conv3x3(h, pk["w1"], hw, bias=pk["b1"], rowbias=emb_out)                                              # :247 + :257
        h = ops.groupnorm_act_tokens(h, pk["gn2_w"], pk["gn2_b"], gn2.num_groups, gn2.eps, silu=True)                 # :258
        if "w_skip" not in pk:
            skip = t
        elif self.use_conv:
            skip = ops.conv3x3(t, pk["w_skip"], hw, bias=pk["b_skip"])
        else:
            skip = ops.proj(t.view(b * hw[0] * hw[1], -1), pk["w_skip"], bias=pk["b_skip"]).view(b, hw[0] * hw[1], -1)
        return ops.conv3x3(h, pk["w2"], hw, bias=pk["b2"], residual=skip)                                             # :258-260

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
        _no_grad_only("Upsample", t)
        up = ops.upsample2x_tokens(t, hw)                                       # :116
        if not self.use_conv:
            return up
        key = _ver(self.conv.weight, self.conv.bias)
        if key != self._pack_key:
            self._pack, self._pack_key = (ops.pack_conv3x3_weight(self.conv.weight), _f32(self.conv.bias)), key
        return ops.conv3x3(up, self._pack[0], (2 * hw[0], 2 * hw[1]), bias=self._pack[1])      # :118

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
        _no_grad_only("Downsample", t)
        key = _ver(self.op.weight, self.op.bias)
        if key != self._pack_key:
            self._pack, self._pack_key = (ops.pack_conv3x3_weight(self.op.weight), _f32(self.op.bias)), key
        return ops.conv3x3(t, self._pack[0], hw, stride=2, bias=self._pack[1])   # :160

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("adaface_b200 Downsample runs on CUDA only (no CPU fallback)")
        hw = tuple(x.shape[2:])
        return _to_nchw(self.forward_tokens(_to_tokens(x), hw), (hw[0] // 2, hw[1] // 2), x.dtype)
