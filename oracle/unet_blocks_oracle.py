"""CPU oracle (torch fp32) for the convolutional blocks of the SD-1.5 U-Net around the attention path.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Functional restatement of
  * ldm/modules/diffusionmodules/openaimodel.py:164-260  (ResBlock.__init__ / _forward, use_scale_shift_norm=False,
    up = down = False: the SD-1.5 configuration)
  * ldm/modules/diffusionmodules/openaimodel.py:92-119   (Upsample: nearest 2x then 3x3 convolution)
  * ldm/modules/diffusionmodules/openaimodel.py:135-161  (Downsample with use_conv: 3x3 convolution, stride 2, padding 1)
  * ldm/modules/diffusionmodules/util.py normalization() = GroupNorm32(32, C): statistics in fp32, eps 1e-5
Pinned against the reference's own modules (tests/golden/make_golden.py, family "unet_blocks").
All tensors are fp32 CPU tensors; weights come in as plain dicts.
"""
import torch
import torch.nn.functional as F


def conv3x3(x, w, b, stride=1):
    """conv_nd(2, cin, cout, 3, padding=1[, stride]) written out as nine shifted channel products, so that the oracle does
    not lean on the library convolution it is used to check (x [B, Cin, H, W], w [Cout, Cin, 3, 3])."""
    B, Cin, H, W = x.shape
    Ho, Wo = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
    xp = F.pad(x, (1, 1, 1, 1))
    out = torch.zeros(B, w.shape[0], Ho, Wo, dtype=x.dtype)
    for ky in range(3):
        for kx in range(3):
            win = xp[:, :, ky:ky + stride * (Ho - 1) + 1:stride, kx:kx + stride * (Wo - 1) + 1:stride]
            out += torch.einsum("bchw,oc->bohw", win, w[:, :, ky, kx])
    return out if b is None else out + b[None, :, None, None]


def group_norm32(x, gamma, beta, eps=1e-5, groups=32):
    """GroupNorm32.forward (util.py): super().forward(x.float()).type(x.dtype), written out."""
    B, C = x.shape[:2]
    xg = x.float().reshape(B, groups, -1)
    mean = xg.mean(dim=2, keepdim=True)
    var = xg.var(dim=2, unbiased=False, keepdim=True)
    xn = ((xg - mean) / torch.sqrt(var + eps)).reshape(x.shape)
    shape = (1, C) + (1,) * (x.dim() - 2)
    return xn * gamma.reshape(shape) + beta.reshape(shape)


def silu(x):
    return x * torch.sigmoid(x)


def res_block(w, x, emb):
    """ResBlock._forward (openaimodel.py:239-260): in_layers (norm, SiLU, conv) -> + emb_layers(emb)[..., None, None]
    -> out_layers (norm, SiLU, dropout (eval: identity), conv) -> skip_connection(x) + h.
    ``w``: gn1_w/b, conv1_w/b, emb_w/b, gn2_w/b, conv2_w/b and, when the channel count changes, skip_w/b
    (1x1, or 3x3 with use_conv)."""
    h = conv3x3(silu(group_norm32(x, w["gn1_w"], w["gn1_b"])), w["conv1_w"], w["conv1_b"])
    emb_out = F.linear(silu(emb), w["emb_w"], w["emb_b"])
    h = h + emb_out[:, :, None, None]
    h = conv3x3(silu(group_norm32(h, w["gn2_w"], w["gn2_b"])), w["conv2_w"], w["conv2_b"])
    if "skip_w" not in w:
        skip = x
    elif w["skip_w"].shape[-1] == 3:
        skip = conv3x3(x, w["skip_w"], w["skip_b"])
    else:
        skip = torch.einsum("bchw,oc->bohw", x, w["skip_w"][:, :, 0, 0]) + w["skip_b"][None, :, None, None]
    return skip + h


def upsample(w, x):
    """Upsample.forward (openaimodel.py:109-119), dims = 2, use_conv."""
    x = x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)        # F.interpolate(scale_factor=2, mode="nearest")
    return conv3x3(x, w["conv_w"], w["conv_b"])


def downsample(w, x):
    """Downsample.forward (openaimodel.py:159-161), use_conv: 3x3, stride 2, padding 1."""
    return conv3x3(x, w["conv_w"], w["conv_b"], stride=2)
