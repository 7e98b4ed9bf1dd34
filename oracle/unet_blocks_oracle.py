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


# ----------------------------------------------------------------------------------------------
# The whole U-Net (openaimodel.py:414-960), driven by a reference-format state dict.
def timestep_embedding(timesteps, dim, max_period=10000):
    """ldm/modules/diffusionmodules/util.py:154-174 (repeat_only=False)."""
    import math
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def unet_layout(cfg):
    """Layer kinds of input_blocks / middle_block / output_blocks as UNetModel.__init__ builds them (:520-684) for
    use_spatial_transformer=True, conv_resample=True, resblock_updown=False: lists of 'conv_in' | 'res' | 'attn' | 'down' | 'up'."""
    nrb, mults, ares = cfg["num_res_blocks"], list(cfg["channel_mult"]), set(cfg["attention_resolutions"])
    inp, ds = [["conv_in"]], 1
    for level in range(len(mults)):
        for _ in range(nrb):
            inp.append(["res", "attn"] if ds in ares else ["res"])
        if level != len(mults) - 1:
            inp.append(["down"])
            ds *= 2
    out = []
    for level in reversed(range(len(mults))):
        for i in range(nrb + 1):
            layers = ["res", "attn"] if ds in ares else ["res"]
            if level and i == nrb:
                layers.append("up")
                ds //= 2
            out.append(layers)
    return inp, ["res", "attn", "res"], out


def _res_weights(sd, p):
    w = {"gn1_w": sd[p + "in_layers.0.weight"], "gn1_b": sd[p + "in_layers.0.bias"], "conv1_w": sd[p + "in_layers.2.weight"],
         "conv1_b": sd[p + "in_layers.2.bias"], "emb_w": sd[p + "emb_layers.1.weight"], "emb_b": sd[p + "emb_layers.1.bias"],
         "gn2_w": sd[p + "out_layers.0.weight"], "gn2_b": sd[p + "out_layers.0.bias"], "conv2_w": sd[p + "out_layers.3.weight"],
         "conv2_b": sd[p + "out_layers.3.bias"]}
    if p + "skip_connection.weight" in sd:
        w["skip_w"], w["skip_b"] = sd[p + "skip_connection.weight"], sd[p + "skip_connection.bias"]
    return w


def _spatial_weights(sd, p):
    b = p + "transformer_blocks.0."
    w = {"gn_w": sd[p + "norm.weight"], "gn_b": sd[p + "norm.bias"], "proj_in_w": sd[p + "proj_in.weight"][:, :, 0, 0],
         "proj_in_b": sd[p + "proj_in.bias"], "proj_out_w": sd[p + "proj_out.weight"][:, :, 0, 0], "proj_out_b": sd[p + "proj_out.bias"],
         "ff_proj_w": sd[b + "ff.net.0.proj.weight"], "ff_proj_b": sd[b + "ff.net.0.proj.bias"], "ff_out_w": sd[b + "ff.net.2.weight"],
         "ff_out_b": sd[b + "ff.net.2.bias"]}
    for i in (1, 2, 3):
        w[f"norm{i}_w"], w[f"norm{i}_b"] = sd[b + f"norm{i}.weight"], sd[b + f"norm{i}.bias"]
    for a in ("attn1", "attn2"):
        w[a] = {"to_q": sd[b + a + ".to_q.weight"], "to_k": sd[b + a + ".to_k.weight"], "to_v": sd[b + a + ".to_v.weight"],
                "to_out_w": sd[b + a + ".to_out.0.weight"], "to_out_b": sd[b + a + ".to_out.0.bias"]}
    return w


def unet_forward(sd, cfg, x, timesteps, context, mask=None, res_hidden_states_gradscale=1.0):
    """UNetModel.forward (openaimodel.py:820-960) without capture: time embedding -> input blocks (skips pushed) -> middle
    block -> output blocks (skip popped and concatenated on the channel axis, :925) -> out (norm, SiLU, conv).
    res_hidden_states_gradscale: the ScaleGrad the training wrapper puts on the skip tensors entering diffusers up_blocks[1:]
    (CrossAttnUpBlock2D_forward_capture, adaface/diffusers_attn_lora_capture.py:382-394; = output blocks from index
    num_res_blocks + 1 on): identity forward, gradient times the factor."""
    from .attn_oracle import spatial_transformer, scale_grad
    heads = cfg["num_heads"]

    def run(layers, prefix, h, emb):
        for j, kind in enumerate(layers):
            p = f"{prefix}{j}."
            if kind == "conv_in":
                h = conv3x3(h, sd[p + "weight"], sd[p + "bias"])
            elif kind == "res":
                h = res_block(_res_weights(sd, p), h, emb)
            elif kind == "attn":
                h = spatial_transformer(_spatial_weights(sd, p), h, context=context, mask=mask, heads=heads)
            elif kind == "down":
                h = downsample({"conv_w": sd[p + "op.weight"], "conv_b": sd[p + "op.bias"]}, h)
            else:
                h = upsample({"conv_w": sd[p + "conv.weight"], "conv_b": sd[p + "conv.bias"]}, h)
        return h

    inp, mid, out = unet_layout(cfg)
    t_emb = timestep_embedding(timesteps, cfg["model_channels"])
    emb = F.linear(silu(F.linear(t_emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    h, hs = x, []
    for i, layers in enumerate(inp):
        h = run(layers, f"input_blocks.{i}.", h, emb)
        hs.append(h)
    h = run(mid, "middle_block.", h, emb)
    first_scaled = cfg["num_res_blocks"] + 1
    for i, layers in enumerate(out):
        skip = hs.pop()
        if i >= first_scaled and res_hidden_states_gradscale != 1.0:
            skip = scale_grad(skip, res_hidden_states_gradscale)
        h = run(layers, f"output_blocks.{i}.", torch.cat([h, skip], dim=1), emb)
    return conv3x3(silu(group_norm32(h, sd["out.0.weight"], sd["out.0.bias"])), sd["out.2.weight"], sd["out.2.bias"])


def lora_dora_conv(x, W, b, A, B, m, scaling):
    """peft ``lora.Conv2d(use_dora=True)`` in eval mode (PARITY UNPINNED: peft is un-vendored; restated from its published
    algorithm, anchored on the reference's call site adaface/diffusers_attn_lora_capture.py:541-591):
        y = bias + m / ||W + s B.A|| * (conv(x, W) + s conv1x1(conv(x, A), B)),  norm over (cin, kh, kw), detached.
    W [cout, cin, k, k], A [r, cin, k, k], B [cout, r, 1, 1], m [cout]; k = 3 (padding 1) or 1."""
    k = W.shape[-1]
    conv = (lambda t, w_: conv3x3(t, w_, None)) if k == 3 else (lambda t, w_: torch.einsum("bchw,oc->bohw", t, w_[:, :, 0, 0]))
    lora = torch.einsum("bchw,oc->bohw", conv(x, A), B[:, :, 0, 0]) * scaling
    comp = (B.flatten(1) @ A.flatten(1)).view_as(W)
    wn = torch.linalg.norm((W + scaling * comp).flatten(1), dim=1).detach()
    y = (m / wn)[None, :, None, None] * (conv(x, W) + lora)
    return y if b is None else y + b[None, :, None, None]
