"""Drop-in mirror of the reference's LDM U-Net (the caller of the attention / ResBlock path).

    UNetModel                 ldm/modules/diffusionmodules/openaimodel.py:414-960  (the SD-1.5 configuration:
                              use_spatial_transformer, conv_resample, no class labels, no scale-shift norm, no resblock_updown)
    TimestepEmbedSequential   ldm/modules/diffusionmodules/openaimodel.py:73-89

Module / parameter names are the reference's (time_embed.0/2, input_blocks.N.M, middle_block.M, output_blocks.N.M, out.0/2),
so the ``model.diffusion_model.*`` weights of an SD-1.5 LDM checkpoint load with ``load_state_dict``.

B200 design: the whole forward runs on NHWC bf16 tokens [B, h*w, C].  The 4-channel latent is transposed (and padded to
8 channels) once on entry and the prediction transposed back once on exit; in between every 3x3 convolution is the
implicit-GEMM tcgen05 kernel, every 1x1 convolution / Linear the projection GEMM, GroupNorm(+SiLU) one pass over tokens, and
the transformer blocks consume / produce the same layout.  SiLU(emb) is evaluated once per forward instead of once per
ResBlock.  The skip concatenations are the only torch data movement left (torch.cat along the channel axis).
Training: the forward is differentiable w.r.t. the prompt context (every block pairs its kernels with backward kernels through
autograd.py; the U-Net's own weights are frozen and get no gradient).  CUDA only, no fallback.
"""
import torch
import torch.nn as nn

from . import ops
from . import autograd as ag
from .attn_processor import gen_gradient_scaler
from .ldm_attention import SpatialTransformer, _bf16, _f32, _ver
from .ldm_unet_blocks import ResBlock, Upsample, Downsample, _to_nchw


class TimestepEmbedSequential(nn.Sequential):
    """Container only (openaimodel.py:73-89): UNetModel.forward walks the layers itself so that activations stay NHWC."""

    def forward(self, x, emb, context=None, mask=None):
        for layer in self:
            if isinstance(layer, ResBlock):
                x = layer(x, emb)
            elif isinstance(layer, SpatialTransformer):
                x = layer(x, context, mask=mask)
            else:
                x = layer(x)
        return x


class UNetModel(nn.Module):
    captured_layer_indices = (22, 23, 24)          # openaimodel.py:853

    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, dropout=0,
                 channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None, use_checkpoint=False, use_fp16=False,
                 num_heads=-1, num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False, resblock_updown=False,
                 use_new_attention_order=False, use_spatial_transformer=False, transformer_depth=1, context_dim=None, n_embed=None,
                 legacy=True):
        super().__init__()
        if (dims != 2 or num_classes is not None or use_scale_shift_norm or resblock_updown or not use_spatial_transformer
                or not conv_resample or n_embed is not None or num_heads == -1 or context_dim is None):
            raise NotImplementedError("UNetModel: only the SD-1.5 configuration of the reference is built (spatial transformers "
                                      "with num_heads, learned up / down-sampling, no class labels)")
        self.in_channels, self.model_channels, self.out_channels = in_channels, model_channels, out_channels
        self.num_res_blocks, self.attention_resolutions, self.channel_mult = num_res_blocks, tuple(attention_resolutions), tuple(channel_mult)
        self.num_heads = num_heads
        ted = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, ted), nn.SiLU(), nn.Linear(ted, ted))

        def st(ch):
            return SpatialTransformer(ch, num_heads, ch // num_heads, depth=transformer_depth, context_dim=context_dim)

        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(nn.Conv2d(in_channels, model_channels, 3, padding=1))])
        chans, ch, ds = [model_channels], model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [ResBlock(ch, ted, dropout, out_channels=mult * model_channels)]
                ch = mult * model_channels
                if ds in self.attention_resolutions:
                    layers.append(st(ch))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, True, out_channels=ch)))
                chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(ResBlock(ch, ted, dropout), st(ch), ResBlock(ch, ted, dropout))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [ResBlock(ch + chans.pop(), ted, dropout, out_channels=model_channels * mult)]
                ch = model_channels * mult
                if ds in self.attention_resolutions:
                    layers.append(st(ch))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, True, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(nn.GroupNorm(32, ch), nn.SiLU(), nn.Conv2d(model_channels, out_channels, 3, padding=1))
        for p in self.out[2].parameters():                         # zero_module (openaimodel.py:690)
            nn.init.zeros_(p)
        self._pack_key, self._pack = None, None

    def load_ldm_state_dict(self, state_dict, prefix="model.diffusion_model.", strict=True):
        """Load the U-Net part of a full LDM / SD-1.5 checkpoint state dict (``LatentDiffusion`` keeps the U-Net under
        ``model.diffusion_model.``, ldm/models/diffusion/ddpm.py): keys outside ``prefix`` are ignored."""
        sd = {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}
        if not sd:
            raise KeyError(f"UNetModel.load_ldm_state_dict: no key starts with {prefix!r}")
        return self.load_state_dict(sd, strict=strict)

    def _weights(self):
        conv_in, l0, l2, gn, conv_out = self.input_blocks[0][0], self.time_embed[0], self.time_embed[2], self.out[0], self.out[2]
        ps = (conv_in.weight, conv_in.bias, l0.weight, l0.bias, l2.weight, l2.bias, gn.weight, gn.bias, conv_out.weight, conv_out.bias)
        key = _ver(*ps)
        if key != self._pack_key:
            with torch.no_grad():
                self._pack = {"w_in": ops.pack_conv3x3_weight(ps[0]), "b_in": _f32(ps[1]), "w_t0": _bf16(ps[2]), "b_t0": _f32(ps[3]),
                              "w_t2": _bf16(ps[4]), "b_t2": _f32(ps[5]), "gn_w": _f32(ps[6]), "gn_b": _f32(ps[7]),
                              "w_out": ops.pack_conv3x3_weight(ps[8]), "b_out": _f32(ps[9])}
            self._pack_key = key
        return self._pack

    def _cross_attn(self, layer_idx):
        n_in = len(self.input_blocks)
        block = (self.input_blocks[layer_idx] if layer_idx < n_in else self.middle_block if layer_idx == n_in
                 else self.output_blocks[layer_idx - n_in - 1])
        return block[1].transformer_blocks[0].attn2

    def _processor_modules(self):
        out = []
        for m in self.modules():
            if isinstance(m, SpatialTransformer):
                for blk in m.transformer_blocks:
                    out += [ca for ca in (blk.attn1, blk.attn2) if ca.processor is not None]
        return out

    def _run(self, block, h, hw, emb, emb_act, context, mask):
        for layer in block:
            if isinstance(layer, ResBlock):
                h = layer.forward_tokens(h, emb, hw, emb_act=emb_act)
            elif isinstance(layer, SpatialTransformer):
                h = layer.forward_tokens(h, hw, context=context, mask=mask)
            elif isinstance(layer, Downsample):
                h = layer.forward_tokens(h, hw)
                hw = (hw[0] // 2, hw[1] // 2)
            elif isinstance(layer, Upsample):
                h = layer.forward_tokens(h, hw)
                hw = (hw[0] * 2, hw[1] * 2)
            else:
                raise RuntimeError(f"UNetModel: unexpected layer {type(layer).__name__}")
        return h, hw

    def forward(self, x, timesteps=None, context=None, y=None, context_in=None, extra_info=None, **kwargs):
        """openaimodel.py:820-960.  x [B, in_channels, h, w], timesteps [B], context [B, S, context_dim] -> [B, out_channels, h, w]
        in x.dtype.  extra_info: 'img_mask' (self-attention key mask) and 'capture_ca_activations' (layers 22-24) as in the
        reference; the captured maps come back in extra_info['ca_layers_activations']."""
        if not x.is_cuda:
            raise RuntimeError("adaface_b200 UNetModel runs on CUDA only (no CPU fallback)")
        if y is not None:
            raise NotImplementedError("UNetModel: class-conditional models are not built")
        if torch.is_grad_enabled() and x.requires_grad:
            raise NotImplementedError("UNetModel: gradients w.r.t. the latent input are not built (stage 2 back-propagates into "
                                      "the prompt context only, ddpm.py:1645-1707); pass x.detach()")
        pk = self._weights()
        B, cin, H, W = x.shape
        hw = (H, W)
        capture = bool(extra_info.get("capture_ca_activations", False)) if extra_info is not None else False
        mask = extra_info.get("img_mask", None) if extra_info is not None else None
        captured = self.captured_layer_indices if capture else ()
        for li in captured:
            if self._cross_attn(li).processor is None:
                self._cross_attn(li).save_cross_attn_vars = True
        # modules routed through an installed AttnProcessor_LoRA_Capture (unet_wrapper.set_up_attn_processors) get this call's
        # cross_attention_kwargs (ddpm.py:4227-4229)
        subj_indices = extra_info.get("subj_indices", None) if extra_info is not None else None
        for ca in self._processor_modules():
            ca.processor_kwargs = {"img_mask": mask, "subj_indices": subj_indices}
        # A7 (dalc:382-394): gradient scale on the skip tensors entering diffusers up_blocks[1:] = output_blocks[num_res_blocks + 1:]
        gradscale = float(extra_info.get("res_hidden_states_gradscale", 1)) if extra_info is not None else 1.0
        cache = self.__dict__.setdefault("_grad_scalers", {})      # one scaler per factor, its alpha moved to the device once
        res_grad_scaler = cache.get(gradscale)                     # (a fresh one per call would stage alpha through the host every time)
        if res_grad_scaler is None:
            res_grad_scaler = cache[gradscale] = gen_gradient_scaler(gradscale)
        acts = {}

        def grab(li, h, hw_):
            if li in captured:
                ca = self._cross_attn(li)
                if ca.processor is not None:               # diffusers surface: the processor's cache (q, q2, k, v, attn, ...; dalc:344-362)
                    acts[li] = dict(ca.processor.cached_activations)
                else:
                    acts[li] = ca.cached_activations
                    ca.cached_activations = None
                acts[li]["outfeat"] = _to_nchw(h, hw_, x.dtype)                   # dalc:438-440 / openaimodel.py:933

        try:
            t_emb = ops.timestep_embedding(timesteps, self.model_channels)                                          # :839
            e = ops.silu(ops.proj(t_emb, pk["w_t0"], bias=pk["b_t0"], out_dtype=torch.float32))
            emb = ops.proj(e, pk["w_t2"], bias=pk["b_t2"], out_dtype=torch.float32)                                 # :840
            emb_act = ops.silu(emb)                       # nn.SiLU() of every ResBlock's emb_layers, evaluated once
            xin = x if x.dtype in (torch.bfloat16, torch.float32) else x.float()
            h = ops.transpose(xin.contiguous().view(B, cin, H * W), out_dtype=torch.bfloat16, pad_to=8)             # NHWC, channels padded to 8
            h = ops.conv3x3(h, pk["w_in"], hw, bias=pk["b_in"])
            hs, layer_idx = [h], 1
            for block in list(self.input_blocks)[1:]:
                h, hw = self._run(block, h, hw, emb, emb_act, context, mask)
                hs.append(h)
                grab(layer_idx, h, hw)
                layer_idx += 1
            h, hw = self._run(self.middle_block, h, hw, emb, emb_act, context, mask)
            grab(layer_idx, h, hw)
            layer_idx += 1
            for bi, block in enumerate(self.output_blocks):
                hw_in = hw
                skip = hs.pop()
                if bi >= self.num_res_blocks + 1 and gradscale != 1.0:      # diffusers up_blocks[1:]
                    skip = res_grad_scaler(skip)
                h = torch.cat([h, skip], dim=2)                                                                    # :925
                h, hw = self._run(block, h, hw_in, emb, emb_act, context, mask)
                if layer_idx in captured:
                    # the captured feature map is the block's output before any trailing Upsample only when there is none:
                    # layers 22-24 of SD-1.5 have no Upsample, so h is the SpatialTransformer output as in the reference
                    grab(layer_idx, h, hw)
                layer_idx += 1
        finally:
            for li in captured:
                self._cross_attn(li).save_cross_attn_vars = False
        if capture:                                                                                                 # :937-941
            keys = ("outfeat", "attn", "attnscore", "q", "attn_out")
            if any(self._cross_attn(li).processor is not None for li in captured):
                keys = ("outfeat", "attn", "attnscore", "q", "q2", "k", "v", "attn_out", "attn_subj", "attn_subj_sum", "attn_sqdiff")
            extra_info["ca_layers_activations"] = {key: {li: acts[li][key] for li in acts if acts[li].get(key) is not None}
                                                   for key in keys}
        gn = self.out[0]
        h = ag.groupnorm_act(h, pk["gn_w"], pk["gn_b"], gn.num_groups, gn.eps, True)                                 # :960
        o = ag.conv3x3(h, pk, "w_out", self.out[2].weight, hw, bias=pk["b_out"], out_dtype=torch.float32)
        return _to_nchw(o, hw, x.dtype)
