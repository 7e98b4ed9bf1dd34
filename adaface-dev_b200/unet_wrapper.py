"""Host glue of the training / LDM-sampling caller of the attention path, over the U-Net mirror.

    set_up_attn_processors        adaface/diffusers_attn_lora_capture.py:451-538
    set_up_ffn_loras              adaface/diffusers_attn_lora_capture.py:541-591   (conv-LoRA: up_blocks.3.resnets.[12].conv*)
    set_lora_and_capture_flags    adaface/diffusers_attn_lora_capture.py:593-629
    get_captured_activations      adaface/diffusers_attn_lora_capture.py:631-661
    DiffusersUNetWrapper          ldm/models/diffusion/ddpm.py:4060-4252           (setup_hooks_and_loras, forward)
    (CrossAttnUpBlock2D_forward_capture, dalc:366-446: the skip-tensor gradient scale and the per-pair output-feature capture
     live inside UNetModel.forward of the mirror -- extra_info['res_hidden_states_gradscale'], 'outfeat')

The reference drives a *diffusers* UNet2DConditionModel (un-vendored, absent here); the mirror underneath is the LDM-surface
``UNetModel`` (same network, LDM module names, openaimodel.py:414-960).  This file therefore also carries the name map between
the two surfaces, so that the flat parameter names the reference's checkpoints use for the trainable set
(``up_blocks_3_attentions_1_transformer_blocks_0_attn2_processor_to_q_lora_lora_A`` ...) are the ones exposed here.
Differences, both deliberate: (1) processors are installed on the three CAPTURED cross-attention modules only -- every other
attention module keeps the fused LDM path, whose arithmetic is the plain processor's (an ``img_mask`` follows the processor's
drop-if-any-instance-is-empty rule, ``CrossAttention.mask_mode = 'processor'``); (2) no autocast: activations are bf16 inside.
"""
import re

import torch
import torch.nn as nn

from .attn_processor import AttnProcessor_LoRA_Capture, LoraDoraLinear
from .ldm_attention import SpatialTransformer
from .ldm_unet_blocks import ResBlock, LoraDoraConv2d


def diffusers_module_names(unet):
    """{diffusers-style name: module} for every attention module and every ResBlock of the LDM-surface mirror, e.g.
    'up_blocks.3.attentions.1.transformer_blocks.0.attn2' -> output_blocks[10][1].transformer_blocks[0].attn2,
    'up_blocks.3.resnets.2' -> output_blocks[11][0]  (SD-1.x layout: two ResBlocks per down block, three per up block)."""
    names = {}

    def add(prefix, res_i, att_i, seq):
        for layer in seq:
            if isinstance(layer, ResBlock):
                names[f"{prefix}.resnets.{res_i}"] = layer
            elif isinstance(layer, SpatialTransformer):
                for bi, blk in enumerate(layer.transformer_blocks):
                    names[f"{prefix}.attentions.{att_i}.transformer_blocks.{bi}.attn1"] = blk.attn1
                    names[f"{prefix}.attentions.{att_i}.transformer_blocks.{bi}.attn2"] = blk.attn2

    nrb = unet.num_res_blocks
    level, j = 0, 0
    for seq in list(unet.input_blocks)[1:]:
        if any(isinstance(layer, ResBlock) for layer in seq):
            add(f"down_blocks.{level}", j, j, seq)
            j += 1
        else:                                     # Downsample closes the level
            level, j = level + 1, 0
    add("mid_block", 0, 0, [unet.middle_block[0], unet.middle_block[1]])
    names["mid_block.resnets.1"] = unet.middle_block[2]
    for k, seq in enumerate(unet.output_blocks):
        add(f"up_blocks.{k // (nrb + 1)}", k % (nrb + 1), k % (nrb + 1), seq)
    return names


def _resnet_convs(rb):
    """diffusers ResnetBlock2D sub-module names -> the LDM ResBlock's convolutions."""
    out = {"conv1": rb.in_layers[2], "conv2": rb.out_layers[3]}
    if isinstance(rb.skip_connection, nn.Conv2d):
        out["conv_shortcut"] = rb.skip_connection
    return out


def set_up_attn_processors(unet, use_attn_lora, attn_lora_layer_names=('q', 'k', 'v', 'out'), lora_rank=192, lora_scale_down=8,
                           q_lora_updates_query=False):
    """dalc:451-538.  Installs one AttnProcessor_LoRA_Capture per cross-attention module of the last up block (diffusers
    'up_blocks.3.*.attn2' = captured layers 22, 23, 24) and returns (attn_capture_procs, attn_opt_modules) with the reference's
    flattened names."""
    names = diffusers_module_names(unet)
    last = max(int(n.split(".")[1]) for n in names if n.startswith("up_blocks."))
    attn_capture_procs, attn_opt_modules = {}, {}
    idx = 0
    for name, mod in names.items():
        if not (name.startswith(f"up_blocks.{last}.") and name.endswith("attn2")):
            if name.endswith(("attn1", "attn2")):
                mod.mask_mode = "processor"                      # img_mask handled as the plain processor would (dalc:254-273)
            continue
        layers = {"q": mod.to_q, "k": mod.to_k, "v": mod.to_v, "out": mod.to_out[0]}
        proc = AttnProcessor_LoRA_Capture(capture_ca_activations=True, enable_lora=use_attn_lora, lora_uses_dora=True,
                                          lora_proj_layers={n: layers[n] for n in attn_lora_layer_names},
                                          lora_rank=lora_rank, lora_alpha=lora_rank // lora_scale_down,
                                          q_lora_updates_query=q_lora_updates_query, attn_proc_idx=idx).to(mod.to_q.weight.device)
        idx += 1
        mod.processor = proc
        flat = (name + ".processor").replace(".", "_")
        attn_capture_procs[flat] = proc
        if use_attn_lora:
            attn_opt_modules[flat + "_cross_attn_scale_factor"] = proc.cross_attn_scale_factor
            for sub, m in proc.named_modules():
                if isinstance(m, LoraDoraLinear):
                    path = flat + "_" + sub.replace(".", "_")
                    attn_opt_modules[path + "_lora_A"] = m.lora_A
                    attn_opt_modules[path + "_lora_B"] = m.lora_B
                    attn_opt_modules[path + "_lora_magnitude_vector"] = m.lora_magnitude_vector
    return attn_capture_procs, attn_opt_modules


def set_up_ffn_loras(unet, target_modules_pat, lora_uses_dora=True, lora_rank=192, lora_alpha=16,
                     adapter_names=("recon_loss", "unet_distill", "comp_distill")):
    """dalc:541-591: conv-LoRA (DoRA) adapters on the ResBlock convolutions whose diffusers name matches ``target_modules_pat``
    (reference: 'up_blocks.3.resnets.[12].conv[a-z0-9_]+'), one set per adapter name.  Returns (ffn_lora_layers,
    ffn_opt_modules) keyed like the reference."""
    ffn_lora_layers, ffn_opt_modules = {}, {}
    if target_modules_pat is None:
        return ffn_lora_layers, ffn_opt_modules
    for name, rb in diffusers_module_names(unet).items():
        if not isinstance(rb, ResBlock):
            continue
        for cname, conv in _resnet_convs(rb).items():
            full = f"{name}.{cname}"
            if not re.search(target_modules_pat, full):
                continue
            lo = LoraDoraConv2d(conv, adapter_names[0], r=lora_rank, lora_alpha=lora_alpha, use_dora=lora_uses_dora)
            for extra in adapter_names[1:]:
                lo.add_adapter(extra)
            rb.conv_loras[cname] = lo
            ffn_lora_layers[full] = lo
            flat = full.replace(".", "_")
            ffn_opt_modules[flat + "_lora_A"] = lo.lora_A
            ffn_opt_modules[flat + "_lora_B"] = lo.lora_B
            if lora_uses_dora:
                ffn_opt_modules[flat + "_lora_magnitude_vector"] = lo.lora_magnitude_vector
    return ffn_lora_layers, ffn_opt_modules


def set_lora_and_capture_flags(unet, unet_lora_modules, attn_capture_procs, outfeat_capture_blocks, res_hidden_states_gradscale_blocks,
                               use_attn_lora, use_ffn_lora, ffn_lora_adapter_name, capture_ca_activations, normalize_cross_attn,
                               mix_attn_mats_in_batch, res_hidden_states_gradscale):
    """dalc:593-629.  ``outfeat_capture_blocks`` / ``res_hidden_states_gradscale_blocks`` are accepted for signature parity; in
    the mirror both behaviours are arguments of UNetModel.forward (extra_info), recorded here on the U-Net object."""
    for proc in attn_capture_procs:
        proc.reset_attn_cache_and_flags(capture_ca_activations, normalize_cross_attn, mix_attn_mats_in_batch, enable_lora=use_attn_lora)
    unet.res_hidden_states_gradscale = res_hidden_states_gradscale
    for m in unet.modules():
        if isinstance(m, ResBlock) and len(m.conv_loras):
            m.ffn_lora_on = bool(use_ffn_lora)
            if use_ffn_lora:
                if ffn_lora_adapter_name is None:
                    raise ValueError("use_ffn_lora=True needs ffn_lora_adapter_name (dalc:612-619)")
                for lo in m.conv_loras.values():
                    lo.set_active_adapter(ffn_lora_adapter_name)
    if unet_lora_modules is not None:                                  # dalc:625-629
        for p in unet_lora_modules.parameters():
            p.requires_grad = True


def get_captured_activations(capture_ca_activations, ca_layers_activations, captured_layer_indices=(22, 23, 24), out_dtype=torch.float32):
    """dalc:631-661 over the dict UNetModel.forward left in extra_info['ca_layers_activations']: {key: {layer: tensor}} cast to
    ``out_dtype``; with capture off, the same keys with empty dicts."""
    keys = ('outfeat', 'attn', 'attnscore', 'q', 'q2', 'k', 'v', 'attn_out')
    out = {k: {} for k in keys}
    if not capture_ca_activations or not ca_layers_activations:
        return out
    for k, per_layer in ca_layers_activations.items():
        out.setdefault(k, {})
        for li in captured_layer_indices:
            if li in per_layer and per_layer[li] is not None:
                out[k][li] = per_layer[li].to(out_dtype)
    return out


class DiffusersUNetWrapper(nn.Module):
    """ddpm.py:4060-4252 over the mirror: owns the U-Net, its capture processors and the trainable LoRA set
    (``unet_lora_modules``: an nn.ParameterDict with the reference's flat names), and runs one U-Net call with the per-call
    flags of ``extra_info``."""

    def __init__(self, unet, use_attn_lora=True, use_ffn_lora=False, lora_rank=192, attn_lora_scale_down=8, ffn_lora_scale_down=8,
                 q_lora_updates_query=False, attn_lora_layer_names=('q', 'k', 'v', 'out')):
        super().__init__()
        self.diffusion_model = unet
        self.use_attn_lora, self.use_ffn_lora, self.lora_rank = use_attn_lora, use_ffn_lora, lora_rank
        self.attn_lora_scale_down, self.ffn_lora_scale_down = attn_lora_scale_down, ffn_lora_scale_down
        self.q_lora_updates_query, self.attn_lora_layer_names = q_lora_updates_query, tuple(attn_lora_layer_names)
        self.setup_hooks_and_loras()

    def setup_hooks_and_loras(self):
        """ddpm.py:4110-4182."""
        procs, attn_opt = set_up_attn_processors(self.diffusion_model, self.use_attn_lora, self.attn_lora_layer_names, self.lora_rank,
                                                 self.attn_lora_scale_down, self.q_lora_updates_query)
        self.attn_capture_procs = list(procs.values())
        self.res_hidden_states_gradscale_blocks, self.outfeat_capture_blocks = [], []      # behaviours live in UNetModel.forward
        for p in self.diffusion_model.parameters():
            p.requires_grad = False
        self.ffn_lora_layers, self.unet_lora_modules = [], None
        if self.use_attn_lora or self.use_ffn_lora:
            # reference: 'up_blocks.3.resnets.[12].conv[a-z0-9_]+' (ddpm.py:4146); "3" = the last up block of SD-1.5
            last = max(int(n.split(".")[1]) for n in diffusers_module_names(self.diffusion_model) if n.startswith("up_blocks."))
            pat = f'up_blocks.{last}.resnets.[12].conv[a-z0-9_]+' if self.use_ffn_lora else None
            ffn_layers, ffn_opt = set_up_ffn_loras(self.diffusion_model, pat, True, self.lora_rank, self.lora_rank // self.ffn_lora_scale_down)
            self.ffn_lora_layers = list(ffn_layers.values())
            mods = {}
            mods.update(attn_opt)
            mods.update(ffn_opt)
            self.unet_lora_modules = nn.ParameterDict(mods)
            for p in self.unet_lora_modules.parameters():
                p.requires_grad = True
                p.data = p.data.to(torch.float32)

    def trainable_parameters(self):
        return [] if self.unet_lora_modules is None else [p for p in self.unet_lora_modules.parameters() if p.requires_grad]

    def forward(self, x, t, cond_context, out_dtype=torch.float32):
        """ddpm.py:4187-4252: x [B,4,h,w], t [B], cond_context = (prompt_emb [B,S,768], prompt_in, extra_info) -> noise prediction
        [B,4,h,w] in out_dtype; captured activations land in extra_info['ca_layers_activations']."""
        prompt_emb, _, extra_info = cond_context
        ei = extra_info if extra_info is not None else {}
        capture = ei.get('capture_ca_activations', False)
        use_attn_lora = ei.get('use_attn_lora', self.use_attn_lora)
        use_ffn_lora = ei.get('use_ffn_lora', self.use_ffn_lora)
        gradscale = ei.get('res_hidden_states_gradscale', 1)
        set_lora_and_capture_flags(self.diffusion_model, self.unet_lora_modules, self.attn_capture_procs, self.outfeat_capture_blocks,
                                   self.res_hidden_states_gradscale_blocks, use_attn_lora, use_ffn_lora, ei.get('ffn_lora_adapter_name', None),
                                   capture, ei.get('normalize_cross_attn', False), ei.get('mix_attn_mats_in_batch', False), gradscale)
        info = {'img_mask': ei.get('img_mask', None), 'subj_indices': ei.get('subj_indices', None), 'capture_ca_activations': capture,
                'res_hidden_states_gradscale': gradscale}
        out = self.diffusion_model(x.float() if x.dtype == torch.float16 else x, t, context=prompt_emb, extra_info=info)
        if extra_info is not None:
            extra_info['ca_layers_activations'] = get_captured_activations(capture, info.get('ca_layers_activations'),
                                                                           self.diffusion_model.captured_layer_indices, out_dtype)
        # restore: capture off, every attention LoRA disabled (ddpm.py:4245-4248)
        set_lora_and_capture_flags(self.diffusion_model, self.unet_lora_modules, self.attn_capture_procs, self.outfeat_capture_blocks,
                                   self.res_hidden_states_gradscale_blocks, False, False, None, False, False, False, gradscale)
        return out.to(out_dtype)
