#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own code on the seeded cases.

Runs only in the authoring container (needs /root/reference, which does not exist on the GPU box).
The reference files are imported VERBATIM from /root/reference:
    adaface/diffusers_attn_lora_capture.py  (AttnProcessor_LoRA_Capture, slow SDPA)
    ldm/modules/attention.py                (CrossAttention, BasicTransformerBlock)
    adaface/arc2face_models.py              (CLIPAttentionMKV, CLIPTextModelWrapper.forward)
    adaface/subj_basis_generator.py         (SubjBasisGenerator.forward / inverse_img_prompt_embs)
behind sys.modules stubs for the packages that are absent here (diffusers, peft, ConsistentID) --
SURVEY.md 8(c).  Two pieces are RESTATED here because the environment cannot run the reference's
third-party dependency (=> "parity unpinned" for exactly these two):
  * peft ``lora.Linear`` + DoRA (package absent, unpinned in requirements.txt:30): ``_PeftLoraLinear``.
  * the transformers-4.44 ``CLIPEncoder.forward`` loop (installed 5.5.0 dropped
    ``causal_attention_mask`` / ``output_hidden_states``): ``_Encoder444``.

Usage:  python tests/golden/make_golden.py            (writes next to this file)
"""
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ADAFACE_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
import cases as C  # noqa: E402


# --------------------------------------------------------------------------------- stubs
def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _PeftLoraLinear(nn.Module):
    """Eval-mode restatement of peft.tuners.lora.Linear(base, 'default', r, lora_alpha, use_dora=True)
    (PARITY UNPINNED: peft is un-vendored).  y = b + m/||W+sBA||_row * (xW^T + s (xA^T)B^T)."""

    def __init__(self, base_layer, adapter_name, r=8, lora_alpha=8, use_dora=True, lora_dropout=0.0, **kw):
        super().__init__()
        self.base_layer = base_layer
        self.r, self.scaling, self.use_dora = r, lora_alpha / r, use_dora
        self.lora_A = nn.ModuleDict({adapter_name: nn.Linear(base_layer.in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({adapter_name: nn.Linear(r, base_layer.out_features, bias=False)})
        nn.init.kaiming_uniform_(self.lora_A[adapter_name].weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B[adapter_name].weight)
        mag = torch.linalg.norm(base_layer.weight.detach(), dim=1)
        self.lora_magnitude_vector = nn.ParameterDict({adapter_name: nn.Parameter(mag.clone())})
        self.adapter = adapter_name

    def forward(self, x):
        A, B = self.lora_A[self.adapter].weight, self.lora_B[self.adapter].weight
        W, b = self.base_layer.weight, self.base_layer.bias
        base = F.linear(x, W)
        lora = F.linear(F.linear(x, A), B) * self.scaling
        if self.use_dora:
            wn = torch.linalg.norm(W + self.scaling * (B @ A), dim=1).detach()
            y = (self.lora_magnitude_vector[self.adapter] / wn) * (base + lora)
        else:
            y = base + lora
        return y if b is None else y + b


def install_stubs():
    class _Dummy:  # placeholder classes only used in annotations / isinstance checks we never hit
        pass

    lg = types.SimpleNamespace(get_logger=lambda *_: types.SimpleNamespace(warning=print, info=print))
    _mod("diffusers", StableDiffusionPipeline=_Dummy, UNet2DConditionModel=_Dummy, DDIMScheduler=_Dummy)
    _mod("diffusers.models")
    _mod("diffusers.models.attention_processor", Attention=_Dummy, AttnProcessor2_0=_Dummy)
    _mod("diffusers.utils", logging=lg, is_torch_version=lambda *a: True, deprecate=lambda *a, **k: None)
    _mod("diffusers.loaders")
    _mod("diffusers.loaders.peft", PeftAdapterMixin=_Dummy)
    lora = _mod("peft.tuners.lora", Linear=_PeftLoraLinear, LoraLayer=_PeftLoraLinear)
    _mod("peft", LoraConfig=_Dummy, get_peft_model=lambda *a, **k: None)
    _mod("peft.tuners", lora=lora)
    _mod("peft.tuners.lora.dora", DoraLinearLayer=_Dummy)
    _mod("peft.tuners.tuners_utils", BaseTunerLayer=_Dummy)
    _mod("omegaconf")                                   # UNetModel.__init__ only compares type(context_dim) with ListConfig
    _mod("omegaconf.listconfig", ListConfig=_Dummy)
    sys.path.insert(0, REF)
    # adaface/util.py drags in diffusers pipelines + the un-vendored ConsistentID package; the two
    # helpers the hot path needs from it are identical copies of dalc:23-67.
    import importlib.util
    import adaface  # noqa: F401  (empty package __init__)
    spec = importlib.util.spec_from_file_location("adaface.diffusers_attn_lora_capture",
                                                  os.path.join(REF, "adaface/diffusers_attn_lora_capture.py"))
    dalc = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = dalc
    spec.loader.exec_module(dalc)
    _mod("adaface.util", gen_gradient_scaler=dalc.gen_gradient_scaler, perturb_tensor=lambda t, *a, **k: t)
    return dalc


# --------------------------------------------------------------------------------- helpers
def T(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a))


def lin(W, b=None):
    m = nn.Linear(W.shape[1], W.shape[0], bias=b is not None)
    m.weight.data = T(W).clone()
    if b is not None:
        m.bias.data = T(b).clone()
    return m


def save(name, case, outs):
    flat = {"input_checksum": np.float64(C.checksum({k: v for k, v in case.items() if k != "spec"}))}
    for k, v in outs.items():
        flat[k] = v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez(path, **flat)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB  keys={list(flat)}")


# --------------------------------------------------------------------------------- processor cases
def run_proc_cases(dalc):
    for name in C.PROC_CASES:
        case = C.build_proc_case(name)
        sp, w = case["spec"], case["w"]
        to_q, to_k, to_v = lin(w["to_q"]), lin(w["to_k"]), lin(w["to_v"])
        to_out = nn.ModuleList([lin(w["to_out_w"], w["to_out_b"]), nn.Dropout(0.0)])
        attn = types.SimpleNamespace(spatial_norm=None, group_norm=None, norm_cross=None, norm_q=None, norm_k=None,
                                     heads=8, to_q=to_q, to_k=to_k, to_v=to_v, to_out=to_out,
                                     residual_connection=False, rescale_output_factor=1.0,
                                     prepare_attention_mask=None)
        r = sp.get("lora_rank", 0)
        layers = {"q": to_q, "k": to_k, "v": to_v, "out": to_out[0]} if r else None
        proc = dalc.AttnProcessor_LoRA_Capture(capture_ca_activations=sp.get("capture", False), enable_lora=bool(r),
                                               lora_uses_dora=True, lora_proj_layers=layers, lora_rank=r or 192,
                                               lora_alpha=sp.get("lora_alpha", 16),
                                               q_lora_updates_query=sp.get("q_upd", False), attn_proc_idx=0)
        if not r:   # reference quirk 2: attribute only exists with LoRA; the slow path reads it always
            proc.cross_attn_scale_factor = nn.Parameter(torch.tensor(0.8))
        for n in (("q", "k", "v", "out") if r else ()):
            A, B, mag = w["lora_" + n]
            mod = getattr(proc, f"to_{n}_lora")
            mod.lora_A["default"].weight.data = T(A).clone()
            mod.lora_B["default"].weight.data = T(B).clone()
            mod.lora_magnitude_vector["default"].data = T(mag).clone()
        proc.reset_attn_cache_and_flags(sp.get("capture", False), sp.get("normalize", False), sp.get("mix", False),
                                        sp.get("enable_lora", False))
        si = case["subj_indices"]
        with torch.no_grad():
            out = proc(attn, T(case["hidden_states"]), encoder_hidden_states=T(case["encoder_hidden_states"]),
                       img_mask=T(case["img_mask"]), subj_indices=None if si is None else (T(si[0]), T(si[1])))
        outs = {"out": out}
        for k, v in proc.cached_activations.items():
            outs["cache_" + k] = v
        save(name, case, outs)


# --------------------------------------------------------------------------------- LDM cases
def run_ldm_cases():
    from ldm.modules.attention import CrossAttention, BasicTransformerBlock

    def load_attn(m, w):
        m.to_q.weight.data, m.to_k.weight.data, m.to_v.weight.data = T(w["to_q"]), T(w["to_k"]), T(w["to_v"])
        m.to_out[0].weight.data, m.to_out[0].bias.data = T(w["to_out_w"]), T(w["to_out_b"])

    for name in C.LDM_CASES:
        case = C.build_ldm_case(name)
        sp, w = case["spec"], case["w"]
        Cc = sp["C"]
        if sp.get("block"):
            blk = BasicTransformerBlock(Cc, 8, Cc // 8, context_dim=768, checkpoint=False).eval()
            load_attn(blk.attn1, w["attn1"])
            load_attn(blk.attn2, w["attn2"])
            for i, ln in enumerate((blk.norm1, blk.norm2, blk.norm3), 1):
                ln.weight.data, ln.bias.data = T(w[f"norm{i}_w"]), T(w[f"norm{i}_b"])
            blk.ff.net[0].proj.weight.data, blk.ff.net[0].proj.bias.data = T(w["ff_proj_w"]), T(w["ff_proj_b"])
            blk.ff.net[2].weight.data, blk.ff.net[2].bias.data = T(w["ff_out_w"]), T(w["ff_out_b"])
            with torch.no_grad():
                out = blk(T(case["x"]), context=T(case["context"]), mask=T(case["mask"]))
            save(name, case, {"out": out})
        else:
            cross = sp.get("cross", False)
            m = CrossAttention(Cc, context_dim=768 if cross else None, heads=8, dim_head=Cc // 8).eval()
            load_attn(m, w)
            m.save_cross_attn_vars = sp.get("save", False)
            with torch.no_grad():
                out = m(T(case["x"]), context=T(case["context"]), mask=T(case["mask"]))
            outs = {"out": out}
            for k, v in (m.cached_activations or {}).items():
                outs["cache_" + k] = v
            save(name, case, outs)


def run_spatial_cases():
    from ldm.modules.attention import SpatialTransformer

    for name in C.SPATIAL_CASES:
        case = C.build_spatial_case(name)
        sp, w = case["spec"], case["w"]
        Cc = sp["C"]
        m = SpatialTransformer(Cc, 8, Cc // 8, depth=1, context_dim=768).eval()
        blk = m.transformer_blocks[0]
        blk.checkpoint = False
        for at, key in ((blk.attn1, "attn1"), (blk.attn2, "attn2")):
            aw = w[key]
            at.to_q.weight.data, at.to_k.weight.data, at.to_v.weight.data = T(aw["to_q"]), T(aw["to_k"]), T(aw["to_v"])
            at.to_out[0].weight.data, at.to_out[0].bias.data = T(aw["to_out_w"]), T(aw["to_out_b"])
        for i, ln in enumerate((blk.norm1, blk.norm2, blk.norm3), 1):
            ln.weight.data, ln.bias.data = T(w[f"norm{i}_w"]), T(w[f"norm{i}_b"])
        blk.ff.net[0].proj.weight.data, blk.ff.net[0].proj.bias.data = T(w["ff_proj_w"]), T(w["ff_proj_b"])
        blk.ff.net[2].weight.data, blk.ff.net[2].bias.data = T(w["ff_out_w"]), T(w["ff_out_b"])
        m.norm.weight.data, m.norm.bias.data = T(w["gn_w"]), T(w["gn_b"])
        m.proj_in.weight.data, m.proj_in.bias.data = T(w["proj_in_w"])[:, :, None, None].clone(), T(w["proj_in_b"])
        m.proj_out.weight.data, m.proj_out.bias.data = T(w["proj_out_w"])[:, :, None, None].clone(), T(w["proj_out_b"])
        with torch.no_grad():
            out = m(T(case["x"]), context=T(case["context"]), mask=T(case["mask"]))
        save(name, case, {"out": out})


def run_unet_block_cases():
    from ldm.modules.diffusionmodules.openaimodel import ResBlock, Upsample, Downsample

    for name in C.UNET_BLOCK_CASES:
        case = C.build_unet_block_case(name)
        sp, w = case["spec"], case["w"]
        if sp["kind"] == "res":
            m = ResBlock(sp["cin"], sp["emb"], 0.0, out_channels=sp["cout"], use_conv=bool(sp.get("skip3")), dims=2,
                         use_checkpoint=False, use_scale_shift_norm=False).eval()
            pairs = [(m.in_layers[0], "gn1"), (m.in_layers[2], "conv1"), (m.emb_layers[1], "emb"), (m.out_layers[0], "gn2"),
                     (m.out_layers[3], "conv2")]
            if "skip_w" in w:
                pairs.append((m.skip_connection, "skip"))
            for mod, key in pairs:
                mod.weight.data, mod.bias.data = T(w[key + "_w"]).clone(), T(w[key + "_b"]).clone()
            with torch.no_grad():
                out = m(T(case["x"]), T(case["emb"]))
        else:
            m = (Upsample if sp["kind"] == "up" else Downsample)(sp["cin"], True, dims=2).eval()
            conv = m.conv if sp["kind"] == "up" else m.op
            conv.weight.data, conv.bias.data = T(w["conv_w"]).clone(), T(w["conv_b"]).clone()
            with torch.no_grad():
                out = m(T(case["x"]))
        save(name, case, {"out": out})


def run_unet_cases():
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    from ldm.modules.attention import BasicTransformerBlock

    for name in C.UNET_CASES:
        if len(sys.argv) > 2 and sys.argv[2] != name:      # `make_golden.py unet unet_small`: one case only
            continue
        case = C.build_unet_case(name)
        sp = case["spec"]
        torch.manual_seed(0)
        m = UNetModel(**sp["cfg"]).eval()
        for mod in m.modules():
            if isinstance(mod, BasicTransformerBlock):
                mod.checkpoint = False
        sd = C.unet_state_dict({k: v.shape for k, v in m.state_dict().items()}, sp["seed"] + 1000)
        m.load_state_dict({k: T(v) for k, v in sd.items()})
        with torch.no_grad():
            out = m(T(case["x"]), T(case["timesteps"]), context=T(case["context"]),
                    extra_info={"img_mask": T(case["mask"]), "capture_ca_activations": False})
        save(name, case, {"out": out})


class _StandInLDM:
    """The 10-line stand-in for LatentDiffusion that DDIMSampler needs: schedule buffers + apply_model."""

    def __init__(self, apply_model):
        from ldm.modules.diffusionmodules.util import make_beta_schedule
        betas = make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.012)
        ac = np.cumprod(1. - betas, axis=0)                                          # ddpm.py:301-303
        self.num_timesteps = 1000
        self.betas = torch.tensor(betas, dtype=torch.float32)
        self.alphas_cumprod = torch.tensor(ac, dtype=torch.float32)
        self.alphas_cumprod_prev = torch.tensor(np.append(1., ac[:-1]), dtype=torch.float32)
        self.device = torch.device("cpu")
        self.apply_model = apply_model


def run_ddim_cases():
    from ldm.models.diffusion.ddim import DDIMSampler
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    from ldm.modules.attention import BasicTransformerBlock

    class CpuDDIMSampler(DDIMSampler):
        def register_buffer(self, name, attr):       # the reference's hard-codes .to("cuda") (ddim.py:21-25): keep tensors where they are
            setattr(self, name, attr)

    for name in C.DDIM_CASES:
        case = C.build_ddim_case(name)
        sp = case["spec"]
        if sp["model"] == "standin":
            apply_model = C.standin_eps
        else:
            m = UNetModel(**C.UNET_CFG_SMALL).eval()
            for mod in m.modules():
                if isinstance(mod, BasicTransformerBlock):
                    mod.checkpoint = False
            sd = C.unet_state_dict({k: v.shape for k, v in m.state_dict().items()}, sp["seed"] + 1000)
            m.load_state_dict({k: T(v) for k, v in sd.items()})
            apply_model = lambda x, t, c, m=m: m(x, t, context=c, extra_info={})
        ldm_model = _StandInLDM(apply_model)
        sampler = CpuDDIMSampler(ldm_model)
        g = sp["guidance"]
        with torch.no_grad():
            img, inter = sampler.sample(sp["steps"], sp["B"], (4, sp["h"], sp["w"]), conditioning=T(case["cond"]), eta=0.,
                                        verbose=False, x_T=T(case["x_T"]), guidance_scale=list(g) if isinstance(g, tuple) else g,
                                        unconditional_conditioning=T(case["uncond"]), log_every_t=1)
        save(name, case, {"x0": img, "pred_x0_last": inter["pred_x0"][-1], "x_after_first": inter["x_inter"][1],
                          "alphas_cumprod": ldm_model.alphas_cumprod, "ddim_timesteps": sampler.ddim_timesteps, "ddim_alphas": sampler.ddim_alphas,
                          "ddim_alphas_prev": np.asarray(sampler.ddim_alphas_prev, dtype=np.float64),
                          "ddim_sigmas": np.asarray(sampler.ddim_sigmas, dtype=np.float64),
                          "ddim_sqrt_one_minus_alphas": sampler.ddim_sqrt_one_minus_alphas})


def run_closs_cases():
    import ldm.util as U

    for name in C.CLOSS_CASES:
        case = C.build_closs_case(name)
        sp = case["spec"]
        subj = (T(case["subj_ib"]), T(case["subj_it"]))
        if sp["kind"] == "bg":
            ca = {22: T(case["attn23"]) * 0 + 1.0 / sp["S"], 23: T(case["attn23"]), 24: T(case["attn24"])}      # layer 22 is ignored
            loss = U.calc_subj_masked_bg_suppress_loss(ca, subj, sp["block"], T(case["fg_mask"]))
            save(name, case, {"loss": loss})
        else:
            acts = {"attn": {23: T(case["attn23"]), 24: T(case["attn24"])}, "k": {23: T(case["k23"]), 24: T(case["k24"])},
                    "v": {23: T(case["v23"]), 24: T(case["v24"])}}
            out = U.calc_sc_rep_attn_distill_loss(acts, subj, T(case["emb_mask"]), T(case["pad_mask"]), sp["fg_percent"])
            save(name, case, {"losses": torch.stack([torch.as_tensor(o, dtype=torch.float32) for o in out])})


# --------------------------------------------------------------------------------- SBG cases
class _EncOut:
    """Minimal stand-in for HF BaseModelOutput: tuple-indexable and attribute-addressable."""

    def __init__(self, last, hs):
        self.last_hidden_state, self.hidden_states, self.attentions = last, hs, None

    def __getitem__(self, i):
        return (self.last_hidden_state, self.hidden_states)[i]


class _Encoder444(nn.Module):
    """transformers-4.44 CLIPEncoder.forward (RESTATED; PARITY UNPINNED for this loop only): run the
    pre-LN CLIPEncoderLayers, handing ``causal_attention_mask`` to self_attn, collecting the 13 hidden
    states.  The layer sub-modules are HF's own (LayerNorm, CLIPMLP) and the reference's CLIPAttentionMKV."""

    def __init__(self, layers):
        super().__init__()
        self.layers = layers

    def forward(self, inputs_embeds, attention_mask=None, causal_attention_mask=None, output_attentions=None,
                output_hidden_states=None, return_dict=None):
        h = inputs_embeds
        hs = (h,)
        for layer in self.layers:
            r = h
            h = layer.layer_norm1(h)
            h = r + layer.self_attn(h, attention_mask, causal_attention_mask, False)[0]
            r = h
            h = r + layer.mlp(layer.layer_norm2(h))
            hs = hs + (h,)
        return _EncOut(h, hs)


def build_ref_text_model(w, mults):
    from transformers import CLIPTextConfig
    from adaface.arc2face_models import CLIPTextModelWrapper, CLIPAttentionMKV
    cfg = CLIPTextConfig(vocab_size=C.VOCAB, hidden_size=C.E, intermediate_size=C.MLP, num_hidden_layers=len(mults),
                         num_attention_heads=C.HEADS, max_position_embeddings=C.NPOS, hidden_act="quick_gelu")
    model = CLIPTextModelWrapper(cfg).eval()
    tm = model.text_model
    tm.embeddings.token_embedding.weight.data.zero_()
    for tid, row in w["token_emb_rows"].items():
        tm.embeddings.token_embedding.weight.data[tid] = T(row)
    tm.embeddings.position_embedding.weight.data = T(w["pos_emb"]).clone()
    for layer, lw, m in zip(tm.encoder.layers, w["layers"], mults):
        layer.self_attn = CLIPAttentionMKV(cfg, multiplier=m)
        load_mkv(layer.self_attn, lw)
        layer.layer_norm1.weight.data, layer.layer_norm1.bias.data = T(lw["ln1_w"]), T(lw["ln1_b"])
        layer.layer_norm2.weight.data, layer.layer_norm2.bias.data = T(lw["ln2_w"]), T(lw["ln2_b"])
        layer.mlp.fc1.weight.data, layer.mlp.fc1.bias.data = T(lw["fc1_w"]), T(lw["fc1_b"])
        layer.mlp.fc2.weight.data, layer.mlp.fc2.bias.data = T(lw["fc2_w"]), T(lw["fc2_b"])
    tm.final_layer_norm.weight.data, tm.final_layer_norm.bias.data = T(w["final_ln_w"]), T(w["final_ln_b"])
    tm.encoder = _Encoder444(tm.encoder.layers)
    return model.eval()


def load_mkv(m, lw):
    for p, mod in (("q", m.q_proj), ("k", m.k_proj), ("v", m.v_proj), ("o", m.out_proj)):
        mod.weight.data, mod.bias.data = T(lw[p + "_w"]).clone(), T(lw[p + "_b"]).clone()


def run_sbg_cases():
    from transformers import CLIPTextConfig
    from transformers.modeling_attn_mask_utils import AttentionMaskConverter
    from adaface.arc2face_models import CLIPAttentionMKV
    import adaface.subj_basis_generator as sbg_mod

    for name in C.SBG_CASES:
        case = C.build_sbg_case(name)
        sp, w = case["spec"], case["w"]
        if sp["layers"] == 0:
            cfg = CLIPTextConfig(hidden_size=C.E, num_attention_heads=C.HEADS)
            m = CLIPAttentionMKV(cfg, multiplier=sp["mult"]).eval()
            load_mkv(m, w)
            x = T(case["x"])
            mask = AttentionMaskConverter._make_causal_mask(x.shape[:2], x.dtype, device=x.device)
            with torch.no_grad():
                out = m(x, None, mask, False)[0]
            save(name, case, {"out": out})
            continue
        # Real SubjBasisGenerator.forward / inverse_img_prompt_embs, constructed without the
        # network-bound __init__ (from_pretrained, subj_basis_generator.py:425-427, 612).
        gen = sbg_mod.SubjBasisGenerator.__new__(sbg_mod.SubjBasisGenerator)
        nn.Module.__init__(gen)
        gen.placeholder_is_bg, gen.dtype = False, torch.float32
        gen.N_ID, gen.N_SFX = 16, sp.get("n_sfx", 0)
        gen.max_prompt_length = C.NPOS
        gen.prompt2token_proj = build_ref_text_model(w, sp["mults"])
        gen.layerwise_proj = nn.Identity()
        gen.initialize_hidden_state_layer_weights("per-layer", "cpu")
        gen.pad_embeddings = T(w["pad_embeddings"])
        gen.static_img_suffix_embs = (nn.Parameter(T(w["static_img_suffix_embs"]))
                                      if sp.get("n_sfx") else None)
        ids = torch.tensor(C.TEMPLATE_IDS)
        gen.tokenizer = lambda prompts, **kw: types.SimpleNamespace(input_ids=ids.unsqueeze(0).repeat(len(prompts), 1))
        gen.eval()
        with torch.no_grad():
            out = gen(T(case["faceid2img_prompt_embs"]), out_id_embs_cfg_scale=sp.get("cfg", 1.0),
                      enable_static_img_suffix_embs=bool(sp.get("n_sfx")))
        save(name, case, {"out": out})


def run_text_cases():
    """the reference's own Arc2Face ID -> image-prompt mapping and its patched SD text-encoder forward."""
    import torch.nn.functional as F2  # noqa: F401
    # ---- Arc2Face_ID2AdaPrompt.map_init_id_to_img_prompt_embs, run verbatim on an instance built without the network-bound __init__
    _mod("ConsistentID")
    _mod("ConsistentID.lib")
    _mod("ConsistentID.lib.pipeline_ConsistentID", ConsistentIDPipeline=type("ConsistentIDPipeline", (), {}))
    _mod("insightface")
    _mod("insightface.app", FaceAnalysis=type("FaceAnalysis", (), {}))
    if "cv2" not in sys.modules:
        try:
            import cv2  # noqa: F401
        except Exception:
            _mod("cv2")
    import adaface.util as au
    for nm in ("pad_image_obj_to_square", "calc_stats", "patch_clip_image_encoder_with_mask"):
        setattr(au, nm, lambda *a, **k: None)
    au.CLIPVisionModelWithMask = type("CLIPVisionModelWithMask", (), {})
    import adaface.face_id_to_ada_prompt as f2a

    case = C.build_text_case("arc2face_id2img")
    w = case["w"]
    rows = dict(w["token_emb_rows"])
    rows.update(case["extra_rows"])
    w2 = dict(w)
    w2["token_emb_rows"] = rows
    enc = f2a.Arc2Face_ID2AdaPrompt.__new__(f2a.Arc2Face_ID2AdaPrompt)
    nn.Module.__init__(enc)
    enc.dtype, enc.id_img_prompt_max_length = torch.float32, 22
    enc.text_to_image_prompt_encoder = build_ref_text_model(w2, [1] * 12)
    ids22 = torch.tensor(C.ARC2FACE_PROMPT_IDS)

    class _Tok:
        def encode(self, text, add_special_tokens=False):
            assert text == "id"
            return [1014]

        def __call__(self, text, **kw):
            assert text == "photo of a id person" and kw.get("max_length") == 22
            return types.SimpleNamespace(input_ids=ids22.unsqueeze(0))
    enc.tokenizer = _Tok()
    with torch.no_grad():
        out = enc.map_init_id_to_img_prompt_embs(T(case["init_id_embs"]))
    save("arc2face_id2img", case, {"out": out})

    # ---- the patched CLIP text model of FrozenCLIPEmbedder (modules.py:180-338): embeddings_forward / encoder_forward /
    #      text_model_forward verbatim.  HF's CLIPEncoderLayer (transformers >= 4.44 API, absent here in that form) is stood in by
    #      a pre-LN residual block over HF's own LayerNorm / CLIPMLP and the reference's CLIPAttentionMKV(multiplier=1), whose
    #      arithmetic at M = 1 is HF CLIPAttention's (RESTATED call signature only).
    import ldm.modules.encoders.modules as M
    case = C.build_text_case("sd_text_encoder")
    w = case["w"]
    ref = build_ref_text_model(w, [1] * 12)
    tm = ref.text_model
    layers = tm.encoder.layers

    class _Layer444(nn.Module):
        def __init__(self, layer):
            super().__init__()
            self.l = layer

        def forward(self, hidden_states, attention_mask, causal_attention_mask, output_attentions=False):
            r = hidden_states
            h = r + self.l.self_attn(self.l.layer_norm1(hidden_states), attention_mask, causal_attention_mask, False)[0]
            return (h + self.l.mlp(self.l.layer_norm2(h)),)

    class _Enc(nn.Module):
        def __init__(self):
            super().__init__()
            self.layers = nn.ModuleList([_Layer444(l_) for l_ in layers])
            self.config = types.SimpleNamespace(output_attentions=False, output_hidden_states=False, use_return_dict=True)
    enc_mod = _Enc()
    enc_mod.forward = types.MethodType(M.encoder_forward, enc_mod)
    emb = tm.embeddings
    if not hasattr(emb, "position_ids"):
        emb.position_ids = torch.arange(C.NPOS).unsqueeze(0)
    emb.forward = types.MethodType(M.embeddings_forward, emb)
    text_model = types.SimpleNamespace(config=enc_mod.config, embeddings=emb, encoder=enc_mod, final_layer_norm=tm.final_layer_norm,
                                       last_layers_skip_weights=[0.5, 0.5])
    ada = T(case["ada"])

    def splice(input_ids, embs):
        embs = embs.clone()
        embs[:, 4:20] = ada
        return embs
    ids = torch.tensor([C.TEMPLATE_IDS] * case["spec"]["B"])
    with torch.no_grad():
        out = M.text_model_forward(text_model, input_ids=ids, embedding_manager=splice)
    save("sd_text_encoder", case, {"out": out})


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_grad_enabled(True)
    dalc = install_stubs()
    only = sys.argv[1] if len(sys.argv) > 1 else ""        # e.g. `make_golden.py spatial`: regenerate one family only
    if only in ("", "proc"):
        run_proc_cases(dalc)
    if only in ("", "ldm"):
        run_ldm_cases()
    if only in ("", "spatial"):
        run_spatial_cases()
    if only in ("", "unet_blocks"):
        run_unet_block_cases()
    if only in ("", "unet"):
        run_unet_cases()
    if only in ("", "ddim"):
        run_ddim_cases()
    if only in ("", "closs"):
        run_closs_cases()
    if only in ("", "sbg"):
        run_sbg_cases()
    if only in ("", "text"):
        run_text_cases()
