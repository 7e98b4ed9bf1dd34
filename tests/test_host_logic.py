"""CPU: host-side logic of the mirrors that does not need the CUDA library -- parameter packing, mask / index
preparation, module surfaces (constructor signatures, state-dict keys, flags), gradient scalers -- held against the
oracle's restatement of the reference (file:line in the oracle docstrings)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cases as C
import oracle
import adaface_dev_b200 as a


def test_img_mask_to_key_mask_follows_dalc_254_273():
    """Nearest resize to sqrt(N) x sqrt(N), key mask, dropped for the WHOLE batch if any instance's mask is all zero."""
    rng = np.random.default_rng(3)
    m = torch.from_numpy(C.img_mask(rng, 2, 64))
    km = a.img_mask_to_key_mask(m, 256)
    ref = F.interpolate(m, size=(16, 16), mode="nearest").reshape(2, -1) != 0
    assert km.dtype == torch.uint8 and torch.equal(km.bool(), ref)
    m0 = torch.from_numpy(C.img_mask(rng, 2, 64, zero_instance=1))
    assert a.img_mask_to_key_mask(m0, 256).min().item() == 1          # dropped: every key attends


def test_processor_surface_and_flags():
    """Constructor / reset_attn_cache_and_flags / attributes of dalc:147-190, incl. the always-present
    cross_attn_scale_factor (reference quirk 2 fixed) and peft's parameter names."""
    attn = a.Attention(320, 768, 8, 40)
    layers = {"q": attn.to_q, "k": attn.to_k, "v": attn.to_v, "out": attn.to_out[0]}
    proc = a.AttnProcessor_LoRA_Capture(capture_ca_activations=True, enable_lora=True, lora_proj_layers=layers, lora_rank=8,
                                        lora_alpha=1, q_lora_updates_query=True, attn_proc_idx=2)
    assert proc.lora_scale == 1 / 8 and proc.attn_proc_idx == 2 and proc.q_lora_updates_query
    names = {n for n, _ in proc.named_parameters()}
    assert "cross_attn_scale_factor" in names and "to_q_lora.lora_A.default.weight" in names
    assert "to_out_lora.lora_magnitude_vector.default.weight" in names and "to_v_lora.lora_B.default.weight" in names
    proc.reset_attn_cache_and_flags(False, True, False, False)
    assert proc.cached_activations == {} and proc.normalize_cross_attn and not proc.enable_lora
    plain = a.AttnProcessor_LoRA_Capture()
    assert float(plain.cross_attn_scale_factor.detach()) == pytest.approx(0.8) and not plain.enable_lora
    with pytest.raises(ValueError):
        a.AttnProcessor_LoRA_Capture(enable_lora=True, lora_proj_layers={"bogus": attn.to_q})
    with pytest.raises(NotImplementedError):
        proc(attn, torch.zeros(1, 4, 320), attention_mask=torch.zeros(1, 4))


def test_gradient_scalers():
    """dalc:23-67: alpha == 1 -> Identity, 0 -> detach, otherwise grad * alpha with an identity forward."""
    x = torch.randn(4, requires_grad=True)
    assert isinstance(a.gen_gradient_scaler(1), torch.nn.Identity)
    assert not a.gen_gradient_scaler(0)(x).requires_grad
    y = a.gen_gradient_scaler(10)(x)
    assert torch.equal(y, x)
    y.sum().backward()
    assert torch.equal(x.grad, torch.full_like(x, 10.0))
    with pytest.raises(ValueError):
        a.gen_gradient_scaler(-1)


def test_ldm_module_state_dict_keys_match_reference_names():
    blk = a.BasicTransformerBlock(320, 8, 40, context_dim=768)
    keys = set(blk.state_dict())
    for k in ("attn1.to_q.weight", "attn1.to_out.0.bias", "attn2.to_k.weight", "ff.net.0.proj.weight", "ff.net.2.bias",
              "norm1.weight", "norm3.bias"):
        assert k in keys, k
    assert blk.attn2.to_k.weight.shape == (320, 768) and blk.attn1.to_q.bias is None
    st = a.SpatialTransformer(320, 8, 40, depth=2, context_dim=768)
    sk = set(st.state_dict())
    assert {"norm.weight", "proj_in.weight", "proj_out.bias", "transformer_blocks.1.attn2.to_v.weight"} <= sk
    assert st.proj_out.weight.abs().max().item() == 0           # zero_module (attention.py:280)


def test_ffn_geglu_packing_is_a_permutation_of_the_reference_rows():
    """FeedForward packs the GEGLU weight rows as [a(64) | gate(64)] per 128-column tile: a pure row permutation."""
    ff = a.FeedForward(320)
    p = ff.net[0].proj
    inner = p.weight.shape[0] // 2
    idx = torch.arange(inner).view(-1, 64)
    perm = torch.cat([idx, idx + inner], dim=1).reshape(-1)
    assert sorted(perm.tolist()) == list(range(2 * inner))
    x = torch.randn(3, 320)
    h = F.linear(x, p.weight[perm], p.bias[perm]).view(3, -1, 2, 64)
    ref = F.linear(x, p.weight, p.bias)
    assert torch.allclose(h[:, :, 0].reshape(3, -1), ref[:, :inner]) and torch.allclose(h[:, :, 1].reshape(3, -1), ref[:, inner:])


def test_sbg_template_and_mkv_weight_surgery():
    """Template ids of subj_basis_generator.py:473-483 and CLIPAttentionMKV.extend / squeeze (arc2face_models.py:82-142)."""
    ids = a.template_ids(16, 77)
    assert len(ids) == 77 and ids[:4] == [49406, 1125, 539, 320] and ids[4:22] == [267] * 18 and set(ids[22:]) == {49407}
    assert ids == oracle.SBG_TEMPLATE_IDS
    at = a.CLIPAttentionMKV(a.CLIPTextConfig(num_hidden_layers=1), 1)
    w0, b0 = at.k_proj.weight.detach().clone(), at.k_proj.bias.detach().clone()
    at.extend_weights(2, perturb_std=0.0)
    assert at.multiplier == 2 and at.k_proj.weight.shape == (1536, 768)
    assert torch.equal(at.k_proj.weight[:768], w0) and torch.equal(at.k_proj.weight[768:], w0)
    at.squeeze_weights(2)
    assert at.multiplier == 1 and torch.allclose(at.k_proj.weight, w0) and torch.allclose(at.k_proj.bias, b0)
    with pytest.raises(ValueError):
        at.squeeze_weights(3)
    gen = a.SubjBasisGenerator(clip_config=a.CLIPTextConfig(num_hidden_layers=2))
    assert gen.N_ID == 16 and gen.hidden_state_layer_weights.shape == (3, 1)
    assert not gen.prompt2token_proj.text_model.embeddings.token_embedding.weight.requires_grad        # :841-853
    assert gen.extend_prompt2token_proj_attention(multiplier=2) == 2 and gen.prompt2token_proj_attention_multipliers == [2, 2]
    enc = a.Arc2FaceID2ImgPrompt(clip_config=a.CLIPTextConfig(num_hidden_layers=1))
    assert enc.PROMPT_IDS == oracle.ARC2FACE_PROMPT_IDS and not any(p.requires_grad for p in enc.parameters())
    with pytest.raises(RuntimeError):
        gen(torch.zeros(1, 16, 768))                            # CPU tensors never fall back


def test_cfg_pair_sharding_indices():
    """parallel.cfg_batch_indices keeps each image's (cond, uncond) pair on one rank."""
    n = 7
    seen = []
    for r in range(3):
        idx = a.parallel.cfg_batch_indices(n, r, 3)
        b, e = a.parallel.shard_range(n, r, 3)
        assert idx.tolist() == list(range(b, e)) + [n + i for i in range(b, e)]
        seen += list(range(b, e))
    assert seen == list(range(n))


def test_conv3x3_weight_pack_layout():
    """ops.pack_conv3x3_weight: K index = (ky*3 + kx) * Kc + ci with Kc = Cin rounded up to 64 and zero padding, i.e. a
    pixel's 3x3 neighbourhood in NHWC order times the packed row reproduces the convolution (oracle conv3x3)."""
    from oracle import unet_blocks_oracle as ub
    torch.manual_seed(1)
    for cin, cout in ((8, 16), (72, 24), (128, 8)):
        w = torch.randn(cout, cin, 3, 3).bfloat16().float()
        wp = a.ops.pack_conv3x3_weight(w)
        kc = (cin + 63) // 64 * 64
        assert wp.dtype == torch.bfloat16 and tuple(wp.shape) == (cout, 9 * kc)
        x = torch.randn(1, cin, 5, 6)
        xp = F.pad(x, (1, 1, 1, 1)).permute(0, 2, 3, 1)                          # NHWC, padded
        rows = []
        for y in range(5):
            for xx in range(6):
                patch = torch.zeros(9, kc)
                patch[:, :cin] = xp[0, y:y + 3, xx:xx + 3].reshape(9, cin)
                rows.append(patch.reshape(-1))
        got = (torch.stack(rows) @ wp.float().T).reshape(5, 6, cout).permute(2, 0, 1)
        assert (got - ub.conv3x3(x, w, None)[0]).abs().max().item() < 1e-4
    with pytest.raises(ValueError):
        a.ops.pack_conv3x3_weight(torch.zeros(4, 4, 1, 1))


def test_unet_mirror_structure_matches_reference_layout():
    """The mirror's module tree (built on the meta device) follows UNetModel.__init__ (openaimodel.py:520-684) as restated by
    the oracle's unet_layout: layer kinds per block, channel counts of the skip concatenations, state-dict key names."""
    from oracle import unet_blocks_oracle as ub
    kinds = {a.ResBlock: "res", a.SpatialTransformer: "attn", a.Downsample: "down", a.Upsample: "up", torch.nn.Conv2d: "conv_in"}
    sd15 = dict(in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2, attention_resolutions=[4, 2, 1],
                channel_mult=(1, 2, 4, 4), num_heads=8, use_spatial_transformer=True, context_dim=768, transformer_depth=1, legacy=False)
    for cfg in (sd15, C.UNET_CFG_SMALL):
        with torch.device("meta"):
            m = a.UNetModel(**cfg)
        inp, mid, out = ub.unet_layout(cfg)
        assert [[kinds[type(l)] for l in b] for b in m.input_blocks] == inp
        assert [kinds[type(l)] for l in m.middle_block] == mid
        assert [[kinds[type(l)] for l in b] for b in m.output_blocks] == out
    with torch.device("meta"):
        m = a.UNetModel(**sd15)
    assert sum(p.numel() for p in m.parameters()) == 859520964            # the SD-1.5 U-Net
    assert [b[0].channels for b in m.output_blocks] == [2560, 2560, 2560, 2560, 2560, 1920, 1920, 1280, 960, 960, 640, 640]
    assert m._cross_attn(24) is m.output_blocks[11][1].transformer_blocks[0].attn2 and m._cross_attn(12) is m.middle_block[1].transformer_blocks[0].attn2
    keys = set(m.state_dict())
    assert {"time_embed.0.weight", "input_blocks.0.0.weight", "input_blocks.3.0.op.weight", "middle_block.1.proj_in.weight",
            "output_blocks.2.1.conv.weight", "output_blocks.11.1.transformer_blocks.0.attn2.to_k.weight", "out.2.bias"} <= keys
    with pytest.raises(NotImplementedError):
        a.UNetModel(4, 320, 4, 2, [4], num_heads=8)                       # no spatial transformer: not the SD-1.5 form
    x = torch.zeros(1, 4, 8, 8)
    with pytest.raises(RuntimeError):
        m(x, torch.zeros(1), context=torch.zeros(1, 77, 768))            # CPU tensors: there is no fallback


def test_unet_loads_the_unet_part_of_an_ldm_checkpoint():
    """LatentDiffusion checkpoints keep the U-Net under 'model.diffusion_model.'; other entries (VAE, text encoder) are ignored."""
    m = a.UNetModel(**C.UNET_CFG_SMALL)
    src = a.UNetModel(**C.UNET_CFG_SMALL)
    with torch.no_grad():
        for p in src.parameters():
            p.normal_()
    ckpt = {"model.diffusion_model." + k: v for k, v in src.state_dict().items()}
    ckpt["first_stage_model.encoder.conv_in.weight"] = torch.zeros(1)
    ckpt["cond_stage_model.transformer.text_model.embeddings.position_ids"] = torch.zeros(1)
    res = m.load_ldm_state_dict(ckpt)
    assert not res.missing_keys and not res.unexpected_keys
    assert all(torch.equal(p, q) for p, q in zip(m.state_dict().values(), src.state_dict().values()))
    with pytest.raises(KeyError):
        m.load_ldm_state_dict({"foo.bar": torch.zeros(1)})
