"""CPU: pin the oracle (oracle/*.py) to the reference's own outputs (tests/golden/*.npz).

The fixtures were produced by tests/golden/make_golden.py, which runs the reference files verbatim from
/root/reference.  Tolerances are fp32 round-off only: the oracle restates the same
arithmetic in a different operation order."""
import os

import numpy as np
import pytest
import torch

import cases as C
import oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ATOL = 2e-5


def load(name, case):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cs = C.checksum({k: v for k, v in case.items() if k != "spec"})
    assert abs(cs - float(g["input_checksum"])) <= 1e-9 * max(1.0, abs(cs)), \
        "seeded inputs drifted from the ones the fixture was generated on (numpy RNG stream changed?)"
    return g


def close(a, ref, atol=ATOL, what=""):
    a = a.detach().numpy() if torch.is_tensor(a) else np.asarray(a)
    assert a.shape == ref.shape, f"{what}: shape {a.shape} vs {ref.shape}"
    fin = np.isfinite(ref)
    assert (np.isfinite(a) == fin).all(), f"{what}: non-finite pattern differs"
    err = np.abs(a[fin] - ref[fin]).max() if fin.any() else 0.0
    assert err <= atol, f"{what}: max-abs {err:.3e} > {atol}"


def run_oracle_proc(case):
    sp = case["spec"]
    t = C.to_torch({k: v for k, v in case.items() if k != "spec"})
    w = t["w"]
    return oracle.processor_forward(
        w, t["hidden_states"], t["encoder_hidden_states"], img_mask=t["img_mask"], subj_indices=t["subj_indices"],
        heads=8, capture_ca_activations=sp.get("capture", False), normalize_cross_attn=sp.get("normalize", False),
        mix_attn_mats_in_batch=sp.get("mix", False), enable_lora=sp.get("enable_lora", False),
        q_lora_updates_query=sp.get("q_upd", False), lora_scaling=float(w.get("lora_scaling", 0.125)))


@pytest.mark.parametrize("name", list(C.PROC_CASES))
def test_processor_oracle_matches_reference(name):
    case = C.build_proc_case(name)
    g = load(name, case)
    out, cache = run_oracle_proc(case)
    close(out, g["out"], what="out")
    gold_keys = {k[6:] for k in g.files if k.startswith("cache_")}
    assert set(cache) == gold_keys
    for k in gold_keys:
        close(cache[k], g["cache_" + k], what=k)


@pytest.mark.parametrize("name", list(C.LDM_CASES))
def test_ldm_oracle_matches_reference(name):
    case = C.build_ldm_case(name)
    g = load(name, case)
    sp = case["spec"]
    t = C.to_torch({k: v for k, v in case.items() if k != "spec"})
    if sp.get("block"):
        out = oracle.basic_transformer_block(t["w"], t["x"], context=t["context"], mask=t["mask"])
        close(out, g["out"], atol=1e-4, what="block out")
    else:
        out, cache = oracle.ldm_cross_attention(t["w"], t["x"], context=t["context"], mask=t["mask"],
                                                save_cross_attn_vars=sp.get("save", False))
        close(out, g["out"], what="out")
        for k in (cache or {}):
            close(cache[k], g["cache_" + k], what=k)


@pytest.mark.parametrize("name", list(C.SBG_CASES))
def test_sbg_oracle_matches_reference(name):
    case = C.build_sbg_case(name)
    g = load(name, case)
    sp = case["spec"]
    t = C.to_torch({k: v for k, v in case.items() if k != "spec"})
    if sp["layers"] == 0:
        out = oracle.clip_mkv_attention(t["w"], t["x"], sp["mult"])
        close(out, g["out"], what="mkv out")
    else:
        out = oracle.sbg_forward(t["w"], t["faceid2img_prompt_embs"], out_id_embs_cfg_scale=sp.get("cfg", 1.0),
                                 enable_static_img_suffix_embs=bool(sp.get("n_sfx")), multipliers=sp["mults"])
        close(out, g["out"], atol=2e-4, what="sbg out")


def test_two_reference_surfaces_agree():
    """The diffusers-processor surface and the LDM CrossAttention surface are the same operator
    (max-abs 1.3e-7); cached q differs only by the C^-1/4 vs d^-1/4 factor (quirk 1)."""
    case = C.build_proc_case("proc_cross_capture")
    t = C.to_torch({k: v for k, v in case.items() if k != "spec"})
    o1, c1 = run_oracle_proc(case)
    o2, c2 = oracle.ldm_cross_attention(t["w"], t["hidden_states"], context=t["encoder_hidden_states"],
                                        save_cross_attn_vars=True)
    assert (o1 - o2).abs().max() < 1e-5
    assert (c1["attn"] - c2["attn"]).abs().max() < 1e-6
    Cc, d = 320, 40
    assert torch.allclose(c1["q"] * (Cc ** 0.25), c2["q"] * (d ** 0.25), atol=1e-5)


def test_dora_identity_at_init():
    """peft DoRA at init (B = 0, m = ||W||_row) is the identity adapter (SURVEY 8c property check)."""
    g = torch.Generator().manual_seed(0)
    W, x = torch.randn(24, 16, generator=g), torch.randn(5, 16, generator=g)
    A, B = torch.randn(4, 16, generator=g), torch.zeros(24, 4)
    y = oracle.lora_dora_linear(x, W, None, A, B, torch.linalg.norm(W, dim=1), 0.125)
    assert torch.allclose(y, x @ W.T, atol=1e-5)
    B = torch.randn(24, 4, generator=g) * 0.1
    m = torch.linalg.norm(W + 0.125 * B @ A, dim=1)          # m == ||W + sBA||  -> plain LoRA
    y = oracle.lora_dora_linear(x, W, None, A, B, m, 0.125)
    assert torch.allclose(y, x @ (W + 0.125 * B @ A).T, atol=1e-5)


def test_causal_truncation_is_exact_in_the_oracle():
    """The product runs the CLIP-shaped encoders on 20 positions instead of 22 / 77: with a causal mask
    the returned positions 4:20 cannot depend on later ones.  Checked on the oracle itself, small 2-layer stack."""
    import oracle
    case = C.build_sbg_case("sbg_m1")
    t = C.to_torch({k: v for k, v in case.items() if k != "spec"})
    w = dict(t["w"])
    w["layers"] = w["layers"][:2]
    g = torch.Generator().manual_seed(3)
    prompt = torch.randn(22, 768, generator=g) * 0.02
    ids = torch.nn.functional.normalize(torch.randn(3, 512, generator=g), dim=-1)
    full = oracle.arc2face_id_to_img_prompt(w, ids, prompt_embs=prompt)
    tok = prompt[:20].unsqueeze(0).repeat(3, 1, 1)
    tok[:, 4] = torch.nn.functional.pad(ids, (0, 256))
    short = oracle.clip_text_wrapper_forward(w, tok, None)[:, 4:20]
    assert tuple(full.shape) == (3, 16, 768)
    assert (full - short).abs().max().item() < 1e-5


@pytest.mark.parametrize("name", list(C.SPATIAL_CASES))
def test_spatial_transformer_oracle_vs_reference(name):
    """SpatialTransformer (ldm/modules/attention.py:287-304): GroupNorm + 1x1 proj_in + block + 1x1 proj_out + residual."""
    case = C.build_spatial_case(name)
    g = load(name, case)
    t = C.to_torch({k: v for k, v in case.items() if k != "spec"})
    out = oracle.spatial_transformer(t["w"], t["x"], context=t["context"], mask=t["mask"])
    close(out, g["out"], atol=5e-5, what=name)


@pytest.mark.parametrize("name", list(C.UNET_BLOCK_CASES))
def test_unet_blocks_oracle_vs_reference(name):
    """ResBlock / Upsample / Downsample (ldm/modules/diffusionmodules/openaimodel.py:92-277) against the reference modules."""
    from oracle import unet_blocks_oracle as ub
    case = C.build_unet_block_case(name)
    g = load(name, case)
    t = C.to_torch({k: v for k, v in case.items() if k != "spec"})
    kind = case["spec"]["kind"]
    if kind == "res":
        out = ub.res_block(t["w"], t["x"], t["emb"])
    elif kind == "up":
        out = ub.upsample(t["w"], t["x"])
    else:
        out = ub.downsample(t["w"], t["x"])
    close(out, g["out"], atol=5e-5, what=name)


@pytest.mark.parametrize("name", [n for n, s_ in C.UNET_CASES.items() if not s_.get("big")])     # SD-1.5 size: GPU parity only
def test_unet_oracle_vs_reference(name):
    """UNetModel.forward (ldm/modules/diffusionmodules/openaimodel.py:820-960) against the reference module on a two-level
    SD-1.5-shaped configuration: time embedding, ResBlocks, SpatialTransformers, Down / Upsample, skip concatenations."""
    from oracle import unet_blocks_oracle as ub
    case = C.build_unet_case(name)
    sp = case["spec"]
    g = load(name, case)
    import adaface_dev_b200 as a      # module structure only (CPU construction on the meta device; no kernels are called)
    with torch.device("meta"):
        shapes = {k: v.shape for k, v in a.UNetModel(**sp["cfg"]).state_dict().items()}
    sd = {k: torch.from_numpy(v) for k, v in C.unet_state_dict(shapes, sp["seed"] + 1000).items()}
    t = C.to_torch({k: v for k, v in case.items() if k != "spec"})
    out = ub.unet_forward(sd, sp["cfg"], t["x"], t["timesteps"], t["context"], mask=t["mask"])
    close(out, g["out"], atol=2e-4, what=name)


@pytest.mark.parametrize("name", list(C.CLOSS_CASES))
def test_capture_consumer_losses_oracle_vs_reference(name):
    """calc_subj_masked_bg_suppress_loss / calc_sc_rep_attn_distill_loss (ldm/util.py:1822-1918, 2047-2121)
    against the reference's own functions -- the oracle the fused capture-consumer kernels will be held to."""
    from oracle import capture_losses_oracle as cl
    case = C.build_closs_case(name)
    sp = case["spec"]
    g = load(name, case)
    t = C.to_torch({k: v for k, v in case.items() if k != "spec"})
    subj = (t["subj_ib"], t["subj_it"])
    if sp["kind"] == "bg":
        loss = cl.subj_masked_bg_suppress_loss({23: t["attn23"], 24: t["attn24"]}, subj, sp["block"], t["fg_mask"])
        assert g["loss"] > 0                       # the hinge is active in the fixture
        close(loss, g["loss"], atol=1e-6, what=name)
    else:
        acts = {"attn": {23: t["attn23"], 24: t["attn24"]}, "k": {23: t["k23"], 24: t["k24"]}, "v": {23: t["v23"], 24: t["v24"]}}
        out = cl.sc_rep_attn_distill_loss(acts, subj, t["emb_mask"], t["pad_mask"], sp["fg_percent"])
        close(torch.stack([torch.as_tensor(o, dtype=torch.float32) for o in out]), g["losses"], atol=1e-5, what=name)
        assert (g["losses"] > 0).all() == (sp["fg_percent"] >= 0.1)      # below FG_THRES every term is zero (:2075)


@pytest.mark.parametrize("name", [n for n, s in C.DDIM_CASES.items() if s["model"] == "standin"])
def test_ddim_oracle_matches_reference_sampler(name):
    """BASELINE config 4: the reference's own DDIMSampler (ldm/models/diffusion/ddim.py:70-302) around an analytic noise predictor
    pins the oracle's schedule tables BIT-exactly and its CFG combine / x0 prediction / x_{t-1} update to fp32 round-off."""
    from oracle import ddim_oracle as dd
    case = C.build_ddim_case(name)
    sp = case["spec"]
    g = load(name, case)
    ac = dd.linear_alphas_cumprod()
    assert np.array_equal(ac.numpy(), g["alphas_cumprod"])
    sched = dd.ddim_schedule(ac, sp["steps"], eta=0.0)
    assert np.array_equal(sched["timesteps"], g["ddim_timesteps"])
    assert np.array_equal(sched["alphas"].numpy(), g["ddim_alphas"])
    assert np.array_equal(np.asarray(sched["alphas_prev"], dtype=np.float64), g["ddim_alphas_prev"])
    assert np.array_equal(np.asarray(sched["sigmas"], dtype=np.float64), g["ddim_sigmas"])
    assert np.array_equal(np.asarray(sched["sqrt_one_minus_alphas"]), g["ddim_sqrt_one_minus_alphas"])
    t = C.to_torch({k: v for k, v in case.items() if k != "spec"})
    x0, pred = dd.ddim_sample(C.standin_eps, ac, t["x_T"], t["cond"], t["uncond"], sp["steps"], guidance_scale=sp["guidance"])
    assert np.array_equal(x0.numpy(), g["x0"]), "same op order as the reference => bit-identical on the CPU"
    assert np.array_equal(pred.numpy(), g["pred_x0_last"])
    if sp["steps"] == 50:
        assert list(sched["timesteps"][:3]) == [1, 21, 41] and sched["timesteps"][-1] == 981        # ddim.py:29-35


def test_arc2face_oracle_matches_reference_function():
    """SURVEY 8f row 3 (first half) PINNED: oracle.arc2face_id_to_img_prompt against the reference's own
    Arc2Face_ID2AdaPrompt.map_init_id_to_img_prompt_embs (adaface/face_id_to_ada_prompt.py:680-724), run verbatim by make_golden.py."""
    case = C.build_text_case("arc2face_id2img")
    g = load("arc2face_id2img", case)
    t = C.to_torch({k: v for k, v in case.items() if k != "spec"})
    w = t["w"]
    rows = {int(k): v for k, v in w["token_emb_rows"].items()}
    rows.update({int(k): v for k, v in t["extra_rows"].items()})
    prompt = torch.stack([rows[i] for i in C.ARC2FACE_PROMPT_IDS])
    assert list(oracle.ARC2FACE_PROMPT_IDS) == list(C.ARC2FACE_PROMPT_IDS)
    out = oracle.arc2face_id_to_img_prompt(w, t["init_id_embs"], prompt_embs=prompt)
    close(out, g["out"], atol=1e-4, what="arc2face id -> image prompt")


def test_sd_text_encoder_oracle_matches_reference_function():
    """SURVEY 8f row 3 (second half) PINNED: the oracle's CLIP loop with the [0.5, 0.5] last-layers weighting against the reference's
    patched text_model_forward / encoder_forward / embeddings_forward (ldm/modules/encoders/modules.py:180-338) with an
    EmbeddingManager-style splice of 16 ada tokens."""
    case = C.build_text_case("sd_text_encoder")
    g = load("sd_text_encoder", case)
    t = C.to_torch({k: v for k, v in case.items() if k != "spec"})
    w = t["w"]
    tok = w["template_embs"].unsqueeze(0).repeat(case["spec"]["B"], 1, 1)
    tok[:, 4:20] = t["ada"]
    out = oracle.clip_text_wrapper_forward(w, tok, torch.tensor([[0.5], [0.5]]))
    close(out, g["out"], atol=1e-4, what="SD text encoder")
