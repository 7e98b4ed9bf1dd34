"""GPU parity of the stage-2 (compositional distillation) machinery around the attention path (BASELINE config 5):
conv-LoRA training (dalc:541-591), the training wrapper over the U-Net mirror (ddpm.py:4110-4252, dalc:451-661) incl. the skip
gradient scale (dalc:382-394), and the four-instance step with the fused capture consumers (ddpm.py:1590-1740, 3470-3485)."""
import math

import numpy as np
import pytest
import torch

import cases as C
from oracle import unet_blocks_oracle as ub
from mirror_utils import _T
from parity_log import record

pytestmark = pytest.mark.gpu
GRAD_TOL = 3e-2


def rel(a, ref):
    a, ref = a.detach().float().cpu(), ref.detach().float().cpu()
    assert a.shape == ref.shape, (a.shape, ref.shape)
    return ((a - ref).abs().max() / ref.abs().max().clamp_min(1e-20)).item()


def _r16(*shape, std=1.0, seed=0):
    return (torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * std).bfloat16().float()


@pytest.mark.parametrize("k,cin,cout,r,B,h,w", [(3, 64, 128, 16, 2, 12, 16), (3, 640, 320, 192, 1, 16, 16), (1, 640, 320, 192, 1, 16, 16)])
def test_conv_lora_forward_backward_vs_oracle(k, cin, cout, r, B, h, w):
    """LoraDoraConv2d training path (ConvLoraFn): y, dX, dA, dB, d magnitude against autograd through the oracle's restatement of
    peft lora.Conv2d + DoRA (PARITY UNPINNED for peft itself, see oracle/unet_blocks_oracle.py:168)."""
    import adaface_dev_b200 as a
    W, bias = _r16(cout, cin, k, k, std=(k * k * cin) ** -0.5, seed=1), _r16(cout, std=0.02, seed=2)
    A, Bm = _r16(r, cin, k, k, std=(k * k * cin) ** -0.5, seed=3), _r16(cout, r, 1, 1, std=0.05, seed=4)
    s = 16 / r
    mag = (torch.linalg.norm((W + s * (Bm.flatten(1) @ A.flatten(1)).view_as(W)).flatten(1), dim=1) * (1 + 0.1 * _r16(cout, seed=5)))
    x = _r16(B, cin, h, w, seed=6)
    gy = _r16(B, cout, h, w, seed=7)
    # reference
    xr, Ar, Br, mr = (t.clone().requires_grad_(True) for t in (x, A, Bm, mag))
    yr = ub.lora_dora_conv(xr, W, bias, Ar, Br, mr, s)
    yr.backward(gy)
    # product
    base = torch.nn.Conv2d(cin, cout, k, padding=k // 2).cuda()
    with torch.no_grad():
        base.weight.copy_(W)
        base.bias.copy_(bias)
    base.requires_grad_(False)
    lo = a.LoraDoraConv2d(base, "default", r=r, lora_alpha=16).cuda()
    with torch.no_grad():
        lo.lora_A["default"].weight.copy_(A)
        lo.lora_B["default"].weight.copy_(Bm)
        lo.lora_magnitude_vector["default"].weight.copy_(mag)
    tok = x.permute(0, 2, 3, 1).reshape(B, h * w, cin).contiguous().bfloat16().cuda().requires_grad_(True)
    y = lo.forward_tokens(tok, (h, w))
    y.backward(gy.permute(0, 2, 3, 1).reshape(B, h * w, cout).contiguous().bfloat16().cuda())
    nchw = lambda t_, c_: t_.float().cpu().reshape(B, h, w, c_).permute(0, 3, 1, 2)
    errs = {"y": rel(nchw(y, cout), yr), "dX": rel(nchw(tok.grad, cin), xr.grad),
            "dA": rel(lo.lora_A["default"].weight.grad, Ar.grad), "dB": rel(lo.lora_B["default"].weight.grad, Br.grad),
            "dm": rel(lo.lora_magnitude_vector["default"].weight.grad, mr.grad)}
    for n_, e in errs.items():
        record("conv_lora", f"k{k}_{cin}to{cout}_r{r}", f"{n_} rel max-abs", e, GRAD_TOL)
    assert all(e < GRAD_TOL for e in errs.values()), errs
    assert base.weight.grad is None


def _small_wrapper(use_ffn_lora=False, seed=61, rank=16):
    import adaface_dev_b200 as a
    from adaface_dev_b200.unet_wrapper import DiffusersUNetWrapper
    unet = a.UNetModel(**C.UNET_CFG_SMALL).cuda().eval()
    sd = C.unet_state_dict({k_: v.shape for k_, v in unet.state_dict().items()}, seed + 1000)
    unet.load_state_dict({k_: torch.from_numpy(v) for k_, v in sd.items()})
    unet.captured_layer_indices = (6, 7, 8)          # the last up block of the small configuration holds layers 7, 8 only -> see below
    return DiffusersUNetWrapper(unet, use_attn_lora=True, use_ffn_lora=use_ffn_lora, lora_rank=rank), sd


def test_wrapper_names_flags_and_identity_at_init():
    """set_up_attn_processors / set_up_ffn_loras expose the reference's flat names; with the adapters at their init (B = 0,
    m = ||W||) the wrapped U-Net predicts what the bare mirror predicts; capture returns the processor surface's keys."""
    from adaface_dev_b200.unet_wrapper import diffusers_module_names
    w, _ = _small_wrapper(use_ffn_lora=True)
    unet = w.diffusion_model
    names = diffusers_module_names(unet)
    assert "up_blocks.1.attentions.1.transformer_blocks.0.attn2" in names and "down_blocks.0.resnets.0" in names
    procs = [m for m in unet._processor_modules()]
    assert len(procs) == 2 and len(w.attn_capture_procs) == 2          # small config: two cross-attention modules in the last up block
    unet.captured_layer_indices = (7, 8)
    keys = set(w.unet_lora_modules.keys())
    assert "up_blocks_1_attentions_1_transformer_blocks_0_attn2_processor_to_q_lora_lora_A" in keys
    assert "up_blocks_1_attentions_0_transformer_blocks_0_attn2_processor_cross_attn_scale_factor" in keys
    assert all(p.requires_grad and p.dtype == torch.float32 for p in w.unet_lora_modules.parameters())
    assert not any(p.requires_grad for n_, p in unet.named_parameters() if "lora" not in n_ and "cross_attn_scale" not in n_)
    case = C.build_unet_case("unet_small")
    x, ts, ctx = _T(case["x"]), torch.from_numpy(case["timesteps"]).cuda(), _T(case["context"])
    with torch.no_grad():
        info = {"capture_ca_activations": True, "use_attn_lora": True, "use_ffn_lora": False}
        out = w(x, ts, (ctx, None, info))
        for p_ in unet._processor_modules():                 # bare mirror: processors off
            p_._saved, p_.processor = p_.processor, None
        bare = unet(x, ts, context=ctx)
        for p_ in procs:
            p_.processor = p_._saved
    assert (out - bare).abs().max().item() < 2e-2
    acts = info["ca_layers_activations"]
    assert set(acts) >= {"outfeat", "attn", "attnscore", "q", "q2", "k", "v", "attn_out"} and set(acts["attn"]) == {7, 8}
    assert tuple(acts["attn"][8].shape) == (2, 8, 256, 77) and acts["attn"][8].dtype == torch.float32
    assert tuple(acts["k"][7].shape) == (2, 320, 77) and tuple(acts["outfeat"][8].shape) == (2, 320, 16, 16)
    assert all(not p_.capture_ca_activations and not p_.enable_lora for p_ in w.attn_capture_procs)      # restored (ddpm.py:4245-4248)


def test_wrapper_context_gradient_with_skip_gradscale_vs_oracle():
    """A7 (dalc:382-394): d loss / d prompt context through the wrapped U-Net with res_hidden_states_gradscale = 0.5 against
    autograd through the CPU oracle with the same ScaleGrad on the skip tensors (adapters at init = identity)."""
    w, sd = _small_wrapper()
    w.diffusion_model.captured_layer_indices = (7, 8)
    case = C.build_unet_case("unet_small")
    t = C.to_torch({k_: v for k_, v in case.items() if k_ != "spec"})
    ctx_r = t["context"].clone().requires_grad_(True)
    sdt = {k_: torch.from_numpy(v) for k_, v in sd.items()}
    ref = ub.unet_forward(sdt, C.UNET_CFG_SMALL, t["x"], t["timesteps"], ctx_r, res_hidden_states_gradscale=0.5)
    g = torch.randn(ref.shape, generator=torch.Generator().manual_seed(0))
    ref.backward(g)
    ref1 = ub.unet_forward(sdt, C.UNET_CFG_SMALL, t["x"], t["timesteps"], t["context"].clone().requires_grad_(True))
    ctx = _T(case["context"]).requires_grad_(True)
    info = {"capture_ca_activations": False, "res_hidden_states_gradscale": 0.5, "use_attn_lora": True}
    out = w(_T(case["x"]), torch.from_numpy(case["timesteps"]).cuda(), (ctx, None, info))
    out.backward(g.cuda())
    e = rel(ctx.grad, ctx_r.grad)
    record("stage2", "unet_small", "d context with skip gradscale 0.5, rel max-abs", e, 6e-2)
    assert e < 6e-2 and (out.detach().cpu() - ref1.detach()).abs().max().item() < 6e-2
    # the LoRA adapters of the captured layers received gradients (B is zero-initialised, so dA = 0 but dB != 0)
    gB = w.attn_capture_procs[0].to_k_lora.lora_B["default"].weight.grad
    assert gB is not None and gB.abs().max().item() > 0


def _step_inputs(S=77):
    g = torch.Generator().manual_seed(11)
    x = torch.randn(4, 4, 16, 16, generator=g).cuda()
    ts = [torch.full((4,), v, dtype=torch.long).cuda() for v in (700, 400)]
    prompt = torch.randn(4, S, 768, generator=g)
    uncond = torch.randn(1, S, 768, generator=g).expand(4, -1, -1).contiguous().cuda()
    si = (torch.zeros(16, dtype=torch.long).cuda(), torch.arange(4, 20).cuda())
    fg = torch.zeros(1, 1, 64, 64)
    fg[0, 0, 10:40, 12:36] = 1
    emb = torch.zeros(4, S, 1)
    emb[:, 1:40] = 1
    pad = torch.zeros(4, S, 1)
    pad[:, 40:] = 1
    return x, ts, prompt, uncond, si, fg.cuda(), emb.cuda(), pad.cuda()


def test_comp_distill_step_fused_equals_unfused():
    """The four-instance step (ss, sc_rep no-grad; sc grad; mc no-grad; uncond) with the capture consumers fused into the kernel
    gives the losses and the gradients (prompt rows, LoRA A / B / magnitude, cross_attn_scale_factor) of the same step run over
    the full probability maps."""
    from adaface_dev_b200.stage2 import CompDistillStep
    x, ts, prompt, uncond, si, fg, emb, pad = _step_inputs()
    res = {}
    for fused in (True, False):
        w, _ = _small_wrapper(use_ffn_lora=True)
        w.diffusion_model.captured_layer_indices = (7, 8)
        with torch.no_grad():                                        # non-trivial adapters so that every gradient is exercised
            gen = torch.Generator().manual_seed(5)
            for n_, p_ in w.unet_lora_modules.named_parameters():        # (A is Kaiming-initialised from the global RNG: seed it too)
                if "lora_B" in n_ or "lora_A" in n_:
                    p_.copy_((torch.randn(p_.shape, generator=gen) * (0.02 if "lora_B" in n_ else p_[0].numel() ** -0.5)).to(p_.device))
        step = CompDistillStep(w, fused_consumers=fused, use_ffn_lora=True)
        step.align_layers = (7, 8)
        pe = prompt.cuda().requires_grad_(True)
        totals = step.step(x, ts, pe, uncond, si, fg, emb, pad, sc_fg_mask_percent=0.3)
        res[fused] = (totals, pe.grad.clone(), {n_: p_.grad.clone() for n_, p_ in w.unet_lora_modules.named_parameters() if p_.grad is not None})
    (tf, gf, pf), (tu, gu, pu) = res[True], res[False]
    for k_ in tf:
        a_, b_ = float(tf[k_]), float(tu[k_])
        assert abs(a_ - b_) <= 2e-2 * abs(b_) + 1e-7, (k_, a_, b_)
        assert b_ > 0 or k_ == "subj_mb_suppress", k_
    assert rel(gf, gu) < GRAD_TOL
    assert gf[2:].abs().max().item() == 0 and gf[0].abs().max().item() == 0        # only the sc instance carries gradient
    assert set(pf) == set(pu) and len(pf) > 10
    errs = {n_: (rel(pf[n_], pu[n_]), pu[n_].abs().max().item()) for n_ in pf if pu[n_].abs().max() > 0}
    worst = max(e_ for e_, _ in errs.values())
    record("stage2", "comp_distill_step_small_unet", "fused vs unfused: worst LoRA-parameter gradient rel", worst, GRAD_TOL)
    assert worst < GRAD_TOL
    assert any("conv1_lora_A" in n_ for n_ in pf) and any("cross_attn_scale_factor" in n_ for n_ in pf)


def test_graphed_training_step_equals_eager():
    """graphs.graphed_step: forward + backward of the four-instance step captured into ONE CUDA graph gives the eager step's
    gradients bit for bit (same kernels, same order), also after the adapters were updated in place (the operand packs of the
    trainable parameters are rebuilt inside the graph)."""
    import adaface_dev_b200 as a
    from adaface_dev_b200.stage2 import CompDistillStep
    x, ts, prompt, uncond, si, fg, emb, pad = _step_inputs()
    w, _ = _small_wrapper(use_ffn_lora=True)
    w.diffusion_model.captured_layer_indices = (7, 8)
    gen = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for n_, p_ in w.unet_lora_modules.named_parameters():
            if "lora_B" in n_ or "lora_A" in n_:
                p_.copy_((torch.randn(p_.shape, generator=gen) * (0.02 if "lora_B" in n_ else p_[0].numel() ** -0.5)).to(p_.device))
    step = CompDistillStep(w, fused_consumers=True, use_ffn_lora=True)
    step.align_layers = (7, 8)
    params = w.trainable_parameters()
    gb = a.parallel.GradBucketer(params, expected_uses=len(ts))
    pe = prompt.cuda()
    pe_grad = torch.zeros_like(pe)

    def fwd_bwd():
        leaf = pe.detach().clone().requires_grad_(True)
        tot = step.step(x, ts, lambda: leaf * 1.0, uncond, si, fg, emb, pad, sc_fg_mask_percent=0.3)
        pe_grad.copy_(leaf.grad)
        return tot

    def run(replay=None):
        gb.zero()
        tot = replay() if replay is not None else fwd_bwd()
        gb.finish()
        return {k_: float(v_) for k_, v_ in tot.items()}, pe_grad.clone(), [None if p_.grad is None else p_.grad.clone() for p_ in params]

    t0, g0, p0 = run()
    gb.defer = True
    gb.zero()
    replay = a.graphed_step(fwd_bwd, w)
    gb.freeze_touched()
    t1, g1, p1 = run(replay)
    assert t0 == t1 and torch.equal(g0, g1)
    assert all((u is None) == (v is None) and (u is None or torch.equal(u, v)) for u, v in zip(p0, p1))
    with torch.no_grad():                                   # an "optimiser step" in place, then both paths again
        for p_ in params:
            if p_.grad is not None:
                p_.add_(p_.grad, alpha=-1e-2)
    t2, g2, p2 = run(replay)
    gb.defer = False
    gb._frozen_touched = None
    t3, g3, p3 = run()
    assert t2 != t1 and t2 == t3 and torch.equal(g2, g3)
    assert all((u is None) == (v is None) and (u is None or torch.equal(u, v)) for u, v in zip(p2, p3))
    gb.close()
