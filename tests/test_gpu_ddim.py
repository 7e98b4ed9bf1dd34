"""GPU parity of the DDIM sampling loop (BASELINE config 4; ldm/models/diffusion/ddim.py:70-302):
  (1) the fused CFG-combine + update kernel against the CPU oracle, BIT-exact (separately rounded fp32 ops in the reference's order);
  (2) DDIMSampler (CUDA-graph replay, micro-batches) around an analytic noise predictor against the fixtures produced by the
      reference's own DDIMSampler -- the only difference is the GPU's tanh;
  (3) DDIMSampler around the U-Net mirror against the reference sampler around the reference U-Net (fixture ddim_unet_small);
  (4) size-independent properties at the full 64 x 64 latent size: micro-batching and CFG-pair sharding never change a sample."""
import os

import numpy as np
import pytest
import torch

import cases as C
from oracle import ddim_oracle as dd
from mirror_utils import _T
from parity_log import record

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


class StandIn:
    """LatentDiffusion stand-in around the analytic predictor of the fixtures."""
    graph_safe = True
    num_timesteps = 1000

    def __init__(self):
        import adaface_dev_b200 as a
        self.betas, self.alphas_cumprod = a.make_linear_alphas_cumprod()
        self.device = torch.device("cuda")

    def apply_model(self, x, t, c):
        return C.standin_eps(x, t, c)


@pytest.mark.parametrize("cfg,noise", [(True, False), (False, False), (True, True)])
def test_ddim_step_kernel_bit_exact_vs_oracle(cfg, noise):
    import adaface_dev_b200 as a
    g = torch.Generator().manual_seed(5)
    B = 3
    x = torch.randn(B, 4, 16, 16, generator=g)
    eps = torch.randn((2 if cfg else 1) * B, 4, 16, 16, generator=g)
    nz = torch.randn(B, 4, 16, 16, generator=g) if noise else None
    sched = dd.ddim_schedule(dd.linear_alphas_cumprod(), 50, eta=0.7 if noise else 0.0)
    smp = a.DDIMSampler(StandIn())
    smp.make_schedule(50, ddim_eta=0.7 if noise else 0.0, verbose=False)
    for index in (0, 17, 49):
        coef = dd.step_coefficients(sched, index)
        e_t = dd.cfg_combine(*eps.chunk(2), 4.5) if cfg else eps
        x_ref, p_ref = dd.ddim_update(x, e_t, coef, noise=nz, temperature=0.9)
        row = smp._coef_rows[index].clone()
        row[0], row[6] = 4.5, 0.9
        x_prev, pred = a.ddim_cfg_step(eps.cuda(), x.cuda(), row.cuda(), has_uncond=cfg, noise=None if nz is None else nz.cuda())
        assert torch.equal(x_prev.cpu(), x_ref) and torch.equal(pred.cpu(), p_ref), index


@pytest.mark.parametrize("name", [n for n, s in C.DDIM_CASES.items() if s["model"] == "standin"])
@pytest.mark.parametrize("graph,mb", [(True, None), (False, 1)])
def test_ddim_sampler_vs_reference_sampler(name, graph, mb):
    import adaface_dev_b200 as a
    case = C.build_ddim_case(name)
    sp = case["spec"]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    smp = a.DDIMSampler(StandIn(), micro_batch=mb, use_cuda_graph=graph)
    gs = list(sp["guidance"]) if isinstance(sp["guidance"], tuple) else sp["guidance"]
    x0, inter = smp.sample(sp["steps"], sp["B"], (4, sp["h"], sp["w"]), conditioning=_T(case["cond"]), eta=0., verbose=False,
                           x_T=_T(case["x_T"]), guidance_scale=gs, unconditional_conditioning=_T(case["uncond"]), log_every_t=1)
    e = (x0.cpu() - torch.from_numpy(g["x0"])).abs().max().item()
    e1 = (inter["x_inter"][1].cpu() - torch.from_numpy(g["x_after_first"])).abs().max().item()
    ep = (smp.last_pred_x0.cpu() - torch.from_numpy(g["pred_x0_last"])).abs().max().item()
    record("ddim", name, f"x0 after {sp['steps']} steps (graph={graph}, mb={mb})", e, 1e-4)
    assert e < 1e-4 and e1 < 1e-5 and ep < 1e-4, (e, e1, ep)       # fp32 throughout; only tanh differs (GPU vs CPU libm)
    assert len(inter["x_inter"]) == sp["steps"] + 1


def _small_unet(seed):
    import adaface_dev_b200 as a
    m = a.UNetModel(**C.UNET_CFG_SMALL).cuda().eval()
    sd = C.unet_state_dict({k: v.shape for k, v in m.state_dict().items()}, seed + 1000)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return m


def test_ddim_unet_sampler_vs_reference():
    """4 DDIM steps with CFG 3 around the U-Net: reference DDIMSampler + reference UNetModel (fp32) vs the mirrors (bf16 U-Net)."""
    import adaface_dev_b200 as a
    name = "ddim_unet_small"
    case = C.build_ddim_case(name)
    sp = case["spec"]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    model = a.UNetDenoiser(_small_unet(sp["seed"]))
    assert np.array_equal(model.alphas_cumprod.cpu().numpy(), g["alphas_cumprod"])
    outs = []
    for graph in (True, False):
        smp = a.DDIMSampler(model, use_cuda_graph=graph)
        x0, inter = smp.sample(sp["steps"], sp["B"], (4, sp["h"], sp["w"]), conditioning=_T(case["cond"]), eta=0., verbose=False,
                               x_T=_T(case["x_T"]), guidance_scale=sp["guidance"], unconditional_conditioning=_T(case["uncond"]), log_every_t=1)
        assert np.array_equal(smp.ddim_timesteps, g["ddim_timesteps"])
        outs.append(x0)
        ref = torch.from_numpy(g["x0"])
        e = (x0.cpu() - ref).abs().max().item()
        e1 = (inter["x_inter"][1].cpu() - torch.from_numpy(g["x_after_first"])).abs().max().item()
        rel = ((x0.cpu() - ref).norm() / ref.norm()).item()
        # The update amplifies the U-Net's bf16 error: eps is multiplied by the CFG scale (3) and by sqrt(1-a_t)/sqrt(a_t) (4.4 at
        # t = 751), and with random weights the latents grow to max-abs 21 (std 5.3).  The bars are therefore RELATIVE: max-abs
        # error over max-abs of the reference, and relative L2, both under north_star's 2e-2.
        ref1 = torch.from_numpy(g["x_after_first"])
        record("ddim", name, f"x0 after 4 U-Net steps, CFG 3 (graph={graph}): max-abs err / max-abs ref", e / ref.abs().max().item(), 2e-2,
               f"abs {e:.3f} on max-abs {ref.abs().max().item():.1f}; rel-L2 {rel:.2e}; after 1 step {e1 / ref1.abs().max().item():.2e}")
        assert e < 2e-2 * ref.abs().max().item() and rel < 2e-2 and e1 < 2e-2 * ref1.abs().max().item(), (e, rel, e1)
    assert torch.equal(outs[0], outs[1])            # graph replay == eager launches, bit for bit


def test_ddim_full_size_batching_and_sharding_invariance():
    """64 x 64 latents: a sample does not depend on which micro-batch / rank it is processed in (no cross-sample op anywhere in
    the U-Net: GroupNorm / LayerNorm are per sample), so sharding the CFG batch needs no collective."""
    import adaface_dev_b200 as a
    model = a.UNetDenoiser(_small_unet(84))
    g = torch.Generator().manual_seed(1)
    n = 4
    xT = torch.randn(n, 4, 64, 64, generator=g).cuda()
    c = (torch.randn(n, 77, 768, generator=g)).bfloat16().cuda()
    u = (torch.randn(n, 77, 768, generator=g)).bfloat16().cuda()
    full, _ = a.DDIMSampler(model).sample(4, n, (4, 64, 64), conditioning=c, x_T=xT, guidance_scale=4.0,
                                          unconditional_conditioning=u, verbose=False)
    mb2, _ = a.DDIMSampler(model, micro_batch=2).sample(4, n, (4, 64, 64), conditioning=c, x_T=xT, guidance_scale=4.0,
                                                        unconditional_conditioning=u, verbose=False)
    parts = []
    for rank in range(2):                                 # what each of two ranks would run
        b, e = a.parallel.shard_range(n, rank, 2)
        x_r, _ = a.DDIMSampler(model).sample(4, e - b, (4, 64, 64), conditioning=c[b:e], x_T=xT[b:e], guidance_scale=4.0,
                                             unconditional_conditioning=u[b:e], verbose=False)
        parts.append(x_r)
    assert torch.isfinite(full).all()
    # same U-Net call shapes (two images x CFG per call) => the SAME kernels and tile schedules => bit-identical samples, whether the
    # pairs are micro-batches of one rank or the shards of two ranks
    assert torch.equal(mb2, torch.cat(parts))
    # a different batch per call changes split-K / tile schedules, i.e. bf16 rounding, which 4 steps at CFG 4 amplify on a
    # random-weight U-Net (latents reach max-abs ~12): equal to bf16 noise, relative to the sample's scale
    scale = full.abs().max().item()
    e = (full - mb2).abs().max().item() / scale
    record("ddim", "batching_invariance_64x64", "whole batch vs micro-batches: max-abs diff / max-abs", e, 8e-2)
    assert e < 8e-2 and ((full - mb2).norm() / full.norm()).item() < 2e-2
