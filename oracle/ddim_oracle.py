"""CPU oracle (numpy / torch fp32) for the DDIM sampling loop around the U-Net (BASELINE config 4).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Functional restatement of
  * ldm/modules/diffusionmodules/util.py:21-43   make_beta_schedule("linear")  (SD-1.5: linear_start 0.00085, linear_end 0.012,
                                                 configs/stable-diffusion/v1-distill-*.yaml:9-11)
  * ldm/models/diffusion/ddpm.py:294-315         register_schedule: alphas_cumprod (float64 cumprod, stored fp32)
  * ldm/modules/diffusionmodules/util.py:46-77   make_ddim_timesteps("uniform"), make_ddim_sampling_parameters
  * ldm/models/diffusion/ddim.py:27-68           DDIMSampler.make_schedule
  * ldm/models/diffusion/ddim.py:133-220         ddim_sampling: reversed time range, guidance annealing
  * ldm/models/diffusion/ddim.py:223-302         p_sample_ddim: CFG batch [cond.., uncond..], combine, x0 prediction, x_{t-1}
Pinned against the reference's own DDIMSampler driven by a stand-in model (tests/golden/make_golden.py, family "ddim";
fixtures ddim_*.npz).  The dtype flow of the reference is kept on purpose (fp32 tensors for a_t, float64 numpy for a_prev
and sigma which torch.full narrows to fp32) so that the coefficient tables agree bit for bit.
"""
import numpy as np
import torch


def linear_alphas_cumprod(n_timestep=1000, linear_start=0.00085, linear_end=0.012):
    """util.py:22-25 + ddpm.py:301-314: betas = linspace(sqrt(s), sqrt(e), T, float64)^2; alphas_cumprod = cumprod(1 - betas) in
    float64, narrowed to an fp32 tensor."""
    betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2).numpy()
    return torch.tensor(np.cumprod(1.0 - betas, axis=0), dtype=torch.float32)


def ddim_timesteps(num_ddim_steps, num_ddpm_steps=1000):
    """util.py:46-60, 'uniform': range(0, T, T // S) + 1 (50 steps -> 1, 21, ..., 981)."""
    c = num_ddpm_steps // num_ddim_steps
    return np.asarray(list(range(0, num_ddpm_steps, c))) + 1


def ddim_schedule(alphas_cumprod, num_ddim_steps, eta=0.0):
    """ddim.py:27-68 + util.py:63-77.  Returns dict(timesteps, alphas [fp32 tensor], alphas_prev [float64 ndarray],
    sigmas, sqrt_one_minus_alphas [fp32 tensor])."""
    ts = ddim_timesteps(num_ddim_steps, alphas_cumprod.shape[0])
    ac = alphas_cumprod.cpu()
    alphas = ac[ts]                                                                   # util.py:66  (fp32 tensor)
    alphas_prev = np.asarray([ac[0]] + ac[ts[:-1]].tolist())                          # util.py:67  (float64 ndarray)
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))     # util.py:72
    return dict(timesteps=ts, alphas=alphas, alphas_prev=alphas_prev, sigmas=sigmas,
                sqrt_one_minus_alphas=np.sqrt(1. - alphas))                           # ddim.py:61


def step_coefficients(sched, index):
    """The four per-step scalars of p_sample_ddim as the fp32 values torch.full produces (ddim.py:275-278)."""
    f32 = lambda v: torch.full((1,), v).float()[0] if not torch.is_tensor(v) else v.float().reshape(())
    return dict(a_t=f32(sched["alphas"][index]), a_prev=f32(sched["alphas_prev"][index]), sigma_t=f32(sched["sigmas"][index]),
                sqrt_one_minus_at=f32(sched["sqrt_one_minus_alphas"][index]))


def cfg_combine(e_cond, e_uncond, guidance_scale):
    """ddim.py:253-255."""
    return e_uncond + guidance_scale * (e_cond - e_uncond)


def ddim_update(x, e_t, coef, noise=None, temperature=1.0):
    """ddim.py:280-301: x0 prediction, direction to x_t, x_{t-1}.  All fp32, one op per line as in the reference."""
    a_t, a_prev, sigma_t, s1m = (coef[k].reshape(1, 1, 1, 1) for k in ("a_t", "a_prev", "sigma_t", "sqrt_one_minus_at"))
    pred_x0 = (x - s1m * e_t) / a_t.sqrt()
    dir_xt = (1. - a_prev - sigma_t ** 2).sqrt() * e_t
    nz = sigma_t * (torch.zeros_like(x) if noise is None else noise) * temperature
    x_prev = a_prev.sqrt() * pred_x0 + dir_xt + nz
    return x_prev, pred_x0


def guidance_schedule(guidance_scale, total_steps):
    """ddim.py:165-186, 213-216: the scale used at loop iteration i (linear annealing from max to min when a pair is given;
    a scalar is clamped to >= 2 and held)."""
    if isinstance(guidance_scale, (list, tuple)):
        max_g, min_g = guidance_scale
    else:
        min_g = max_g = max(2.0, guidance_scale)
    max_anneal = total_steps - 1
    delta = (max_g - min_g) / max_anneal
    out, g = [], max_g
    for i in range(total_steps):
        out.append(g)
        g = g - delta if i <= max_anneal else 1
    return out


def ddim_sample(apply_model, alphas_cumprod, x_T, cond, uncond, num_steps, guidance_scale=1.0, eta=0.0):
    """ddim.py:133-220 (mask / x0 / correctors absent, as at every reference call site on this path).
    apply_model(x [2B or B,4,h,w], t [.] long, context) -> eps.  Returns (x_0 estimate after the last step, last pred_x0)."""
    sched = ddim_schedule(alphas_cumprod, num_steps, eta)
    time_range = np.flip(sched["timesteps"])
    total = time_range.shape[0]
    gs = guidance_schedule(guidance_scale, total)
    img, pred_x0 = x_T, x_T
    b = x_T.shape[0]
    for i, step in enumerate(time_range):
        index = total - i - 1
        ts = torch.full((b,), int(step), dtype=torch.long)
        if uncond is None or gs[i] == 1.:
            e_t = apply_model(img, ts, cond)
        else:                                                                        # ddim.py:231-255
            e_c, e_u = apply_model(torch.cat([img] * 2), torch.cat([ts] * 2), torch.cat([cond, uncond])).chunk(2)
            e_t = cfg_combine(e_c, e_u, gs[i])
        img, pred_x0 = ddim_update(img, e_t, step_coefficients(sched, index))
    return img, pred_x0
