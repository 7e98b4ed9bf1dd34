"""Drop-in mirror of the reference's DDIM sampler around the U-Net (BASELINE config 4).

    DDIMSampler     ldm/models/diffusion/ddim.py:11-302   (make_schedule, sample, ddim_sampling, p_sample_ddim)
    UNetDenoiser    the slice of LatentDiffusion the sampler touches: register_schedule (ldm/models/diffusion/ddpm.py:294-315),
                    apply_model (ddpm.py -> DiffusionWrapper / DiffusersUNetWrapper.forward -> UNetModel.forward), num_timesteps,
                    betas, alphas_cumprod(_prev), device

B200 design.  A sampling run is `steps` x (one U-Net forward over the CFG-doubled batch + a handful of scalar-times-tensor
ops).  Here the CFG combine, the x0 prediction and the x_{t-1} update are ONE kernel (adaface_ddim_cfg_step) whose per-step
scalars live in a device coefficient row; it writes x_{t-1} straight into both halves of the next step's CFG batch.  U-Net +
update are captured into ONE CUDA graph per (micro-batch, cfg) shape and replayed for every step -- the 50-step loop issues
two 32-byte device copies and one graph launch per step, no host synchronisation, and only latents (32 KB / image) ever
cross PCIe.  Images are independent, so the batch is processed in micro-batches (and sharded across GPUs by the caller with
parallel.shard_range: no collective).  CUDA only, no fallback.
"""
import numpy as np
import torch

from . import _lib, ops


def make_linear_alphas_cumprod(timesteps=1000, linear_start=0.00085, linear_end=0.012):
    """make_beta_schedule('linear') (ldm/modules/diffusionmodules/util.py:22-25) + register_schedule (ddpm.py:301-314): float64
    betas / cumprod on the host, stored fp32.  Defaults: the SD-1.5 configuration (configs/stable-diffusion/v1-*.yaml:9-11)."""
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2
    alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
    return betas.float(), alphas_cumprod.float()


class UNetDenoiser(torch.nn.Module):
    """What DDIMSampler needs from the reference's LatentDiffusion: the noise schedule and ``apply_model``."""

    graph_safe = True          # apply_model launches only stream-ordered kernels: may be captured into a CUDA graph

    def __init__(self, unet, timesteps=1000, linear_start=0.00085, linear_end=0.012):
        super().__init__()
        self.model = unet
        self.parameterization = "eps"
        self.register_schedule(timesteps=timesteps, linear_start=linear_start, linear_end=linear_end)

    def register_schedule(self, given_betas=None, beta_schedule="linear", timesteps=1000, linear_start=1e-4, linear_end=2e-2,
                          cosine_s=8e-3):
        if given_betas is not None:
            betas = torch.as_tensor(given_betas, dtype=torch.float64)
            ac = torch.cumprod(1.0 - betas, dim=0).float()
            betas = betas.float()
        elif beta_schedule == "linear":
            betas, ac = make_linear_alphas_cumprod(timesteps, linear_start, linear_end)
        else:
            raise NotImplementedError(f"beta schedule {beta_schedule!r}: SD-1.5 uses 'linear'")
        self.num_timesteps = int(betas.shape[0])
        self.linear_start, self.linear_end = linear_start, linear_end
        for name, val in (("betas", betas), ("alphas_cumprod", ac), ("alphas_cumprod_prev", torch.cat([torch.ones(1), ac[:-1]]))):
            if name in self._buffers:
                self._buffers[name] = val
            else:
                self.register_buffer(name, val, persistent=False)

    @property
    def device(self):
        return next(self.model.parameters()).device

    def apply_model(self, x_noisy, t, cond, **kwargs):
        """cond: the context tensor [B, S, 768] or the reference's (context, prompt_in, extra_info) tuple (ddim.py:236-250)."""
        extra_info = None
        if isinstance(cond, (tuple, list)):
            cond, _, extra_info = cond
        return self.model(x_noisy, t, context=cond, extra_info=extra_info)


def ddim_cfg_step(eps, x, coef, *, has_uncond, noise=None, x_prev=None, x_dup=None, pred_x0=None):
    """adaface_ddim_cfg_step: eps fp32 [(2)B, ...] (cond half first), x fp32 [B, ...], coef device fp32[8] -> (x_prev, pred_x0)."""
    for t, nm in ((eps, "eps"), (x, "x"), (coef, "coef")):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise ValueError(f"ddim_cfg_step: `{nm}` must be a contiguous CUDA fp32 tensor")
    B = x.shape[0]
    per = x.numel() // B
    if eps.shape[0] != (2 * B if has_uncond else B) or eps.numel() != eps.shape[0] * per or coef.numel() < 8:
        raise ValueError(f"ddim_cfg_step: inconsistent shapes eps{tuple(eps.shape)} x{tuple(x.shape)} coef{tuple(coef.shape)}")
    x_prev = torch.empty_like(x) if x_prev is None else x_prev
    pred_x0 = torch.empty_like(x) if pred_x0 is None else pred_x0
    for t in (noise, x_prev, x_dup, pred_x0):
        if t is not None and not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == x.numel()):
            raise ValueError("ddim_cfg_step: noise / outputs must be contiguous CUDA fp32 tensors of x's size")
    p = lambda t: None if t is None else t.data_ptr()
    _lib.call("adaface_ddim_cfg_step", p(eps), B, per, 1 if has_uncond else 0, p(x), p(coef), p(noise), p(x_prev), p(x_dup), p(pred_x0),
              ops._stream())
    return x_prev, pred_x0


class DDIMSampler:
    def __init__(self, model, schedule="linear", micro_batch=None, use_cuda_graph=True):
        """model: exposes num_timesteps, alphas_cumprod, device, apply_model (UNetDenoiser or the reference's LatentDiffusion).
        micro_batch: images per U-Net call (None = the whole local batch); use_cuda_graph: capture U-Net + update once per shape."""
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        self.micro_batch = micro_batch
        self.use_cuda_graph = bool(use_cuda_graph) and getattr(model, "graph_safe", False)
        self._graphs = {}

    # ------------------------------------------------------------------------------------------------ schedule (host)
    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        """ddim.py:27-68 with util.py:46-77.  Tiny host arithmetic, same dtype flow as the reference (a_t from the fp32 table,
        a_prev / sigma through float64) so the fp32 coefficients the kernel reads equal the reference's bit for bit."""
        if ddim_discretize != "uniform":
            raise NotImplementedError("only the 'uniform' DDIM discretisation is used on this path (ddim.py:37)")
        T = self.ddpm_num_timesteps
        c = T // ddim_num_steps
        self.ddim_timesteps = np.asarray(list(range(0, T, c))) + 1                                 # util.py:48-57
        if self.ddim_timesteps[-1] >= T:
            # e.g. 3 steps of 1000: range(0, 1000, 333) + 1 ends at 1000.  The reference indexes alphas_cumprod[1000] there and dies
            # with an IndexError (util.py:66); same condition, a message that says why
            raise ValueError(f"ddim_num_steps={ddim_num_steps} does not divide the {T} training steps evenly enough: the last DDIM "
                             f"timestep would be {int(self.ddim_timesteps[-1])} >= {T} (the reference fails at util.py:66)")
        ac = self.model.alphas_cumprod.detach().float().cpu()
        if ac.shape[0] != T:
            raise ValueError("alphas_cumprod have to be defined for each timestep")
        alphas = ac[self.ddim_timesteps]                                                            # fp32
        alphas_prev = np.asarray([float(ac[0])] + ac[self.ddim_timesteps[:-1]].tolist())           # float64
        # util.py:72 verbatim in its operand types (numpy float64 array against an fp32 tensor), so eta > 0 rounds identically
        sigmas = np.asarray(ddim_eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev)), dtype=np.float64)
        self.ddim_alphas, self.ddim_alphas_prev, self.ddim_sigmas = alphas, alphas_prev, sigmas
        self.ddim_sqrt_one_minus_alphas = (1. - alphas).sqrt()                                      # ddim.py:61
        # the per-step coefficient rows of adaface_ddim_cfg_step (fp32, as torch.full narrows them, ddim.py:275-278)
        a_prev32 = torch.tensor(alphas_prev, dtype=torch.float64).float()
        sig32 = torch.tensor(np.asarray(sigmas, dtype=np.float64)).float()
        rows = torch.zeros(len(self.ddim_timesteps), 8, dtype=torch.float32)
        rows[:, 1] = self.ddim_sqrt_one_minus_alphas
        rows[:, 2] = alphas.sqrt()
        rows[:, 3] = a_prev32.sqrt()
        rows[:, 4] = (1. - a_prev32 - sig32 ** 2).sqrt()
        rows[:, 5] = sig32
        self._coef_rows = rows
        if verbose:
            print(f"Selected timesteps for ddim sampler: {self.ddim_timesteps}")

    # ------------------------------------------------------------------------------------------------ public entry
    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, guidance_scale=1., unconditional_conditioning=None,
               **kwargs):
        """ddim.py:70-131.  Returns (samples [B, C, H, W] fp32 on the device, intermediates)."""
        if quantize_x0 or mask is not None or x0 is not None or score_corrector is not None or noise_dropout:
            raise NotImplementedError("DDIMSampler: mask / x0 / quantize_x0 / score_corrector / noise_dropout are never set on this "
                                      "path by the reference's callers and are not built")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        return self.ddim_sampling(conditioning, (batch_size, C, H, W), callback=callback, img_callback=img_callback,
                                  temperature=temperature, x_T=x_T, log_every_t=log_every_t, guidance_scale=guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning, **kwargs)

    @staticmethod
    def guidance_schedule(guidance_scale, total_steps):
        """ddim.py:165-186, 213-216: a (max, min) pair anneals linearly; a scalar is clamped to >= 2 and held."""
        if isinstance(guidance_scale, (list, tuple)):
            max_g, min_g = guidance_scale
        else:
            min_g = max_g = max(2.0, guidance_scale)
        max_anneal = total_steps - 1
        delta = (max_g - min_g) / max_anneal if max_anneal > 0 else 0.0
        out, g = [], max_g
        for i in range(total_steps):
            out.append(g)
            g = g - delta if i <= max_anneal else 1
        return out

    @torch.no_grad()
    def ddim_sampling(self, cond_context, shape, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                      img_callback=None, log_every_t=100, temperature=1., guidance_scale=1., unconditional_conditioning=None,
                      generator=None, **kwargs):
        """ddim.py:133-220."""
        if ddim_use_original_steps or timesteps is not None:
            raise NotImplementedError("DDIMSampler: ddim_use_original_steps / timesteps subsets are not used on this path")
        dev = self.model.device
        b = shape[0]
        img = torch.randn(shape, device=dev, generator=generator) if x_T is None else x_T.to(dev, torch.float32)
        img = img.contiguous().clone()
        time_range = np.flip(self.ddim_timesteps)                                                   # 981, 961, ..., 1
        total = int(time_range.shape[0])
        gs = self.guidance_schedule(guidance_scale, total)
        extra = None
        if isinstance(cond_context, (tuple, list)):                                                 # (c, prompt_in, extra_info)
            cond, prompt_in_c, extra = cond_context
            unc = None if unconditional_conditioning is None else unconditional_conditioning[0]
            prompt_in_u = None if unconditional_conditioning is None else unconditional_conditioning[1]
        else:
            cond, unc, prompt_in_c, prompt_in_u = cond_context, unconditional_conditioning, None, None
        # device tables: timestep of every loop iteration and its coefficient row (row index = total - i - 1, ddim.py:192)
        coef_tab = self._coef_rows.flip(0).clone()
        coef_tab[:, 0] = torch.tensor(gs, dtype=torch.float32)
        coef_tab[:, 6] = float(temperature)
        coef_tab = coef_tab.to(dev)
        ts_tab = torch.tensor(np.ascontiguousarray(time_range), dtype=torch.long, device=dev)
        sigma_any = bool((self._coef_rows[:, 5] != 0).any())
        inter = {"x_inter": [img.clone()], "pred_x0": [img.clone()]}
        pred_all = torch.empty_like(img)
        mb = b if not self.micro_batch else max(1, min(int(self.micro_batch), b))
        log_idx = [i for i in range(total) if (total - i - 1) % log_every_t == 0 or (total - i - 1) == total - 1]
        logs = {i: (torch.empty_like(img), torch.empty_like(img)) for i in log_idx}
        for lo in range(0, b, mb):
            hi = min(b, lo + mb)
            n = hi - lo
            c_mb = cond[lo:hi]
            u_mb = None if unc is None else unc[lo:hi]
            st = self._state(n, tuple(shape[1:]), c_mb, u_mb, dev)
            st["x2"][:n].copy_(img[lo:hi])
            if u_mb is not None:
                st["x2"][n:].copy_(img[lo:hi])
                st["ctx2"][:n].copy_(c_mb)
                st["ctx2"][n:].copy_(u_mb)
            else:
                st["ctx2"].copy_(c_mb)
            for i in range(total):
                cfg = u_mb is not None and gs[i] != 1.
                st["ts2"].copy_(ts_tab[i].expand_as(st["ts2"]))
                st["coef"].copy_(coef_tab[i])
                noise = None
                if sigma_any:
                    noise = torch.randn((n,) + tuple(shape[1:]), device=dev, generator=generator)   # noise_like (:288)
                self._step(st, n, cfg, extra, prompt_in_c, prompt_in_u, lo, hi, noise)
                if i in logs:
                    logs[i][0][lo:hi].copy_(st["x2"][:n])
                    logs[i][1][lo:hi].copy_(st["pred"])
                if lo == 0 and callback:
                    callback(i)
                if img_callback:
                    img_callback(st["pred"], i)
            img[lo:hi].copy_(st["x2"][:n])
            pred_all[lo:hi].copy_(st["pred"])
        for i in log_idx:
            inter["x_inter"].append(logs[i][0])
            inter["pred_x0"].append(logs[i][1])
        self.last_pred_x0 = pred_all
        return img, inter

    # ------------------------------------------------------------------------------------------------ one step
    def _state(self, n, chw, c_mb, u_mb, dev):
        """Static buffers of one micro-batch shape: the CFG-doubled latent batch, timesteps, contexts, coefficient row."""
        key = (n, chw, tuple(c_mb.shape[1:]), c_mb.dtype, u_mb is not None)
        st = self._graphs.get(key)
        if st is None:
            k = 2 if u_mb is not None else 1
            st = {"x2": torch.zeros((k * n,) + chw, device=dev, dtype=torch.float32),
                  "ts2": torch.zeros(k * n, device=dev, dtype=torch.long),
                  "ctx2": torch.zeros((k * n,) + tuple(c_mb.shape[1:]), device=dev, dtype=c_mb.dtype),
                  "coef": torch.zeros(8, device=dev, dtype=torch.float32),
                  "pred": torch.zeros((n,) + chw, device=dev, dtype=torch.float32), "graph": {}}
            self._graphs[key] = st
        return st

    def _body(self, st, n, cfg, extra, noise):
        """p_sample_ddim (ddim.py:223-302) on the static buffers: U-Net over [cond.., uncond..] (or the cond half only when
        guidance is off), then the fused CFG combine + update, written back into BOTH halves of the latent batch."""
        k = 2 if st["x2"].shape[0] == 2 * n else 1
        if cfg:
            x_in, t_in, c_in = st["x2"], st["ts2"], st["ctx2"]
        else:
            x_in, t_in, c_in = st["x2"][:n], st["ts2"][:n], st["ctx2"][:n]
        cond = c_in if extra is None else (c_in, None, extra)
        eps = self.model.apply_model(x_in, t_in, cond)
        eps = eps.float().contiguous()
        ddim_cfg_step(eps, st["x2"][:n], st["coef"], has_uncond=cfg, noise=noise, x_prev=st["x2"][:n],
                      x_dup=st["x2"][n:] if k == 2 else None, pred_x0=st["pred"])

    def _step(self, st, n, cfg, extra, prompt_in_c, prompt_in_u, lo, hi, noise):
        graphable = self.use_cuda_graph and noise is None and (extra is None or not extra.get("capture_ca_activations", False))
        if not graphable:
            return self._body(st, n, cfg, extra, noise)
        g = st["graph"].get(cfg)
        if g is None:
            keep = {k: st[k].clone() for k in ("x2", "pred")}
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):                      # warm-up: weight packs, workspaces and kernel attributes exist before capture
                    self._body(st, n, cfg, extra, None)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._body(st, n, cfg, extra, None)
            for k_, v in keep.items():                  # warm-up and capture advanced the latents: restore them
                st[k_].copy_(v)
            st["graph"][cfg] = g
        g.replay()

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False, temperature=1.,
                      noise_dropout=0., score_corrector=None, corrector_kwargs=None, guidance_scale=1., unconditional_conditioning=None):
        """ddim.py:223-302 for callers that drive single steps themselves (eager, no graph).  Returns (x_prev, pred_x0)."""
        if use_original_steps or quantize_denoised or noise_dropout or score_corrector is not None:
            raise NotImplementedError("p_sample_ddim: only the options the reference's sampler loop uses are built")
        b = x.shape[0]
        x = x.float().contiguous()
        cfg = unconditional_conditioning is not None and guidance_scale != 1.
        if cfg:
            if isinstance(c, (tuple, list)):
                c_in = (torch.cat([c[0], unconditional_conditioning[0]]), sum([c[1], unconditional_conditioning[1]], []), c[2])
            else:
                c_in = torch.cat([c, unconditional_conditioning])
            eps = self.model.apply_model(torch.cat([x] * 2), torch.cat([t] * 2), c_in)
        else:
            eps = self.model.apply_model(x, t, c)
        coef = self._coef_rows[index].clone()
        coef[0], coef[6] = float(guidance_scale), float(temperature)
        noise = None
        if float(coef[5]) != 0:
            noise = torch.randn_like(x[:1]).expand_as(x).contiguous() if repeat_noise else torch.randn_like(x)
        return ddim_cfg_step(eps.float().contiguous(), x, coef.to(x.device), has_uncond=cfg, noise=noise)
