"""The stage-2 (compositional distillation) step as the hot path sees it (BASELINE config 5).

    guided_denoise, batch_part_has_grad='subject-compos'   ldm/models/diffusion/ddpm.py:1590-1740
    comp_distill_multistep_denoise                         ldm/models/diffusion/ddpm.py:2421-2430 (4 denoising steps per iteration)
    loss assembly from the captured activations            ldm/models/diffusion/ddpm.py:3470-3485, 3704-3709

One denoising step = four SLICED U-Net calls at B = 1 on the (ss, sc, sc_rep, mc) instances of one subject -- ss and sc_rep
under no_grad, then sc WITH grad (attention LoRA on, normalize_cross_attn), then mc under no_grad without LoRA -- all with
capture on layers 22-24, plus one no-grad call on the 4 unconditional prompts for the CFG'd reconstruction.  The loss is the sum
of the two consumers of the captured maps (subject-mass background suppression, sc vs sc_rep distillation): its backward runs
through the sc instance into the LoRA / DoRA adapters, ``cross_attn_scale_factor`` and -- through rows 4:20 of the prompt -- the
SubjBasisGenerator.

B200 specifics: the sc call's probability maps of layers 23 / 24 are reduced INSIDE the capture kernel against the sc_rep maps
(``set_capture_consumers``), the ss / mc calls do not write their maps at all (no consumer reads them), the U-Net weights are
frozen bf16 packs, and the data-parallel gradient all-reduce (parallel.GradBucketer, NCCL) overlaps the SubjBasisGenerator's
backward with the LoRA buckets.  Everything the reference's iteration does OUTSIDE this path -- face detection, ArcFace
alignment, the flow-based fg / bg preserve loss, optimiser, logging -- is out of scope.
"""
import torch

from . import capture_losses as closs
from .attn_processor import gen_gradient_scaler  # noqa: F401  (re-export for callers that build their own steps)


class CompDistillStep:
    def __init__(self, wrapper, fused_consumers=True, normalize_cross_attn=True, res_hidden_states_gradscale=0.5,
                 use_ffn_lora=False, ffn_lora_adapter_name="comp_distill", cfg_scale=2.5):
        """wrapper: unet_wrapper.DiffusersUNetWrapper.  res_hidden_states_gradscale: ddpm.py:140 (0.5)."""
        self.w = wrapper
        self.fused = fused_consumers
        self.normalize = normalize_cross_attn
        self.gradscale = res_hidden_states_gradscale
        self.use_ffn_lora, self.ffn_adapter = use_ffn_lora, ffn_lora_adapter_name
        self.cfg_scale = cfg_scale
        self.layers = tuple(wrapper.diffusion_model.captured_layer_indices)
        self.align_layers = self.layers[-2:]                      # 23, 24: the layers whose maps the two losses read (ldm/util.py:1839, 2058)

    def _consumers(self, mode, ref=None):
        """Per-call configuration of the three capture processors.  mode: 'maps' (write the full maps: the sc_rep instance),
        'none' (no consumer reads this instance's maps: ss, mc), 'sc' (reduce against ``ref`` = {layer: sc_rep map})."""
        for proc, li in zip(self.w.attn_capture_procs, self.layers):
            if not self.fused or mode == "maps":
                proc.set_capture_consumers()
            elif mode == "none" or li not in self.align_layers:
                proc.set_capture_consumers(subj_sum=True)          # [1,8,N] instead of two [1,8,N,S] maps
            else:
                proc.set_capture_consumers(subj_sum=True, ref_attn=ref[li])

    def _call(self, x, t, prompt_emb, idx, grad, subj_indices, normalize, use_attn_lora, use_ffn_lora):
        """sliced_apply_model (ddpm.py:1572-1587)."""
        info = {"capture_ca_activations": True, "normalize_cross_attn": normalize, "mix_attn_mats_in_batch": False,
                "subj_indices": subj_indices, "res_hidden_states_gradscale": self.gradscale, "use_attn_lora": use_attn_lora,
                "use_ffn_lora": use_ffn_lora, "ffn_lora_adapter_name": self.ffn_adapter if use_ffn_lora else None, "img_mask": None}
        with torch.set_grad_enabled(grad):
            sl = slice(idx, idx + 1)             # (a slice, not a list index: no host index tensor, CUDA-graph-capturable)
            out = self.w(x[sl], t[sl], (prompt_emb[sl], None, info))
        return out, info["ca_layers_activations"]

    def denoise(self, x_noisy, t, prompt_emb, uncond_emb, subj_indices_1b):
        """One 'subject-compos' denoising step.  x_noisy [4,4,h,w], t [4], prompt_emb [4,S,768] ordered (ss, sc, sc_rep, mc),
        uncond_emb [4,S,768].  Returns (noise_pred [4,4,h,w], acts = {'ss','sc','sr','mc': ca_layers_activations}, uncond pred)."""
        ss, sc, sr, mc = 0, 1, 2, 3
        self._consumers("none")
        n_ss, a_ss = self._call(x_noisy, t, prompt_emb, ss, False, subj_indices_1b, False, True, self.use_ffn_lora)
        self._consumers("maps")
        n_sr, a_sr = self._call(x_noisy, t, prompt_emb, sr, False, subj_indices_1b, self.normalize, True, self.use_ffn_lora)
        self._consumers("sc", ref=a_sr["attn"])
        n_sc, a_sc = self._call(x_noisy, t, prompt_emb, sc, True, subj_indices_1b, self.normalize, True, self.use_ffn_lora)
        self._consumers("none")
        n_mc, a_mc = self._call(x_noisy, t, prompt_emb, mc, False, subj_indices_1b, False, False, False)
        self._consumers("maps")
        with torch.no_grad():                                       # CFG partner of the pixel reconstruction (ddpm.py:1741-1760)
            n_un = self.w(x_noisy, t, (uncond_emb, None, {"capture_ca_activations": False, "use_attn_lora": False, "use_ffn_lora": False}))
        noise_pred = torch.cat([n_ss, n_sc, n_sr, n_mc], dim=0)
        return noise_pred, {"ss": a_ss, "sc": a_sc, "sr": a_sr, "mc": a_mc}, n_un

    def losses(self, acts, subj_indices_1b, sc_fg_mask, prompt_emb_mask_4b, prompt_pad_mask_4b, sc_fg_mask_percent):
        """The two consumers of the captured maps (ddpm.py:3470-3485, 3704-3709) on this step's activations."""
        a_ss, a_sc, a_sr, a_mc = acts["ss"], acts["sc"], acts["sr"], acts["mc"]
        L = self.align_layers
        cat4 = lambda key: {li: torch.cat([a_ss[key][li], a_sc[key][li], a_sr[key][li], a_mc[key][li]], dim=0) for li in L}
        if self.fused:
            sums = {li: a_sc["attn_subj_sum"][li] for li in L}
            sq = {li: a_sc["attn_sqdiff"][li] for li in L}
            shp = {li: tuple(a_sr["attn"][li].shape[1:]) for li in L}
        else:                                                       # the same reductions over the full maps (comparison path)
            ib, it = subj_indices_1b
            sums, sq, shp = {}, {}, {}
            for li in L:
                p_sc, p_sr = a_sc["attn"][li], a_sr["attn"][li]
                flag = torch.zeros(p_sc.shape[0], p_sc.shape[3], device=p_sc.device)
                flag.index_put_((ib.long(), it.long()), torch.ones(ib.numel(), device=p_sc.device))
                sums[li] = (p_sc * flag[:, None, None, :]).sum(-1)
                sq[li] = ((p_sc - p_sr.detach()) ** 2).sum().reshape(1)
                shp[li] = tuple(p_sc.shape[1:])
        lws = {li: 1.0 for li in L}
        l_bg = closs.calc_subj_masked_bg_suppress_loss(sums, subj_indices_1b, 1, sc_fg_mask, layer_weights=lws)
        l_attn, l_sk, l_nk, l_sv, l_nv = closs.calc_sc_rep_attn_distill_loss(sq, shp, cat4("k"), cat4("v"), subj_indices_1b,
                                                                            prompt_emb_mask_4b, prompt_pad_mask_4b, sc_fg_mask_percent,
                                                                            layer_weights=lws)
        return {"subj_mb_suppress": l_bg, "rep_distill_attn": l_attn, "rep_distill_subj_k": l_sk, "rep_distill_nonsubj_k": l_nk,
                "rep_distill_subj_v": l_sv, "rep_distill_nonsubj_v": l_nv}

    def step(self, x_noisy, ts, prompt_emb, uncond_emb, subj_indices_1b, sc_fg_mask, prompt_emb_mask_4b, prompt_pad_mask_4b,
             sc_fg_mask_percent=0.3, loss_weights=None):
        """The whole iteration on this path: ``len(ts)`` denoising steps (reference: 4), each followed by the backward of its
        losses (gradients accumulate in the trainable parameters and in whatever ``prompt_emb`` was computed from).  ts: list of
        [4] timestep tensors.  Returns the summed loss terms (detached floats on the device)."""
        lw = loss_weights or {}
        totals = {}
        for t in ts:
            # ``prompt_emb`` may be a callable returning this step's [4,S,768] prompt (so that a prompt assembled from a
            # trainable encoder's output gets a fresh, shallow graph per denoising step and no graph is retained)
            pe = prompt_emb() if callable(prompt_emb) else prompt_emb
            _, acts, _ = self.denoise(x_noisy, t, pe, uncond_emb, subj_indices_1b)
            terms = self.losses(acts, subj_indices_1b, sc_fg_mask, prompt_emb_mask_4b, prompt_pad_mask_4b, sc_fg_mask_percent)
            loss = sum(v * lw.get(k, 1.0) for k, v in terms.items() if torch.is_tensor(v))
            if torch.is_tensor(loss) and loss.requires_grad:
                loss.backward(retain_graph=(not callable(prompt_emb)) and len(ts) > 1 and pe.requires_grad and not pe.is_leaf)
            for k, v in terms.items():
                totals[k] = totals.get(k, 0) + (v.detach() if torch.is_tensor(v) else v)
        return totals
