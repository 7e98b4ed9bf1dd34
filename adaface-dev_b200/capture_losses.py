"""Consumers of the captured cross-attention activations on the REDUCED quantities the fused capture kernel delivers
(AttnProcessor_LoRA_Capture.set_capture_consumers / ops.attention_cross_consume).

    calc_subj_masked_bg_suppress_loss   ldm/util.py:1822-1918  -- from attn_subj_sum {layer: [B, H, N]} instead of the [B,H,N,S] maps
    calc_sc_rep_attn_distill_loss       ldm/util.py:2047-2121  -- probability term from attn_sqdiff {layer: [1]}, k / v terms from
                                                                  the captured k, v [4, C, S] (a few KB: plain tensor code)

The O(B * H * N * S) work -- the subject-column sum and the sc vs sc_rep squared difference -- happened inside the attention
kernel (forward and backward); what is left here is O(B * H * N) masking / hinge arithmetic and O(C * S) MSEs, written as
tensor expressions so that autograd carries the upstream scalars back to ``attn_subj_sum`` / ``attn_sqdiff`` / k / v.
"""
import torch
import torch.nn.functional as F

ALIGN_LAYERS = (23, 24)          # attn_align_layer_weights = subj_comp_rep_distill_layer_weights = {23: 1, 24: 1}, normalised (:1839, :2058)


def _layer_weights(layer_weights=None):
    """normalize_dict_values (ldm/util.py:1088-1095) of the per-layer weights; default: the reference's {23: 1, 24: 1}."""
    lw = {li: 1.0 for li in ALIGN_LAYERS} if layer_weights is None else dict(layer_weights)
    tot = float(sum(lw.values()))
    return {li: v / tot for li, v in lw.items()}


def resize_mask_to_target_size(mask, area):
    """ldm/util.py:1333-1360 with mode 'nearest|bilinear' on square maps: element-wise max of the two interpolations."""
    side = int(area ** 0.5)
    near = F.interpolate(mask.float(), size=(side, side), mode="nearest")
    bil = F.interpolate(mask.float(), size=(side, side), mode="bilinear", align_corners=False)
    return torch.maximum(near, bil)


def calc_subj_masked_bg_suppress_loss(attn_subj_sum, subj_indices, BLOCK_SIZE, fg_mask, bg_attn_tolerance=0.02, layer_weights=None):
    """ldm/util.py:1822-1918 on the kernel-reduced maps.  attn_subj_sum: {layer: [>= BLOCK_SIZE, H, N]} = probability mass on each
    instance's subject columns (the reference computes it as sel_emb_attns_by_indices(..., do_sum=True) from the full map)."""
    if subj_indices is None or len(subj_indices) == 0 or fg_mask is None:
        return 0
    # The reference's early exits (:1845 no fg / bg split; :1881-1888 an instance without foreground or background pixels skips
    # the layer) are evaluated ON THE DEVICE as 0 / 1 factors, so the function never synchronises (CUDA-graph-capturable).
    run = (fg_mask.chunk(4)[0].float().mean() < 0.998).float()
    loss, lws = 0, _layer_weights(layer_weights)
    for li, lw in lws.items():
        if li not in attn_subj_sum:
            continue
        subj_attn = attn_subj_sum[li][:BLOCK_SIZE]                                      # [block, H, N]
        fg = resize_mask_to_target_size(fg_mask, subj_attn.shape[-1]).reshape(BLOCK_SIZE, 1, -1).to(subj_attn.device)
        fg3 = (fg.expand(-1, subj_attn.shape[1], -1) > 1e-6).float()                    # :1875-1877
        bg3 = 1 - fg3
        keep = ((fg3.sum(dim=(1, 2)) > 0).all() & (bg3.sum(dim=(1, 2)) > 0).all()).float()
        excess = subj_attn * bg3 - bg_attn_tolerance                                    # :1907
        pos = (excess > 0).float()
        loss = loss + keep * (excess * pos).sum() / torch.clamp(pos.sum(), min=1e-6) * lw      # masked_mean (:1910)
    return run.to(loss.device) * loss if torch.is_tensor(loss) else loss


def masked_l2_loss(pred, target, mask):
    """ldm/util.py:1215-1239."""
    l2 = (pred - target) ** 2 * mask
    dims = tuple(range(1, mask.ndim))
    msum = mask.sum(dim=dims) * pred.shape[1:].numel() / mask.shape[1:].numel()
    return (l2.sum(dim=dims) / (msum + 1e-8)).mean()


def calc_sc_rep_attn_distill_loss(attn_sqdiff, attn_shape, ca_k, ca_v, subj_indices_1b, prompt_emb_mask_4b, prompt_pad_mask_4b,
                                  sc_fg_mask_percent, FG_THRES=0.1, layer_weights=None):
    """ldm/util.py:2047-2121.  attn_sqdiff: {layer: [1] = sum over (h, i, j) of (sc_attn - sc_rep_attn)^2} from the fused kernel,
    attn_shape: {layer: (H, N, S)} of the map it was reduced over; ca_k / ca_v: {layer: [4, C, S]} for (ss, sc, sc_rep, mc).
    Returns (attn, subj_k, nonsubj_k, subj_v, nonsubj_v) losses."""
    z = 0
    if sc_fg_mask_percent < FG_THRES:                                                   # :2075
        return z, z, z, z, z
    _, sc_emb, _, _ = prompt_emb_mask_4b.squeeze(2).chunk(4)
    _, sc_pad, _, _ = prompt_pad_mask_4b.squeeze(2).chunk(4)
    nonsubj = sc_emb.clone()
    ib, it = subj_indices_1b                                                            # :2068 (value as a device tensor: no host staging)
    nonsubj.index_put_((ib.to(nonsubj.device).long(), it.to(nonsubj.device).long()), torch.zeros(ib.numel(), device=nonsubj.device, dtype=nonsubj.dtype))
    nonsubj = torch.logical_or(nonsubj, sc_pad).unsqueeze(1)                            # [1, 1, S]
    l_attn = l_sk = l_nk = l_sv = l_nv = 0
    for li, lw in _layer_weights(layer_weights).items():
        if li not in attn_sqdiff:
            continue
        H, N, S = attn_shape[li]
        scale = S * 10                                                                  # :2081 (taken before the permute)
        l_attn = l_attn + attn_sqdiff[li].reshape(-1)[0] / float(H * N * S) * scale * lw      # F.mse_loss = mean of squares (:2085-2089)
        ss_k, sc_k, _, mc_k = ca_k[li].chunk(4)
        ss_v, sc_v, _, mc_v = ca_v[li].chunk(4)
        pick = lambda t: t.permute(0, 2, 1)[subj_indices_1b]
        l_sk = l_sk + F.mse_loss(pick(sc_k), pick(ss_k).detach()) * lw
        l_sv = l_sv + F.mse_loss(pick(sc_v), pick(ss_v).detach()) * lw
        l_nk = l_nk + masked_l2_loss(sc_k, mc_k.detach(), nonsubj.to(sc_k.device)) * lw
        l_nv = l_nv + masked_l2_loss(sc_v, mc_v.detach(), nonsubj.to(sc_v.device)) * lw
    return l_attn, l_sk, l_nk, l_sv, l_nv
