"""CPU oracle (torch fp32) for the consumers of the captured cross-attention activations.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Functional restatement of
  * ldm/util.py:1822-1918  calc_subj_masked_bg_suppress_loss  (subject-column sum, fg / bg masking, tolerance hinge)
  * ldm/util.py:2047-2121  calc_sc_rep_attn_distill_loss      (sc vs sc_rep probability MSE, subject / non-subject k, v MSEs)
and of the helpers they call: sel_emb_attns_by_indices (:1398-1423), resize_mask_to_target_size (:1333-1360),
masked_mean (:1194-1210), masked_l2_loss (:1215-1239), normalize_dict_values (:1088-1095).
Pinned against the reference's own functions (tests/golden/make_golden.py, family "closs"; fixtures closs_*.npz).
These are the reductions the next round fuses into the capture kernel so that the [B, 8, N, S] fp32 maps need not reach HBM;
no kernel consumes this module yet.  All tensors are fp32 CPU tensors.
"""
import torch
import torch.nn.functional as F

ALIGN_LAYERS = (23, 24)                     # attn_align_layer_weights / subj_comp_rep_distill_layer_weights: {23: 1, 24: 1}


def _layer_weights():
    return {li: 1.0 / len(ALIGN_LAYERS) for li in ALIGN_LAYERS}            # normalize_dict_values (:1088-1095)


def resize_mask(mask, area):
    """resize_mask_to_target_size (:1333-1360), mode 'nearest|bilinear', square maps: max of the two interpolations."""
    side = int(area ** 0.5)
    near = F.interpolate(mask.float(), size=(side, side), mode="nearest")
    bil = F.interpolate(mask.float(), size=(side, side), mode="bilinear", align_corners=False)
    return torch.maximum(near, bil)


def subject_column_sum(attn, subj_indices):
    """sel_emb_attns_by_indices(attn.permute(0, 3, 1, 2), subj_indices, do_sum=True) (:1398-1423, :1868):
    attn [B, H, N, S] -> [n_instances, H, N], the probability mass on each instance's subject columns."""
    ib, it = subj_indices
    out = []
    for b in torch.unique(ib):
        cols = it[ib == b]
        out.append(attn[b][:, :, cols].sum(dim=-1))
    return torch.stack(out, dim=0)


def subj_masked_bg_suppress_loss(ca_attn, subj_indices, block_size, fg_mask, bg_attn_tolerance=0.02):
    """calc_subj_masked_bg_suppress_loss (:1822-1918).  ca_attn: {layer: [B, H, N, S]} probabilities; subj_indices:
    (LongTensor[K], LongTensor[K]); fg_mask [B, 1, 64, 64].  Returns a scalar."""
    if subj_indices is None or len(subj_indices) == 0 or fg_mask is None or fg_mask.chunk(4)[0].float().mean() >= 0.998:
        return torch.tensor(0.0)
    k_subj = len(subj_indices[0]) // len(torch.unique(subj_indices[0]))
    subj_indices = (subj_indices[0][:block_size * k_subj], subj_indices[1][:block_size * k_subj])      # :1851
    losses = []
    for li, lw in _layer_weights().items():
        if li not in ca_attn:
            continue
        subj_attn = subject_column_sum(ca_attn[li], subj_indices)                                      # [block, H, N]
        fg2 = resize_mask(fg_mask, subj_attn.shape[-1])
        fg2 = fg2.reshape(block_size, 1, -1).repeat(1, subj_attn.shape[1], 1)
        fg3 = (fg2 > 1e-6).float()                                                                     # :1875-1877
        bg3 = 1 - fg3
        if (fg3.sum(dim=(1, 2)) == 0).any() or (bg3.sum(dim=(1, 2)) == 0).any():                      # :1881-1888
            continue
        excess = subj_attn * bg3 - bg_attn_tolerance                                                   # :1907
        pos = (excess > 0).float()
        losses.append((excess * pos).sum() / torch.clamp(pos.sum(), min=1e-6) * lw)                    # masked_mean (:1910)
    return sum(losses) if losses else torch.tensor(0.0)


def masked_l2_loss(pred, target, mask):
    """masked_l2_loss (:1215-1239): per-instance masked mean of squared differences, then the batch mean."""
    l2 = (pred - target) ** 2 * mask
    dims = tuple(range(1, mask.ndim))
    per = l2.sum(dim=dims)
    msum = mask.sum(dim=dims) * pred.shape[1:].numel() / mask.shape[1:].numel()
    return (per / (msum + 1e-8)).mean()


def sc_rep_attn_distill_loss(acts, subj_indices_1b, prompt_emb_mask_4b, prompt_pad_mask_4b, sc_fg_mask_percent, fg_thres=0.1):
    """calc_sc_rep_attn_distill_loss (:2047-2121).  acts: {'attn': {layer: [4, H, N, S]}, 'k' / 'v': {layer: [4, C, S]}} for the
    instances (ss, sc, sc_rep, mc); masks [4, S, 1].  Returns (attn, subj_k, nonsubj_k, subj_v, nonsubj_v) losses."""
    z = torch.tensor(0.0)
    l_attn, l_sk, l_sv, l_nk, l_nv = z, z, z, z, z
    _, sc_emb, _, _ = prompt_emb_mask_4b.squeeze(2).chunk(4)
    _, sc_pad, _, _ = prompt_pad_mask_4b.squeeze(2).chunk(4)
    nonsubj = sc_emb.clone()
    nonsubj[subj_indices_1b] = 0                                                                       # :2068
    nonsubj = torch.logical_or(nonsubj, sc_pad).unsqueeze(1)                                           # :2071-2073  [1, 1, S]
    if sc_fg_mask_percent < fg_thres:
        return l_attn, l_sk, l_nk, l_sv, l_nv
    for li, lw in _layer_weights().items():
        if li not in acts["attn"]:
            continue
        scale = acts["attn"][li].shape[3] * 10                 # :2081, taken BEFORE the permute: S * 10 ("distributed over 77 tokens")
        ca = acts["attn"][li].permute(0, 3, 1, 2)                                                      # [4, S, H, N]
        _, sc_attn, rep_attn, _ = ca.chunk(4)
        l_attn = l_attn + F.mse_loss(sc_attn, rep_attn.detach()) * scale * lw                          # :2085-2089
        ss_k, sc_k, _, mc_k = acts["k"][li].chunk(4)
        ss_v, sc_v, _, mc_v = acts["v"][li].chunk(4)
        l_sk = l_sk + F.mse_loss(sc_k.permute(0, 2, 1)[subj_indices_1b], ss_k.permute(0, 2, 1)[subj_indices_1b].detach()) * lw
        l_sv = l_sv + F.mse_loss(sc_v.permute(0, 2, 1)[subj_indices_1b], ss_v.permute(0, 2, 1)[subj_indices_1b].detach()) * lw
        l_nk = l_nk + masked_l2_loss(sc_k, mc_k.detach(), nonsubj) * lw
        l_nv = l_nv + masked_l2_loss(sc_v, mc_v.detach(), nonsubj) * lw
    return l_attn, l_sk, l_nk, l_sv, l_nv
