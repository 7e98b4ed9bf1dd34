"""CPU oracle (torch fp32) for the SD-1.5 attention operator of AdaFace.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Functional restatement of
  * adaface/diffusers_attn_lora_capture.py:79-139   (slow SDPA with normalize / mix)
  * adaface/diffusers_attn_lora_capture.py:192-364  (AttnProcessor_LoRA_Capture.__call__)
  * peft lora.Linear + DoraLinearLayer, eval form   (parity unpinned)
  * ldm/modules/attention.py:168-222                (CrossAttention.forward)
  * ldm/modules/attention.py:31-58, 242-252         (GEGLU feed-forward, BasicTransformerBlock._forward)
All tensors are fp32 CPU tensors; weights come in as plain dicts.
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# gradient scaler (dalc:23-67): identity forward, grad * alpha backward.
class _ScaleGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, alpha):
        ctx.alpha = alpha
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.alpha, None


def scale_grad(x, alpha):
    """dalc:59-67 gen_gradient_scaler: alpha==1 -> identity, alpha==0 -> detach."""
    if alpha == 1:
        return x
    if alpha == 0:
        return x.detach()
    return _ScaleGrad.apply(x, alpha)


# ----------------------------------------------------------------------------------------------
def lora_dora_linear(x, W, bias, lora_A, lora_B, magnitude, scaling, use_dora=True):
    """peft ``lora.Linear(base, r, lora_alpha, use_dora=True)`` in eval mode (dropout inactive
    because the U-Net stays in eval(), ddpm.py:637-638); call sites dalc:171-181, 242, 281, 286, 329.

        y = b + m / ||W + s*B@A||_row  *  (x W^T + s * (x A^T) B^T)          

    The row norm is detached (peft DoraLinearLayer.forward).  Without DoRA: y = b + x W^T + s (x A^T) B^T.
    """
    base = F.linear(x, W)
    lora = F.linear(F.linear(x, lora_A), lora_B) * scaling
    if use_dora:
        w_norm = torch.linalg.norm(W + scaling * (lora_B @ lora_A), dim=1).detach()
        col = (magnitude / w_norm).view(*([1] * (x.dim() - 1)), -1)
        y = col * (base + lora)
    else:
        y = base + lora
    if bias is not None:
        y = y + bias
    return y


# ----------------------------------------------------------------------------------------------
def slow_sdpa(query, key, value, cross_attn_scale_factor, attn_mask=None, subj_indices=None,
              normalize_cross_attn=False, mix_attn_mats_in_batch=False):
    """dalc:79-139.  query [B,H,L,d], key/value [B,H,S,d].  Returns (out, score_after_edit, prob)."""
    B, L, S = query.size(0), query.size(-2), key.size(-2)
    scale_factor = 1.0 / math.sqrt(query.size(-1))                                  # :85
    attn_bias = torch.zeros(B, 1, L, S, dtype=query.dtype)                          # :87
    if attn_mask is not None:                                                       # :94-98
        if attn_mask.dtype == torch.bool:
            attn_bias = attn_bias.masked_fill(attn_mask.logical_not(), float("-inf"))
        else:
            attn_bias = attn_bias + attn_mask
    score = query @ key.transpose(-2, -1) * scale_factor                            # :104
    score = score + attn_bias                                                       # :107
    if mix_attn_mats_in_batch:                                                      # :108-118
        if score.shape[0] % 2 != 0:
            raise ValueError("mix_attn_mats_in_batch needs an even batch [sc.., mc..]")
        sc, mc = score.chunk(2, dim=0)
        score = ((sc + mc.detach()) / 2).repeat(2, 1, 1, 1)
    elif normalize_cross_attn:                                                      # :119-133
        if subj_indices is None:
            raise ValueError("normalize_cross_attn needs subj_indices")
        ib, in_ = subj_indices
        subj = score[ib, :, :, in_]                                                 # [K,H,L]
        subj = subj - subj.mean(dim=2, keepdim=True).detach()                       # :126
        subj = subj * scale_grad(cross_attn_scale_factor, 10)                       # :129-130
        score2 = score.clone()
        score2[ib, :, :, in_] = subj
        score = score2
    prob = torch.softmax(score, dim=-1)                                             # :136
    out = prob @ value                                                              # :138
    return out, score, prob


def _heads(t, B, H):
    d = t.shape[-1] // H
    return t.view(B, -1, H, d).transpose(1, 2)                                      # dalc:299-303


def processor_forward(w, hidden_states, encoder_hidden_states=None, img_mask=None, subj_indices=None,
                      heads=8, capture_ca_activations=False, normalize_cross_attn=False,
                      mix_attn_mats_in_batch=False, enable_lora=False, q_lora_updates_query=False,
                      lora_scaling=0.125, use_dora=True):
    """AttnProcessor_LoRA_Capture.__call__ for an SD-1.5 ``Attention`` (all norms None, no
    residual, rescale 1; dalc:192-364).

    ``w``: dict with to_q, to_k, to_v, to_out_w, to_out_b, cross_attn_scale_factor (scalar tensor) and,
    per adapted projection p in {q,k,v,out}, optional ``lora_p = (A, B, magnitude)``.
    Returns (out [B,N,C], cached_activations dict (empty unless capturing)).
    """
    x = hidden_states
    B = x.shape[0]

    def proj(name, inp, W, b=None):
        lw = w.get("lora_" + name)
        if enable_lora and lw is not None:
            return lora_dora_linear(inp, W, b, lw[0], lw[1], lw[2], lora_scaling, use_dora)
        return F.linear(inp, W, b)

    query = F.linear(x, w["to_q"])                                                  # :235
    if enable_lora and w.get("lora_q") is not None:                                 # :239-249
        query2 = proj("q", x, w["to_q"])
        if q_lora_updates_query:
            query = query2
    else:
        query2 = query
    scale = 1.0 / math.sqrt(query.size(-1))                                         # :251 (inner dim!)

    is_cross = encoder_hidden_states is not None
    attention_mask = None
    if (not is_cross) and img_mask is not None:                                     # :254-273
        ms = int(math.sqrt(x.shape[-2]))
        m = F.interpolate(img_mask, size=(ms, ms), mode="nearest")
        if not (m.sum(dim=(2, 3)) == 0).any():
            attention_mask = m.reshape(B, -1).bool()[:, None, None, :]
    ctx = x if not is_cross else encoder_hidden_states                              # :275-278
    key = proj("k", ctx, w["to_k"])                                                 # :280-283
    value = proj("v", ctx, w["to_v"])                                               # :285-288

    q, q2, k, v = (_heads(t, B, heads) for t in (query, query2, key, value))
    if is_cross and (capture_ca_activations or normalize_cross_attn):               # :309-315
        o, score, prob = slow_sdpa(q, k, v, w["cross_attn_scale_factor"], attn_mask=attention_mask,
                                   subj_indices=subj_indices, normalize_cross_attn=normalize_cross_attn,
                                   mix_attn_mats_in_batch=mix_attn_mats_in_batch)
    else:                                                                           # :320-322
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0)
        score = prob = None
    o = o.transpose(1, 2).reshape(B, -1, q.shape[1] * q.shape[-1])                  # :324
    out = proj("out", o, w["to_out_w"], w["to_out_b"])                              # :328-331

    cache = {}
    if is_cross and capture_ca_activations:                                         # :344-362
        f = math.sqrt(scale)

        def chan_major(t):                                                          # 'b h n d -> b (h d) n'
            return t.permute(0, 1, 3, 2).reshape(B, -1, t.shape[2]).contiguous() * f

        cache = {"q": chan_major(q), "q2": chan_major(q2), "k": chan_major(k), "v": chan_major(v),
                 "attn": prob, "attnscore": score, "attn_out": out.permute(0, 2, 1).contiguous()}
    return out, cache


# ----------------------------------------------------------------------------------------------
def ldm_cross_attention(w, x, context=None, mask=None, heads=8, save_cross_attn_vars=False):
    """ldm/modules/attention.py:168-222.  ``w``: to_q,to_k,to_v (no bias), to_out_w, to_out_b.
    ``mask`` [B,1,h,w] is a *key* mask filled with -finfo.max (:185-194)."""
    B = x.shape[0]
    q = F.linear(x, w["to_q"])
    ctx = x if context is None else context
    k = F.linear(ctx, w["to_k"])
    v = F.linear(ctx, w["to_v"])
    d = q.shape[-1] // heads
    scale = d ** -0.5                                                               # :152
    qh, kh, vh = (_heads(t, B, heads) for t in (q, k, v))
    score = qh @ kh.transpose(-1, -2) * scale                                       # :181
    if mask is not None:
        km = mask.reshape(B, -1).bool()[:, None, None, :]
        score = score.masked_fill(~km, -torch.finfo(score.dtype).max)               # :188-194
    attn = score.softmax(dim=-1)                                                    # :200
    o = (attn @ vh).transpose(1, 2).reshape(B, -1, heads * d)                       # :202-204
    out = F.linear(o, w["to_out_w"], w["to_out_b"])                                 # :205
    cache = None
    if save_cross_attn_vars:                                                        # :207-220
        f = math.sqrt(scale)
        cache = {"q": qh.permute(0, 1, 3, 2).reshape(B, heads * d, -1).contiguous() * f,
                 "attn": attn, "attnscore": score, "attn_out": out.permute(0, 2, 1).contiguous()}
    return out, cache


def geglu_feed_forward(w, x):
    """ldm/modules/attention.py:31-58 with glu=True: Linear(C,8C) -> a * gelu(gate) -> Linear(4C,C)."""
    h = F.linear(x, w["ff_proj_w"], w["ff_proj_b"])
    a, gate = h.chunk(2, dim=-1)
    return F.linear(a * F.gelu(gate), w["ff_out_w"], w["ff_out_b"])


def basic_transformer_block(w, x, context=None, mask=None, heads=8):
    """ldm/modules/attention.py:242-252.  ``w``: dicts attn1, attn2, plus norm{1,2,3}_{w,b}, ff_*.
    The mask only reaches the self-attention (:244-247)."""
    C = x.shape[-1]
    ln = lambda t, i: F.layer_norm(t, (C,), w[f"norm{i}_w"], w[f"norm{i}_b"], 1e-5)
    x1 = ldm_cross_attention(w["attn1"], ln(x, 1), mask=mask, heads=heads)[0] + x
    x2 = x1 + ldm_cross_attention(w["attn2"], ln(x1, 2), context=context, heads=heads)[0]
    return geglu_feed_forward(w, ln(x2, 3)) + x2


def spatial_transformer(w, x, context=None, mask=None, heads=8):
    """ldm/modules/attention.py:287-304 with depth 1: GroupNorm(32, eps 1e-6) (:70-71) -> 1x1 proj_in -> 'b c h w -> b (h w) c'
    -> BasicTransformerBlock (mask nearest-resized to the map, :298) -> back to NCHW -> 1x1 proj_out -> + input.
    ``w``: the block's dict plus gn_w, gn_b, proj_in_w/b [C,C], proj_out_w/b."""
    B, C, h, wd = x.shape
    t = F.group_norm(x, 32, w["gn_w"], w["gn_b"], 1e-6)
    t = F.conv2d(t, w["proj_in_w"][:, :, None, None], w["proj_in_b"])
    t = t.permute(0, 2, 3, 1).reshape(B, h * wd, -1)
    m2 = F.interpolate(mask, size=(h, wd), mode="nearest") if mask is not None else None
    t = basic_transformer_block(w, t, context=context, mask=m2, heads=heads)
    t = t.reshape(B, h, wd, -1).permute(0, 3, 1, 2)
    return F.conv2d(t, w["proj_out_w"][:, :, None, None], w["proj_out_b"]) + x
