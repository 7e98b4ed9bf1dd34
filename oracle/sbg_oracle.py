"""CPU oracle (torch fp32) for SubjBasisGenerator's face path.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Functional restatement of
  * adaface/arc2face_models.py:145-231   CLIPAttentionMKV.forward (K/V widened x multiplier)
  * adaface/arc2face_models.py:262-306   CLIPTextModelWrapper.forward (embeddings + causal mask +
                                          12 pre-LN layers + sum-normalised last-3 mix + final LN)
  * HF transformers 4.44 CLIPEncoderLayer / CLIPMLP (external; quick-GELU, LN eps 1e-5)
  * adaface/subj_basis_generator.py:443-522, 692-770  inverse_img_prompt_embs / forward, face path
"""
import torch
import torch.nn.functional as F

# Token ids of "photo of a" + ", " * 18 padded to 77 (subj_basis_generator.py:473-483, N_ID = 16):
# BOS 49406, photo 1125, of 539, a 320, "," 267 x18, EOS/pad 49407 x55   (SURVEY.md 8c).
SBG_TEMPLATE_IDS = [49406, 1125, 539, 320] + [267] * 18 + [49407] * 55


def clip_mkv_attention(w, x, multiplier=1, causal=True):
    """arc2face_models.py:145-231.  ``w``: q_w,q_b [E,E]; k_w,k_b,v_w,v_b [E*M, E]; o_w,o_b.
    Each token contributes M consecutive keys per head (:74-79, 156-165); the causal mask is
    broadcast over M (:192)."""
    B, T, E = x.shape
    H = w["num_heads"]
    d = E // H
    q = F.linear(x, w["q_w"], w["q_b"]) * (d ** -0.5)                               # :156
    k = F.linear(x, w["k_w"], w["k_b"]).view(B, -1, H, d).transpose(1, 2)           # :159 [B,H,T*M,d]
    v = F.linear(x, w["v_w"], w["v_b"]).view(B, -1, H, d).transpose(1, 2)           # :160
    q = q.view(B, T, H, d).transpose(1, 2)
    score = q @ k.transpose(-1, -2)                                                 # :170 [B,H,T,T*M]
    if causal:
        neg = torch.finfo(score.dtype).min
        cm = torch.full((T, T), neg, dtype=score.dtype).triu(1)                      # _make_causal_mask
        score = (score.view(B, H, T, T, multiplier) + cm[None, None, :, :, None]).view(B, H, T, T * multiplier)
    p = score.softmax(dim=-1)                                                       # :203
    o = (p @ v).transpose(1, 2).reshape(B, T, E)                                    # :217-227
    return F.linear(o, w["o_w"], w["o_b"])                                          # :229


def clip_encoder_layer(w, h, multiplier=1):
    """HF CLIPEncoderLayer (pre-LN residual block) with CLIPAttentionMKV as self_attn and quick-GELU MLP."""
    E = h.shape[-1]
    r = h
    h = F.layer_norm(h, (E,), w["ln1_w"], w["ln1_b"], 1e-5)
    h = r + clip_mkv_attention(w, h, multiplier)
    r = h
    h = F.layer_norm(h, (E,), w["ln2_w"], w["ln2_b"], 1e-5)
    h = F.linear(h, w["fc1_w"], w["fc1_b"])
    h = h * torch.sigmoid(1.702 * h)                                                # QuickGELU
    h = F.linear(h, w["fc2_w"], w["fc2_b"])
    return r + h


def clip_text_wrapper_forward(w, input_token_embs, hidden_state_layer_weights=None, multipliers=None):
    """arc2face_models.py:262-306.  ``w``: pos_emb [77,E], layers [list of dicts], final_ln_w/b.
    hidden_state_layer_weights [3,1]: last = sum_l (w_l / sum w) h_l over the last 3 of the 13 hidden
    states, THEN the final layer norm (:291-306)."""
    T = input_token_embs.shape[1]
    h = input_token_embs + w["pos_emb"][:T]                                         # :268
    hs = [h]
    for i, lw in enumerate(w["layers"]):
        h = clip_encoder_layer(lw, h, 1 if multipliers is None else multipliers[i])
        hs.append(h)
    if hidden_state_layer_weights is None:
        last = h
    else:
        n = len(hidden_state_layer_weights)
        lwts = hidden_state_layer_weights / hidden_state_layer_weights.sum(dim=0, keepdim=True)   # :299
        last = (torch.stack(hs[-n:], dim=0) * lwts[:, None, None, :]).sum(dim=0)    # :304
    E = last.shape[-1]
    return F.layer_norm(last, (E,), w["final_ln_w"], w["final_ln_b"], 1e-5)         # :306


def sbg_forward(w, faceid2img_prompt_embs, out_id_embs_cfg_scale=1.0, enable_static_img_suffix_embs=False,
                multipliers=None, template_ids=None):
    """SubjBasisGenerator.forward, face path (subj_basis_generator.py:692-770) via
    inverse_img_prompt_embs (:443-522).  ``w``: token_emb [V,E], pos_emb, layers, final_ln_*,
    hidden_state_layer_weights [3,1], optional static_img_suffix_embs [1,N_SFX,E], pad_embeddings [77,E]."""
    BS, N_ID, E = faceid2img_prompt_embs.shape
    if "template_embs" in w:                                                        # precomputed lookup
        tok = w["template_embs"].unsqueeze(0).repeat(BS, 1, 1)
    else:
        ids = torch.tensor(SBG_TEMPLATE_IDS if template_ids is None else template_ids)
        tok = w["token_emb"][ids].unsqueeze(0).repeat(BS, 1, 1)                      # :492
    ID_END = 4 + N_ID
    tok[:, 4:ID_END] = faceid2img_prompt_embs                                       # :495
    n_sfx = 0
    if enable_static_img_suffix_embs and w.get("static_img_suffix_embs") is not None:
        n_sfx = w["static_img_suffix_embs"].shape[1]
        tok[:, ID_END:ID_END + n_sfx] = w["static_img_suffix_embs"]                  # :500-502
    pe = clip_text_wrapper_forward(w, tok, w.get("hidden_state_layer_weights"), multipliers)   # :505-510
    core = pe[:, 4:ID_END + (n_sfx if enable_static_img_suffix_embs else 0)]        # :519-522
    out = core.clone()
    if out_id_embs_cfg_scale != 1:                                                  # :761-768
        pad = w["pad_embeddings"][4:4 + N_ID].unsqueeze(0)
        out[:, :N_ID] = core[:, :N_ID] * out_id_embs_cfg_scale + pad * (1 - out_id_embs_cfg_scale)
    return out


ARC2FACE_PROMPT_IDS = [49406, 1125, 539, 320, 1014, 2533] + [49407] * 16       # "photo of a id person" padded to 22


def arc2face_id_to_img_prompt(w, init_id_embs, prompt_embs=None):
    """Arc2Face_ID2AdaPrompt.map_init_id_to_img_prompt_embs (adaface/face_id_to_ada_prompt.py:680-724): the 512-d
    ArcFace embedding, zero-padded to 768 (:703), replaces the token embedding of "id" (position 4, :709) in the
    22-token prompt; frozen CLIP text encoder (CLIPTextModelWrapper.forward without layer weights, :711-715);
    positions 4:20 are returned (:723).  ``w``: token_emb [V,E] or precomputed ``prompt_embs`` [22,E], pos_emb, layers,
    final_ln_*."""
    N = init_id_embs.shape[0]
    base = prompt_embs if prompt_embs is not None else w["token_emb"][torch.tensor(ARC2FACE_PROMPT_IDS)]
    tok = base.unsqueeze(0).repeat(N, 1, 1)
    E = tok.shape[-1]
    tok[:, 4] = F.pad(init_id_embs, (0, E - init_id_embs.shape[-1]))
    return clip_text_wrapper_forward(w, tok, None)[:, 4:20]
