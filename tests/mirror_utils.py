"""Build the product's host-side mirrors (adaface_dev_b200) from the seeded golden cases and run them on cuda:0."""
import torch
import torch.nn as nn

import cases as C


a_processor_hook = None


def _T(a, dtype=torch.float32):
    return torch.from_numpy(a).to("cuda", dtype) if a is not None else None


def make_attention(w, C_, ctx_dim, heads=8, processor=None):
    import adaface_dev_b200 as a
    attn = a.Attention(C_, ctx_dim, heads, C_ // heads, processor=processor, device="cuda")
    with torch.no_grad():
        attn.to_q.weight.copy_(_T(w["to_q"]))
        attn.to_k.weight.copy_(_T(w["to_k"]))
        attn.to_v.weight.copy_(_T(w["to_v"]))
        attn.to_out[0].weight.copy_(_T(w["to_out_w"]))
        attn.to_out[0].bias.copy_(_T(w["to_out_b"]))
    return attn


def run_mirror_proc(case, in_dtype=torch.bfloat16, train=False):
    """AttnProcessor_LoRA_Capture mirror on a processor case -> (out, cached_activations).
    train=True: run the autograd-enabled path with leaf inputs -> (out, cache, handles) where handles holds the
    leaf tensors / modules whose .grad the backward tests read."""
    import adaface_dev_b200 as a
    sp, w = case["spec"], case["w"]
    cross = sp.get("cross", False)
    attn = make_attention(w, sp["C"], 768 if cross else None)
    r = sp.get("lora_rank", 0)
    layers = {"q": attn.to_q, "k": attn.to_k, "v": attn.to_v, "out": attn.to_out[0]} if r else None
    proc = a.AttnProcessor_LoRA_Capture(capture_ca_activations=sp.get("capture", False), enable_lora=bool(r),
                                        lora_proj_layers=layers, lora_rank=r or 192, lora_alpha=sp.get("lora_alpha", 16),
                                        q_lora_updates_query=sp.get("q_upd", False), attn_proc_idx=0).cuda()
    with torch.no_grad():
        for n in (("q", "k", "v", "out") if r else ()):
            A, B, mag = w["lora_" + n]
            mod = getattr(proc, f"to_{n}_lora")
            mod.lora_A["default"].weight.copy_(_T(A))
            mod.lora_B["default"].weight.copy_(_T(B))
            mod.lora_magnitude_vector["default"].weight.copy_(_T(mag))
    proc.reset_attn_cache_and_flags(sp.get("capture", False), sp.get("normalize", False), sp.get("mix", False),
                                    sp.get("enable_lora", False))
    if a_processor_hook is not None:          # tests may configure the processor further (e.g. fused capture consumers)
        a_processor_hook(proc)
    attn.set_processor(proc)
    si = case["subj_indices"]
    kw = {}
    if case["img_mask"] is not None:
        kw["img_mask"] = _T(case["img_mask"])
    if si is not None:
        kw["subj_indices"] = (torch.from_numpy(si[0]).cuda(), torch.from_numpy(si[1]).cuda())
    hs, ehs = _T(case["hidden_states"], in_dtype), _T(case["encoder_hidden_states"], in_dtype)
    if train:
        hs.requires_grad_(True)
        if ehs is not None:
            ehs.requires_grad_(True)
        out = attn(hs, encoder_hidden_states=ehs, **kw)
        return out, proc.cached_activations, dict(hidden_states=hs, encoder_hidden_states=ehs, proc=proc, attn=attn)
    with torch.no_grad():
        out = attn(hs, encoder_hidden_states=ehs, **kw)
    return out, proc.cached_activations


def load_ldm_attn(m, w):
    with torch.no_grad():
        m.to_q.weight.copy_(_T(w["to_q"]))
        m.to_k.weight.copy_(_T(w["to_k"]))
        m.to_v.weight.copy_(_T(w["to_v"]))
        m.to_out[0].weight.copy_(_T(w["to_out_w"]))
        m.to_out[0].bias.copy_(_T(w["to_out_b"]))


def run_mirror_ldm(case, train=False):
    with torch.set_grad_enabled(train):
        return _run_mirror_ldm(case, train)


def _run_mirror_ldm(case, train):
    import adaface_dev_b200 as a
    sp, w = case["spec"], case["w"]
    Cc = sp["C"]
    x, ctx, mask = _T(case["x"], torch.bfloat16), _T(case["context"], torch.bfloat16), _T(case["mask"])
    if train:
        x.requires_grad_(True)
        if ctx is not None:
            ctx.requires_grad_(True)
        case["_leaves"] = (x, ctx)
    if sp.get("block"):
        blk = a.BasicTransformerBlock(Cc, 8, Cc // 8, context_dim=768).cuda()
        load_ldm_attn(blk.attn1, w["attn1"])
        load_ldm_attn(blk.attn2, w["attn2"])
        with torch.no_grad():
            for i, ln in enumerate((blk.norm1, blk.norm2, blk.norm3), 1):
                ln.weight.copy_(_T(w[f"norm{i}_w"]))
                ln.bias.copy_(_T(w[f"norm{i}_b"]))
            blk.ff.net[0].proj.weight.copy_(_T(w["ff_proj_w"]))
            blk.ff.net[0].proj.bias.copy_(_T(w["ff_proj_b"]))
            blk.ff.net[2].weight.copy_(_T(w["ff_out_w"]))
            blk.ff.net[2].bias.copy_(_T(w["ff_out_b"]))
        return blk(x, context=ctx, mask=mask), None
    cross = sp.get("cross", False)
    m = a.CrossAttention(Cc, context_dim=768 if cross else None, heads=8, dim_head=Cc // 8).cuda()
    load_ldm_attn(m, w)
    m.save_cross_attn_vars = sp.get("save", False)
    return m(x, context=ctx, mask=mask), m.cached_activations


def make_sbg(w, mults, n_sfx=0):
    import adaface_dev_b200 as a
    gen = a.SubjBasisGenerator(num_static_img_suffix_embs=n_sfx,
                               clip_config=a.CLIPTextConfig(num_hidden_layers=len(mults))).cuda()
    tm = gen.prompt2token_proj.text_model
    with torch.no_grad():
        tm.embeddings.token_embedding.weight.zero_()
        for tid, row in w["token_emb_rows"].items():
            tm.embeddings.token_embedding.weight[tid] = _T(row)
        tm.embeddings.position_embedding.weight.copy_(_T(w["pos_emb"]))
        for layer, lw, m in zip(tm.encoder.layers, w["layers"], mults):
            at = layer.self_attn
            if m != 1:
                at.extend_weights(m, 0.0)
            for p, mod in (("q", at.q_proj), ("k", at.k_proj), ("v", at.v_proj), ("o", at.out_proj)):
                mod.weight.copy_(_T(lw[p + "_w"]))
                mod.bias.copy_(_T(lw[p + "_b"]))
            layer.layer_norm1.weight.copy_(_T(lw["ln1_w"])); layer.layer_norm1.bias.copy_(_T(lw["ln1_b"]))
            layer.layer_norm2.weight.copy_(_T(lw["ln2_w"])); layer.layer_norm2.bias.copy_(_T(lw["ln2_b"]))
            layer.mlp.fc1.weight.copy_(_T(lw["fc1_w"])); layer.mlp.fc1.bias.copy_(_T(lw["fc1_b"]))
            layer.mlp.fc2.weight.copy_(_T(lw["fc2_w"])); layer.mlp.fc2.bias.copy_(_T(lw["fc2_b"]))
        tm.final_layer_norm.weight.copy_(_T(w["final_ln_w"])); tm.final_layer_norm.bias.copy_(_T(w["final_ln_b"]))
        if n_sfx:
            gen.static_img_suffix_embs.copy_(_T(w["static_img_suffix_embs"]))
    gen.pad_embeddings = _T(w["pad_embeddings"])
    gen.prompt2token_proj_attention_multipliers = list(mults)
    return gen.eval()
