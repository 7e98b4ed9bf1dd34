"""Drop-in mirror of the reference's SubjBasisGenerator face path (surface 3).

    SubjBasisGenerator.forward / inverse_img_prompt_embs   adaface/subj_basis_generator.py:443-522, 692-770
    CLIPTextModelWrapper.forward                           adaface/arc2face_models.py:236-338
    CLIPAttentionMKV (+ extend / squeeze weight surgery)   adaface/arc2face_models.py:51-231

The module tree and parameter names are those of HF ``CLIPTextModel`` (text_model.embeddings.*,
text_model.encoder.layers.N.{self_attn.{q,k,v,out}_proj, layer_norm1, mlp.fc1, mlp.fc2, layer_norm2},
text_model.final_layer_norm) so reference checkpoints load with ``load_state_dict``.  The 12 pre-LN layers run
as kernels of libadaface_b200.so: LayerNorm -> fused Q|K|V GEMM(+bias) -> causal multi-K/V flash attention ->
out-proj GEMM(+bias +residual) -> LayerNorm -> fc1 GEMM(+bias, quick-GELU) -> fc2 GEMM(+bias +residual), then the
sum-normalised mix of the last three hidden states + final LayerNorm in one kernel.  The residual stream is fp32.

Because the encoder is causal and only positions 4 .. 4+N_ID(+N_SFX) are returned (subj_basis_generator.py:519-522),
positions past the last returned one cannot influence the output: the encoder runs on T = 4 + N_ID (+ N_SFX)
tokens instead of 77 -- exact, not an approximation.
"""
import math

import torch
import torch.nn as nn

from . import ops
from . import autograd as ag
from .attn_processor import gen_gradient_scaler

# "photo of a" + ", " * (N_ID + 2) tokenised by the CLIP BPE vocab and padded to 77
# (subj_basis_generator.py:473-483; ids per SURVEY 8c -- no tokenizer files are needed at run time).
BOS, EOS, COMMA = 49406, 49407, 267
TEMPLATE_PREFIX = [BOS, 1125, 539, 320]          # <bos> photo of a


def template_ids(n_id=16, max_len=77):
    ids = TEMPLATE_PREFIX + [COMMA] * (n_id + 2)
    return ids + [EOS] * (max_len - len(ids))


class CLIPTextConfig:
    def __init__(self, vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                 num_attention_heads=12, max_position_embeddings=77, layer_norm_eps=1e-5, attention_dropout=0.0):
        self.vocab_size, self.hidden_size, self.intermediate_size = vocab_size, hidden_size, intermediate_size
        self.num_hidden_layers, self.num_attention_heads = num_hidden_layers, num_attention_heads
        self.max_position_embeddings, self.layer_norm_eps = max_position_embeddings, layer_norm_eps
        self.attention_dropout = attention_dropout


class CLIPAttentionMKV(nn.Module):
    """arc2face_models.py:51-143: K/V projections widened x multiplier; q/out as stock CLIP."""

    def __init__(self, config, multiplier=1):
        super().__init__()
        self.config = config
        self.embed_dim, self.num_heads = config.hidden_size, config.num_attention_heads
        self.head_dim = self.embed_dim // self.num_heads
        self.scale = self.head_dim ** -0.5
        self.multiplier = multiplier
        self.k_proj = nn.Linear(self.embed_dim, self.embed_dim * multiplier)
        self.v_proj = nn.Linear(self.embed_dim, self.embed_dim * multiplier)
        self.q_proj = nn.Linear(self.embed_dim, self.embed_dim)
        self.out_proj = nn.Linear(self.embed_dim, self.embed_dim)

    def extend_weights(self, multiplier, perturb_std=0.0):
        """arc2face_models.py:82-127: repeat K/V rows `multiplier` times; the extra copies may be perturbed
        (relative Gaussian noise) so that they diverge under training."""
        if multiplier == 1:
            return
        E0 = self.k_proj.weight.shape[0]
        for proj in (self.k_proj, self.v_proj):
            w = proj.weight.data.repeat(multiplier, 1)
            if perturb_std > 0:
                w[E0:] += torch.randn_like(w[E0:]) * perturb_std * w[E0:].std()
            proj.weight = nn.Parameter(w)
            proj.bias = nn.Parameter(proj.bias.data.repeat(multiplier))
            proj.out_features = w.shape[0]
        self.multiplier *= multiplier

    def squeeze_weights(self, divisor):
        """arc2face_models.py:129-142: average groups of `divisor` K/V copies."""
        if self.multiplier % divisor:
            raise ValueError(f"multiplier {self.multiplier} is not divisible by {divisor}")
        for proj in (self.k_proj, self.v_proj):
            proj.weight = nn.Parameter(proj.weight.data.reshape(divisor, -1, proj.weight.shape[1]).mean(dim=0))
            proj.bias = nn.Parameter(proj.bias.data.reshape(divisor, -1).mean(dim=0))
            proj.out_features = proj.weight.shape[0]
        self.multiplier //= divisor


class CLIPMLP(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.fc1 = nn.Linear(config.hidden_size, config.intermediate_size)
        self.fc2 = nn.Linear(config.intermediate_size, config.hidden_size)


class CLIPEncoderLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.self_attn = CLIPAttentionMKV(config, 1)
        self.layer_norm1 = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.mlp = CLIPMLP(config)
        self.layer_norm2 = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self._pack_key, self._pack = None, None

    def invalidate(self):
        """Force the next pack() to rebuild the bf16 operands from the live parameters (see graphs.graphed_step)."""
        self._pack_key = None

    def pack(self):
        a, m = self.self_attn, self.mlp
        ps = (a.q_proj.weight, a.q_proj.bias, a.k_proj.weight, a.k_proj.bias, a.v_proj.weight, a.v_proj.bias,
              a.out_proj.weight, a.out_proj.bias, m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias,
              self.layer_norm1.weight, self.layer_norm1.bias, self.layer_norm2.weight, self.layer_norm2.bias)
        key = tuple((t.data_ptr(), t._version) for t in ps)
        if key != self._pack_key:
            with torch.no_grad():
                b16 = lambda t: t.detach().to(torch.bfloat16).contiguous()
                f32 = lambda t: t.detach().float().contiguous()
                self._pack = {
                    "wqkv": torch.cat([b16(ps[0]), b16(ps[2]), b16(ps[4])], dim=0),
                    "bqkv": torch.cat([f32(ps[1]), f32(ps[3]), f32(ps[5])]),
                    "wo": b16(ps[6]), "bo": f32(ps[7]), "w1": b16(ps[8]), "b1": f32(ps[9]), "w2": b16(ps[10]),
                    "b2": f32(ps[11]), "ln1": (f32(ps[12]), f32(ps[13])), "ln2": (f32(ps[14]), f32(ps[15]))}
            self._pack_key = key
        return self._pack

    def run(self, h, BS, T):
        """h: fp32 residual stream [BS*T, E] -> next hidden state (HF CLIPEncoderLayer, pre-LN)."""
        pk, a = self.pack(), self.self_attn
        E, M = a.embed_dim, a.multiplier
        x = ops.layernorm(h, *pk["ln1"], eps=self.layer_norm1.eps)
        qkv = ops.proj(x, pk["wqkv"], bias=pk["bqkv"]).view(BS, T, E * (1 + 2 * M))
        q, k, v = qkv[:, :, :E], qkv[:, :, E:E + E * M], qkv[:, :, E + E * M:]
        o = ops.attention(q, k, v, a.num_heads, a.scale, causal_mult=M)          # arc2face_models.py:156-217
        h = ops.proj(o.view(BS * T, E), pk["wo"], bias=pk["bo"], residual=h, out_dtype=torch.float32)
        x = ops.layernorm(h, *pk["ln2"], eps=self.layer_norm2.eps)
        f = ops.proj(x, pk["w1"], bias=pk["b1"], act=ops.ACT_QUICK_GELU)
        return ops.proj(f, pk["w2"], bias=pk["b2"], residual=h, out_dtype=torch.float32)


    def run_train(self, h, BS, T):
        """``run`` with every kernel paired with its backward: gradients to the residual stream and to every
        parameter of the layer (q/k/v/out projections, fc1, fc2, both LayerNorms)."""
        pk, a, m = self.pack(), self.self_attn, self.mlp
        E, M = a.embed_dim, a.multiplier
        x = ag.LayerNormFn.apply(h, self.layer_norm1.weight, self.layer_norm1.bias, self.layer_norm1.eps, torch.bfloat16)
        qkv = ag.linear(x, pk, "wqkv", "bqkv", params=(a.q_proj.weight, a.q_proj.bias, a.k_proj.weight, a.k_proj.bias,
                                                        a.v_proj.weight, a.v_proj.bias)).view(BS, T, E * (1 + 2 * M))
        o = ag.attention(qkv, qkv, qkv, (0, E, E + E * M), E, E * M, a.num_heads, a.scale, None, M)
        h = ag.linear(o.view(BS * T, E), pk, "wo", "bo", params=(a.out_proj.weight, a.out_proj.bias), residual=h,
                      out_dtype=torch.float32)
        x = ag.LayerNormFn.apply(h, self.layer_norm2.weight, self.layer_norm2.bias, self.layer_norm2.eps, torch.bfloat16)
        u = ag.linear(x, pk, "w1", "b1", params=(m.fc1.weight, m.fc1.bias))
        f = ag.ActFn.apply(u, ops.ACT_QUICK_GELU)
        return ag.linear(f, pk, "w2", "b2", params=(m.fc2.weight, m.fc2.bias), residual=h, out_dtype=torch.float32)


class CLIPEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layers = nn.ModuleList([CLIPEncoderLayer(config) for _ in range(config.num_hidden_layers)])


class CLIPTextEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.token_embedding = nn.Embedding(config.vocab_size, config.hidden_size)
        self.position_embedding = nn.Embedding(config.max_position_embeddings, config.hidden_size)


class CLIPTextTransformer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.embeddings = CLIPTextEmbeddings(config)
        self.encoder = CLIPEncoder(config)
        self.final_layer_norm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class CLIPTextModelWrapper(nn.Module):
    """arc2face_models.py:233-382 (forward: precomputed token embeddings in, weighted last-3 mix + final LN out)."""

    def __init__(self, config=None):
        super().__init__()
        self.config = config or CLIPTextConfig()
        self.text_model = CLIPTextTransformer(self.config)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.normal_(p, std=0.02)

    def forward(self, input_ids=None, input_token_embs=None, hidden_state_layer_weights=None, return_token_embs=False,
                num_positions=None, **unused):
        tm = self.text_model
        if return_token_embs:                                                      # arc2face_models.py:249-250
            return tm.embeddings.token_embedding(input_ids)
        if input_token_embs is None:
            input_token_embs = tm.embeddings.token_embedding(input_ids)
        if not input_token_embs.is_cuda:
            raise RuntimeError("adaface_b200 CLIPTextModelWrapper runs on CUDA only (no CPU fallback)")
        BS, T, E = input_token_embs.shape
        T_run = T if num_positions is None else min(T, num_positions)           # causal-exact truncation
        h = (input_token_embs[:, :T_run].float() + tm.embeddings.position_embedding.weight[:T_run].float())   # :268
        h = h.reshape(BS * T_run, E).contiguous()
        train = torch.is_grad_enabled() and (h.requires_grad or any(p.requires_grad for p in tm.encoder.parameters()))
        hs = [h]
        for layer in tm.encoder.layers:                                           # :278-286
            h = layer.run_train(h, BS, T_run) if train else layer.run(h, BS, T_run)
            hs.append(h)
        fl = tm.final_layer_norm
        if train:
            if isinstance(hidden_state_layer_weights, (list, tuple)):
                hidden_state_layer_weights = torch.tensor(hidden_state_layer_weights, device=h.device, dtype=torch.float32).view(-1, 1)
            if hidden_state_layer_weights is None:
                tail, wl_t = [h], torch.ones(1, device=h.device)
            else:
                if hidden_state_layer_weights.numel() != hidden_state_layer_weights.shape[0]:
                    raise NotImplementedError("per-channel hidden_state_layer_weights ([3,768]) are not used by the face path")
                w = hidden_state_layer_weights.float().reshape(-1)
                wl_t = w / w.sum()                                                # :291-306 (3-element host-side glue)
                tail = hs[-w.numel():]
            out = ag.SbgHeadFn.apply(wl_t, fl.weight, fl.bias, fl.eps, *tail)
            return (out.view(BS, T_run, E),)
        if hidden_state_layer_weights is None:                                    # :291-306
            tail, wl = [h], [1.0]
        elif isinstance(hidden_state_layer_weights, (list, tuple)):               # host weights: no device read at all
            tot = float(sum(hidden_state_layer_weights))
            wl = [float(v) / tot for v in hidden_state_layer_weights]
            tail = hs[-len(wl):]
        else:
            w = hidden_state_layer_weights.detach().float().reshape(-1)
            if hidden_state_layer_weights.numel() != hidden_state_layer_weights.shape[0]:
                raise NotImplementedError("per-channel hidden_state_layer_weights ([3,768]) are not used by the face path")
            # host copy of the 3 normalised weights, cached per tensor OBJECT and version: the cache keeps the tensor alive, so
            # its identity cannot be recycled by another tensor (a data_ptr can).  A parameter that is stepped in place bumps
            # _version and is re-read; steady-state inference does no device sync and can be captured into a CUDA graph.
            cache = getattr(self, "_wl_cache", None)
            if cache is None or cache[0] is not hidden_state_layer_weights or cache[1] != hidden_state_layer_weights._version:
                if w.is_cuda and torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("hidden_state_layer_weights: pass a list of floats (or run once outside the capture) -- a "
                                       "device tensor cannot be read during CUDA-graph capture")
                cache = (hidden_state_layer_weights, hidden_state_layer_weights._version, (w / w.sum()).tolist())
                self._wl_cache = cache
            wl = cache[2]
            tail = hs[-len(wl):]
        out = ops.sbg_head(tail, wl, fl.weight.detach().float(), fl.bias.detach().float(), fl.eps)
        return (out.view(BS, T_run, E),)

    def extend_clip_attention_MKV_multiplier(self, multipliers, perturb_std=0.1):
        n = 0
        for layer, m in zip(self.text_model.encoder.layers, multipliers):
            if m != 1:
                layer.self_attn.extend_weights(m, perturb_std)
                n += 1
        return n

    def squeeze_clip_attention_MKV_divisor(self, divisors):
        n = 0
        for layer, dv in zip(self.text_model.encoder.layers, divisors):
            if dv != 1:
                layer.self_attn.squeeze_weights(dv)
                n += 1
        return n


class Arc2FaceID2ImgPrompt(nn.Module):
    """The stage in front of SubjBasisGenerator: ArcFace 512-d ID embedding -> 16 x 768 image-prompt
    embeddings, ``Arc2Face_ID2AdaPrompt.map_init_id_to_img_prompt_embs`` (adaface/face_id_to_ada_prompt.py:680-724).

    "photo of a id person" is tokenised and padded to 22 tokens, the 512-d embedding is zero-padded to 768 and written
    over the token embedding of "id" (position 4), the FROZEN Arc2Face CLIP text encoder (CLIPTextModelWrapper, stock
    K/V multiplier 1, plain final LayerNorm) runs, and positions 4:20 are returned.  The encoder is causal and nothing
    past position 19 is returned, so it runs on 20 positions instead of 22 -- exact.  Same kernels as SubjBasisGenerator."""

    PROMPT_IDS = [BOS, 1125, 539, 320, 1014, 2533] + [EOS] * 16          # <bos> photo of a id person <eos> + padding -> 22
    ARCFACE_POS = 4

    def __init__(self, clip_config=None, dtype=torch.float32):
        super().__init__()
        self.dtype = dtype
        self.id_img_prompt_max_length = 22
        self.text_to_image_prompt_encoder = CLIPTextModelWrapper(clip_config)
        for p in self.text_to_image_prompt_encoder.parameters():                 # face_id_to_ada_prompt.py:639-640
            p.requires_grad = False
        self.register_buffer("input_ids", torch.tensor(self.PROMPT_IDS), persistent=False)

    def map_init_id_to_img_prompt_embs(self, init_id_embs, clip_features=None, called_for_neg_img_prompt=False):
        if init_id_embs.dim() != 2 or init_id_embs.shape[1] > 768:
            raise ValueError(f"init_id_embs must be [N, 512], got {tuple(init_id_embs.shape)}")
        enc = self.text_to_image_prompt_encoder
        N, E = init_id_embs.shape[0], enc.config.hidden_size
        ids = self.input_ids[:20].to(init_id_embs.device)
        tok = enc(input_ids=ids, return_token_embs=True).float().unsqueeze(0).repeat(N, 1, 1)     # :708
        tok[:, self.ARCFACE_POS, :init_id_embs.shape[1]] = init_id_embs.float()                   # :705-709 (zero-padded)
        tok[:, self.ARCFACE_POS, init_id_embs.shape[1]:] = 0
        prompt_embeds = enc(input_token_embs=tok, num_positions=20)[0]                            # :711-715
        return prompt_embeds[:, 4:20].to(self.dtype)                                              # :718-723

    forward = map_init_id_to_img_prompt_embs


class FrozenCLIPTextEncoder(nn.Module):
    """The SD-1.5 prompt encoder around the path (second half): the patched HF CLIP text model of
    ``FrozenCLIPEmbedder`` (ldm/modules/encoders/modules.py:180-338) -- token embeddings, optional EmbeddingManager
    splice of the ada tokens (``embedding_manager(input_ids, inputs_embeds)``, :196-197), position embeddings, 12
    causal pre-LN layers, weighted sum of the last hidden states (``last_layers_skip_weights`` = [0.5, 0.5], the
    NovelAI trick, :318-327) and the final LayerNorm.  Frozen; same kernels as SubjBasisGenerator; HF parameter names
    under ``transformer.text_model.*`` so SD-1.5 text-encoder checkpoints load as they are."""

    def __init__(self, clip_config=None, last_layers_skip_weights=(0.5, 0.5), max_length=77):
        super().__init__()
        self.transformer = CLIPTextModelWrapper(clip_config)
        self.max_length = max_length
        self.last_layers_skip_weights = None if last_layers_skip_weights is None else list(last_layers_skip_weights)
        for p in self.parameters():
            p.requires_grad = False

    def forward(self, input_ids, embedding_manager=None):
        if input_ids.dim() != 2:
            raise ValueError("input_ids must be [B, T]")
        tok = self.transformer(input_ids=input_ids, return_token_embs=True)
        if embedding_manager is not None:                                            # modules.py:196-197
            tok = embedding_manager(input_ids, tok)
        w = None
        if self.last_layers_skip_weights is not None:
            sw = self.last_layers_skip_weights
            if abs(sum(sw) - 1.0) > 1e-6:
                raise NotImplementedError("last_layers_skip_weights must sum to 1 (the reference uses [0.5, 0.5])")
            w = [float(v) for v in sw]        # host list: no per-call device tensor, no sync, graph-capturable
        return self.transformer(input_token_embs=tok, hidden_state_layer_weights=w)[0]    # modules.py:300-330


class SubjBasisGenerator(nn.Module):
    """Face-ID image-prompt embeddings [BS,16,768] -> ada prompt embeddings [BS,16(+N_SFX),768]
    (subj_basis_generator.py:564-770, face path; the bg / object branches :733-756 are out of scope)."""

    def __init__(self, dtype=torch.float32, num_id_vecs=16, num_static_img_suffix_embs: int = 0, output_dim=768,
                 placeholder_is_bg: bool = False, learnable_hidden_state_weights_scheme: str = 'per-layer',
                 clip_config=None, **unused):
        super().__init__()
        if placeholder_is_bg:
            raise NotImplementedError("background SubjBasisGenerator is out of scope (SURVEY.md section 2)")
        self.dtype, self.placeholder_is_bg = dtype, False
        self.N_ID, self.N_SFX = num_id_vecs, num_static_img_suffix_embs
        self.num_out_embs, self.output_dim = self.N_ID + self.N_SFX, output_dim
        self.max_prompt_length = 77
        self.prompt2token_proj = CLIPTextModelWrapper(clip_config)
        self.layerwise_proj = nn.Identity()
        self.prompt2token_proj_attention_multipliers = [1] * len(self.prompt2token_proj.text_model.encoder.layers)
        self.static_img_suffix_embs = (nn.Parameter(torch.randn(1, self.N_SFX, output_dim)) if self.N_SFX > 0 else None)
        self.initialize_hidden_state_layer_weights(learnable_hidden_state_weights_scheme, 'cpu')
        self.register_buffer("input_ids", torch.tensor(template_ids(self.N_ID, self.max_prompt_length)), persistent=False)
        self.pad_embeddings = None
        self.freeze_prompt2token_proj()

    # -- reference helpers kept for callers (face_id_to_ada_prompt.py:125-162)
    def initialize_hidden_state_layer_weights(self, scheme, device):
        if scheme == 'none':
            self.hidden_state_layer_weights = None
            self.hidden_state_layer_weights_grad_scaler = gen_gradient_scaler(1)
        elif scheme == 'per-layer':
            self.hidden_state_layer_weights = nn.Parameter(torch.tensor([[1.0], [2.0], [4.0]], device=device))
            self.hidden_state_layer_weights_grad_scaler = gen_gradient_scaler(5)
        else:
            raise ValueError(f"unknown learnable_hidden_state_weights_scheme '{scheme}'")

    def initialize_text_components(self):
        """pad_embeddings = embeddings(pad ids) incl. positions (subj_basis_generator.py:431-438)."""
        emb = self.prompt2token_proj.text_model.embeddings
        with torch.no_grad():
            pad = emb.token_embedding.weight[EOS].unsqueeze(0) + emb.position_embedding.weight[:self.max_prompt_length]
        self.pad_embeddings = pad.detach().to(self.dtype)

    def freeze_prompt2token_proj(self):
        for p in self.prompt2token_proj.text_model.embeddings.parameters():
            p.requires_grad = False

    def extend_prompt2token_proj_attention(self, prompt2token_proj_attention_multipliers=None, begin_layer_idx=-1,
                                           end_layer_idx=-1, multiplier=1, perturb_std=0.1):
        n_layers = len(self.prompt2token_proj_attention_multipliers)
        b = 0 if begin_layer_idx == -1 else begin_layer_idx
        e = n_layers - 1 if end_layer_idx == -1 else end_layer_idx
        if prompt2token_proj_attention_multipliers is None:
            if multiplier == 1:
                return 0
            prompt2token_proj_attention_multipliers = [multiplier if b <= i <= e else 1 for i in range(n_layers)]
        n = self.prompt2token_proj.extend_clip_attention_MKV_multiplier(prompt2token_proj_attention_multipliers, perturb_std)
        for i in range(b, e + 1):
            self.prompt2token_proj_attention_multipliers[i] *= prompt2token_proj_attention_multipliers[i]
        return n

    def squeeze_prompt2token_proj_attention(self, prompt2token_proj_attention_divisors=None, begin_layer_idx=-1,
                                            end_layer_idx=-1, divisor=1):
        n_layers = len(self.prompt2token_proj_attention_multipliers)
        b = 0 if begin_layer_idx == -1 else begin_layer_idx
        e = n_layers - 1 if end_layer_idx == -1 else end_layer_idx
        if prompt2token_proj_attention_divisors is None:
            if divisor == 1:
                return 0
            prompt2token_proj_attention_divisors = [divisor if b <= i <= e else 1 for i in range(n_layers)]
        n = self.prompt2token_proj.squeeze_clip_attention_MKV_divisor(prompt2token_proj_attention_divisors)
        for i in range(b, e + 1):
            self.prompt2token_proj_attention_multipliers[i] //= prompt2token_proj_attention_divisors[i]
        return n

    # -- the hot path
    def inverse_img_prompt_embs(self, face_prompt_embs, hidden_state_layer_weights=None, enable_static_img_suffix_embs=False):
        """subj_basis_generator.py:443-522 with list_extra_words=None, return_emb_types=['core']."""
        BS = face_prompt_embs.shape[0]
        orig_dtype = face_prompt_embs.dtype
        ID_END = 4 + self.N_ID
        use_sfx = enable_static_img_suffix_embs and self.N_SFX > 0
        T_run = ID_END + (self.N_SFX if enable_static_img_suffix_embs else 0)
        ids = self.input_ids[:T_run].to(face_prompt_embs.device)
        tok = self.prompt2token_proj(input_ids=ids, return_token_embs=True).float().unsqueeze(0).repeat(BS, 1, 1)   # :492
        tok[:, 4:ID_END] = face_prompt_embs.float()                                  # :495
        if use_sfx:
            tok[:, ID_END:ID_END + self.N_SFX] = self.static_img_suffix_embs[:, :self.N_SFX].float()   # :500-502
        prompt_embeds = self.prompt2token_proj(input_token_embs=tok, hidden_state_layer_weights=hidden_state_layer_weights,
                                               num_positions=T_run)[0]              # :505-510
        return prompt_embeds[:, 4:T_run].to(orig_dtype)                             # :519-522

    def forward(self, faceid2img_prompt_embs, clip_features=None, raw_id_embs=None, out_id_embs_cfg_scale=1.0,
                is_face=True, enable_static_img_suffix_embs=False):
        if not is_face:
            raise NotImplementedError("only the face path is in scope (SURVEY.md section 2: object branch is dead for faces)")
        if faceid2img_prompt_embs is None:
            raise ValueError("faceid2img_prompt_embs is required on the face path (subj_basis_generator.py:705)")
        hw = self.hidden_state_layer_weights_grad_scaler(self.hidden_state_layer_weights)   # :716
        ada_id_embs = self.inverse_img_prompt_embs(faceid2img_prompt_embs, hw, enable_static_img_suffix_embs)
        adaface_out_embs = self.layerwise_proj(ada_id_embs)
        if out_id_embs_cfg_scale != 1:                                              # :761-768
            if self.pad_embeddings is None:
                self.initialize_text_components()
            pad = self.pad_embeddings[4:4 + self.N_ID].unsqueeze(0).to(ada_id_embs.device, ada_id_embs.dtype)
            adaface_out_embs = adaface_out_embs.clone()
            adaface_out_embs[:, :self.N_ID] = ada_id_embs[:, :self.N_ID] * out_id_embs_cfg_scale \
                + pad * (1 - out_id_embs_cfg_scale)
        return adaface_out_embs

    def __repr__(self):
        return f"subj SubjBasisGenerator (adaface_b200): num_out_embs={self.num_out_embs}"
