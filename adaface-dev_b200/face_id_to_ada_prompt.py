"""Host caller of SubjBasisGenerator: face-ID embeddings -> ada prompt tokens.

    FaceID2AdaPrompt.get_img_prompt_embs            adaface/face_id_to_ada_prompt.py:367-466
    FaceID2AdaPrompt.get_batched_img_prompt_embs    adaface/face_id_to_ada_prompt.py:474-489
    FaceID2AdaPrompt.generate_adaface_embeddings    adaface/face_id_to_ada_prompt.py:503-578
    Arc2Face_ID2AdaPrompt.map_init_id_to_img_prompt_embs   :680-724  (-> subj_basis_generator.Arc2FaceID2ImgPrompt)
    perturb_tensor                                  adaface/util.py:30-53

Scope: the path from a 512-d ArcFace ID embedding onwards.  The stages in front of it -- reading images, face detection
(insightface FaceAnalysis) and the CLIP image encoder that only the ConsistentID variant uses -- are outside the hot path
(SURVEY 2): passing ``image_paths`` / ``image_objs`` raises NotImplementedError instead of falling back to anything.
Averaging, perturbation, batch repetition, truncation to ``num_id_vecs`` and the batch squeeze at inference are the
reference's, evaluated on the device.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .subj_basis_generator import SubjBasisGenerator, Arc2FaceID2ImgPrompt


def perturb_tensor(ts, perturb_std, perturb_std_is_relative=True, keep_norm=False, std_dim=-1, norm_dim=-1, verbose=False,
                   generator=None):
    """adaface/util.py:30-53: Gaussian perturbation, optionally relative to the mean per-row std and norm-preserving."""
    if perturb_std_is_relative:
        perturb_std = perturb_std * ts.std(dim=std_dim).mean().detach()
    noise = torch.randn(ts.shape, device=ts.device, dtype=ts.dtype, generator=generator) * perturb_std
    if keep_norm:
        orig_norm = ts.norm(dim=norm_dim, keepdim=True)
        ts = ts + noise
        new_norm = ts.norm(dim=norm_dim, keepdim=True).detach()
        return ts * orig_norm / (new_norm + 1e-8)
    return ts + noise


class Arc2Face_ID2AdaPrompt(nn.Module):
    name = "arc2face"
    num_id_vecs0 = 16

    def __init__(self, subj_basis_generator=None, id2img_prompt_encoder=None, num_id_vecs=-1, num_static_img_suffix_embs=0,
                 out_id_embs_cfg_scale=1.0, is_training=False, enable_static_img_suffix_embs=False, dtype=torch.float32):
        super().__init__()
        self.dtype = dtype
        self.num_id_vecs = self.num_id_vecs0 if num_id_vecs <= 0 else min(num_id_vecs, self.num_id_vecs0)
        self.num_static_img_suffix_embs = num_static_img_suffix_embs
        self.default_enable_static_img_suffix_embs = bool(enable_static_img_suffix_embs)
        self.out_id_embs_cfg_scale = out_id_embs_cfg_scale
        self.is_training = is_training
        self.use_clip_embs = False                # Arc2Face never consumes CLIP image features (face_id_to_ada_prompt.py:676-679)
        self.gen_neg_img_prompt = False
        self.text_to_image_prompt_encoder = id2img_prompt_encoder if id2img_prompt_encoder is not None else Arc2FaceID2ImgPrompt(dtype=dtype)
        self.subj_basis_generator = subj_basis_generator if subj_basis_generator is not None else SubjBasisGenerator(
            num_id_vecs=self.num_id_vecs0, num_static_img_suffix_embs=num_static_img_suffix_embs)
        for p in self.text_to_image_prompt_encoder.parameters():                                   # :639-642
            p.requires_grad = False
        for p in self.subj_basis_generator.parameters():
            p.requires_grad = p.requires_grad and is_training

    @property
    def device(self):
        return next(self.subj_basis_generator.parameters()).device

    def map_init_id_to_img_prompt_embs(self, init_id_embs, clip_features=None, called_for_neg_img_prompt=False):
        return self.text_to_image_prompt_encoder.map_init_id_to_img_prompt_embs(init_id_embs)

    def get_img_prompt_embs(self, init_id_embs, pre_clip_features, image_paths, image_objs, id_batch_size, skip_non_faces=True,
                            avg_at_stage=None, perturb_at_stage=None, perturb_std=0.0, verbose=False, generator=None):
        """:367-466.  Returns (face_image_count, faceid_embeds [BS,512], pos_prompt_embs [BS,16,768], neg_prompt_embs | None)."""
        if image_paths is not None or image_objs is not None:
            raise NotImplementedError("extracting face embeddings from images (insightface, :400-414) is outside the hot path: pass "
                                      "init_id_embs (ArcFace 512-d embeddings)")
        dev = self.device
        if init_id_embs is None:                                                                   # random identities (:385-393)
            faceid_embeds = torch.randn(id_batch_size, 512, device=dev, generator=generator)
        else:
            faceid_embeds = init_id_embs.to(dev)
            if faceid_embeds.shape[0] == 1:                                                        # :424-427
                faceid_embeds = faceid_embeds.repeat(id_batch_size, 1)
        if perturb_at_stage == 'id_emb' and perturb_std > 0:                                       # :435-440
            faceid_embeds = perturb_tensor(faceid_embeds, perturb_std, True, keep_norm=True, generator=generator)
        faceid_embeds = F.normalize(faceid_embeds.float(), p=2, dim=-1)                            # :442
        with torch.no_grad():                                                                      # :445-449 (frozen teacher)
            pos_prompt_embs = self.map_init_id_to_img_prompt_embs(faceid_embeds, pre_clip_features)
        if avg_at_stage == 'img_prompt_emb':                                                       # :451-455
            pos_prompt_embs = pos_prompt_embs.mean(dim=0, keepdim=True)
            faceid_embeds = faceid_embeds.mean(dim=0, keepdim=True)
        if perturb_at_stage == 'img_prompt_emb' and perturb_std > 0:                               # :457-462
            pos_prompt_embs = perturb_tensor(pos_prompt_embs, perturb_std, True, keep_norm=True, generator=generator)
        return 0, faceid_embeds, pos_prompt_embs, None

    def get_batched_img_prompt_embs(self, batch_size, init_id_embs, pre_clip_features=None):
        """:474-489: the training form -- never average across instances."""
        return self.get_img_prompt_embs(init_id_embs, pre_clip_features, None, None, batch_size, skip_non_faces=False, avg_at_stage=None)

    def generate_adaface_embeddings(self, image_paths, face_id_embs=None, img_prompt_embs=None, p_dropout=0,
                                    return_zero_embs_for_dropped_encoders=True, avg_at_stage='id_emb', perturb_at_stage=None,
                                    perturb_std=0, enable_static_img_suffix_embs=None):
        """:503-578.  Returns (adaface_subj_embs [BS,16,768] -- or [16,768] when averaging --, img_prompt_embs, lens_subj_emb_segments)."""
        if enable_static_img_suffix_embs is None:
            enable_static_img_suffix_embs = self.default_enable_static_img_suffix_embs
        lens_subj_emb_segments = [self.num_id_vecs + enable_static_img_suffix_embs * self.num_static_img_suffix_embs]
        stage = None if (avg_at_stage is None or str(avg_at_stage).lower() == 'none') else avg_at_stage
        if img_prompt_embs is None:
            if stage is not None:
                id_batch_size = 1
            elif face_id_embs is not None:
                id_batch_size = face_id_embs.shape[0]
            else:
                id_batch_size = 1
            if stage == 'id_emb' and face_id_embs is not None and face_id_embs.shape[0] > 1:
                # extract_init_id_embeds_from_images(calc_avg=True) averages the per-image embeddings and re-normalises (:351-354)
                face_id_embs = F.normalize(face_id_embs.float().mean(dim=0, keepdim=True), p=2, dim=-1)
            _, _, img_prompt_embs, _ = self.get_img_prompt_embs(face_id_embs, None, image_paths, None, id_batch_size,
                                                               perturb_at_stage=perturb_at_stage, perturb_std=perturb_std,
                                                               avg_at_stage=stage)
        elif stage is not None:                                                                    # :556-558
            img_prompt_embs = img_prompt_embs.mean(dim=0, keepdim=True)
        adaface_subj_embs = self.subj_basis_generator(img_prompt_embs, clip_features=None, raw_id_embs=None,
                                                      out_id_embs_cfg_scale=self.out_id_embs_cfg_scale, is_face=True,
                                                      enable_static_img_suffix_embs=enable_static_img_suffix_embs)      # :561-567
        if self.num_id_vecs < self.num_id_vecs0:                                                   # :569-570
            adaface_subj_embs = adaface_subj_embs[:, :self.num_id_vecs, :]
        if stage is not None:                                                                      # :574-576
            adaface_subj_embs = adaface_subj_embs.squeeze(0)
        return adaface_subj_embs, img_prompt_embs, lens_subj_emb_segments
