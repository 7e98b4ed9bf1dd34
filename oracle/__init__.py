"""CPU oracle for the AdaFace hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, in plain PyTorch fp32 on the CPU, the arithmetic of the
reference functions on the hot path.  Every function
cites the reference file:line it follows.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline.  Nothing under ``adaface-dev_b200/`` imports this package:
the product path fails loudly when the CUDA library is missing.

Pinning (see DESIGN.md "Oracle"):
  * attention processor / LDM attention / BasicTransformerBlock / CLIPAttentionMKV:
    pinned against outputs of the reference's own code, imported verbatim from
    /root/reference behind sys.modules stubs (tests/golden/make_golden.py),
    committed as fixtures under tests/golden/*.npz.
  * ResBlock / Upsample / Downsample (unet_blocks_oracle.py): pinned the same way (fixture unet_blocks.npz).
  * capture-consumer losses (capture_losses_oracle.py): pinned the same way (fixtures closs_*.npz); no
    kernel consumes them yet.
  * LoRA/DoRA linear (peft, un-vendored, unpinned in requirements.txt:30) and the
    HF-4.44 CLIP encoder loop (transformers>=4.44.2, installed here: 5.5.0 whose
    CLIPEncoder no longer accepts causal_attention_mask): PARITY UNPINNED -- the
    reference ships no tests or vectors for them and the packages are absent;
    restated from the published algorithm, anchored on the reference call sites.
"""

from .attn_oracle import (  # noqa: F401
    lora_dora_linear,
    slow_sdpa,
    processor_forward,
    ldm_cross_attention,
    basic_transformer_block,
    geglu_feed_forward,
    spatial_transformer,
)
from . import unet_blocks_oracle, capture_losses_oracle  # noqa: F401
from .sbg_oracle import (  # noqa: F401
    clip_mkv_attention,
    clip_encoder_layer,
    clip_text_wrapper_forward,
    sbg_forward,
    arc2face_id_to_img_prompt,
    SBG_TEMPLATE_IDS,
    ARC2FACE_PROMPT_IDS,
)
