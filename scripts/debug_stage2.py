import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import torch
import test_gpu_stage2 as T
from adaface_dev_b200.stage2 import CompDistillStep
x, ts, prompt, uncond, si, fg, emb, pad = T._step_inputs()
w, _ = T._small_wrapper(use_ffn_lora=True)
w.diffusion_model.captured_layer_indices = (7, 8)
step = CompDistillStep(w, fused_consumers=True, use_ffn_lora=True)
step.align_layers = (7, 8)
pe = prompt.cuda().requires_grad_(True)
_, acts, _ = step.denoise(x, ts[0], pe, uncond, si)
for k, v in acts["sc"].items():
    print(k, {li: (tuple(t.shape), t.requires_grad) for li, t in v.items()})
terms = step.losses(acts, si, fg, emb, pad, 0.3)
print({k: (v.requires_grad if torch.is_tensor(v) else v) for k, v in terms.items()})
loss = sum(v for v in terms.values() if torch.is_tensor(v))
loss.backward()
print("pe.grad", None if pe.grad is None else pe.grad.abs().sum(dim=(1, 2)))
