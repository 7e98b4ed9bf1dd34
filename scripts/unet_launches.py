"""One eager SD-1.5 U-Net forward between two 256 MB marker fills, for an `ncu --metrics gpu__time_duration.sum` launch list
(per-kernel shares of the forward).  Usage: ncu ... python scripts/unet_launches.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = dict(in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2, attention_resolutions=[4, 2, 1], channel_mult=(1, 2, 4, 4),
           num_heads=8, use_spatial_transformer=True, context_dim=768, transformer_depth=1, legacy=False)
with torch.device("meta"):
    unet = a.UNetModel(**cfg)
unet = unet.to_empty(device="cuda").eval()
with torch.no_grad():
    for k, p in unet.named_parameters():
        if p.dim() >= 2:
            p.copy_(torch.randn(p.shape, device="cuda") * p[0].numel() ** -0.5)
        elif k.endswith("weight"):
            p.fill_(1.0)
        else:
            p.zero_()
x, ts, ctx = torch.randn(B, 4, 64, 64, device="cuda"), torch.randint(0, 1000, (B,), device="cuda"), torch.randn(B, 77, 768, device="cuda").bfloat16()
marker = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with torch.no_grad():
    unet(x, ts, context=ctx)          # packs the weights
    torch.cuda.synchronize()
    marker.fill_(1)
    y = unet(x, ts, context=ctx)
    marker.fill_(2)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
