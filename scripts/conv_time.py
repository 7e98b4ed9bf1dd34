"""Device times (CUDA events, L2 flushed between iterations) of the implicit-GEMM 3x3 convolution, the GroupNorm + SiLU pass
and a whole ResBlock at the SD-1.5 U-Net sizes.  Usage: python scripts/conv_time.py [filter]"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a

ops = a.ops
BF = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
only = sys.argv[1] if len(sys.argv) > 1 else ""


def timeit(name, fn, bytes_=None, flops=None, iters=10):
    if only and only not in name:
        return
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = statistics.median(ts)
    extra = ""
    if bytes_:
        extra += f"  {bytes_ / ms / 1e6:8.0f} GB/s"
    if flops:
        extra += f"  {flops / ms / 1e9:8.1f} TFLOP/s"
    print(f"{name:58s} {ms * 1e3:9.1f} us{extra}", flush=True)


def rn(*s, dt=BF, scale=1.0):
    return (scale * torch.randn(*s, device="cuda")).to(dt)


for B in (2, 8):
    for side, cin, cout in ((64, 320, 320), (64, 640, 320), (32, 640, 640), (32, 1280, 640), (16, 1280, 1280), (16, 2560, 1280), (8, 1280, 1280)):
        x = rn(B, side * side, cin)
        wp = ops.pack_conv3x3_weight(rn(cout, cin, 3, 3, scale=(9 * cin) ** -0.5))
        bias = torch.zeros(cout, device="cuda")
        y = torch.empty(B, side * side, cout, device="cuda", dtype=BF)
        fl = 2.0 * B * side * side * cout * 9 * cin
        timeit(f"conv3x3 B={B} {side}x{side} {cin}->{cout}", lambda: ops.conv3x3(x, wp, (side, side), bias=bias, out=y), flops=fl)
    x = rn(B, 4096, 320)
    wp = ops.pack_conv3x3_weight(rn(320, 320, 3, 3, scale=0.02))
    timeit(f"conv3x3 stride 2 B={B} 64x64 320->320", lambda: ops.conv3x3(x, wp, (64, 64), stride=2), flops=2.0 * B * 1024 * 320 * 9 * 320)
    for hw, c in ((4096, 320), (1024, 640), (256, 1280)):
        x = rn(B, hw, c)
        g, b_ = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
        timeit(f"groupnorm+silu tokens B={B} HW={hw} C={c}", lambda: ops.groupnorm_act_tokens(x, g, b_), bytes_=3 * x.numel() * 2)
    for side, c in ((64, 320), (32, 640), (16, 1280)):
        m = a.ResBlock(c, 1280, 0.0).cuda().eval()
        torch.nn.init.normal_(m.out_layers[3].weight, std=0.02)
        t, emb = rn(B, side * side, c), rn(B, 1280)
        with torch.no_grad():
            timeit(f"ResBlock tokens B={B} {side}x{side} C={c}", lambda: m.forward_tokens(t, emb, (side, side)), flops=2 * 2.0 * B * side * side * c * 9 * c)


# ---- the whole SD-1.5 U-Net forward (random weights), one CUDA graph: BASELINE config 1 (B = 2) and config 3 (B = 8)
if not only or "unet" in only:
    cfg = dict(in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2, attention_resolutions=[4, 2, 1], channel_mult=(1, 2, 4, 4),
               num_heads=8, use_spatial_transformer=True, context_dim=768, transformer_depth=1, legacy=False)
    with torch.device("meta"):
        unet = a.UNetModel(**cfg)
    unet = unet.to_empty(device="cuda").eval()
    with torch.no_grad():
        for k, p in unet.named_parameters():
            if p.dim() >= 2:
                p.normal_(std=p[0].numel() ** -0.5)
            elif k.endswith("weight"):
                p.fill_(1.0)
            else:
                p.zero_()
    for B in (2, 8):
        x, ts, ctx = torch.randn(B, 4, 64, 64, device="cuda"), torch.randint(0, 1000, (B,), device="cuda"), rn(B, 77, 768)
        launches0 = a._lib.launch_count()
        with torch.no_grad():
            y = unet(x, ts, context=ctx)
        n_launch = a._lib.launch_count() - launches0
        assert torch.isfinite(y).all()
        g = a.graphed(lambda x_, t_, c_: unet(x_, t_, context=c_), x, ts, ctx)
        # 0.80 TFLOP per sample: convolutions 55 %, attention blocks 45 %
        timeit(f"unet SD-1.5 forward, CUDA graph, B={B} ({n_launch} kernels)", lambda: g(x, ts, ctx), iters=10)
        with torch.no_grad():
            timeit(f"unet SD-1.5 forward, eager launches, B={B}", lambda: unet(x, ts, context=ctx), iters=5)
