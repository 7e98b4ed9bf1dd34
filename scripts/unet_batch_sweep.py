"""U-Net forward (CUDA graph) at several batch sizes: picks the micro-batch of the DDIM bench (samples/s vs batch)."""
import os, sys, json, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
import bench

dev = torch.device("cuda")
unet = bench.build_unet(torch, a, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}
with torch.no_grad():
    for B in [int(v) for v in (sys.argv[1:] or ["8", "16", "32", "64"])]:
        x, t, c = torch.randn(B, 4, 64, 64, device=dev), torch.randint(0, 1000, (B,), device=dev), torch.randn(B, 77, 768, device=dev).bfloat16()
        fn = a.graphed(lambda x_, t_, c_: unet(x_, t_, context=c_), x, t, c)
        us = bench._time_us(torch, flush, lambda: fn(x, t, c), iters=5)
        res[B] = {"ms": us / 1e3, "samples_per_s": B / us * 1e6, "tflops": B * 8.0327e11 / us / 1e6}
        print(B, res[B], flush=True)
        del fn
        torch.cuda.empty_cache()
print(json.dumps(res))
