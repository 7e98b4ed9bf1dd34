import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
for M, N, K in ((32768, 320, 320), (32768, 960, 320), (8192, 640, 640)):
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    for _ in range(2): a.ops.proj(x, w)
torch.cuda.synchronize()
