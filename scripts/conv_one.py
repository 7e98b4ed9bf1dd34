"""Runs the implicit-GEMM 3x3 convolution a few times at two U-Net sizes (for ncu captures): level A 320 -> 320 and
level B 1280 -> 640 (the widest up-path convolution), B = 8."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a
B = 8
for side, cin, cout in ((64, 320, 320), (32, 1280, 640)):
    x = torch.randn(B, side * side, cin, device="cuda").to(torch.bfloat16)
    wp = a.ops.pack_conv3x3_weight(torch.randn(cout, cin, 3, 3, device="cuda") * (9 * cin) ** -0.5)
    bias = torch.zeros(cout, device="cuda")
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
        a.ops.conv3x3(x, wp, (side, side), bias=bias)
torch.cuda.synchronize()
