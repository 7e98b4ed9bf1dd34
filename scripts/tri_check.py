"""Three-tile attention kernel (ADAFACE_ATTN_TRI=1) vs an fp32 torch reference: level A (B=8, 4096 tokens, 8 x 40) and ragged shapes; then time.
Run one process per configuration (the switch is read once): python scripts/tri_check.py [time-only]."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adaface_dev_b200 as a

torch.manual_seed(0)
H = 8


def ref(q, k, v, scale):
    B, N, C = q.shape
    d = C // H
    qf, kf, vf = (t.float().view(B, -1, H, d).transpose(1, 2) for t in (q, k, v))
    p = torch.softmax(qf @ kf.transpose(-1, -2) * scale, dim=-1)
    return (p @ vf).transpose(1, 2).reshape(B, N, C)


if len(sys.argv) < 2:
    for (B, N, Lk, amp) in ((2, 4096, 4096, 1.0), (1, 1024, 1024, 3.0), (1, 1152, 1000, 1.0), (3, 1280, 4096, 2.0)):
        C = 320
        q = (torch.randn(B, N, C, device="cuda") * amp).to(torch.bfloat16)
        k = (torch.randn(B, Lk, C, device="cuda") * amp).to(torch.bfloat16)
        v = torch.randn(B, Lk, C, device="cuda").to(torch.bfloat16)
        lse = torch.empty(B, H, N, device="cuda", dtype=torch.float32)
        o = a.ops.attention(q, k, v, H, 40 ** -0.5, lse=lse)
        torch.cuda.synchronize()
        r = ref(q, k, v, 40 ** -0.5)
        err = (o.float() - r).abs().max().item()
        s = (q.float().view(B, N, H, 40).transpose(1, 2) @ k.float().view(B, Lk, H, 40).transpose(1, 2).transpose(-1, -2)) * (40 ** -0.5)
        lse_ref = torch.logsumexp(s, dim=-1) * 1.4426950408889634
        lerr = (lse - lse_ref).abs().max().item()
        print(f"B={B} N={N} Lk={Lk} amp={amp}: max |o - ref| = {err:.4e}  max |lse - ref| = {lerr:.4e}  launches {a._lib.launch_count()}", flush=True)

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
if os.environ.get("TRI_PDL") is not None:
    a._lib.set_pdl(int(os.environ["TRI_PDL"]))
noflush = os.environ.get("TRI_NOFLUSH") == "1"
B, N, C = 8, 4096, 320
qkv = torch.randn(B, N, 3 * C, device="cuda").to(torch.bfloat16)
q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
for _ in range(3):
    a.ops.attention(q, k, v, H, 40 ** -0.5)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    if not noflush:
        flush.fill_(1)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); a.ops.attention(q, k, v, H, 40 ** -0.5); e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ms = statistics.median(ts)
print(f"level-A self-attention: {ms*1e3:.1f} us  {4.0*B*N*N*C/ms/1e9:.1f} TFLOP/s  (min {min(ts)*1e3:.1f})", flush=True)
