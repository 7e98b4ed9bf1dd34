"""GPU parity of the U-Net's convolutional blocks: the implicit-GEMM 3x3 convolution on tcgen05, GroupNorm +
SiLU over NHWC tokens, nearest 2x, and the ResBlock / Upsample / Downsample mirrors against
  (1) the committed golden fixtures = outputs of the reference's own modules (tests/golden/unet_*.npz),
  (2) the CPU oracle (oracle/unet_blocks_oracle.py) on seeded inputs, up to the level-A size of BASELINE.json.
Tolerance: max-abs 2e-2 on bf16 outputs of O(1) magnitude (north_star), tighter where the output is fp32."""
import os

import numpy as np
import pytest
import torch

import cases as C
from oracle import unet_blocks_oracle as ub
from mirror_utils import _T
from parity_log import record, log_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def serr(a, ref):
    """max-abs error per unit of output scale, err / max(1, max|ref|) -- see tests/test_gpu_parity.py::serr."""
    r = torch.from_numpy(ref) if isinstance(ref, np.ndarray) else ref
    return err(a, ref) / max(1.0, r.float().abs().max().item())


def err(a, ref):
    ref = torch.from_numpy(ref) if isinstance(ref, np.ndarray) else ref
    e_ = (a.detach().float().cpu() - ref.float()).abs().max().item()
    log_err(e_, ref)
    return e_


def rnd(shape, seed, scale=1.0):
    return (scale * torch.randn(*shape, generator=torch.Generator().manual_seed(seed))).bfloat16().float()


def nhwc(x):                      # [B, C, h, w] fp32 -> bf16 tokens [B, h*w, C] on the GPU
    b, c, h, w = x.shape
    return x.permute(0, 2, 3, 1).reshape(b, h * w, c).contiguous().bfloat16().cuda()


def nchw(t, hw):                  # tokens [B, h*w, C] -> fp32 [B, C, h, w] on the CPU
    b, _, c = t.shape
    return t.float().cpu().reshape(b, hw[0], hw[1], c).permute(0, 3, 1, 2)


# (B, h, w, cin, cout, stride): levels A-D of SD-1.5, odd batches with several images per tile, ragged image rows
# (w = 40 -> 120-pixel tiles; h = 12 with 8-row tiles), channel counts that are not multiples of 64, stride 2
CONV_SHAPES = [
    (2, 64, 64, 320, 320, 1), (2, 32, 32, 640, 320, 1), (3, 16, 16, 1280, 640, 1), (3, 8, 8, 1280, 1280, 1),
    (1, 12, 32, 64, 128, 1), (2, 24, 40, 72, 100, 1), (5, 4, 4, 64, 64, 1), (1, 8, 128, 64, 64, 1),
    (2, 64, 64, 320, 320, 2), (3, 16, 16, 1280, 1280, 2), (1, 8, 8, 64, 64, 2), (2, 24, 80, 72, 100, 2),
]


@pytest.mark.parametrize("shape", CONV_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv3x3_vs_oracle(shape):
    import adaface_dev_b200 as a
    B, h, w, cin, cout, stride = shape
    x, wt, bias = rnd((B, cin, h, w), 1), rnd((cout, cin, 3, 3), 2, (9 * cin) ** -0.5), rnd((cout,), 3, 0.1).float()
    ref = ub.conv3x3(x, wt, bias, stride=stride)
    wp = a.ops.pack_conv3x3_weight(wt.cuda())
    y = a.ops.conv3x3(nhwc(x), wp, (h, w), stride=stride, bias=bias.cuda(), out_dtype=torch.float32)
    ho, wo = h // stride, w // stride
    assert tuple(y.shape) == (B, ho * wo, cout)
    assert err(nchw(y, (ho, wo)), ref) < 2e-3          # fp32 accumulation of exact bf16 products; fp32 output
    yb = a.ops.conv3x3(nhwc(x), wp, (h, w), stride=stride, bias=bias.cuda())
    assert yb.dtype == torch.bfloat16 and err(nchw(yb, (ho, wo)), ref) < 2e-2


def test_conv3x3_epilogue_terms():
    """Per-image bias (the ResBlock's time-embedding term), residual (its skip connection) and a conv-LoRA tail."""
    import adaface_dev_b200 as a
    B, h, w, cin, cout, R = 3, 8, 8, 128, 192, 16
    x, wt = rnd((B, cin, h, w), 11), rnd((cout, cin, 3, 3), 12, (9 * cin) ** -0.5)
    bias, rowb, res = rnd((cout,), 13, 0.1), rnd((B, cout), 14, 0.5), rnd((B, cout, h, w), 15)
    wa, wb = rnd((R, cin, 3, 3), 16, (9 * cin) ** -0.5), rnd((cout, R), 17, R ** -0.5)
    wp = a.ops.pack_conv3x3_weight(wt.cuda())
    base = ub.conv3x3(x, wt, bias) + rowb[:, :, None, None] + res
    y = a.ops.conv3x3(nhwc(x), wp, (h, w), bias=bias.cuda(), rowbias=rowb.cuda(), residual=nhwc(res), out_dtype=torch.float32)
    assert err(nchw(y, (h, w)), base) < 2e-3
    # conv-LoRA: lora_A is a 3x3 convolution to R channels, lora_B a 1x1 back to cout (dalc:541-591)
    t = a.ops.conv3x3(nhwc(x), a.ops.pack_conv3x3_weight(wa.cuda()), (h, w))                 # [B, hw, R] bf16
    t_ref = t.float().cpu()
    lora = torch.einsum("bnr,or->bno", t_ref, wb)
    assert err(nchw(t, (h, w)), ub.conv3x3(x, wa, None)) < 2e-2
    y2 = a.ops.conv3x3(nhwc(x), wp, (h, w), bias=bias.cuda(), t=t.view(B * h * w, R), bs=wb.bfloat16().cuda(), out_dtype=torch.float32)
    ref2 = ub.conv3x3(x, wt, bias) + lora.reshape(B, h, w, cout).permute(0, 3, 1, 2)
    assert err(nchw(y2, (h, w)), ref2) < 2e-3


def test_conv3x3_split_k_epilogue_terms():
    """A small-M, long-K convolution (level D of the U-Net: 2 x 8 x 8 pixels, K = 9 x 1280) runs split-K: partial fp32 tiles +
    the reduce kernel, which applies bias, per-image bias and residual."""
    import adaface_dev_b200 as a
    B, h, w, cin, cout = 2, 8, 8, 1280, 1280
    x, wt = rnd((B, cin, h, w), 41), rnd((cout, cin, 3, 3), 42, (9 * cin) ** -0.5)
    bias, rowb, res = rnd((cout,), 43, 0.1), rnd((B, cout), 44, 0.5), rnd((B, cout, h, w), 45)
    ref = ub.conv3x3(x, wt, bias) + rowb[:, :, None, None] + res
    wp = a.ops.pack_conv3x3_weight(wt.cuda())
    y = a.ops.conv3x3(nhwc(x), wp, (h, w), bias=bias.cuda(), rowbias=rowb.cuda(), residual=nhwc(res), out_dtype=torch.float32)
    assert err(nchw(y, (h, w)), ref) < 2e-3
    yb = a.ops.conv3x3(nhwc(x), wp, (h, w), bias=bias.cuda(), rowbias=rowb.cuda(), residual=nhwc(res))
    assert yb.dtype == torch.bfloat16 and err(nchw(yb, (h, w)), ref) < 4e-2          # |y| up to ~6: bf16 half-ulp 1.6e-2
    y1 = a.ops.conv3x3(nhwc(x), wp, (h, w), bias=bias.cuda(), rowbias=rowb.cuda(), residual=nhwc(res), out_dtype=torch.float32)
    assert torch.equal(y, y1)                                                        # fixed summation order: deterministic


def test_conv3x3_rejects_bad_input():
    import adaface_dev_b200 as a
    wp = a.ops.pack_conv3x3_weight(torch.zeros(64, 64, 3, 3).cuda())
    x = torch.zeros(1, 64, 64, dtype=torch.bfloat16).cuda()
    with pytest.raises(ValueError):
        a.ops.conv3x3(x, wp, (4, 8))                       # hw does not match the token count
    with pytest.raises(RuntimeError):
        a.ops.conv3x3(torch.zeros(1, 9, 64, dtype=torch.bfloat16).cuda(), wp, (3, 3), stride=2)    # odd size under stride 2
    with pytest.raises(RuntimeError):
        a.ops.conv3x3(torch.zeros(1, 256, 64, dtype=torch.bfloat16), wp, (16, 16))                 # CPU tensor: no fallback


@pytest.mark.parametrize("shape", [(2, 4096, 320), (3, 64, 1280), (1, 100, 64), (2, 1024, 960)], ids=lambda s: "x".join(map(str, s)))
def test_groupnorm_act_tokens_vs_oracle(shape):
    import adaface_dev_b200 as a
    B, HW, Cc = shape
    x = rnd((B, HW, Cc), 21) * 2 + 0.5
    gam, bet = 1 + 0.1 * torch.randn(Cc, generator=torch.Generator().manual_seed(22)), 0.1 * torch.randn(Cc, generator=torch.Generator().manual_seed(23))
    xc = x.bfloat16().float().permute(0, 2, 1).reshape(B, Cc, HW, 1)
    n = ub.group_norm32(xc, gam, bet)
    for silu in (True, False):
        y = a.ops.groupnorm_act_tokens(x.bfloat16().cuda(), gam.cuda(), bet.cuda(), 32, 1e-5, silu=silu)
        ref = (ub.silu(n) if silu else n).reshape(B, Cc, HW).permute(0, 2, 1)
        assert y.dtype == torch.bfloat16 and err(y, ref) < 2e-2


def test_silu_and_upsample_kernels():
    import adaface_dev_b200 as a
    e = rnd((3, 1280), 31, 2.0)
    assert err(a.ops.silu(e.cuda()), ub.silu(e)) < 2e-2          # bf16 output, |y| < 8: half an ulp = 1.6e-2
    assert err(a.ops.silu(e.bfloat16().cuda()), ub.silu(e)) < 2e-2
    x = rnd((2, 64, 6, 10), 32)
    up = a.ops.upsample2x_tokens(nhwc(x), (6, 10))
    ref = x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    assert err(nchw(up, (12, 20)), ref) == 0.0             # a copy: bit-exact


def _load_res(m, w):
    pairs = [(m.in_layers[0], "gn1"), (m.in_layers[2], "conv1"), (m.emb_layers[1], "emb"), (m.out_layers[0], "gn2"), (m.out_layers[3], "conv2")]
    if "skip_w" in w:
        pairs.append((m.skip_connection, "skip"))
    with torch.no_grad():
        for mod, key in pairs:
            mod.weight.copy_(_T(w[key + "_w"]))
            mod.bias.copy_(_T(w[key + "_b"]))


@pytest.mark.parametrize("name", list(C.UNET_BLOCK_CASES))
def test_unet_block_vs_reference_golden(name):
    """The NCHW drop-in mirrors against the outputs of the reference's own ResBlock / Upsample / Downsample modules."""
    import adaface_dev_b200 as a
    case = C.build_unet_block_case(name)
    sp, w = case["spec"], case["w"]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    if sp["kind"] == "res":
        m = a.ResBlock(sp["cin"], sp["emb"], 0.0, out_channels=sp["cout"], use_conv=bool(sp.get("skip3"))).cuda().eval()
        _load_res(m, w)
        with torch.no_grad():
            out = m(_T(case["x"]), _T(case["emb"]))
        assert {"in_layers.0.weight", "in_layers.2.bias", "emb_layers.1.weight", "out_layers.0.bias", "out_layers.3.weight"} <= set(m.state_dict())
    else:
        m = (a.Upsample if sp["kind"] == "up" else a.Downsample)(sp["cin"], True).cuda().eval()
        conv = m.conv if sp["kind"] == "up" else m.op
        with torch.no_grad():
            conv.weight.copy_(_T(w["conv_w"]))
            conv.bias.copy_(_T(w["conv_b"]))
            out = m(_T(case["x"]))
    assert out.dtype == torch.float32 and tuple(out.shape) == g["out"].shape
    assert serr(out, g["out"]) < 2e-2         # outputs reach |5.3|: bf16 half-ulp 1.6e-2 there


def test_resblock_tokens_entry_and_full_size():
    """NHWC-resident entry at the level-A size of BASELINE.json (B = 2, 64 x 64, 320 channels) against the oracle."""
    import adaface_dev_b200 as a
    B, h, wd, Cc, E = 2, 64, 64, 320, 1280
    rng = np.random.default_rng(7)
    f = lambda shape, s=1.0: torch.from_numpy((s * rng.standard_normal(shape)).astype(np.float32)).bfloat16().float()
    w = {"gn1_w": 1 + f((Cc,), 0.1), "gn1_b": f((Cc,), 0.05), "conv1_w": f((Cc, Cc, 3, 3), (9 * Cc) ** -0.5), "conv1_b": f((Cc,), 0.02),
         "emb_w": f((Cc, E), E ** -0.5), "emb_b": f((Cc,), 0.02), "gn2_w": 1 + f((Cc,), 0.1), "gn2_b": f((Cc,), 0.05),
         "conv2_w": f((Cc, Cc, 3, 3), (9 * Cc) ** -0.5), "conv2_b": f((Cc,), 0.02)}
    x, emb = f((B, Cc, h, wd)), f((B, E))
    ref = ub.res_block(w, x, emb)
    m = a.ResBlock(Cc, E, 0.0).cuda().eval()
    _load_res(m, {k: v.numpy() for k, v in w.items()})
    with torch.no_grad():
        out = m.forward_tokens(nhwc(x), emb.cuda(), (h, wd))
    assert out.dtype == torch.bfloat16 and serr(nchw(out, (h, wd)), ref) < 2e-2


def test_timestep_embedding_vs_oracle():
    import adaface_dev_b200 as a
    t = torch.tensor([0, 1, 37, 500, 999], dtype=torch.int64)
    for dim in (320, 64, 33):
        got = a.ops.timestep_embedding(t.cuda(), dim)
        assert got.dtype == torch.bfloat16 and tuple(got.shape) == (5, dim)
        assert err(got, ub.timestep_embedding(t, dim)) < 5e-3        # values in [-1, 1], bf16 output


def _unet_from_case(case):
    import adaface_dev_b200 as a
    sp = case["spec"]
    m = a.UNetModel(**sp["cfg"]).cuda().eval()
    sd = C.unet_state_dict({k: v.shape for k, v in m.state_dict().items()}, sp["seed"] + 1000)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return m


@pytest.mark.parametrize("name", list(C.UNET_CASES))
def test_unet_vs_reference_golden(name):
    """The whole U-Net mirror (NHWC-resident forward) against the output of the reference's UNetModel on the same state dict."""
    case = C.build_unet_case(name)
    g = np.load(os.path.join(GOLD, name + ".npz"))
    sp_big = case["spec"].get("big", False)
    m = _unet_from_case(case)
    x, ts, ctx, mask = _T(case["x"]), torch.from_numpy(case["timesteps"]).cuda(), _T(case["context"]), _T(case["mask"])
    with torch.no_grad():
        out = m(x, ts, context=ctx, extra_info={"img_mask": mask})
    assert out.dtype == torch.float32 and tuple(out.shape) == g["out"].shape
    ref = torch.from_numpy(g["out"])
    rel = ((out.cpu() - ref).norm() / ref.norm()).item()
    record("unet", name, "eps [B,4,h,w]: max-abs err / max(1, max|ref|)", serr(out, ref), 2e-2, f"abs {err(out, ref):.3f}; rel-L2 {rel:.2e}; ref max-abs {ref.abs().max().item():.2f}")
    # eight (small) / 25 (SD-1.5) bf16 blocks deep (the reference runs fp16 under autocast): output std 0.56, max 2.4
    assert serr(out, ref) < 2e-2 and rel < 2e-2, (err(out, ref), rel)
    if sp_big:
        # SD-1.5 itself (BASELINE config 3's size): captured layers are the reference's own 22, 23, 24 (openaimodel.py:853)
        info = {"capture_ca_activations": True}
        with torch.no_grad():
            out2 = m(x, ts, context=ctx, extra_info=info)
        assert err(out2, out.cpu()) < 3e-2
        acts = info["ca_layers_activations"]
        assert set(acts["attn"]) == {22, 23, 24} and tuple(acts["attn"][24].shape) == (2, 8, 4096, 77)
        assert tuple(acts["outfeat"][22].shape) == (2, 320, 64, 64)
        assert (acts["attn"][23].sum(-1) - 1).abs().max().item() < 1e-3
        return
    # capture plumbing (openaimodel.py:849-941): same prediction, maps of the captured cross-attention layers handed back
    m.captured_layer_indices = (7, 8)              # the two full-resolution output blocks of this small configuration
    info = {"img_mask": mask, "capture_ca_activations": True}
    with torch.no_grad():
        out2 = m(x, ts, context=ctx, extra_info=info)
    assert err(out2, out.cpu()) < 3e-2
    acts = info["ca_layers_activations"]
    assert set(acts) == {"outfeat", "attn", "attnscore", "q", "attn_out"} and set(acts["attn"]) == {7, 8}
    assert tuple(acts["attn"][8].shape) == (2, 8, 256, 77) and tuple(acts["outfeat"][8].shape) == (2, 320, 16, 16)
    assert (acts["attn"][8].sum(-1) - 1).abs().max().item() < 1e-3
    assert all(not m._cross_attn(li).save_cross_attn_vars for li in (7, 8))


@pytest.mark.parametrize("k", [3, 1])
def test_lora_dora_conv_vs_oracle(k):
    """Conv-LoRA / DoRA adapter (dalc:541-591: r = 192, alpha = 16 on up_blocks.3.resnets.[12].conv1 / conv2 / conv_shortcut):
    640 -> 320 at 16 x 16, adapter moved off its identity initialisation."""
    import adaface_dev_b200 as a
    torch.manual_seed(5)
    B, h, w, cin, cout = 2, 16, 16, 640, 320
    base = torch.nn.Conv2d(cin, cout, k, padding=k // 2)
    with torch.no_grad():
        base.weight.copy_(rnd((cout, cin, k, k), 51, (k * k * cin) ** -0.5))
        base.bias.copy_(rnd((cout,), 52, 0.05))
    W0, b0 = base.weight.detach().clone(), base.bias.detach().clone()        # CPU copies: .cuda() below moves `base` in place
    m = a.LoraDoraConv2d(base, r=192, lora_alpha=16)
    with torch.no_grad():
        m.lora_A["default"].weight.copy_(rnd((192, cin, k, k), 53, (k * k * cin) ** -0.5))
        m.lora_B["default"].weight.copy_(rnd((cout, 192, 1, 1), 54, 0.3))
        m.lora_magnitude_vector["default"].weight.mul_(1.1)
    x = rnd((B, cin, h, w), 55)
    ref = ub.lora_dora_conv(x, W0, b0, m.lora_A["default"].weight.detach(),
                            m.lora_B["default"].weight.detach(), m.lora_magnitude_vector["default"].weight.detach(), m.scaling)
    m = m.cuda().eval()
    with torch.no_grad():
        out = m(x.cuda())
    assert out.dtype == torch.float32 and tuple(out.shape) == (B, cout, h, w)
    assert serr(out, ref) < 2e-2                     # bf16 adapter branch T = conv(x, A) feeding the rank-192 tail
    assert {"base_layer.weight", "lora_A.default.weight", "lora_B.default.weight", "lora_magnitude_vector.default.weight"} <= set(m.state_dict())
    # identity at init (peft: B = 0, magnitude = ||W||): the adapter reproduces the base convolution
    fresh = a.LoraDoraConv2d(base, r=192, lora_alpha=16).cuda().eval()
    with torch.no_grad():
        out0 = fresh(x.cuda())
    ref0 = ub.conv3x3(x, W0, b0) if k == 3 else torch.einsum("bchw,oc->bohw", x, W0[:, :, 0, 0]) + b0[None, :, None, None]
    assert err(out0, ref0) < 2e-2


# ------------------------------------------------------------------------------------------------ backward (training path)
def gerr(got, ref):
    """bf16-grade gradient bar of tests/test_gpu_backward.py: max-abs error relative to the max-abs of the reference gradient."""
    ref = ref.float()
    return (got.detach().float().cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-12)


@pytest.mark.parametrize("shape", [(2, 256, 320), (1, 64, 1280), (2, 100, 64), (1, 1024, 960)], ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("silu", [True, False])
def test_groupnorm_act_tokens_bwd_vs_oracle(shape, silu):
    import adaface_dev_b200 as a
    B, HW, Cc = shape
    x = (rnd((B, HW, Cc), 61) * 1.5 + 0.3).bfloat16().float().requires_grad_(True)
    dy = rnd((B, HW, Cc), 62)
    gam, bet = 1 + 0.1 * torch.randn(Cc, generator=torch.Generator().manual_seed(63)), 0.1 * torch.randn(Cc, generator=torch.Generator().manual_seed(64))
    n = ub.group_norm32(x.permute(0, 2, 1).reshape(B, Cc, HW, 1), gam, bet)
    y = (ub.silu(n) if silu else n).reshape(B, Cc, HW).permute(0, 2, 1)
    (ref,) = torch.autograd.grad(y, x, dy)
    got = a.ops.groupnorm_act_tokens_bwd(x.detach().bfloat16().cuda(), dy.bfloat16().cuda(), gam.cuda(), bet.cuda(), 32, 1e-5, silu=silu)
    assert got.dtype == torch.bfloat16 and gerr(got, ref) < 2e-2


@pytest.mark.parametrize("shape", [(2, 16, 16, 128, 64, 1), (2, 16, 16, 64, 128, 2), (1, 8, 8, 1280, 1280, 1), (3, 12, 32, 64, 64, 1)],
                         ids=lambda s: "x".join(map(str, s)))
def test_conv3x3_input_gradient_vs_oracle(shape):
    """dX of the frozen-weight convolution = the same kernel over flipped / transposed weights (stride 2: after zero-insertion)."""
    import adaface_dev_b200 as a
    from adaface_dev_b200 import autograd as ag
    B, h, w, cin, cout, stride = shape
    x = rnd((B, cin, h, w), 71).requires_grad_(True)
    wt, bias = rnd((cout, cin, 3, 3), 72, (9 * cin) ** -0.5), rnd((cout,), 73, 0.1)
    dy = rnd((B, cout, h // stride, w // stride), 74)
    (ref,) = torch.autograd.grad(ub.conv3x3(x, wt, bias, stride=stride), x, dy)
    xt = nhwc(x.detach()).requires_grad_(True)
    pack = {"w": a.ops.pack_conv3x3_weight(wt.cuda())}
    y = ag.conv3x3(xt, pack, "w", wt.cuda(), (h, w), stride=stride, bias=bias.cuda())
    y.backward(nhwc(dy))
    assert xt.grad.dtype == torch.bfloat16 and gerr(nchw(xt.grad, (h, w)), ref) < 2e-2
    assert "w_dx" in pack                                   # the dX operand is packed once and cached next to the forward pack


def test_upsample_and_resblock_backward_vs_oracle():
    import adaface_dev_b200 as a
    from adaface_dev_b200 import autograd as ag
    x = rnd((2, 64, 6, 10), 81)
    t = nhwc(x).requires_grad_(True)
    g = rnd((2, 64, 12, 20), 82)
    ag.upsample2x(t, (6, 10)).backward(nhwc(g))
    ref = g.reshape(2, 64, 6, 2, 10, 2).sum(dim=(3, 5))
    assert gerr(nchw(t.grad, (6, 10)), ref) < 1e-2
    for name in ("unet_res_a", "unet_res_skip", "unet_res_rect"):
        case = C.build_unet_block_case(name)
        sp, w = case["spec"], case["w"]
        wt = {k: torch.from_numpy(v) for k, v in w.items()}
        xr = torch.from_numpy(case["x"]).requires_grad_(True)
        emb = torch.from_numpy(case["emb"])
        out_ref = ub.res_block(wt, xr, emb)
        G = rnd(tuple(out_ref.shape), 83)
        (ref,) = torch.autograd.grad(out_ref, xr, G)
        m = a.ResBlock(sp["cin"], sp["emb"], 0.0, out_channels=sp["cout"], use_conv=bool(sp.get("skip3"))).cuda().eval()
        _load_res(m, w)
        hw = (sp["h"], sp["w"])
        tt = nhwc(xr.detach()).requires_grad_(True)
        out = m.forward_tokens(tt, emb.cuda(), hw)
        assert err(nchw(out, hw), out_ref.detach()) < 3e-2
        out.backward(nhwc(G))
        assert gerr(nchw(tt.grad, hw), ref) < 3e-2, name
        assert all(p.grad is None for p in m.parameters())          # frozen U-Net weights receive no gradient


def test_unet_context_gradient_vs_oracle():
    """Stage-2 direction of the whole U-Net mirror (ddpm.py:1645-1707): d loss / d prompt context through every ResBlock,
    SpatialTransformer, Down / Upsample and skip concatenation, against autograd through the CPU oracle."""
    case = C.build_unet_case("unet_small")
    sp = case["spec"]
    m = _unet_from_case(case)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    x, ts = torch.from_numpy(case["x"]), torch.from_numpy(case["timesteps"])
    ctx_ref = torch.from_numpy(case["context"]).requires_grad_(True)
    mask = torch.from_numpy(case["mask"])
    out_ref = ub.unet_forward(sd, sp["cfg"], x, ts, ctx_ref, mask=mask)
    G = rnd(tuple(out_ref.shape), 91)
    (ref,) = torch.autograd.grad(out_ref, ctx_ref, G)
    ctx = torch.from_numpy(case["context"]).cuda().requires_grad_(True)
    out = m(x.cuda(), ts.cuda(), context=ctx, extra_info={"img_mask": mask.cuda()})
    assert out.requires_grad and err(out, out_ref.detach()) < 6e-2
    out.backward(G.cuda())
    assert tuple(ctx.grad.shape) == tuple(ref.shape) and gerr(ctx.grad, ref) < 6e-2
    assert all(p.grad is None for p in m.parameters())
    with pytest.raises(NotImplementedError):
        m(x.cuda().requires_grad_(True), ts.cuda(), context=ctx)


def test_spatial_transformer_context_gradient_vs_oracle():
    """SpatialTransformer.forward (NCHW surface) with a context that requires grad runs the differentiable tokens path:
    output against the reference fixture, d context against autograd through the oracle."""
    import adaface_dev_b200 as a
    import oracle
    from mirror_utils import load_ldm_attn
    name = "ldm_spatial_d80"
    case = C.build_spatial_case(name)
    sp, w = case["spec"], case["w"]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    Cc = sp["C"]
    m = a.SpatialTransformer(Cc, 8, Cc // 8, depth=1, context_dim=768).cuda()
    blk = m.transformer_blocks[0]
    load_ldm_attn(blk.attn1, w["attn1"])
    load_ldm_attn(blk.attn2, w["attn2"])
    with torch.no_grad():
        for i, ln in enumerate((blk.norm1, blk.norm2, blk.norm3), 1):
            ln.weight.copy_(_T(w[f"norm{i}_w"])); ln.bias.copy_(_T(w[f"norm{i}_b"]))
        blk.ff.net[0].proj.weight.copy_(_T(w["ff_proj_w"])); blk.ff.net[0].proj.bias.copy_(_T(w["ff_proj_b"]))
        blk.ff.net[2].weight.copy_(_T(w["ff_out_w"])); blk.ff.net[2].bias.copy_(_T(w["ff_out_b"]))
        m.norm.weight.copy_(_T(w["gn_w"])); m.norm.bias.copy_(_T(w["gn_b"]))
        m.proj_in.weight.copy_(_T(w["proj_in_w"])[:, :, None, None]); m.proj_in.bias.copy_(_T(w["proj_in_b"]))
        m.proj_out.weight.copy_(_T(w["proj_out_w"])[:, :, None, None]); m.proj_out.bias.copy_(_T(w["proj_out_b"]))
    t = C.to_torch({k: v for k, v in case.items() if k != "spec"})
    ctx_ref = t["context"].clone().requires_grad_(True)
    out_ref = oracle.spatial_transformer(t["w"], t["x"], context=ctx_ref, mask=t["mask"])
    G = rnd(tuple(out_ref.shape), 95)
    (ref,) = torch.autograd.grad(out_ref, ctx_ref, G)
    ctx = _T(case["context"]).requires_grad_(True)
    out = m(_T(case["x"]), context=ctx, mask=_T(case["mask"]))
    assert out.requires_grad and err(out, g["out"]) < 3e-2
    out.backward(G.cuda())
    assert gerr(ctx.grad, ref) < 3e-2


def test_conv_dora_pack_matches_peft_formula():
    """LoraDoraConv2d.pack(): packed W / A, s*B and colscale = m / ||W + s B.A|| reproduce the oracle's lora_dora_conv when
    applied as y = colscale o (conv(x, W) + conv1x1(conv(x, A), sB)) + b; identity at initialisation."""
    from oracle import unet_blocks_oracle as ub
    import adaface_dev_b200 as a
    torch.manual_seed(0)
    for k in (3, 1):
        base = torch.nn.Conv2d(16, 24, k, padding=k // 2).cuda()
        lora = a.LoraDoraConv2d(base, r=8, lora_alpha=4).cuda()
        wp0, ap0, bs0, cs0, b0 = lora.pack()
        assert bs0.abs().max().item() == 0 and (cs0 - 1).abs().max().item() < 1e-5 and lora.pack()[0] is wp0
        with torch.no_grad():
            lora.lora_B["default"].weight.normal_(std=0.2)
            lora.lora_magnitude_vector["default"].weight.mul_(1.2)
        wp, ap, bs, cs, b = lora.pack()
        assert wp is not wp0 and wp.dtype == torch.bfloat16 and cs.dtype == torch.float32
        assert tuple(wp.shape) == ((24, 9 * 64) if k == 3 else (24, 16)) and tuple(ap.shape) == ((8, 9 * 64) if k == 3 else (8, 16))
        x = torch.randn(2, 16, 5, 4)
        c_ = lambda t_: t_.detach().float().cpu()
        A, B = c_(lora.lora_A["default"].weight), c_(lora.lora_B["default"].weight)
        conv = (lambda t, w: ub.conv3x3(t, w, None)) if k == 3 else (lambda t, w: torch.einsum("bchw,oc->bohw", t, w[:, :, 0, 0]))
        y = c_(cs)[None, :, None, None] * (conv(x, c_(base.weight)) + torch.einsum("bchw,oc->bohw", conv(x, A), c_(bs))) + c_(b)[None, :, None, None]
        ref = ub.lora_dora_conv(x, c_(base.weight), c_(base.bias), A, B, c_(lora.lora_magnitude_vector["default"].weight), lora.scaling)
        assert (y - ref).abs().max().item() < 2e-2          # bf16 rounding of s*B only
    with pytest.raises(NotImplementedError):
        a.LoraDoraConv2d(torch.nn.Conv2d(8, 8, 3, stride=2, padding=1))
