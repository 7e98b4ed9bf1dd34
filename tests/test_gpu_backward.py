"""GPU parity of the BACKWARD kernels (north-star item 5): every gradient kernel against autograd
over an fp32 restatement, then the three mirrors' training paths against autograd through the CPU oracle on the golden
cases.  Gradients are bf16-grade, so the bar is relative: max-abs error <= GRAD_TOL * max-abs of the reference gradient
(3e-2; the forward bar of 2e-2 on O(1) outputs is the same relative accuracy)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cases as C
import oracle
from mirror_utils import run_mirror_proc, run_mirror_ldm, make_sbg, _T

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
GRAD_TOL = 3e-2


def ops():
    import adaface_dev_b200 as a
    return a.ops


def ag():
    import adaface_dev_b200.autograd as m
    return m


def rnd(*shape, std=1.0, seed=0, dtype=BF):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * std).to(dtype).cuda()


def rel(a, ref):
    ref = torch.from_numpy(ref) if isinstance(ref, np.ndarray) else ref
    a, ref = a.detach().float().cpu(), ref.detach().float().cpu()
    assert a.shape == ref.shape, (a.shape, ref.shape)
    return ((a - ref).abs().max() / ref.abs().max().clamp_min(1e-20)).item()


# ------------------------------------------------------------------------------------------- flash backward
def ref_attn_grads(q, k, v, do, H, scale, key_mask=None, causal_mult=0):
    B, Lq, Cq = q.shape
    d = Cq // H
    q, k, v = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    qh = q.view(B, Lq, H, d).transpose(1, 2)
    kh = k.reshape(B, -1, H, d).transpose(1, 2)
    vh = v.reshape(B, -1, H, d).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) * scale
    Lk = kh.shape[2]
    if key_mask is not None:
        s = s.masked_fill(~key_mask.bool()[:, None, None, :], float("-inf"))
    if causal_mult:
        i = torch.arange(Lq, device=s.device)[:, None]
        j = torch.arange(Lk, device=s.device)[None, :]
        s = s.masked_fill((j // causal_mult) > i, float("-inf"))
    o = (s.softmax(-1) @ vh).transpose(1, 2).reshape(B, Lq, Cq)
    o.backward(do.float())
    return o.detach(), q.grad, k.grad, v.grad


@pytest.mark.parametrize("d,Lq,Lk", [(40, 256, 256), (40, 1000, 1000), (40, 77, 300), (80, 192, 130), (160, 100, 77), (64, 20, 20),
                                     (40, 4096, 77), (160, 64, 64)])
def test_attention_bwd(d, Lq, Lk):
    B, H = 2, 8 if d != 64 else 12
    Cc = H * d
    q, k, v, do = rnd(B, Lq, Cc, seed=1), rnd(B, Lk, Cc, seed=2), rnd(B, Lk, Cc, seed=3), rnd(B, Lq, Cc, seed=4)
    lse = torch.empty(B, H, Lq, device="cuda")
    o = ops().attention(q, k, v, H, d ** -0.5, lse=lse)
    ro, rq, rk, rv = ref_attn_grads(q, k, v, do, H, d ** -0.5)
    assert rel(o, ro) < 2e-2
    rlse = torch.logsumexp((q.float().view(B, Lq, H, d).transpose(1, 2) @ k.float().view(B, Lk, H, d).transpose(1, 2).transpose(-1, -2))
                           * d ** -0.5, dim=-1) * math.log2(math.e)
    assert (lse - rlse).abs().max().item() < 2e-2
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ops().attention_bwd(q, k, v, o, do, lse, H, d ** -0.5, dq, dk, dv)
    assert rel(dq, rq) < GRAD_TOL and rel(dk, rk) < GRAD_TOL and rel(dv, rv) < GRAD_TOL


def test_attention_bwd_fused_views_and_key_mask():
    B, N, H, d = 2, 320, 8, 40
    Cc = H * d
    qkv, do = rnd(B, N, 3 * Cc, seed=1), rnd(B, N, Cc, seed=2)
    km = (torch.rand(B, N, generator=torch.Generator().manual_seed(5)) > 0.3).to(torch.uint8).cuda()
    q, k, v = qkv[:, :, :Cc], qkv[:, :, Cc:2 * Cc], qkv[:, :, 2 * Cc:]
    lse = torch.empty(B, H, N, device="cuda")
    o = ops().attention(q, k, v, H, d ** -0.5, key_mask=km, lse=lse)
    _, rq, rk, rv = ref_attn_grads(q, k, v, do, H, d ** -0.5, key_mask=km)
    dqkv = torch.empty_like(qkv)
    ops().attention_bwd(q, k, v, o, do, lse, H, d ** -0.5, dqkv[:, :, :Cc], dqkv[:, :, Cc:2 * Cc], dqkv[:, :, 2 * Cc:], key_mask=km)
    assert rel(dqkv[:, :, :Cc], rq) < GRAD_TOL and rel(dqkv[:, :, Cc:2 * Cc], rk) < GRAD_TOL and rel(dqkv[:, :, 2 * Cc:], rv) < GRAD_TOL
    # masked keys receive exactly zero gradient
    dead = (km == 0)[:, :, None].expand(B, N, Cc)
    assert dqkv[:, :, Cc:2 * Cc][dead].abs().max().item() == 0 and dqkv[:, :, 2 * Cc:][dead].abs().max().item() == 0


@pytest.mark.parametrize("d,N", [(40, 1024), (80, 512), (40, 4096)])
def test_attention_bwd_key_mask_on_tcgen05(d, N):
    """img_mask self-attention backward (dalc:254-273) on the tensor-core kernels: random masks, a masked leading block, a single kept
    key, and one unmasked instance in the same batch; masked keys receive exactly zero dK / dV."""
    B, H = 4, 8
    Cc = H * d
    qkv, do = rnd(B, N, 3 * Cc, seed=11), rnd(B, N, Cc, seed=12)
    km = (torch.rand(B, N, generator=torch.Generator().manual_seed(6)) > 0.4).to(torch.uint8)
    km[1, :200] = 0
    km[2] = 0
    km[2, 333] = 1
    km[3] = 1
    km = km.cuda()
    q, k, v = qkv[:, :, :Cc], qkv[:, :, Cc:2 * Cc], qkv[:, :, 2 * Cc:]
    lse = torch.empty(B, H, N, device="cuda")
    o = ops().attention(q, k, v, H, d ** -0.5, key_mask=km, lse=lse)
    ro, rq, rk, rv = ref_attn_grads(q, k, v, do, H, d ** -0.5, key_mask=km)
    assert rel(o, ro) < 2e-2
    dqkv = torch.empty_like(qkv)
    ops().attention_bwd(q, k, v, o, do, lse, H, d ** -0.5, dqkv[:, :, :Cc], dqkv[:, :, Cc:2 * Cc], dqkv[:, :, 2 * Cc:], key_mask=km)
    assert rel(dqkv[:, :, :Cc], rq) < GRAD_TOL and rel(dqkv[:, :, Cc:2 * Cc], rk) < GRAD_TOL and rel(dqkv[:, :, 2 * Cc:], rv) < GRAD_TOL
    dead = (km == 0)[:, :, None].expand(B, N, Cc)
    assert dqkv[:, :, Cc:2 * Cc][dead].abs().max().item() == 0 and dqkv[:, :, 2 * Cc:][dead].abs().max().item() == 0


@pytest.mark.parametrize("mult,T", [(1, 20), (2, 20), (4, 24), (2, 77)])
def test_attention_bwd_causal_multi_kv(mult, T):
    """CLIPAttentionMKV (arc2face_models.py:145-231): token t carries its M keys back to back."""
    BS, H, d = 3, 12, 64
    E = H * d
    qkv, do = rnd(BS, T, E * (1 + 2 * mult), seed=1), rnd(BS, T, E, seed=2)
    q, k, v = qkv[:, :, :E], qkv[:, :, E:E + E * mult], qkv[:, :, E + E * mult:]
    lse = torch.empty(BS, H, T, device="cuda")
    o = ops().attention(q, k, v, H, d ** -0.5, causal_mult=mult, lse=lse)
    _, rq, rk, rv = ref_attn_grads(q, k, v, do, H, d ** -0.5, causal_mult=mult)
    dqkv = torch.empty_like(qkv)
    ops().attention_bwd(q, k, v, o, do, lse, H, d ** -0.5, dqkv[:, :, :E], dqkv[:, :, E:E + E * mult], dqkv[:, :, E + E * mult:],
                        causal_mult=mult)
    assert rel(dqkv[:, :, :E], rq) < GRAD_TOL
    assert rel(dqkv[:, :, E:E + E * mult], rk) < GRAD_TOL and rel(dqkv[:, :, E + E * mult:], rv) < GRAD_TOL


# ------------------------------------------------------------------------------------------- capture backward
def _heads(t, H):
    B, L, Cc = t.shape
    return t.view(B, L, H, Cc // H).transpose(1, 2)


@pytest.mark.parametrize("d,Lq,S,dt,mode", [(40, 300, 77, torch.float32, "normalize"), (40, 256, 97, BF, "plain"),
                                            (80, 100, 77, torch.float32, "normalize"), (40, 4096, 77, torch.float32, "plain"),
                                            (40, 200, 77, torch.float32, "mix"), (40, 64, 128, BF, "mix")])
def test_cross_capture_bwd(d, Lq, S, dt, mode):
    B, H = 2, 8
    Cc = H * d
    q, k, v = rnd(B, Lq, Cc, seed=1, dtype=dt), rnd(B, S, Cc, seed=2, dtype=dt), rnd(B, S, Cc, seed=3, dtype=dt)
    do = rnd(B, Lq, Cc, seed=4)
    dprob, dscore = rnd(B, H, Lq, S, seed=5, dtype=torch.float32), rnd(B, H, Lq, S, seed=6, std=0.05, dtype=torch.float32)
    ca = torch.tensor(0.8, device="cuda")
    si = C.subj_indices(B) if mode == "normalize" else None
    col_flag = qm = None
    if si is not None:
        col_flag = torch.zeros(B, S, dtype=torch.uint8, device="cuda")
        col_flag[torch.from_numpy(si[0]).cuda(), torch.from_numpy(si[1]).cuda()] = 1
        qm = ops().qmean(q)
    dq, dk, dv, dca = ops().attention_cross_capture_bwd(q, k, v, do, H, d ** -0.5, dprob=dprob, dscore=dscore, col_flag=col_flag,
                                                        qmean=qm, ca_scale=ca.reshape(1), mix=mode == "mix", dca_mul=10.0,
                                                        dkv_dtype=dt)
    # reference: autograd through the oracle's slow SDPA (dalc:79-139) on the CPU
    qc, kc, vc = (t.float().cpu().requires_grad_(True) for t in (q, k, v))
    cac = ca.cpu().clone().requires_grad_(True)
    sic = None if si is None else (torch.from_numpy(si[0]), torch.from_numpy(si[1]))
    out, score, prob = oracle.slow_sdpa(_heads(qc, H), _heads(kc, H), _heads(vc, H), cac, subj_indices=sic,
                                        normalize_cross_attn=mode == "normalize", mix_attn_mats_in_batch=mode == "mix")
    loss = (out.transpose(1, 2).reshape(B, Lq, Cc) * do.float().cpu()).sum() + (prob * dprob.cpu()).sum() + (score * dscore.cpu()).sum()
    loss.backward()
    assert rel(dq, qc.grad) < GRAD_TOL and rel(dk, kc.grad) < GRAD_TOL and rel(dv, vc.grad) < GRAD_TOL
    if mode == "normalize":
        assert abs(dca.item() - cac.grad.item()) < GRAD_TOL * abs(cac.grad.item())
    if mode == "mix":       # the mc half is detached (dalc:117)
        assert dq[B // 2:].abs().max().item() == 0 and dk[B // 2:].float().abs().max().item() == 0


# ------------------------------------------------------------------------------------------- HBM-bound helpers
def test_transpose_colsum():
    x = rnd(3, 200, 77, seed=1, dtype=torch.float32)
    cs, rs = rnd(77, seed=2, dtype=torch.float32), rnd(200, seed=3, dtype=torch.float32)
    y = ops().transpose(x, out_dtype=BF, alpha=0.5, colscale=cs, rowscale=rs, pad_to=8)
    ref = (0.5 * x * cs[None, None, :] * rs[None, :, None]).transpose(1, 2)
    assert tuple(y.shape) == (3, 77, 200) and rel(y, ref) < 1e-2
    y2 = ops().transpose(x[0, :, :75].to(BF), pad_to=8)                 # odd sizes: padded to 8 with zeros
    assert tuple(y2.shape) == (75, 200)
    y3 = ops().transpose(rnd(45, 64, seed=4), pad_to=8)
    assert tuple(y3.shape) == (64, 48) and y3[:, 45:].abs().max().item() == 0
    a, b = rnd(1000, 333, seed=5), rnd(1000, 333, seed=6, dtype=torch.float32)
    bias, cm = rnd(333, seed=7, dtype=torch.float32), rnd(333, seed=8, dtype=torch.float32)
    assert rel(ops().colsum(a), a.float().sum(0)) < 1e-3
    assert rel(ops().colsum(a, b=b, bias=bias, colmul=cm), (a.float() * (b - bias)).sum(0) * cm) < 1e-3


@pytest.mark.parametrize("Cc", [320, 768, 1280])
@pytest.mark.parametrize("xdt", [BF, torch.float32])
def test_layernorm_bwd(Cc, xdt):
    M = 300
    x, dy = rnd(M, Cc, seed=1, dtype=xdt), rnd(M, Cc, seed=2)
    w, b = (1 + 0.1 * rnd(Cc, seed=3, dtype=torch.float32)), rnd(Cc, seed=4, std=0.1, dtype=torch.float32)
    xr, wr, br = x.float().clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    F.layer_norm(xr, (Cc,), wr, br, 1e-5).backward(dy.float())
    dx, dw, db = ops().layernorm_bwd(x, dy, w, 1e-5, want_wgrad=True)
    assert dx.dtype == xdt and rel(dx, xr.grad) < (2e-2 if xdt == BF else 1e-4)
    assert rel(dw, wr.grad) < 1e-3 and rel(db, br.grad) < 1e-3
    dx2, dw2, _ = ops().layernorm_bwd(x, dy, w, 1e-5)
    assert dw2 is None and torch.equal(dx2, dx)


def test_activations_fwd_bwd():
    o = ops()
    M, I = 200, 1280
    u, dh = rnd(M, I, seed=1), rnd(M, I, seed=2)
    ur = u.float().clone().requires_grad_(True)
    hr = ur * torch.sigmoid(1.702 * ur)
    hr.backward(dh.float())
    assert rel(o.act_fwd(u, o.ACT_QUICK_GELU), hr) < 1e-2 and rel(o.act_bwd(u, dh, o.ACT_QUICK_GELU), ur.grad) < 1e-2
    # GEGLU on packed [a(64) | gate(64)] tiles
    up = rnd(M, 2 * I, seed=3)
    t = up.float().view(M, I // 64, 2, 64)
    a_, g_ = t[:, :, 0].reshape(M, I).clone().requires_grad_(True), t[:, :, 1].reshape(M, I).clone().requires_grad_(True)
    hr = a_ * F.gelu(g_)
    hr.backward(dh.float())
    assert rel(o.act_fwd(up, o.ACT_GEGLU), hr) < 1e-2
    du = o.act_bwd(up, dh, o.ACT_GEGLU).float().view(M, I // 64, 2, 64)
    assert rel(du[:, :, 0].reshape(M, I), a_.grad) < 1e-2 and rel(du[:, :, 1].reshape(M, I), g_.grad) < 1e-2


def test_sbg_head_bwd():
    M, E = 100, 768
    hs = [rnd(M, E, seed=i, dtype=torch.float32) for i in range(3)]
    wl = [1 / 7, 2 / 7, 4 / 7]
    w, b = (1 + 0.1 * rnd(E, seed=5, dtype=torch.float32)), rnd(E, seed=6, std=0.1, dtype=torch.float32)
    dout = rnd(M, E, seed=7, dtype=torch.float32)
    hr = [h.clone().requires_grad_(True) for h in hs]
    wlr, wr, br = torch.tensor(wl, device="cuda", requires_grad=True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    F.layer_norm(sum(wlr[i] * hr[i] for i in range(3)), (E,), wr, br, 1e-5).backward(dout)
    dhs, dwl, dw, db = ops().sbg_head_bwd(hs, wl, w, dout, 1e-5)
    for i in range(3):
        assert rel(dhs[i], hr[i].grad) < 1e-4
    assert rel(dwl, wlr.grad) < 1e-3 and rel(dw, wr.grad) < 1e-3 and rel(db, br.grad) < 1e-3


# ------------------------------------------------------------------------------------------- linear Functions
@pytest.mark.parametrize("M,N,K,R", [(512, 320, 320, 8), (154, 320, 768, 16), (4100, 320, 320, 192)])
def test_lora_linear_grads(M, N, K, R):
    """LoraLinearFn against autograd through the oracle's peft restatement."""
    import adaface_dev_b200 as a
    base = torch.nn.Linear(K, N, device="cuda")
    with torch.no_grad():
        base.weight.copy_(rnd(N, K, std=K ** -0.5, seed=1).float())
        base.bias.copy_(rnd(N, std=0.02, seed=2).float())
    lora = a.LoraDoraLinear(base, r=R, lora_alpha=R / 8).cuda()
    with torch.no_grad():
        lora.lora_A["default"].weight.copy_(rnd(R, K, std=K ** -0.5, seed=3).float())
        lora.lora_B["default"].weight.copy_(rnd(N, R, std=0.02, seed=4).float())
        lora.lora_magnitude_vector["default"].weight.mul_(1 + 0.1 * rnd(N, seed=5, dtype=torch.float32))
    x = rnd(M, K, seed=6).requires_grad_(True)
    dy = rnd(M, N, seed=7)
    pack = {"w": base.weight.detach().to(BF).contiguous(), "b": base.bias.detach().float().contiguous()}
    y = ag().linear(x, pack, "w", "b", lora=lora)
    y.backward(dy)
    A, Bm, mag = (lora.lora_A["default"].weight, lora.lora_B["default"].weight, lora.lora_magnitude_vector["default"].weight)
    xr = x.detach().float().cpu().requires_grad_(True)
    Ar, Br, mr = (t.detach().cpu().clone().requires_grad_(True) for t in (A, Bm, mag))
    yr = oracle.lora_dora_linear(xr, base.weight.detach().cpu(), base.bias.detach().cpu(), Ar, Br, mr, lora.scaling)
    yr.backward(dy.float().cpu())
    assert rel(y, yr) < 2e-2
    assert rel(x.grad, xr.grad) < GRAD_TOL
    assert rel(A.grad, Ar.grad) < GRAD_TOL and rel(Bm.grad, Br.grad) < GRAD_TOL and rel(mag.grad, mr.grad) < GRAD_TOL


@pytest.mark.parametrize("M", [40, 1280])
def test_train_linear_grads(M):
    """TrainLinearFn (fused rows of several nn.Linear, fp32 residual) against F.linear autograd."""
    K, Ns = 768, (768, 1536, 1536)
    lins = [torch.nn.Linear(K, n, device="cuda") for n in Ns]
    pack = {"w": torch.cat([l.weight.detach().to(BF) for l in lins]).contiguous(),
            "b": torch.cat([l.bias.detach().float() for l in lins]).contiguous()}
    x = rnd(M, K, seed=1).requires_grad_(True)
    res = rnd(M, sum(Ns), seed=2, dtype=torch.float32).requires_grad_(True)
    dy = rnd(M, sum(Ns), seed=3, dtype=torch.float32)
    params = sum(((l.weight, l.bias) for l in lins), ())
    y = ag().linear(x, pack, "w", "b", params=params, residual=res, out_dtype=torch.float32)
    y.backward(dy)
    xr = x.detach().float().requires_grad_(True)
    wr = [l.weight.detach().to(BF).float().requires_grad_(True) for l in lins]
    br = [l.bias.detach().clone().requires_grad_(True) for l in lins]
    yr = torch.cat([F.linear(xr, w_, b_) for w_, b_ in zip(wr, br)], dim=1) + res.detach()
    yr.backward(dy)
    assert rel(y, yr) < 1e-2 and rel(x.grad, xr.grad) < GRAD_TOL and torch.equal(res.grad, dy)
    for l, w_, b_ in zip(lins, wr, br):
        assert rel(l.weight.grad, w_.grad) < GRAD_TOL and rel(l.bias.grad, b_.grad) < GRAD_TOL


# ------------------------------------------------------------------------------------------- mirrors vs the oracle
def _loss_weights(shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


TRAIN_PROC_CASES = ["proc_cross_norm_lora", "proc_cross_norm_lora_qupd", "proc_cross_mix_lora", "proc_cross_capture",
                    "proc_cross_fast", "proc_self_mask", "proc_self_d160", "proc_cross_capture_d80"]


@pytest.mark.parametrize("name", TRAIN_PROC_CASES)
def test_processor_training_step_vs_oracle(name):
    """One forward + backward through AttnProcessor_LoRA_Capture in training mode: the loss touches the output and every
    cached activation the stage-2 losses consume (attn, attnscore, k, v, q2, attn_out: ldm/util.py:1822-1918, 2047-2121)."""
    case = C.build_proc_case(name)
    sp = case["spec"]
    t = C.to_torch({k: v for k, v in case.items() if k != "spec"})
    w = t["w"]
    leaves = {"hidden_states": t["hidden_states"].requires_grad_(True)}
    if t["encoder_hidden_states"] is not None:
        leaves["encoder_hidden_states"] = t["encoder_hidden_states"].requires_grad_(True)
    w["cross_attn_scale_factor"] = torch.tensor(float(w["cross_attn_scale_factor"]), requires_grad=True)
    for n in ("q", "k", "v", "out"):
        if "lora_" + n in w:
            w["lora_" + n] = tuple(p.clone().requires_grad_(True) for p in w["lora_" + n])
    ref_out, ref_cache = oracle.processor_forward(
        w, t["hidden_states"], t["encoder_hidden_states"], img_mask=t["img_mask"], subj_indices=t["subj_indices"],
        capture_ca_activations=sp.get("capture", False), normalize_cross_attn=sp.get("normalize", False),
        mix_attn_mats_in_batch=sp.get("mix", False), enable_lora=sp.get("enable_lora", False),
        q_lora_updates_query=sp.get("q_upd", False), lora_scaling=float(w.get("lora_scaling", 0.125)))
    out, cache, h = run_mirror_proc(case, train=True)
    gw = {"out": _loss_weights(ref_out.shape, 1)}
    for i, key in enumerate(("attn", "attnscore", "k", "v", "q2", "attn_out")):
        if key in ref_cache:
            gw[key] = _loss_weights(ref_cache[key].shape, 10 + i) * (0.05 if key == "attnscore" else 1.0)
    ref_loss = (ref_out * gw["out"]).sum() + sum((ref_cache[k] * gw[k]).sum() for k in gw if k != "out")
    ref_loss.backward()
    loss = (out.float() * gw["out"].cuda()).sum() + sum((cache[k] * gw[k].cuda()).sum() for k in gw if k != "out")
    loss.backward()
    assert rel(out, ref_out) < 2e-2
    assert rel(h["hidden_states"].grad, leaves["hidden_states"].grad) < GRAD_TOL
    if "encoder_hidden_states" in leaves:
        assert rel(h["encoder_hidden_states"].grad, leaves["encoder_hidden_states"].grad) < GRAD_TOL
    proc = h["proc"]
    if sp.get("enable_lora"):
        for n in ("q", "k", "v", "out"):
            mod = getattr(proc, f"to_{n}_lora")
            A, Bm, mag = w["lora_" + n]
            if A.grad is None:      # the q adapter is unused on the self-attention path etc.
                continue
            assert rel(mod.lora_A["default"].weight.grad, A.grad) < GRAD_TOL, n
            assert rel(mod.lora_B["default"].weight.grad, Bm.grad) < GRAD_TOL, n
            assert rel(mod.lora_magnitude_vector["default"].weight.grad, mag.grad) < GRAD_TOL, n
    if sp.get("normalize"):
        g, gr = proc.cross_attn_scale_factor.grad.item(), w["cross_attn_scale_factor"].grad.item()
        assert abs(g - gr) < GRAD_TOL * abs(gr)
    at = h["attn"]
    for mod in (at.to_q, at.to_k, at.to_v, at.to_out[0]):        # the frozen base weights never receive a gradient
        assert all(p.grad is None for p in mod.parameters())


@pytest.mark.parametrize("name", ["ldm_block", "ldm_block_d80", "ldm_cross_save", "ldm_self_mask", "ldm_self_mask_empty"])
def test_ldm_training_step_vs_oracle(name):
    case = C.build_ldm_case(name)
    sp = case["spec"]
    t = C.to_torch({k: v for k, v in case.items() if k != "spec"})
    x = t["x"].requires_grad_(True)
    ctx = t["context"].requires_grad_(True) if t["context"] is not None else None
    if sp.get("block"):
        ref, ref_cache = oracle.basic_transformer_block(t["w"], x, context=ctx, mask=t["mask"]), None
    else:
        ref, ref_cache = oracle.ldm_cross_attention(t["w"], x, context=ctx, mask=t["mask"], save_cross_attn_vars=sp.get("save", False))
    out, cache = run_mirror_ldm(case, train=True)
    mx, mctx = case["_leaves"]
    gw = _loss_weights(ref.shape, 1)
    ref_loss, loss = (ref * gw).sum(), (out.float() * gw.cuda()).sum()
    if ref_cache:
        for i, key in enumerate(("attn", "q", "attn_out")):
            g = _loss_weights(ref_cache[key].shape, 20 + i)
            ref_loss = ref_loss + (ref_cache[key] * g).sum()
            loss = loss + (cache[key] * g.cuda()).sum()
    ref_loss.backward()
    loss.backward()
    assert rel(out, ref) < 2e-2
    assert rel(mx.grad, x.grad) < GRAD_TOL
    if ctx is not None:
        assert rel(mctx.grad, ctx.grad) < GRAD_TOL


@pytest.mark.parametrize("name", ["sbg_m1", "sbg_mixed_sfx"])
def test_sbg_training_step_vs_oracle(name):
    """SubjBasisGenerator forward + backward (stage-2 trains it through the prompt context): gradients of the input
    embeddings, of representative encoder parameters and of hidden_state_layer_weights (x5 GradientScaler,
    subj_basis_generator.py:786) against autograd through the oracle at the dense T = 77."""
    case = C.build_sbg_case(name)
    sp = case["spec"]
    n_sfx = sp.get("n_sfx", 0)
    t = C.to_torch({k: v for k, v in case.items() if k != "spec"})
    w = t["w"]
    x = t["faceid2img_prompt_embs"].requires_grad_(True)
    # (k_proj.bias is not probed: softmax is invariant to it, its true gradient is 0 and only rounding noise remains)
    probe = [(0, "q_w"), (0, "v_b"), (2, "k_w"), (5, "v_w"), (5, "o_w"), (11, "fc1_w"), (11, "fc2_b"), (3, "ln1_w"), (7, "ln2_b")]
    for li, key in probe:
        w["layers"][li][key].requires_grad_(True)
    w["final_ln_w"].requires_grad_(True)
    w["hidden_state_layer_weights"].requires_grad_(True)
    if n_sfx:
        w["static_img_suffix_embs"].requires_grad_(True)
    ref = oracle.sbg_forward(w, x, enable_static_img_suffix_embs=bool(n_sfx), multipliers=sp["mults"])
    gw = _loss_weights(ref.shape, 3)
    (ref * gw).sum().backward()

    gen = make_sbg(case["w"], sp["mults"], n_sfx)
    xm = _T(case["faceid2img_prompt_embs"]).requires_grad_(True)
    out = gen(xm, enable_static_img_suffix_embs=bool(n_sfx))
    (out * gw.cuda()).sum().backward()
    assert rel(out, ref) < 3e-2
    tol = 5e-2                          # 12 layers of bf16 GEMM gradients
    assert rel(xm.grad, x.grad) < tol
    layers = gen.prompt2token_proj.text_model.encoder.layers
    name_of = {"q_w": lambda l: l.self_attn.q_proj.weight, "v_b": lambda l: l.self_attn.v_proj.bias,
               "k_w": lambda l: l.self_attn.k_proj.weight,
               "v_w": lambda l: l.self_attn.v_proj.weight, "o_w": lambda l: l.self_attn.out_proj.weight,
               "fc1_w": lambda l: l.mlp.fc1.weight, "fc2_b": lambda l: l.mlp.fc2.bias,
               "ln1_w": lambda l: l.layer_norm1.weight, "ln2_b": lambda l: l.layer_norm2.bias}
    for li, key in probe:
        assert rel(name_of[key](layers[li]).grad, w["layers"][li][key].grad) < tol, (li, key)
    assert rel(gen.prompt2token_proj.text_model.final_layer_norm.weight.grad, w["final_ln_w"].grad) < tol
    assert rel(gen.hidden_state_layer_weights.grad, 5.0 * w["hidden_state_layer_weights"].grad) < tol
    if n_sfx:
        assert rel(gen.static_img_suffix_embs.grad, w["static_img_suffix_embs"].grad) < tol
    assert gen.prompt2token_proj.text_model.embeddings.token_embedding.weight.grad is None      # frozen (:841-853)
