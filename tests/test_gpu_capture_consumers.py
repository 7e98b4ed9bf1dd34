"""GPU parity of the FUSED CAPTURE CONSUMERS: the subject-column sum and the sc vs sc_rep squared difference
reduced inside the capture kernel (forward and backward), and the two stage-2 losses built on them
(ldm/util.py:1822-1918 calc_subj_masked_bg_suppress_loss, :2047-2121 calc_sc_rep_attn_distill_loss), against
  (1) the CPU oracle's slow SDPA (dalc:79-139) + plain tensor reductions of its probability map, values and gradients;
  (2) the committed closs_* fixtures = outputs of the reference's own loss functions (bar 1e-3 relative);
  (3) the un-fused capture path of the same processor (the full map, then reduced)."""
import os

import numpy as np
import pytest
import torch

import cases as C
import oracle
from oracle import capture_losses_oracle as cl
from mirror_utils import run_mirror_proc, _T
from parity_log import record

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
H = 8


def _rel(a, ref):
    a, ref = a.detach().float().cpu(), ref.detach().float().cpu()
    return ((a - ref).abs().max() / ref.abs().max().clamp_min(1e-20)).item()


def _inputs(B, N, S, Cc=320, seed=0):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, N, Cc, generator=g)
    k = torch.randn(B, S, Cc, generator=g)
    v = torch.randn(B, S, Cc, generator=g)
    ref = torch.softmax(torch.randn(B, H, N, S, generator=g) * 2, dim=-1)
    flag = torch.zeros(B, S, dtype=torch.uint8)
    for b in range(B):
        flag[b, 4 + b:20 + b] = 1
    return q, k, v, ref, flag


def _oracle_prob(q, k, v, flag, normalize, ca):
    B, N, Cc = q.shape
    hd = lambda t: t.view(B, -1, H, Cc // H).transpose(1, 2)
    si = None
    if normalize:
        ib, in_ = flag.nonzero(as_tuple=True)
        si = (ib, in_)
    out, _, prob = oracle.slow_sdpa(hd(q), hd(k), hd(v), ca, subj_indices=si, normalize_cross_attn=normalize)
    return out.transpose(1, 2).reshape(B, N, Cc), prob


@pytest.mark.parametrize("B,N,S,normalize", [(2, 200, 77, False), (1, 256, 97, True), (3, 64, 77, True)])
def test_consume_kernel_forward_vs_oracle(B, N, S, normalize):
    import adaface_dev_b200 as a
    q, k, v, ref, flag = _inputs(B, N, S, seed=B + N)
    ca = torch.tensor(0.8)
    out_ref, prob = _oracle_prob(q, k, v, flag, normalize, ca)
    sum_ref = (prob * flag[:, None, None, :].float()).sum(-1)
    sq_ref = ((prob - ref) ** 2).sum(dim=(1, 2, 3))
    qc, kc, vc = q.cuda(), k.cuda(), v.cuda()
    kw = {}
    if normalize:
        kw = dict(col_flag=flag.cuda(), qmean=a.ops.qmean(qc), ca_scale=ca.cuda().reshape(1))
    out, ssum, sqd, p_full = a.ops.attention_cross_consume(qc, kc, vc, H, 40 ** -0.5, sum_flag=flag.cuda(), ref_prob=ref.cuda(),
                                                           want_prob=True, **kw)
    e_sum = (ssum.cpu() - sum_ref).abs().max().item()
    e_sq = _rel(sqd, sq_ref)
    record("capture_consumers", f"B{B}_N{N}_S{S}_norm{int(normalize)}", "subj_sum max-abs", e_sum, 1e-3)
    record("capture_consumers", f"B{B}_N{N}_S{S}_norm{int(normalize)}", "sqdiff rel", e_sq, 1e-3)
    assert e_sum < 1e-3 and e_sq < 1e-3 and (out.float().cpu() - out_ref).abs().max().item() < 2e-2
    assert (p_full.cpu() - prob).abs().max().item() < 1e-3
    # the reductions are those of the map the same kernel would write: agree to fp32 summation order
    assert (ssum - (p_full * flag.cuda()[:, None, None, :].float()).sum(-1)).abs().max().item() < 1e-5
    assert _rel(sqd, ((p_full - ref.cuda()) ** 2).sum(dim=(1, 2, 3))) < 1e-5
    # either consumer alone, and none of the map written
    o2, s2, q2, p2 = a.ops.attention_cross_consume(qc, kc, vc, H, 40 ** -0.5, sum_flag=flag.cuda(), **kw)
    assert q2 is None and p2 is None and torch.equal(s2, ssum) and torch.equal(o2, out)
    o3, s3, q3, _ = a.ops.attention_cross_consume(qc, kc, vc, H, 40 ** -0.5, ref_prob=ref.cuda(), **kw)
    assert s3 is None and torch.equal(q3, sqd)


@pytest.mark.parametrize("B,N,S,normalize", [(1, 256, 97, True), (2, 200, 77, False)])
def test_consume_backward_vs_autograd_through_oracle(B, N, S, normalize):
    import adaface_dev_b200.autograd as ag
    q, k, v, ref, flag = _inputs(B, N, S, seed=7 * B + N)
    g = torch.Generator().manual_seed(3)
    w_out, w_sum, w_sq = torch.randn(B, N, 320, generator=g), torch.randn(B, H, N, generator=g), torch.rand(B, generator=g) + 0.5
    # -- reference: autograd through the oracle
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
    ca_r = torch.tensor(0.8, requires_grad=True)
    out_r, prob = _oracle_prob(qr, kr, vr, flag, normalize, ca_r)
    loss_r = (out_r * w_out).sum() + ((prob * flag[:, None, None, :].float()).sum(-1) * w_sum).sum() \
        + (((prob - ref) ** 2).sum(dim=(1, 2, 3)) * w_sq).sum() * 50
    loss_r.backward()
    # -- fused path
    qc, kc, vc = (t.cuda().requires_grad_(True) for t in (q, k, v))
    ca = torch.nn.Parameter(torch.tensor(0.8, device="cuda"))
    out, ssum, sqd = ag.CrossConsumeFn.apply(qc, kc, vc, ca, H, 40 ** -0.5, flag.cuda() if normalize else None, flag.cuda(), ref.cuda(), 10.0)
    loss = (out.float() * w_out.cuda()).sum() + (ssum * w_sum.cuda()).sum() + (sqd * w_sq.cuda()).sum() * 50
    loss.backward()
    errs = {n: _rel(a_.grad, b_.grad) for n, a_, b_ in (("dq", qc, qr), ("dk", kc, kr), ("dv", vc, vr))}
    for n, e in errs.items():
        record("capture_consumers", f"bwd_B{B}_N{N}_S{S}_norm{int(normalize)}", f"{n} rel max-abs", e, 3e-2)
    assert all(e < 3e-2 for e in errs.values()), errs
    if normalize:
        assert abs(ca.grad.item() - ca_r.grad.item()) < 3e-2 * abs(ca_r.grad.item()) + 1e-6


@pytest.mark.parametrize("name", list(C.CLOSS_CASES))
def test_fused_losses_vs_reference_fixtures(name):
    """The loss arithmetic on top of the reduced quantities against the reference's own functions (fixtures closs_*): the
    reduced inputs are formed here from the fixture's probability maps exactly as the kernel forms them."""
    import adaface_dev_b200.capture_losses as fl
    case = C.build_closs_case(name)
    sp = case["spec"]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    t = {k_: (_T(v_) if isinstance(v_, np.ndarray) and v_.dtype.kind == "f" else v_) for k_, v_ in case.items() if k_ != "spec"}
    ib, it = torch.from_numpy(case["subj_ib"]).cuda(), torch.from_numpy(case["subj_it"]).cuda()
    if sp["kind"] == "bg":
        sums = {}
        for li, key in ((23, "attn23"), (24, "attn24")):
            B, S = t[key].shape[0], t[key].shape[3]
            flag = torch.zeros(B, S, device="cuda")
            flag[ib, it] = 1
            sums[li] = (t[key] * flag[:, None, None, :]).sum(-1)
        loss = fl.calc_subj_masked_bg_suppress_loss(sums, (ib, it), sp["block"], t["fg_mask"])
        e = abs(float(loss) - float(g["loss"])) / abs(float(g["loss"]))
        record("capture_consumers", name, "bg-suppress loss rel", e, 1e-3)
        assert e < 1e-3
    else:
        sq = {li: ((t[key][1:2] - t[key][2:3]) ** 2).sum().reshape(1) for li, key in ((23, "attn23"), (24, "attn24"))}
        shp = {li: tuple(t[key].shape[1:]) for li, key in ((23, "attn23"), (24, "attn24"))}
        out = fl.calc_sc_rep_attn_distill_loss(sq, shp, {23: t["k23"], 24: t["k24"]}, {23: t["v23"], 24: t["v24"]}, (ib, it),
                                               t["emb_mask"], t["pad_mask"], sp["fg_percent"])
        got = torch.stack([torch.as_tensor(o, dtype=torch.float32).cpu() for o in out])
        ref = torch.from_numpy(g["losses"])
        e = ((got - ref).abs() / ref.abs().clamp_min(1e-12)).max().item() if (ref != 0).any() else got.abs().max().item()
        record("capture_consumers", name, "distill losses rel (5 terms)", e, 1e-3)
        assert e < 1e-3, (got, ref)


def test_processor_consumers_match_full_capture_and_train():
    """Through the drop-in processor (dalc:192-364 surface): set_capture_consumers() yields the same numbers as capturing the
    full map and reducing it, both without and with autograd, and gradients reach the hidden states, the prompt and the LoRAs."""
    case = C.build_proc_case("proc_cross_norm_lora")
    sp = case["spec"]
    out_full, cache_full = run_mirror_proc(case)
    si = case["subj_indices"]
    flag = torch.zeros(sp["B"], sp["S"], device="cuda")
    flag[torch.from_numpy(si[0]).cuda(), torch.from_numpy(si[1]).cuda()] = 1
    sum_ref = (cache_full["attn"] * flag[:, None, None, :]).sum(-1)
    ref_map = torch.softmax(torch.randn(cache_full["attn"].shape, generator=torch.Generator().manual_seed(1)) * 2, -1).cuda()
    sq_ref = ((cache_full["attn"] - ref_map) ** 2).sum(dim=(1, 2, 3))

    def hook(proc):
        proc.set_capture_consumers(subj_sum=True, ref_attn=ref_map)
    import mirror_utils
    orig = mirror_utils.a_processor_hook if hasattr(mirror_utils, "a_processor_hook") else None
    mirror_utils.a_processor_hook = hook
    try:
        out_c, cache_c = run_mirror_proc(case)
        out_t, cache_t, hd = run_mirror_proc(case, train=True)
    finally:
        mirror_utils.a_processor_hook = orig
    assert torch.equal(out_c, out_full)
    assert cache_c.get("attn") is None and (cache_c["attn_subj_sum"] - sum_ref).abs().max().item() < 1e-5
    assert _rel(cache_c["attn_sqdiff"], sq_ref) < 1e-5
    assert (cache_t["attn_subj_sum"] - sum_ref).abs().max().item() < 2e-3 and _rel(cache_t["attn_sqdiff"], sq_ref) < 5e-3
    (cache_t["attn_subj_sum"].sum() + cache_t["attn_sqdiff"].sum() + out_t.float().sum()).backward()
    assert hd["hidden_states"].grad is not None and hd["encoder_hidden_states"].grad.abs().max().item() > 0
    # (the q adapter only feeds the cached q2 unless q_lora_updates_query, dalc:239-249: probe the k / v / out adapters)
    for n in ("k", "v", "out"):
        assert getattr(hd["proc"], f"to_{n}_lora").lora_A["default"].weight.grad.abs().max().item() > 0, n
