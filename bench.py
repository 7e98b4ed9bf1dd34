#!/usr/bin/env python
"""bench.py -- SD-1.5 attention stack throughput on B200 (BASELINE.json metric) + roofline + CPU baseline.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config.workload = "sd15_attn_stack_512"): the 32 attention modules of one SD-1.5 U-Net denoising step
at 512x512 (64x64 latent) -- 16 self- and 16 cross-attention modules incl. their Q/K/V/out projections, levels
A (4096 tok, 320 ch) x5, B (1024, 640) x5, C (256, 1280) x5, D (64, 1280) x1, 8 heads, 77-token context with the
16 ada tokens spliced into rows 4:20 -- for a batch of 8 (BASELINE config 3), driven through
the reference-facing operator ``AttnProcessor_LoRA_Capture.__call__``.  One "step" = one pass over that batch.
value = algorithmic TFLOP/s (closed forms of SURVEY 8d: self 8NC^2 + 4N^2C, cross 4NC^2 + 4*S*768*C + 4NSC per
sample and block), whole job over all ranks (weak scaling: every rank processes its own batch of 8).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LEVELS = [("A", 4096, 320, 5), ("B", 1024, 640, 5), ("C", 256, 1280, 5), ("D", 64, 1280, 1)]
HEADS, S_CTX, CTX_DIM, BATCH = 8, 77, 768, 8
METRIC, UNIT = "SD1.5 attn-block TFLOP/s", "TFLOP/s"


def flops_per_sample(levels=LEVELS):
    f = 0.0
    for _, N, C, nblk in levels:
        self_f = 8 * N * C * C + 4 * N * N * C
        cross_f = 4 * N * C * C + 4 * S_CTX * CTX_DIM * C + 4 * N * S_CTX * C
        f += nblk * (self_f + cross_f)
    return f


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update({k: m[k] for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in m})
        p["source"] = "measured"
    except Exception:
        pass
    return p


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_sample_step(torch, oracle, weights, xs, ctx):
    """Bounded sample of the workload for the CPU arms: ONE sample (B=1), ONE block per level (4 of the 16
    blocks): self-attention + cross-attention through the oracle's restatement of the processor."""
    for (name, N, C, _), w, x in zip(LEVELS, weights, xs):
        y, _ = oracle.processor_forward(w["self"], x, None, heads=HEADS)
        oracle.processor_forward(w["cross"], y, ctx, heads=HEADS)


def cpu_setup(torch):
    import oracle
    g = torch.Generator().manual_seed(0)
    weights, xs = [], []
    for _, N, C, _ in LEVELS:
        mk = lambda o, i: torch.randn(o, i, generator=g) / i ** 0.5
        weights.append({kind: {"to_q": mk(C, C), "to_k": mk(C, cd), "to_v": mk(C, cd), "to_out_w": mk(C, C),
                               "to_out_b": torch.zeros(C), "cross_attn_scale_factor": torch.tensor(0.8)}
                        for kind, cd in (("self", C), ("cross", CTX_DIM))})
        xs.append(torch.randn(1, N, C, generator=g))
    ctx = torch.randn(1, S_CTX, CTX_DIM, generator=g)
    sample_levels = [(n, N, C, 1) for n, N, C, _ in LEVELS]
    return oracle, weights, xs, ctx, flops_per_sample(sample_levels)


def run_cpu(torch, steps, warmup, min_seconds=0.0):
    """Times `steps` sample steps (or, with min_seconds, as many as fit in about that much CPU time)."""
    torch.set_num_threads(os.cpu_count())
    oracle, weights, xs, ctx, fl = cpu_setup(torch)
    with torch.no_grad():
        for _ in range(warmup):
            cpu_sample_step(torch, oracle, weights, xs, ctx)
        t0 = time.perf_counter()
        n = 0
        while n < steps or (time.perf_counter() - t0) < min_seconds:
            cpu_sample_step(torch, oracle, weights, xs, ctx)
            n += 1
        dt = (time.perf_counter() - t0) / n
    return fl / dt / 1e12, dt, fl, n


def run_reference_ldm(torch, steps, warmup):
    """The UNMODIFIED reference operator from baseline/_ref (pip-installed copy of /root/reference):
    ``ldm.modules.attention.CrossAttention`` (ldm/modules/attention.py:146-222), the LDM surface of the same attention
    operator, on the same bounded sample as the port.  Returns None when baseline/_ref is absent or not importable."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "ldm")):
        return None
    sys.path.insert(0, ref)
    try:
        from ldm.modules.attention import CrossAttention
    except Exception:
        return None
    finally:
        sys.path.remove(ref)
    torch.manual_seed(0)
    mods, xs = [], []
    for _, N, C, _ in LEVELS:
        mods.append((CrossAttention(C, None, HEADS, C // HEADS).eval(), CrossAttention(C, CTX_DIM, HEADS, C // HEADS).eval()))
        xs.append(torch.randn(1, N, C))
    ctx = torch.randn(1, S_CTX, CTX_DIM)
    fl = flops_per_sample([(n, N, C, 1) for n, N, C, _ in LEVELS])

    def step():
        for (a1, a2), x in zip(mods, xs):
            a2(a1(x), context=ctx)
    with torch.no_grad():
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = (time.perf_counter() - t0) / steps
    return {"value": fl / dt / 1e12, "unit": UNIT, "ms_per_step": dt * 1e3, "kind": "reference",
            "what": "baseline/_ref ldm.modules.attention.CrossAttention, verbatim, same sample (einsum path: [B*8,N,N] scores)"}


def run_reference_unet(torch):
    """The UNMODIFIED reference U-Net from baseline/_ref (``ldm.modules.diffusionmodules.openaimodel.UNetModel``, SD-1.5
    configuration, random init, fp32, all host threads): ONE forward of ONE sample -- the CPU comparator of the GPU arm's
    secondary ``unet_forward`` line.  ``omegaconf`` (absent here; only its ListConfig type is compared in the constructor)
    is stubbed.  Returns None when baseline/_ref is absent or not importable."""
    import types
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "ldm")):
        return None
    sys.path.insert(0, ref)
    try:
        if "omegaconf" not in sys.modules:
            lc = types.ModuleType("omegaconf.listconfig")
            lc.ListConfig = type("ListConfig", (), {})
            sys.modules["omegaconf"], sys.modules["omegaconf.listconfig"] = types.ModuleType("omegaconf"), lc
        from ldm.modules.diffusionmodules.openaimodel import UNetModel
    except Exception:
        return None
    finally:
        sys.path.remove(ref)
    torch.manual_seed(0)
    unet = UNetModel(in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2, attention_resolutions=[4, 2, 1],
                     channel_mult=(1, 2, 4, 4), num_heads=8, use_spatial_transformer=True, context_dim=CTX_DIM, legacy=False).eval()
    x, ts, ctx = torch.randn(1, 4, 64, 64), torch.tensor([500]), torch.randn(1, S_CTX, CTX_DIM)
    with torch.no_grad():
        t0 = time.perf_counter()
        unet(x, ts, context=ctx, extra_info={})
        dt = time.perf_counter() - t0
    return {"ms_per_sample": dt * 1e3, "samples_per_s": 1 / dt, "kind": "reference", "cores": torch.get_num_threads(),
            "what": "baseline/_ref UNetModel.forward, verbatim, SD-1.5 config, B=1, 64x64 latents, fp32, one un-warmed forward"}


def main_reference(args):
    """--impl reference: the reference's CPU implementation of the path (the oracle port -- /root/reference does
    not exist on the GPU box and has no compiled code on this path), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    val, dt, fl, _ = run_cpu(torch, args.steps, args.warmup)
    sample = "B=1, one block per level (4 of 16 blocks: self+cross attn at 4096x320, 1024x640, 256x1280, 64x1280), fp32"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "sd15_attn_stack_512", "sample": sample, "flops_per_step": fl},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    # The headline CPU arm is the oracle's restatement of the diffusers processor (F.scaled_dot_product_attention: the
    # FASTER of the reference's two attention paths, so the conservative baseline).  When the pip-installed reference
    # is present its own LDM module is timed verbatim as well and reported next to it.
    ldm = run_reference_ldm(torch, max(1, min(args.steps, 3)), 1)
    if ldm is not None:
        line["reference_verbatim_ldm"] = ldm
    if os.environ.get("ADAFACE_BENCH_EXTRAS", "1") != "0":
        un = run_reference_unet(torch)
        if un is not None:
            line["reference_verbatim_unet"] = un
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v == "Active":
                    reasons.add(n)
        os.unlink(self.f.name)
        if sm:
            hi = sorted(sm)[len(sm) // 2:]          # upper half = samples taken under load
            out = {"sm_mhz": statistics.median(hi), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def bind_host_memory_to_gpu_node(torch, local):
    """Multi-rank runs: every rank's pinned host buffers should live on the NUMA node its GPU hangs off (round 1: eight ranks on node 0
    lost a third of the end-to-end rate to cross-socket copies).  Best effort: set this thread's memory policy to PREFERRED(node of the
    GPU) before the pinned allocations, and move the thread to that node's CPUs when the cpuset allows it.  Returns what was done."""
    import ctypes
    info = {"gpu_node": None, "mempolicy": "unchanged", "cpus": "unchanged"}
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        info["gpu_node"] = node
        if node < 0:
            return info
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(238, 1, ctypes.byref(mask), 65)       # set_mempolicy(MPOL_PREFERRED, {node})  (x86-64 syscall number)
        info["mempolicy"] = f"preferred node {node}" if rc == 0 else f"refused (errno {ctypes.get_errno()})"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = f"{len(allowed)} CPUs of node {node}"
        else:
            info["cpus"] = f"node {node} is outside this process's cpuset"
    except Exception as ex:      # placement is an optimisation, never a failure
        info["error"] = f"{type(ex).__name__}: {ex}"
    return info


def build_stack(torch, a, device):
    g = torch.Generator().manual_seed(0)
    mods, xs = [], []
    for _, N, C, nblk in LEVELS:
        blocks = []
        for _ in range(nblk):
            pair = []
            for ctx_dim in (None, CTX_DIM):
                m = a.Attention(C, ctx_dim, HEADS, C // HEADS, processor=a.AttnProcessor_LoRA_Capture(), device=device)
                with torch.no_grad():
                    for p in m.parameters():
                        if p.dim() > 1:
                            p.copy_((torch.randn(p.shape, generator=g) / p.shape[1] ** 0.5).to(device))
                        else:
                            p.zero_()
                pair.append(m)
            blocks.append(pair)
        mods.append(blocks)
        xs.append(torch.randn(BATCH, N, C, generator=g).to(torch.bfloat16))
    ctx = torch.randn(BATCH, S_CTX, CTX_DIM, generator=g)
    ctx[:, 4:20] = torch.randn(BATCH, 16, CTX_DIM, generator=g) * 0.5      # ada tokens spliced into the prompt
    return mods, xs, ctx.to(torch.bfloat16)


# kernels of libadaface_b200.so per step: 16 blocks x (self: QKV GEMM, attention, out GEMM; cross: q GEMM, kv GEMM,
# attention, out GEMM) = 112 + the tail launch of the 5 level-A self-attention calls; COUNTED LIVE in main_gpu on one eager step.
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the roofline kernel, from the ncu --set full capture
# summarised in profiles/r02_ncu_full_attn_quad.txt (algorithmic: 4 * B * N * C * 2 = 83.9 MB; the 21 MB of output leave L2 after the kernel)
ROOFLINE_TRAFFIC_BYTES = 67349760   # bulk 55.08 MB read + 2.94 MB written, tail 9.32 MB read, inside the two launches (outputs are still in L2)
ROOFLINE_TRAFFIC_SOURCE = "ncu --set full capture, profiles/r02_ncu_full_attn_quad.txt (dram__bytes_read.sum + dram__bytes_write.sum; not measurable inside bench.py)"


def run_stack(mods, xs, ctx):
    outs = []
    for blocks, x in zip(mods, xs):
        y = x
        for attn1, attn2 in blocks:
            y1 = attn1(x)
            y = attn2(y1, encoder_hidden_states=ctx)
        outs.append(y)
    return outs


def check_step_output(torch, mods, xs_cpu, ctx_cpu, outs):
    """Ties the timed step to correctness inside the run: the outputs the timed CUDA graph left behind for levels D and C
    (last block: self-attention then cross-attention module) against the CPU oracle on sample 0 (bf16-rounded weights,
    fp32 arithmetic).  Bar = north_star's max-abs 2e-2 on bf16 block outputs."""
    import oracle
    errs = {}
    for li in (3, 2):
        attn1, attn2 = mods[li][-1]
        r = lambda p: p.detach().float().cpu().bfloat16().float()
        wd = lambda m: {"to_q": r(m.to_q.weight), "to_k": r(m.to_k.weight), "to_v": r(m.to_v.weight), "to_out_w": r(m.to_out[0].weight),
                        "to_out_b": r(m.to_out[0].bias), "cross_attn_scale_factor": torch.tensor(0.8)}
        x, c = xs_cpu[li][:1].float(), ctx_cpu[:1].float()
        with torch.no_grad():
            y1, _ = oracle.processor_forward(wd(attn1), x, None, heads=HEADS)
            y, _ = oracle.processor_forward(wd(attn2), y1.bfloat16().float(), c, heads=HEADS)
        errs[LEVELS[li][0]] = (outs[li][:1].float().cpu() - y).abs().max().item()
    return {"what": "output of the timed graph, levels D and C (last block, sample 0) vs the CPU oracle", "max_abs_err": errs,
            "tol": 2e-2, "ok": all(v < 2e-2 for v in errs.values())}


def _time_us(torch, flush, fn, iters=10, warm=3):
    for _ in range(warm):
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
        ts.append(s.elapsed_time(e) * 1e3)
    return statistics.median(ts)


def secondary_measurements(torch, a, dev, flush, pk, unet):
    """Secondary lines (not the headline): the HBM-bound cross-attention / capture kernels against the measured copy
    bandwidth (SURVEY 8d #1 byte counts), SubjBasisGenerator at BASELINE config 2, and one stage-2-style
    forward + backward through a captured level-A cross-attention module (config 5, scaled to the hot path)."""
    ops = a.ops
    out = {}
    H, S, N, C = HEADS, S_CTX, 4096, 320
    d = C // H
    hbm = pk["hbm_gbs"]
    with torch.no_grad():
        # -- cross-attention core, fast path, level A, B = 8: read Q + K,V, write O (bf16)
        B = BATCH
        # q as the processor now hands it over: plain [B, N, C] rows from the q projection (the head-major padded detour is off by default)
        q = torch.randn(B, N, C, device=dev).to(torch.bfloat16)
        kv = torch.randn(B, S, 2 * C, device=dev).to(torch.bfloat16)
        o_c = torch.empty(B, N, C, device=dev, dtype=torch.bfloat16)
        us = _time_us(torch, flush, lambda: ops.attention(q, kv[:, :, :C], kv[:, :, C:], H, d ** -0.5, out=o_c))
        by = 2 * B * N * C * 2 + 2 * B * S * C * 2
        out["cross_attn_fast"] = {"kernel": "attn_cross_tc_kernel<40, 80> (persistent, K / V resident, Q boxes over whole rows)", "shape": f"B={B} N={N} C={C} S={S}", "us": us,
                                  "algorithmic_bytes": by, "achieved_gbs": by / us / 1e3, "peak_gbs": hbm,
                                  "frac": by / us / 1e3 / hbm}
        # -- img_mask self-attention at level A (dalc:254-273): key mask on the four-tile tcgen05 kernel (was the warp-MMA kernel)
        qkv = torch.randn(B, N, 3 * C, device=dev).to(torch.bfloat16)
        km = (torch.rand(B, N, device=dev) > 0.3).to(torch.uint8)
        o_ = torch.empty(B, N, C, device=dev, dtype=torch.bfloat16)
        fl = 4.0 * B * H * N * N * d
        for nm, mask in (("self_attn_levelA_unmasked", None), ("self_attn_levelA_key_mask", km)):
            us = _time_us(torch, flush, lambda: ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], H, d ** -0.5, key_mask=mask, out=o_))
            out[nm] = {"kernel": "attn_fwd_tcgen05_quad_kernel<40>" + (" + key mask in the spare K column" if mask is not None else ""),
                       "shape": f"B={B} N={N} H={H} d={d}", "us": us, "tflops": fl / us / 1e6, "peak_tflops": pk["bf16_tflops"], "frac": fl / us / 1e6 / pk["bf16_tflops"]}
        del qkv, km, o_
        # -- projection GEMMs as the step graph sees them: 20 back-to-back launches in one CUDA graph over rotating buffers
        #    (inputs recently written by another launch, no host launch cost), us per launch
        gl = {}
        for nm, M_, N_, K_ in (("A_qkv", 32768, 960, 320), ("A_out", 32768, 320, 320), ("B_qkv", 8192, 1920, 640), ("C_qkv", 2048, 3840, 1280)):
            xs_ = [torch.randn(M_, K_, device=dev).to(torch.bfloat16) for _ in range(4)]
            w_ = (torch.randn(N_, K_, device=dev) * K_ ** -0.5).to(torch.bfloat16)
            b_ = torch.zeros(N_, device=dev)
            ys_ = [torch.empty(M_, N_, device=dev, dtype=torch.bfloat16) for _ in range(4)]
            for i in range(3):
                ops.proj(xs_[i], w_, bias=b_, out=ys_[i])
            torch.cuda.synchronize()
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                for i in range(20):
                    ops.proj(xs_[i % 4], w_, bias=b_, out=ys_[i % 4])
            g_.replay()
            torch.cuda.synchronize()
            ts_ = []
            for _ in range(5):
                s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s_.record(); g_.replay(); e_.record(); torch.cuda.synchronize()
                ts_.append(s_.elapsed_time(e_) * 1e3 / 20)
            us = statistics.median(ts_)
            gl[nm] = {"M": M_, "N": N_, "K": K_, "us_per_launch": us, "tflops": 2.0 * M_ * N_ * K_ / us / 1e6, "frac": 2.0 * M_ * N_ * K_ / us / 1e6 / pk["bf16_tflops"]}
            del xs_, ys_, g_
        out["proj_gemm_in_graph"] = {"kernel": "gemm_tn_tcgen05_kernel<BN, LEAN> (bias, bf16 out through the store warp)", "peak_tflops": pk["bf16_tflops"], **gl}
        # -- capture path, config 1 (B = 2, fp32 q/k/v in, bf16 O + fp32 prob [+ score] out)
        B = 2
        qf, kf, vf = (torch.randn(B, n_, C, device=dev) for n_ in (N, S, S))
        core = B * N * C * (4 + 2) + 2 * B * S * C * 4
        maps = B * H * N * S * 4
        for nm, kw, by in (("capture_prob", dict(want_score=False), core + maps), ("capture_prob_score", {}, core + 2 * maps)):
            us = _time_us(torch, flush, lambda: ops.attention_cross_capture(qf, kf, vf, H, d ** -0.5, **kw))
            out[nm] = {"kernel": "attn_cross_stream_kernel<40,fp32 hi/lo>", "shape": f"B={B} N={N} C={C} S={S}", "us": us,
                       "algorithmic_bytes": by, "achieved_gbs": by / us / 1e3, "peak_gbs": hbm, "frac": by / us / 1e3 / hbm}
        # -- the same capture with FUSED CONSUMERS -- the [B,H,N,S] map is reduced in registers, not written.
        #    "subj_sum": mass on the 16 subject columns -> [B,H,N] fp32; "sqdiff": sum (prob - ref)^2 against a resident reference map
        sflag = torch.zeros(B, S, device=dev, dtype=torch.uint8)
        sflag[:, 4:20] = 1
        refp = torch.softmax(torch.randn(B, H, N, S, device=dev), dim=-1)
        for nm, kw, by in (("capture_fused_subj_sum", dict(sum_flag=sflag), core + B * H * N * 4),
                           ("capture_fused_sqdiff", dict(ref_prob=refp), core + maps),
                           ("capture_fused_both", dict(sum_flag=sflag, ref_prob=refp), core + maps + B * H * N * 4)):
            us = _time_us(torch, flush, lambda: ops.attention_cross_consume(qf, kf, vf, H, d ** -0.5, **kw))
            out[nm] = {"kernel": "attn_cross_stream_kernel<40,fp32 hi/lo> + fused consumers", "shape": f"B={B} N={N} C={C} S={S}", "us": us,
                       "algorithmic_bytes": by, "achieved_gbs": by / us / 1e3, "peak_gbs": hbm, "frac": by / us / 1e3 / hbm,
                       "bytes_not_written_vs_capture_prob": maps if "ref_prob" not in kw else 0,
                       "note": "same job as capture_prob + a separate reduction pass over the 20 MB map; compare `us`, the fraction falls "
                               "because the bytes do"}
        del refp
        # -- BASELINE config 2: ArcFace 512-d ID embedding -> 16 ada prompt tokens, batch 64, random init:
        #    Arc2Face ID -> image-prompt encoder (12 CLIP layers) + SubjBasisGenerator (12 layers, K/V multiplier 1)
        gen = a.SubjBasisGenerator().to(dev).eval()
        id2img = a.Arc2FaceID2ImgPrompt().to(dev).eval()
        ids = torch.nn.functional.normalize(torch.randn(64, 512, device=dev), dim=-1)
        x = torch.randn(64, 16, 768, device=dev) * 0.5
        fn = a.graphed(lambda t: gen(t), x)
        us = _time_us(torch, flush, lambda: fn(x))
        fn2 = a.graphed(lambda t: gen(id2img(t)), ids)
        us2 = _time_us(torch, flush, lambda: fn2(ids))
        T = 20
        fl = 64 * 12 * (24 * T * 768 * 768 + 4 * T * T * 768)
        out["subj_basis_generator"] = {"shape": "BS=64, N_ID=16, T_run=20 (causal-exact truncation of 77)", "us": us,
                                       "flops_executed": fl, "tflops": fl / us / 1e6, "samples_per_s": 64 / us * 1e6}
        out["arcface_to_ada_tokens"] = {"shape": "BS=64: [64,512] -> Arc2Face CLIP encoder -> SubjBasisGenerator -> [64,16,768]",
                                        "us": us2, "flops_executed": 2 * fl, "tflops": 2 * fl / us2 / 1e6,
                                        "samples_per_s": 64 / us2 * 1e6}
        # -- the implicit-GEMM 3x3 convolution (tensor-bound: K = 9 Cin) at the U-Net's sizes, B = 8
        tf_peak = pk["bf16_tflops"]
        conv = {}
        for side, cin, cout in ((64, 320, 320), (32, 640, 640), (16, 1280, 1280)):
            xt = torch.randn(BATCH, side * side, cin, device=dev).to(torch.bfloat16)
            wp = ops.pack_conv3x3_weight(torch.randn(cout, cin, 3, 3, device=dev) * (9 * cin) ** -0.5)
            bias = torch.zeros(cout, device=dev)
            yt = torch.empty(BATCH, side * side, cout, device=dev, dtype=torch.bfloat16)
            us = _time_us(torch, flush, lambda: ops.conv3x3(xt, wp, (side, side), bias=bias, out=yt))
            fl = 2.0 * BATCH * side * side * cout * 9 * cin
            conv[f"{side}x{side}_{cin}to{cout}"] = {"us": us, "tflops": fl / us / 1e6, "peak_tflops": tf_peak, "frac": fl / us / 1e6 / tf_peak}
        out["conv3x3_implicit_gemm"] = {"kernel": "gemm_tn_tcgen05_kernel<BN, CONV>", "batch": BATCH, "bound": "tensor", **conv}
        # -- the caller of the whole path: one SD-1.5 U-Net forward (random weights), one CUDA graph per batch size
        un = {"what": "UNetModel.forward (openaimodel.py:820-960), 64x64 latents, 77-token context, bf16 NHWC-resident, CUDA graph",
              "params": sum(p_.numel() for p_ in unet.parameters())}
        for B in (2, BATCH):
            xl, ts_, cx = torch.randn(B, 4, 64, 64, device=dev), torch.randint(0, 1000, (B,), device=dev), torch.randn(B, S, CTX_DIM, device=dev).to(torch.bfloat16)
            n0 = a._lib.launch_count()
            unet(xl, ts_, context=cx)
            nk = a._lib.launch_count() - n0
            gfn = a.graphed(lambda x_, t_, c_: unet(x_, t_, context=c_), xl, ts_, cx)
            us = _time_us(torch, flush, lambda: gfn(xl, ts_, cx), iters=5)
            un[f"B{B}"] = {"ms": us / 1e3, "samples_per_s": B / us * 1e6, "kernels": int(nk)}
            del gfn
        out["unet_forward"] = un
    # -- stage-2 direction through the same mirror: forward + backward w.r.t. the prompt context (frozen U-Net), B = 2, eager
    xl, ts_ = torch.randn(2, 4, 64, 64, device=dev), torch.randint(0, 1000, (2,), device=dev)
    cx = torch.randn(2, 97, CTX_DIM, device=dev, requires_grad=True)
    gy = torch.randn(2, 4, 64, 64, device=dev)

    def unet_step():
        cx.grad = None
        unet(xl, ts_, context=cx).backward(gy)
    n0 = a._lib.launch_count()
    unet_step()
    nk = a._lib.launch_count() - n0
    us = _time_us(torch, flush, unet_step, iters=3, warm=1)
    out["unet_forward"]["train_fwd_bwd_context_B2"] = {"ms": us / 1e3, "kernels": int(nk), "what": "forward + backward w.r.t. a 97-token "
                                                      "context through every block (frozen weights), launches issued from Python"}
    # -- training: forward + backward through one captured cross-attention module (level A, B = 1, S = 97, DoRA r = 192
    #    on q/k/v/out, normalize_cross_attn, loss on out + captured attn) and through one self-attention module
    S2 = 97
    attn = a.Attention(C, CTX_DIM, H, d, device=dev)
    layers = {"q": attn.to_q, "k": attn.to_k, "v": attn.to_v, "out": attn.to_out[0]}
    proc = a.AttnProcessor_LoRA_Capture(capture_ca_activations=True, enable_lora=True, lora_proj_layers=layers, lora_rank=192,
                                        lora_alpha=16).to(dev)
    proc.reset_attn_cache_and_flags(True, True, False, True)
    attn.set_processor(proc)
    hs = torch.randn(1, N, C, device=dev, dtype=torch.bfloat16, requires_grad=True)
    ehs = torch.randn(1, S2, CTX_DIM, device=dev, dtype=torch.bfloat16, requires_grad=True)
    si = (torch.zeros(16, dtype=torch.long, device=dev), torch.arange(4, 20, device=dev))
    gout = torch.randn(1, N, C, device=dev, dtype=torch.bfloat16)
    gp = torch.randn(1, H, N, S2, device=dev)

    def cross_step():
        o = attn(hs, encoder_hidden_states=ehs, subj_indices=si)
        torch.autograd.backward([o, proc.cached_activations["attn"]], [gout, gp])
    n0 = a._lib.launch_count()
    cross_step()
    n_cross = a._lib.launch_count() - n0
    us_c = _time_us(torch, flush, cross_step, iters=5, warm=2)
    sattn = a.Attention(C, None, H, d, device=dev)

    def self_step():
        sattn(hs).backward(gout)
    us_s = _time_us(torch, flush, self_step, iters=5, warm=2)
    out["train_fwd_bwd"] = {"cross_capture_lora_r192_levelA_B1_us": us_c, "cross_kernel_launches": int(n_cross),
                            "self_attn_levelA_B1_us": us_s,
                            "self_attn_tflops": 3.5 * (4.0 * N * N * C) / us_s / 1e6}
    return out


def build_unet(torch, a, dev):
    """SD-1.5 U-Net mirror, random init (no checkpoints offline): weights ~ N(0, 1/fan_in), norms 1 / 0."""
    with torch.device("meta"):
        unet = a.UNetModel(in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2, attention_resolutions=[4, 2, 1],
                           channel_mult=(1, 2, 4, 4), num_heads=8, use_spatial_transformer=True, context_dim=CTX_DIM)
    unet = unet.to_empty(device=dev).eval()
    g = torch.Generator(device=dev).manual_seed(0)
    with torch.no_grad():
        for k_, p_ in unet.named_parameters():
            if p_.dim() >= 2:
                p_.normal_(std=p_[0].numel() ** -0.5, generator=g)
            elif k_.endswith("weight"):
                p_.fill_(1.0)
            else:
                p_.zero_()
    return unet


DDIM_IMAGES, DDIM_STEPS, DDIM_GUIDANCE = 64, 50, 4.0


def ddim_measurement(torch, a, dev, rank, world, barrier, unet=None):
    """BASELINE config 4: 50-step DDIM at 512^2, 64 images x CFG = U-Net batch 128 in total, the images sharded
    across the ranks (STRONG scaling, no collective: every image's (cond, uncond) pair stays on one GPU).  Metric: denoise
    image-steps / s = 50 * 64 / time.  `value`: latents and prompts already resident in HBM; `e2e`: pinned host buffers in
    (x_T fp32, cond / uncond prompt embeddings bf16), host latents out, copies inside the timed region."""
    import torch.distributed as dist
    unet = unet if unet is not None else build_unet(torch, a, dev)
    model = a.UNetDenoiser(unet)
    b, e = a.parallel.shard_range(DDIM_IMAGES, rank, world)
    n_loc = e - b
    mb = int(os.environ.get("ADAFACE_BENCH_DDIM_MB", "16"))
    mb = max(1, min(mb, n_loc))
    g = torch.Generator().manual_seed(1234)
    xT_all = torch.randn(DDIM_IMAGES, 4, 64, 64, generator=g)
    cond_all = torch.randn(DDIM_IMAGES, S_CTX, CTX_DIM, generator=g)
    cond_all[:, 4:20] = torch.randn(DDIM_IMAGES, 16, CTX_DIM, generator=g) * 0.5          # ada tokens spliced into the prompt
    unc_all = torch.randn(1, S_CTX, CTX_DIM, generator=g).expand(DDIM_IMAGES, -1, -1)
    xT_h = xT_all[b:e].contiguous().pin_memory()
    cond_h = cond_all[b:e].to(torch.bfloat16).contiguous().pin_memory()
    unc_h = unc_all[b:e].to(torch.bfloat16).contiguous().pin_memory()
    out_h = torch.empty(n_loc, 4, 64, 64).pin_memory()
    xT_d, cond_d, unc_d = xT_h.to(dev), cond_h.to(dev), unc_h.to(dev)
    smp = a.DDIMSampler(model, micro_batch=mb)
    kw = dict(eta=0., verbose=False, guidance_scale=DDIM_GUIDANCE)
    n0 = a._lib.launch_count()
    smp.sample(2, n_loc, (4, 64, 64), conditioning=cond_d, x_T=xT_d, unconditional_conditioning=unc_d, **kw)     # captures the step graph(s)
    launches_warm = a._lib.launch_count() - n0
    torch.cuda.synchronize()

    def timed(fn):
        barrier()
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = fn()
        t.record()
        barrier()
        return s.elapsed_time(t), r

    t_dev, x0 = timed(lambda: smp.sample(DDIM_STEPS, n_loc, (4, 64, 64), conditioning=cond_d, x_T=xT_d,
                                         unconditional_conditioning=unc_d, **kw)[0])

    def e2e():
        xd, cd, ud = xT_h.to(dev, non_blocking=True), cond_h.to(dev, non_blocking=True), unc_h.to(dev, non_blocking=True)
        x0_ = smp.sample(DDIM_STEPS, n_loc, (4, 64, 64), conditioning=cd, x_T=xd, unconditional_conditioning=ud, **kw)[0]
        out_h.copy_(x0_, non_blocking=True)
        return x0_
    t_e2e, x0e = timed(e2e)
    torch.cuda.synchronize()
    same = bool(torch.equal(x0e, x0)) and bool(torch.equal(out_h, x0.cpu()))
    finite = bool(torch.isfinite(x0).all())
    tt = torch.tensor([t_dev, t_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = tt.tolist()
    work = DDIM_STEPS * DDIM_IMAGES
    n_mb = (n_loc + mb - 1) // mb
    return {"metric": "UNet denoise steps/s @512^2 (50-step DDIM, 64 images x CFG, batch sharded)", "unit": "image-steps/s",
            "value": work / (t_dev * 1e-3), "ms_total": t_dev, "scaling": "strong", "n_gpus": world,
            "e2e": {"value": work / (t_e2e * 1e-3), "unit": "image-steps/s", "ms_total": t_e2e,
                    "h2d_bytes_per_run": int(xT_h.numel() * 4 + cond_h.numel() * 2 + unc_h.numel() * 2), "d2h_bytes_per_run": int(out_h.numel() * 4)},
            "unet_samples_per_s_per_gpu": 2 * n_loc * DDIM_STEPS / (t_dev * 1e-3),
            "unet_tflops_per_gpu": 2 * n_loc * DDIM_STEPS * 8.0327e11 / (t_dev * 1e-3) / 1e12,
            "config": {"global_images": DDIM_IMAGES, "unet_batch_global": 2 * DDIM_IMAGES, "images_per_gpu": n_loc,
                       "micro_batch_images": mb, "unet_batch_per_call": 2 * mb, "ddim_steps": DDIM_STEPS, "guidance_scale": DDIM_GUIDANCE,
                       "eta": 0.0, "timesteps": "1, 21, ..., 981 (ddim.py:29-35)", "latent": "4x64x64", "ctx": "77x768, ada tokens in rows 4:20",
                       "launch": f"one CUDA graph (U-Net + fused CFG/DDIM update) replayed {DDIM_STEPS * n_mb} times per rank; no collective"},
            "graph_replays_per_rank": DDIM_STEPS * n_mb, "kernels_in_two_warmup_steps_eager": int(launches_warm),
            "checks": {"finite": finite, "e2e_equals_device_run_bitwise": same}}


def stage2_measurement(torch, a, dev, rank, world, barrier, unet, steps=4, warm=2):
    """BASELINE config 5, scaled to what the hot path sees: one stage-2 (compositional distillation) iteration =
    4 denoising steps x [ss, sc_rep no-grad; sc WITH grad; mc no-grad -- four sliced B = 1 U-Net calls with capture on layers
    22-24, S = 97, attention LoRA r = 192 (DoRA) on q/k/v/out + conv-LoRA on up_blocks.3.resnets.[12], normalize_cross_attn on
    sc / sc_rep -- plus one no-grad B = 4 unconditional call], loss = subject-mass background suppression + sc-vs-sc_rep
    distillation (the consumers of the captured maps, fused into the capture kernel), backward through the sc instance into
    the adapters and -- through rows 4:20 of the prompt -- the SubjBasisGenerator, then the data-parallel all-reduce of the
    trainable gradients (GradBucketer over NCCL, overlapped with the SubjBasisGenerator's backward).  WEAK scaling: every rank
    runs its own subject (as the reference seeds per rank, ldm/util.py:524-530).  No optimiser step (out of scope)."""
    import torch.distributed as dist
    from adaface_dev_b200.unet_wrapper import DiffusersUNetWrapper
    from adaface_dev_b200.stage2 import CompDistillStep
    S2 = 97
    w = DiffusersUNetWrapper(unet, use_attn_lora=True, use_ffn_lora=True, lora_rank=192)
    gen = torch.Generator().manual_seed(7)
    with torch.no_grad():                              # adapters off their identity init so that every gradient path is live
        for n_, p_ in w.unet_lora_modules.named_parameters():
            if "lora_B" in n_:
                p_.copy_((torch.randn(p_.shape, generator=gen) * 0.02).to(dev))
    sbg = a.SubjBasisGenerator().to(dev)
    sbg.train()
    lora_params = w.trainable_parameters()
    sbg_params = [p_ for p_ in sbg.parameters() if p_.requires_grad]
    gb_lora = a.parallel.GradBucketer(lora_params, expected_uses=steps)
    gb_sbg = a.parallel.GradBucketer(sbg_params, expected_uses=1)
    step = CompDistillStep(w, fused_consumers=os.environ.get("ADAFACE_BENCH_STAGE2_FUSED", "1") != "0", use_ffn_lora=True)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(4, 4, 64, 64, generator=g).to(dev)
    ts = [torch.full((4,), v, dtype=torch.long, device=dev) for v in (800, 600, 400, 200)][:steps]
    base_prompt = torch.randn(4, S2, CTX_DIM, generator=g).to(dev)
    uncond = torch.randn(1, S2, CTX_DIM, generator=g).expand(4, -1, -1).contiguous().to(dev)
    id_embs = (torch.randn(1, 16, CTX_DIM, generator=g) * 0.5).to(dev)
    si = (torch.zeros(16, dtype=torch.long, device=dev), torch.arange(4, 20, device=dev))
    fg = torch.zeros(1, 1, 64, 64, device=dev)
    fg[0, 0, 12:44, 16:40] = 1
    emb_mask = torch.zeros(4, S2, 1, device=dev)
    emb_mask[:, 1:40] = 1
    pad_mask = torch.zeros(4, S2, 1, device=dev)
    pad_mask[:, 40:] = 1

    def fwd_bwd():
        """Forward + backward of the whole iteration (no collective, no host synchronisation: CUDA-graph-capturable)."""
        ada = sbg(id_embs)                                           # [1, 16, 768], differentiable
        ada_leaf = ada.detach().requires_grad_(True)

        def prompt():
            pe = base_prompt.clone()
            pe[0:3, 4:20] = ada_leaf.to(pe.dtype)                    # ss, sc, sc_rep carry the subject tokens; mc is the class prompt
            return pe
        totals = step.step(x, ts, prompt, uncond, si, fg, emb_mask, pad_mask, sc_fg_mask_percent=0.3)
        ada.backward(ada_leaf.grad)                                  # SubjBasisGenerator backward: its buckets go out as they fill
        return totals

    replay = None

    def iteration(comm=True):
        gb_lora.zero()
        gb_sbg.zero()
        if not comm:
            gb_lora.world = gb_sbg.world = 1
        totals = replay() if replay is not None else fwd_bwd()
        gb_lora.finish()
        gb_sbg.finish()
        gb_lora.world = gb_sbg.world = world
        return totals

    def timed(fn, n):
        barrier()
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        for _ in range(n):
            r_ = fn()
        e_.record()
        barrier()
        return s_.elapsed_time(e_) / n, r_

    for _ in range(warm):
        iteration()
    n0 = a._lib.launch_count()
    iteration()
    launches = a._lib.launch_count() - n0
    n_it = 3
    t_full, totals = timed(iteration, n_it)
    t_nocomm, _ = timed(lambda: iteration(comm=False), n_it) if world > 1 else (t_full, None)

    def allreduce_only():
        for gb in (gb_lora, gb_sbg):
            for b_ in gb.buckets:
                dist.all_reduce(b_["flat"])
    t_ar = timed(allreduce_only, 3)[0] if world > 1 else 0.0
    n_par = sum(p_.numel() for p_ in lora_params + sbg_params)
    # ---- the same iteration as ONE CUDA graph (forward + backward of all four denoising steps and the SubjBasisGenerator);
    #      the all-reduce runs after the replay.  Gradients are compared with the eager iteration's before timing.
    t_graph = t_graph_nocomm = None
    graph_check = None
    if os.environ.get("ADAFACE_BENCH_STAGE2_GRAPH", "1") != "0":
        iteration(comm=False)
        ref_grads = [p_.grad.clone() for p_ in (lora_params + sbg_params) if p_.grad is not None][:40]
        gb_lora.defer = gb_sbg.defer = True
        gb_lora.zero()
        gb_sbg.zero()
        replay = a.graphed_step(fwd_bwd, w, sbg)
        gb_lora.freeze_touched()
        gb_sbg.freeze_touched()
        iteration(comm=False)
        got = [p_.grad for p_ in (lora_params + sbg_params) if p_.grad is not None][:40]
        graph_check = max(((g_ - r_).abs().max() / r_.abs().max().clamp_min(1e-20)).item() for g_, r_ in zip(got, ref_grads))
        t_graph, totals = timed(iteration, n_it)
        t_graph_nocomm, _ = timed(lambda: iteration(comm=False), n_it) if world > 1 else (t_graph, None)
    tt = torch.tensor([t_full, t_nocomm, t_ar, t_graph or 0.0, t_graph_nocomm or 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_full, t_nocomm, t_ar, tg, tgn = tt.tolist()
    if t_graph is not None:
        t_graph, t_graph_nocomm = tg, tgn
    finite = all(bool(torch.isfinite(torch.as_tensor(v)).all()) for v in totals.values())
    gb_lora.close()
    gb_sbg.close()
    best = t_graph if t_graph is not None else t_full
    best_nocomm = t_graph_nocomm if t_graph is not None else t_nocomm
    return {"metric": "stage-2 compositional-distillation iterations/s (4 denoising steps, capture + backward, DDP all-reduce)",
            "unit": "iterations/s", "value": world * 1e3 / best, "ms_per_iteration": best, "scaling": "weak", "n_gpus": world,
            "ms_per_iteration_without_collectives": best_nocomm, "allreduce_alone_ms": t_ar,
            "allreduce_exposed_share": max(0.0, (best - best_nocomm) / best) if world > 1 else 0.0,
            "eager": {"ms_per_iteration": t_full, "ms_per_iteration_without_collectives": t_nocomm,
                      "what": "the same iteration with every kernel launched from Python; its all-reduce overlaps the backward (GradBucketer hooks)"},
            "cuda_graph": None if t_graph is None else {"ms_per_iteration": t_graph, "max_rel_grad_diff_vs_eager": graph_check,
                                                        "what": "forward + backward of the whole iteration replayed as one CUDA graph; all-reduce after the replay"},
            "allreduce_bytes": int(n_par * 4), "trainable_params": {"lora": int(sum(p_.numel() for p_ in lora_params)),
                                                                      "subj_basis_generator": int(sum(p_.numel() for p_ in sbg_params))},
            "buckets": len(gb_lora.buckets) + len(gb_sbg.buckets), "kernels_per_iteration": int(launches),
            "config": {"denoising_steps": steps, "instances": "ss, sc, sc_rep, mc (B=1 slices) + uncond B=4", "ctx_tokens": S2, "lora_rank": 192,
                       "captured_layers": [22, 23, 24], "fused_capture_consumers": step.fused,
                       "launch": "one CUDA graph per iteration" if t_graph is not None else "eager (Python launches)"},
            "losses": {k_: float(v_) for k_, v_ in totals.items()}, "checks": {"finite": finite}}


def main_gpu(args):
    import torch
    import torch.distributed as dist
    import adaface_dev_b200 as a
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    a._lib.load()
    # only multi-rank runs are placed (at N = 1 the CPU baseline of this same process keeps every core)
    host_numa = bind_host_memory_to_gpu_node(torch, local) if world > 1 and os.environ.get("ADAFACE_BENCH_NUMA", "1") != "0" else None
    sampler = ClockSampler(local)      # nvidia-smi needs ~1 s to start: launch it before the set-up work

    mods, xs_cpu, ctx_cpu = build_stack(torch, a, dev)
    xs = [x.to(dev) for x in xs_cpu]
    ctx = ctx_cpu.to(dev)
    xs_pin = [x.pin_memory() for x in xs_cpu]
    ctx_pin = ctx_cpu.pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    use_graph = os.environ.get("ADAFACE_BENCH_GRAPH", "1") != "0"
    with torch.no_grad():
        for _ in range(max(3, args.warmup)):
            run_stack(mods, xs, ctx)
        n0 = a._lib.launch_count()
        run_stack(mods, xs, ctx)
        launches_per_step = a._lib.launch_count() - n0          # counted live on one eager step (the graph replays the same kernels)
        if use_graph:
            # one U-Net step of the attention stack = 112 short launches: replay them as one CUDA graph
            step_fn = a.graphed(lambda *t: run_stack(mods, list(t[:-1]), t[-1]), *xs, ctx)
            step = lambda xs_, ctx_: step_fn(*xs_, ctx_)
            for _ in range(max(3, args.warmup)):
                step(xs, ctx)
        else:
            step = lambda xs_, ctx_: run_stack(mods, xs_, ctx_)
        barrier()
        n0 = a._lib.launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for s, e in ev:
            flush.fill_(1)                 # evict L2 between timed iterations (outside the timed interval)
            s.record()
            step(xs, ctx)
            e.record()
        barrier()
        launches = (a._lib.launch_count() - n0) if not use_graph else launches_per_step * args.steps
        t_ms = sum(s.elapsed_time(e) for s, e in ev)
        check = check_step_output(torch, mods, xs_cpu, ctx_cpu, step(xs, ctx)) if rank == 0 else None

        # ---- end to end: host buffers in, host results out, through the same public operator.
        # The step is cut into independent units (levels D, C, B, A -- the four levels of the stack do not depend on each
        # other), each its own CUDA graph.  Three streams: H2D of unit i+1 and D2H of unit i-1 overlap the kernels of unit i; the small levels go
        # first so that level A's 21 MB input is in flight behind them.  Exposed: the first unit's H2D and the last D2H.
        # Level A's 21 MB output would otherwise only start its D2H after the whole level: its LAST block is run per batch
        # slice (samples are independent), so each slice's D2H overlaps the next slice's kernels.
        tail_split = int(os.environ.get("ADAFACE_BENCH_A_TAIL_SPLIT", "4"))     # measured: off 3.51 ms, 2: 3.43 ms, 4: 3.41 ms
        specs = [dict(blocks=mods[li], li=li, sl=slice(0, BATCH), up=None, out=True) for li in (3, 2, 1)]
        if tail_split > 1 and len(mods[0]) > 1:
            specs.append(dict(blocks=mods[0][:-1], li=0, sl=slice(0, BATCH), up=None, out=False))
            sb = BATCH // tail_split
            specs += [dict(blocks=mods[0][-1:], li=0, sl=slice(i * sb, (i + 1) * sb), up=3, out=True) for i in range(tail_split)]
        else:
            specs.append(dict(blocks=mods[0], li=0, sl=slice(0, BATCH), up=None, out=True))
        prev_pdl = a._lib.set_pdl(int(os.environ.get("ADAFACE_BENCH_E2E_PDL", "3")))   # programmatic dependent launch mask inside the unit graphs
        units = []
        for sp in specs:
            u = dict(sp)
            if sp["up"] is None:
                x_u, c_u = xs[sp["li"]][sp["sl"]].contiguous(), ctx[sp["sl"]].contiguous()
                u["host_in"] = (xs_pin[sp["li"]][sp["sl"]], ctx_pin[sp["sl"]])
            else:                                   # chained unit: its input is (a batch slice of) an earlier unit's output
                up = units[sp["up"]]
                x_u, c_u = up["dev_out"][sp["sl"]].contiguous(), up["dev_in"][1][sp["sl"]].contiguous()
                u["host_in"] = None
            body = (lambda x_, c_, blocks=sp["blocks"]: run_stack([blocks], [x_], c_)[0])
            u["fn"] = a.graphed(body, x_u, c_u) if use_graph else body
            u["dev_in"] = tuple(u["fn"].static_inputs[:2]) if use_graph else (torch.empty_like(x_u), torch.empty_like(c_u))
            u["dev_out"] = u["fn"](*u["dev_in"]) if use_graph else torch.empty_like(x_u)      # graphs return their static output
            u["host_out"] = torch.empty(x_u.shape, dtype=torch.bfloat16).pin_memory() if sp["out"] else None
            units.append(u)
        a._lib.set_pdl(prev_pdl)
        h2d = sum(x.numel() * 2 + c.numel() * 2 for x, c in (u["host_in"] for u in units if u["host_in"] is not None))
        d2h = sum(u["host_out"].numel() * 2 for u in units if u["host_out"] is not None)
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        comp = torch.cuda.current_stream()
        torch.cuda.synchronize()

        diag_nocopy = os.environ.get("ADAFACE_BENCH_E2E_NOCOPY") == "1"     # diagnosis only: unit graphs without the copies

        def e2e_step():
            s_in.wait_stream(comp)          # the previous step has finished reading the input buffers
            ev_in = {}
            if not diag_nocopy:
                with torch.cuda.stream(s_in):
                    for ui, u in enumerate(units):
                        if u["host_in"] is not None:
                            u["dev_in"][0].copy_(u["host_in"][0], non_blocking=True)
                            u["dev_in"][1].copy_(u["host_in"][1], non_blocking=True)
                            ev_in[ui] = torch.cuda.Event()
                            ev_in[ui].record(s_in)
            done = []
            for ui, u in enumerate(units):
                if ui in ev_in:
                    comp.wait_event(ev_in[ui])
                if u["up"] is None:
                    out = u["fn"](*u["dev_in"])          # graph replay (inputs already in the graph's buffers) or eager launches
                else:                                    # device-to-device hand-over of the upstream unit's output slice
                    up = units[u["up"]]
                    out = u["fn"](up["cur_out"][u["sl"]], up["dev_in"][1][u["sl"]])
                u["cur_out"] = out
                if u["host_out"] is not None and not diag_nocopy:
                    ev = torch.cuda.Event()
                    ev.record(comp)
                    done.append((u, out, ev))
            with torch.cuda.stream(s_out):
                for u, out, ev in done:
                    s_out.wait_event(ev)
                    u["host_out"].copy_(out, non_blocking=True)
            comp.wait_stream(s_out)         # the closing event on `comp` then covers the last D2H

        e2e_step()
        barrier()
        e2e_steps = max(3, args.steps // 2)
        ee = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(e2e_steps)]
        for s, e in ee:
            flush.fill_(1)
            s.record()
            e2e_step()
            e.record()
        barrier()
        t_e2e_ms = sum(s.elapsed_time(e) for s, e in ee) / e2e_steps
        e2e_how = f"{len(units)} CUDA-graph units (levels D, C, B, A; the last level-A block per batch slice so that its D2H overlaps); H2D / kernels / D2H on three streams"

        # ---- the same step as ONE CUDA graph: the pinned-memory copies become memcpy nodes on two side branches of the graph (forked from and
        #      joined to the kernel branch by events), so the host launches one graph per step instead of eight graphs plus the copies and
        #      events between them.  Same units, same order, same buffers; kept only if it captures, reproduces the multi-graph output and is not slower.
        if use_graph and os.environ.get("ADAFACE_BENCH_E2E_ONEGRAPH", "1") != "0":
            prev = a._lib.set_pdl(int(os.environ.get("ADAFACE_BENCH_E2E_PDL", "3")))
            try:
                bodies = [(lambda x_, c_, blocks=sp["blocks"]: run_stack([blocks], [x_], c_)[0]) for sp in specs]
                head_split = int(os.environ.get("ADAFACE_BENCH_A_HEAD_SPLIT", "1"))     # measured: 1 (off) 3.13 ms, 2: 3.21, 4: 3.27, 8: 3.69
                head_ui = -1
                if head_split > 1 and BATCH % head_split == 0:
                    cand = [ui for ui, u in enumerate(units) if u["li"] == 0 and u["up"] is None and u["host_out"] is None and len(u["blocks"]) > 1]
                    head_ui = cand[0] if cand else -1
                hb = BATCH // head_split if head_ui >= 0 else BATCH
                tail_full_self = os.environ.get("ADAFACE_BENCH_TAIL_FULL_SELF", "0") != "0"      # measured: on 3.30 ms, off 3.12 ms (the slices' D2H then has nothing to hide behind)

                def one_graph_body():
                    cs = torch.cuda.current_stream()
                    s_in.wait_stream(cs)
                    ev_in = {}
                    ev_head = {}
                    with torch.cuda.stream(s_in):
                        for ui, u in enumerate(units):
                            if u["host_in"] is not None:
                                u["dev_in"][1].copy_(u["host_in"][1], non_blocking=True)
                                if ui == head_ui:      # level A's 21 MB input arrives per batch slice: its first block starts on slice 0
                                    for hi in range(head_split):
                                        sl = slice(hi * hb, (hi + 1) * hb)
                                        u["dev_in"][0][sl].copy_(u["host_in"][0][sl], non_blocking=True)
                                        ev_head[hi] = torch.cuda.Event()
                                        ev_head[hi].record(s_in)
                                else:
                                    u["dev_in"][0].copy_(u["host_in"][0], non_blocking=True)
                                ev_in[ui] = torch.cuda.Event()
                                ev_in[ui].record(s_in)
                    outs = {}
                    tail_y1 = {}
                    forked_out = False
                    for ui, u in enumerate(units):
                        if ui == head_ui:
                            x_full, c_full = u["dev_in"]
                            for hi in range(head_split):
                                sl = slice(hi * hb, (hi + 1) * hb)
                                cs.wait_event(ev_head[hi])
                                run_stack([u["blocks"][:1]], [x_full[sl]], c_full[sl])
                            out = run_stack([u["blocks"][1:]], [x_full], c_full)[0]
                            outs[ui] = out
                            continue
                        if ui in ev_in:
                            cs.wait_event(ev_in[ui])
                        if u["up"] is None:
                            out = bodies[ui](*u["dev_in"])
                        else:
                            # tail units (the level's last block, one per batch slice): the self-attention module runs ONCE on the whole
                            # batch (the four-tile kernel wants >= 3 waves of tiles), only the cheap cross-attention module runs per slice
                            up = units[u["up"]]
                            if tail_full_self and len(u["blocks"]) == 1:
                                if u["up"] not in tail_y1:
                                    tail_y1[u["up"]] = u["blocks"][0][0](outs[u["up"]])
                                out = u["blocks"][0][1](tail_y1[u["up"]][u["sl"]], encoder_hidden_states=up["dev_in"][1][u["sl"]])
                            else:
                                out = bodies[ui](outs[u["up"]][u["sl"]], up["dev_in"][1][u["sl"]])
                        outs[ui] = out
                        if u["host_out"] is not None:
                            ev = torch.cuda.Event()
                            ev.record(cs)
                            with torch.cuda.stream(s_out):
                                s_out.wait_event(ev)
                                u["host_out"].copy_(out, non_blocking=True)
                            forked_out = True
                    cs.wait_stream(s_in)
                    if forked_out:
                        cs.wait_stream(s_out)
                    return outs

                torch.cuda.synchronize()
                g1 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1):
                    one_graph_body()
                for _ in range(3):
                    g1.replay()
                barrier()
                ee = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(e2e_steps)]
                for s, e in ee:
                    flush.fill_(1)
                    s.record()
                    g1.replay()
                    e.record()
                barrier()
                t_one = sum(s.elapsed_time(e) for s, e in ee) / e2e_steps
                # the graph's results must be what the unit path produced (same inputs): compare the last unit's host buffer
                chk_unit = [u for u in units if u["host_out"] is not None][-1]
                ref_host = chk_unit["host_out"].clone()
                e2e_step()
                torch.cuda.synchronize()
                # (bf16 outputs of O(1) values; the whole-batch self-attention may sum a row's keys in another order than the per-slice launch)
                same = bool((ref_host.float() - chk_unit["host_out"].float()).abs().max().item() <= 2e-2)
                if same and t_one < t_e2e_ms:
                    e2e_how = (f"ONE CUDA graph per step: {len(units)} kernel units on the main branch, the pinned-memory H2D / D2H copies as memcpy nodes on two side "
                               f"branches" + (f"; level A's input arrives in {head_split} batch slices and its first block runs per slice" if head_ui >= 0 else "") +
                               f" (the multi-graph form of the same step: {t_e2e_ms:.3f} ms)")
                    t_e2e_ms = t_one
                else:
                    e2e_how += f"; the one-graph form measured {t_one:.3f} ms (same output: {same})"
            except Exception as ex:
                e2e_how += f"; one-graph capture failed ({type(ex).__name__}: {ex})"
            finally:
                a._lib.set_pdl(prev)

        # ---- what the box's host links give all ranks AT ONCE (plain pinned-memory copies, both directions concurrently): the ceiling of
        #      the end-to-end number at this N.  128 MB each way per rank, three rounds, barrier in front.
        link = None
        try:
            hb = torch.empty(128 << 20, dtype=torch.uint8).pin_memory()
            hb2 = torch.empty(128 << 20, dtype=torch.uint8).pin_memory()
            db, db2 = torch.empty(128 << 20, dtype=torch.uint8, device=dev), torch.empty(128 << 20, dtype=torch.uint8, device=dev)
            lt = []
            for _ in range(4):
                barrier()
                s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record()
                with torch.cuda.stream(s_in):
                    s_in.wait_event(s0)
                    db.copy_(hb, non_blocking=True)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(s0)
                    hb2.copy_(db2, non_blocking=True)
                comp.wait_stream(s_in)
                comp.wait_stream(s_out)
                e0.record()
                torch.cuda.synchronize()
                lt.append(s0.elapsed_time(e0))
            lms = torch.tensor([min(lt[1:])], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(lms, op=dist.ReduceOp.MAX)
            link = {"what": "128 MB H2D + 128 MB D2H per rank, all ranks at once, pinned memory", "ms": lms.item(),
                    "aggregate_gbs_both_directions": world * 2 * (128 << 20) / (lms.item() * 1e-3) / 1e9,
                    "step_needs_gb": world * (h2d + d2h) / 1e9}
            link["copy_floor_ms_per_step"] = link["step_needs_gb"] / link["aggregate_gbs_both_directions"] * 1e3
            del hb, hb2, db, db2
        except Exception as ex:
            link = {"error": f"{type(ex).__name__}: {ex}"}

        # ---- roofline of the dominant kernel: level-A self-attention core, in the layout the processor feeds it
        #      (q/k/v = column slices of the fused [B, N, 3C] projection buffer), timed alone with L2 flushed
        _, N, C, _ = LEVELS[0]
        dh = C // HEADS
        qkv = torch.randn(BATCH, N, 3 * C, device=dev).to(torch.bfloat16)
        kq, kk, kv_ = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
        ko = torch.empty(BATCH, N, C, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            a.ops.attention(kq, kk, kv_, HEADS, dh ** -0.5, out=ko)
        torch.cuda.synchronize()
        kt = []
        for _ in range(10):
            flush.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            a.ops.attention(kq, kk, kv_, HEADS, dh ** -0.5, out=ko)
            e.record()
            torch.cuda.synchronize()
            kt.append(s.elapsed_time(e))
        k_ms = statistics.mean(kt)
        k_flops = 4.0 * BATCH * N * N * C
        clocks = sampler.stop()
    extra = ddim = unet = None
    want_ddim = os.environ.get("ADAFACE_BENCH_DDIM", "1") != "0"
    want_extra = world == 1 and os.environ.get("ADAFACE_BENCH_EXTRAS", "1") != "0"
    if want_ddim or want_extra:
        unet = build_unet(torch, a, dev)
    if want_ddim:
        # BASELINE config 4 on every N: strong scaling of the 50-step DDIM run (global batch 128 sharded, no collective)
        try:
            ddim = ddim_measurement(torch, a, dev, rank, world, barrier, unet)
        except Exception as ex:
            if world > 1:
                raise                  # a rank that drops out would dead-lock the others in the next collective
            ddim = {"error": f"{type(ex).__name__}: {ex}"}
    if want_extra:
        try:
            extra = secondary_measurements(torch, a, dev, flush, peaks(), unet)
        except Exception as ex:      # secondary lines never take the headline down with them
            extra = {"error": f"{type(ex).__name__}: {ex}"}
    stage2 = None
    if os.environ.get("ADAFACE_BENCH_STAGE2", "1") != "0":
        if unet is None:
            unet = build_unet(torch, a, dev)
        try:                           # last: the wrapper installs processors / adapters on the U-Net and freezes it
            stage2 = stage2_measurement(torch, a, dev, rank, world, barrier, unet)
        except Exception as ex:
            if world > 1:
                raise
            stage2 = {"error": f"{type(ex).__name__}: {ex}"}
    del unet

    tt = torch.tensor([t_ms, t_e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms, t_e2e_ms = tt.tolist()
    fl_step = flops_per_sample() * BATCH
    pk = peaks()
    if rank == 0:
        ms_per_step = t_ms / args.steps
        value = world * fl_step / (ms_per_step * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "sd15_attn_stack_512", "batch_per_gpu": BATCH, "global_batch": BATCH * world,
                       "modules": 32, "ctx_tokens": S_CTX, "heads": HEADS, "flops_per_step_per_gpu": fl_step,
                       "parallelism": f"dp{world} (batch sharded, no collective)",
                       "l2": "256 MB flush written between timed iterations; per-step working set > 1 GB",
                       "launch": f"CUDA graph replay of the {launches_per_step}-kernel step" if use_graph else "eager Python launches",
                       "chaining": "within a level every block reads the level's input x (blocks are independent: same FLOPs, same shapes "
                                   "as the U-Net's, no data dependence between blocks of a level); attn2 consumes attn1's output"},
            "frac_of_bf16_peak": value / world / pk["bf16_tflops_sustained"],
            "e2e": {"value": world * fl_step / (t_e2e_ms * 1e-3) / 1e12, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": t_e2e_ms,
                    "host_link": link,
                    "how": e2e_how},
            "gpu_launches": int(launches), "gpu_launches_per_step": int(launches_per_step),
            "check": check,
            "clocks": clocks,
            "host_numa": host_numa,
            "roofline": {"kernel": "attn_fwd_tcgen05_quad_kernel<40> (level-A self-attention core, B=8, 4096 tok, 8x40; bulk + tail launch)",
                         "bound": "tensor", "achieved": k_flops / (k_ms * 1e-3) / 1e12, "peak": pk["bf16_tflops"],
                         "unit": "TFLOP/s", "frac": k_flops / (k_ms * 1e-3) / 1e12 / pk["bf16_tflops"],
                         "traffic": ROOFLINE_TRAFFIC_BYTES, "traffic_source": ROOFLINE_TRAFFIC_SOURCE, "peak_source": pk["source"] + " (burst: kernel timed alone)",
                         "ms_per_launch": k_ms, "flops_per_launch": k_flops},
        }
        if ddim is not None:
            line["unet_steps_per_s"] = ddim
        if stage2 is not None:
            line["stage2_step"] = stage2
        if extra is not None:
            line["secondary"] = extra
        if world == 1:
            val, dt, _, n = run_cpu(torch, 3, 1, min_seconds=12.0)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"B=1, one block per level (4 of 16 blocks), fp32 oracle, {n} runs in ~12 s"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_gpu(args)
