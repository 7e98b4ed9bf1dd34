"""Multi-GPU plumbing of the hot path: one process per GPU.

Inference shards the denoising batch -- every image x CFG branch is independent through the whole U-Net -- so there
is NO data-path collective: each rank takes a contiguous range of images and keeps each image's (cond, uncond) pair
local so that the CFG combine (ldm/models/diffusion/ddim.py:253-255) stays on the rank.  Training is data parallel
as in the reference (Lightning DDP, main.py:618): one bucketed all-reduce of the trainable gradients
(SubjBasisGenerator + attention LoRA) per optimizer step, NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) range of `n_items` for `rank` (first n_items % world ranks get one extra)."""
    if not (0 <= rank < world) or n_items < 0:
        raise ValueError(f"bad shard request n_items={n_items} rank={rank} world={world}")
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def cfg_batch_indices(n_images: int, rank: int, world: int) -> torch.Tensor:
    """Rows of the CFG-doubled batch this rank owns.  The LDM sampler stacks the batch as [cond_0..cond_{B-1},
    uncond_0..uncond_{B-1}] (ddim.py:239-248): image i lives at rows i and B+i, and both go to the same rank."""
    b, e = shard_range(n_images, rank, world)
    idx = torch.arange(b, e)
    return torch.cat([idx, idx + n_images])


def shard_cfg_batch(tensors: Sequence[torch.Tensor], n_images: int, rank: int, world: int) -> List[torch.Tensor]:
    """Slice [2*n_images, ...] tensors (latents, contexts, timesteps) down to this rank's images, pairs kept together;
    the local result is again ordered [cond.., uncond..] so subj_indices need no offset."""
    idx = cfg_batch_indices(n_images, rank, world)
    out = []
    for t in tensors:
        if t.shape[0] != 2 * n_images:
            raise ValueError(f"expected leading dim {2 * n_images}, got {tuple(t.shape)}")
        out.append(t.index_select(0, idx.to(t.device)))
    return out


def cfg_combine(eps: torch.Tensor, scale: float) -> torch.Tensor:
    """e_t = e_uncond + scale * (e_cond - e_uncond) on a local [cond.., uncond..] batch (ddim.py:253-255)."""
    e_c, e_u = eps.chunk(2, dim=0)
    return e_u + scale * (e_c - e_u)


def gather_images(local: torch.Tensor, n_images: int, group=None) -> torch.Tensor:
    """Optional final gather of per-rank results [n_local, ...] into [n_images, ...] on every rank (the only
    communication of a sampling run: 32 KB per image for 4x64x64 latents)."""
    world = dist.get_world_size(group)
    sizes = [shard_range(n_images, r, world) for r in range(world)]
    n_max = max(e - b for b, e in sizes)                       # ragged shards: pad to the largest, trim after
    padded = local.new_zeros((n_max,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:e - b] for p, (b, e) in zip(parts, sizes)], dim=0)


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None, bucket_bytes: int = 64 << 20,
                        average: bool = True) -> int:
    """Bucketed gradient all-reduce of the trainable parameters (what DDP does at the accumulation boundary,
    main.py:618 / yaml:149).  NVSwitch makes the cost launch-latency bound, hence few large flat buckets.

    Every rank must issue the same collectives, so a has-gradient flag per parameter is all-reduced (MAX) first: a parameter
    whose gradient is None on EVERY rank (e.g. LoRA modules switched off for this iteration by reset_attn_cache_and_flags
    while requires_grad stays True) is left out and keeps ``.grad = None`` -- exactly what the reference's DDP step leaves
    behind, so Adam / AdamW neither update its moments nor decay it.  A parameter with a gradient on at least one rank
    contributes zeros from the ranks that have none.  Returns the number of data buckets."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    world = dist.get_world_size(group)
    dev = params[0].device
    has = torch.tensor([0 if p.grad is None else 1 for p in params], device=dev, dtype=torch.int32)
    dist.all_reduce(has, op=dist.ReduceOp.MAX, group=group)
    params = [p for p, h in zip(params, has.tolist()) if h]
    n_buckets, i = 0, 0
    while i < len(params):
        bucket, size = [], 0
        while i < len(params) and (not bucket or size + params[i].numel() * 4 <= bucket_bytes):
            bucket.append(params[i])
            size += params[i].numel() * 4
            i += 1
        flat = torch.empty(size // 4, device=dev, dtype=torch.float32)
        off = 0
        for p in bucket:
            dst = flat[off:off + p.numel()]
            if p.grad is None:
                dst.zero_()
            else:
                dst.copy_(p.grad.reshape(-1))
            off += p.numel()
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= world
        off = 0
        for p in bucket:
            g = flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                p.grad = g.to(p.dtype, copy=True)
            else:
                p.grad.copy_(g)
            off += p.numel()
        n_buckets += 1
    return n_buckets


class GradBucketer:
    """Overlapped data-parallel gradient all-reduce for the stage-2 trainable set (LoRA / DoRA adapters + SubjBasisGenerator,
    ~1e8 fp32 values; the reference relies on Lightning DDP for this, main.py:618).

    Parameters are packed, in the order their gradients become ready (reverse registration order), into flat fp32 buckets and
    every ``p.grad`` is made a VIEW of its bucket, so autograd accumulates straight into the communication buffer (no copy in,
    no copy out).  A post-accumulate hook counts ready parameters; when a bucket is complete its all-reduce is launched at once
    on a side stream (NCCL over NVLink / NVSwitch), overlapping the rest of the backward pass.  ``finish()`` launches whatever
    is left, waits, averages, and -- DDP semantics -- leaves ``.grad = None`` on parameters that received no gradient on ANY
    rank this iteration so that Adam / AdamW skip them.  ``zero()`` re-arms the buckets for the next iteration."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None, bucket_bytes: int = 64 << 20, average: bool = True,
                 expected_uses: int = 1):
        self.group, self.average, self.expected = group, average, max(1, int(expected_uses))
        self.params = [p for p in params if p.requires_grad]
        if any(p.dtype != torch.float32 for p in self.params):
            raise ValueError("GradBucketer: trainable parameters are kept in fp32 (ddpm.py:4175-4177)")
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buckets, self._where = [], {}
        cur, size = [], 0
        for p in reversed(self.params):
            if cur and size + p.numel() * 4 > bucket_bytes:
                self._add_bucket(cur)
                cur, size = [], 0
            cur.append(p)
            size += p.numel() * 4
        if cur:
            self._add_bucket(cur)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._touched = torch.zeros(len(self.params), dtype=torch.int32)
        self._index = {id(p): i for i, p in enumerate(self.params)}
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.defer = False                       # True: hooks only record which parameters received a gradient (graph capture)
        self._frozen_touched = None
        self.zero()

    def _add_bucket(self, plist):
        n = sum(p.numel() for p in plist)
        flat = torch.zeros(n, device=plist[0].device, dtype=torch.float32)
        b = {"flat": flat, "params": plist, "views": [], "pending": 0, "work": None, "launched": False}
        off = 0
        for p in plist:
            b["views"].append(flat[off:off + p.numel()].view_as(p))
            self._where[id(p)] = b
            off += p.numel()
        self.buckets.append(b)

    def freeze_touched(self):
        """After capturing a training step into a CUDA graph (hooks do not run on replay): remember which parameters the captured
        step gives a gradient to; zero() then restores that set every iteration."""
        self._frozen_touched = self._touched.clone()

    def zero(self):
        """Zero the buckets and point every .grad back at its view (call instead of optimizer.zero_grad())."""
        self._touched.zero_()
        if self._frozen_touched is not None:
            self._touched.copy_(self._frozen_touched)
        self._count = {}
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"], b["work"], b["launched"] = len(b["params"]), None, False
            for p, v in zip(b["params"], b["views"]):
                p.grad = v

    def _on_grad(self, p):
        self._touched[self._index[id(p)]] = 1
        if self.defer:                           # CUDA-graph capture / replay mode: nothing is launched from the hooks
            return
        c = self._count.get(id(p), 0) + 1
        self._count[id(p)] = c
        if c != self.expected:                   # a parameter used several times per iteration (e.g. 4 denoising steps) is ready
            return                               # only after its last accumulation
        b = self._where[id(p)]
        b["pending"] -= 1
        if b["pending"] == 0 and not b["launched"]:
            self._launch(b)

    def _launch(self, b):
        b["launched"] = True
        if self.world == 1:
            return
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group)

    def finish(self) -> int:
        """Complete the iteration's reduction.  Returns the number of buckets."""
        for b in self.buckets:                   # buckets holding parameters without a gradient on this rank: same collectives everywhere
            if not b["launched"]:
                self._launch(b)
        for b in self.buckets:
            if b["work"] is not None:
                b["work"].wait()
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        if self.world > 1:
            if self.average:
                for b in self.buckets:
                    b["flat"] /= self.world
            dev = self.params[0].device
            t = self._touched.to(dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            touched = t.cpu()
        else:
            touched = self._touched
        for p, used in zip(self.params, touched.tolist()):
            if not used:
                p.grad = None                    # no rank produced a gradient: the optimiser must skip it (DDP semantics)
        return len(self.buckets)

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
