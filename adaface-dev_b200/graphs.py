"""CUDA-graph capture of a hot-path callable.

The SD-1.5 attention stack is 112 kernel launches per U-Net step, most of them a few microseconds long: driven
from Python the step is launch-bound.  ``graphed(fn, *example_inputs)`` warms the callable up (so that every kernel
variant has been configured and every weight pack / workspace exists), captures one invocation into a CUDA graph
and returns a function that copies new inputs into the captured input buffers and replays the graph.  Tensor maps
and all other kernel arguments are baked into the graph; outputs live in the graph's private pool and are
overwritten by the next replay (clone them if they must survive).
"""
import torch


def graphed(fn, *example_inputs, warmup=3):
    static_in = [x.clone() if torch.is_tensor(x) else x for x in example_inputs]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        for _ in range(warmup):
            fn(*static_in)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.graph(g):
        static_out = fn(*static_in)

    def replay(*inputs):
        for dst, src in zip(static_in, inputs):
            if torch.is_tensor(dst) and src is not dst:
                dst.copy_(src, non_blocking=True)
        g.replay()
        return static_out

    replay.graph = g
    replay.static_inputs = static_in
    return replay


def invalidate_trainable_packs(*modules):
    """Every module that caches bf16 operand packs of TRAINABLE parameters (LoRA / DoRA adapters, SubjBasisGenerator layers)
    exposes ``invalidate()``; calling it makes the next forward rebuild the pack from the live parameters."""
    n = 0
    for root in modules:
        for m in root.modules():
            inv = getattr(m, "invalidate", None)
            if callable(inv):
                inv()
                n += 1
    return n


def graphed_step(fn, *trainable_modules, warmup=2):
    """CUDA-graph capture of a whole TRAINING step: ``fn()`` runs forward + backward and accumulates into pre-existing ``.grad``
    buffers (parallel.GradBucketer keeps them at fixed addresses).  The stage-2 iteration is ~14 000 kernel launches issued from
    Python (35 us each on average); replayed as one graph it is bound by the kernels instead.

    Requirements on ``fn``: no host synchronisation (no .item() / bool(tensor) / .tolist()), no collective (run the gradient
    all-reduce after the replay), the same control flow every iteration.  The operand packs of trainable parameters are
    invalidated right before the capture so that their rebuild is PART of the graph: a replay after an optimiser step computes
    with the updated weights.  Returns ``replay() -> fn's captured return value`` (tensors in the graph's private pool)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(warmup):
            invalidate_trainable_packs(*trainable_modules)
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    invalidate_trainable_packs(*trainable_modules)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()

    def replay():
        g.replay()
        return out

    replay.graph = g
    return replay
