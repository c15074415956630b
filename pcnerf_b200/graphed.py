"""CUDA-graph capture of a whole training step (render -> losses -> backward -> all-reduce-free optimizer step).

The path launches thousands of small kernels per step (one GEMM + one fold kernel per layer per BatchNorm chunk, or
the float64 parameter-sized algebra of the closed-form mode); replaying them as ONE graph removes the per-launch host
cost.  Everything on the path is capture-safe: kernels are enqueued on torch's current stream through the C ABI, random
numbers come from torch's graph-safe Philox generator, memory comes from the graph's private pool.  Data-dependent
shapes are not capturable, so the caller must keep the number of rays fixed (`ops.aabb_pack_train(..., compact=False)`).
"""
import torch

from . import ops


class GraphedStep:
    """Capture `fn()` (no arguments; it reads / writes fixed tensors) after `warmup` eager calls on a side stream.
    Calling the object replays the graph.  `fn` may call `loss.backward()` and `optimizer.step()` (capturable=True)."""

    def __init__(self, fn, warmup=3):
        self.fn = fn
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph):
            self.out = fn()
        torch.cuda.synchronize()

    def __call__(self):
        self.graph.replay()
        ops.note_param_write()      # the replayed kernels write parameters / running statistics behind torch's back
        return self.out
