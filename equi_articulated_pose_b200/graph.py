"""A whole training step (forward + backward + gradient exchange + optimiser) captured once in a CUDA graph.

The step of the classic backbone is ~380 kernel launches issued from ~600 Python-level calls; enqueueing them costs the
host ~14 ms per step, about as long as the GPU needs to execute them, so the host would bound the step as soon as the
kernels get faster.  All shapes of the path are static (per-sample convolution over fixed-size clouds), which makes the
step a textbook CUDA-graph candidate: capture once, replay with new inputs copied into the static input buffers.

Requirements on `step_fn`: no host synchronisation (no .item(), no data-dependent Python control flow), optimiser built
with `capturable=True`, every tensor it allocates comes from torch's caching allocator (the capture gets a private
pool).  The ctypes entry points of libvgtkb200 launch on torch's current stream, i.e. the capture stream.
Nothing may keep an autograd graph of an EARLIER eager step alive when the capture starts (return `loss.detach()` from the
step, drop outputs): its AccumulateGrad nodes are bound to the stream they were created on, and the captured backward
would synchronise with that stream, which invalidates the capture.
"""
import torch

from . import ops as _ops


class CapturedStep:
    def __init__(self, step_fn, example_inputs, warmup=3):
        """step_fn(*tensors) -> tensor or tuple of tensors; example_inputs: CUDA tensors with the static shapes."""
        self.static_in = [t.detach().clone() for t in example_inputs]
        cur = torch.cuda.current_stream()
        # warm-up and capture run on ONE side stream: autograd's AccumulateGrad nodes remember the stream they were created
        # on, and a node left over from a different stream would make the captured backward synchronise with it
        self.stream = torch.cuda.Stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):               # lazy init, kernel loading, allocator warm-up
            for _ in range(warmup):
                step_fn(*self.static_in)
        cur.wait_stream(self.stream)
        torch.cuda.synchronize()
        _ops.clear_planes()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.static_out = step_fn(*self.static_in)
        _ops.clear_planes()                                # the table only matters while the step is being recorded

    def __call__(self, *inputs):
        for s, t in zip(self.static_in, inputs):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.static_out
