"""A training step of the loss as a CUDA graph: capture once, replay every step.

An eager step of the public API costs ~0.14 ms of host time (ctypes calls 0.03, argument checks and allocations 0.02,
the autograd engine's hand-over to its device thread 0.09; profiles/r2_host_profile.txt) -- more than the kernels of a
small batch take, and half of what BASELINE's batch takes.  The reference's call sites run the loss on tensors of the
same shape step after step (run/ctc/cnn/train.py:191-200 inside a fixed-batch-size loop), which is what a CUDA graph
wants: ``GraphedStep`` captures ``loss = fn(); loss.backward()`` once -- the fork/join of the side stream inside
``b200ctc_forward`` is captured with it -- and ``replay()`` re-runs the kernels on the same buffers with one launch.
New data goes in by copying into the captured input tensors (``x.copy_(...)``, labels and lengths likewise); the
gradients appear in the captured tensors' ``.grad``.  bench.py's headline figure is measured through this class.
"""
import torch


class GraphedStep(object):
    def __init__(self, fn, inputs, warmup=1, capture_error_mode=None):
        """fn: () -> scalar loss tensor, calling this package's loss on fixed tensors.  inputs: the leaf tensors whose
        ``.grad`` the step produces (reset to None before the capture, so that the graph owns the gradient buffers)."""
        self.inputs = list(inputs)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # first calls outside the capture: lazy initialisations,
            for _ in range(max(1, int(warmup))):            # the library's side stream, the allocator's pools
                for t in self.inputs:
                    t.grad = None
                fn().backward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for t in self.inputs:
            t.grad = None
        self.graph = torch.cuda.CUDAGraph()
        kw = {} if capture_error_mode is None else {"capture_error_mode": capture_error_mode}
        with torch.cuda.graph(self.graph, **kw):
            self.loss = fn()
            self.loss.backward()

    def replay(self):
        """Runs the captured forward + backward; returns the (static) loss tensor, valid in stream order."""
        self.graph.replay()
        return self.loss
