"""CUDA-graph capture of a whole training step (forward + loss + backward through the libvtb200 kernels).

A Swin-S step is ~3700 kernel launches of 5-50 us each; issued one by one from Python the GPU waits on the host.
Capturing the step once and replaying it removes every per-launch host cost ("CUDA streams and graphs instead of a
tracing compiler").  Everything the step launches goes to torch's current stream, which is the capturing stream
inside `torch.cuda.graph`, so the ctypes C-ABI calls are captured like any other kernel; TMA descriptors are
passed by value as kernel parameters and the graph's private memory pool keeps every address stable across replays.
DropPath masks stay random: torch's CUDA generator is graph-aware (philox offsets advance per replay).
"""
import torch


class GraphedStep:
    """step_fn(*static_inputs) -> tensor (e.g. the loss); parameters' .grad are produced inside the graph.

    Usage:
        g = GraphedStep(step_fn, (x_static, y_static), warmup=3)
        x_static.copy_(x); y_static.copy_(y); loss = g.replay()
    """

    def __init__(self, step_fn, static_inputs, warmup=3):
        self.static_inputs = tuple(static_inputs)
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step_fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph):
            self.output = step_fn(*self.static_inputs)

    def replay(self):
        self.graph.replay()
        return self.output
