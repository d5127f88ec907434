"""Inference through a captured CUDA graph (the late InternImage stages are launch
bound -- 22 of the 30 InternImage-T DCNv3 layers run kernels of 20-60 us between cuBLAS / cuDNN calls of similar
size, and under the reference every layer is its own XLA cluster launch, intern_image_layer.py:122-174).

`GraphedInference(module)` captures `module(x)` once per input shape / dtype and replays the graph afterwards:
all launches of a forward -- the DCNv3 kernels with their programmatic-dependent-launch attributes included --
are submitted as one graph.  Works for a whole `InternImage`, a single `InternImageBlock` (one stage) or any
module whose forward allocates no cross-call state; gradients are not recorded (inference only).
"""
import torch


class GraphedInference(torch.nn.Module):
    def __init__(self, module, warmup=2):
        super().__init__()
        self.module = module
        self.warmup = warmup
        self._graphs = {}

    def _capture(self, x):
        static_in = x.clone()
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(self.warmup):  # lazy initialisation (cuBLAS handles, kernel attributes) outside the capture
                self.module(static_in)
            side.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                static_out = self.module(static_in)
        torch.cuda.current_stream(x.device).wait_stream(side)
        return graph, static_in, static_out

    def forward(self, x):
        if self.module.training or torch.is_grad_enabled():
            return self.module(x)  # training / autograd recording: eager (call under torch.no_grad() for the graph)
        key = (tuple(x.shape), x.dtype, x.device)
        if key not in self._graphs:
            self._graphs[key] = self._capture(x)
        graph, static_in, static_out = self._graphs[key]
        static_in.copy_(x)
        graph.replay()
        clone = lambda t: t.clone()  # noqa: E731  (the static outputs are overwritten by the next replay)
        if isinstance(static_out, (list, tuple)):
            return type(static_out)(clone(t) for t in static_out)
        return clone(static_out)
