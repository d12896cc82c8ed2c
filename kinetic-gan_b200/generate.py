"""Inference-only generator pass, the compute of the reference's `generate.py` (generate.py:57-67 model setup in eval
mode, :85-93 latent / label construction and `generator(z, labels, trunc)`), restructured for B200:

  * no autograd graph (the reference builds one: it has no `no_grad`);
  * the whole forward pass - mapping network, 7 blocks, per-block noise draws - is captured ONCE into a CUDA graph
    fed from static (z, labels) buffers, so a call is two small copies and one graph launch;
  * W-space truncation (generate.py default `--trunc_mode w`) runs the 1000 mapping passes of generator.py:97-108 as one
    batched pass (models/generator.py: Generator.truncate).

File outputs (.npy / labels .pkl, generate.py:105-123) are host I/O and out of scope."""
import numpy as np
import torch

from . import ops


class GeneratorRunner:
    """`runner(z, labels) -> (N, C, T, V)` in eval mode.  `z`: (N, latent) float32, `labels`: (N,) int64; host (pinned)
    or device tensors.  The returned tensor is the graph's static output buffer: copy it (or `to_host`) before the next call."""

    def __init__(self, generator, batch, latent_dim=512, trunc=None, graphs=True, device=None):
        self.G = generator.eval()
        self.device = device or next(generator.parameters()).device
        # W-space truncation draws its 1000 latents from the HOST RNG on every call (generator.py:98): a captured graph would freeze
        # them (and a pageable H2D copy cannot be captured), so that mode launches eagerly
        self.batch, self.trunc, self.graphs = batch, trunc, graphs and trunc is None
        self.z = torch.zeros(batch, latent_dim, device=self.device)
        self.labels = torch.zeros(batch, dtype=torch.long, device=self.device)
        self.out = None
        self._graph = None
        self.launches_per_call = 0
        self._host_out = None

    def _forward(self):
        with torch.no_grad():
            return self.G(self.z, self.labels, self.trunc)

    def capture(self):
        if self._graph is not None or not self.graphs:
            return
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                # warm-up: builds geometry tables, packed weights, descriptors
            for _ in range(2):
                self._forward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        ops.clear_temporary_packs()
        g = torch.cuda.CUDAGraph()
        l0 = ops.launches
        with torch.cuda.graph(g):
            self.out = self._forward()
        self.launches_per_call = ops.launches - l0
        ops.clear_temporary_packs()
        self._graph = g

    def __call__(self, z, labels):
        assert z.shape == self.z.shape and labels.shape == self.labels.shape
        if z is not self.z:
            self.z.copy_(z, non_blocking=True)
        if labels is not self.labels:
            self.labels.copy_(labels, non_blocking=True)
        if self.graphs:
            self.capture()
            self._graph.replay()
            ops.launches += self.launches_per_call
        else:
            self.out = self._forward()
        return self.out

    def to_host(self):
        """Asynchronous copy of the last result into a pinned host buffer (what generate.py's `.cpu()` at :95 does)."""
        if self._host_out is None:
            self._host_out = torch.empty(self.out.shape, dtype=self.out.dtype).pin_memory()
        self._host_out.copy_(self.out, non_blocking=True)
        return self._host_out


def class_conditioned_batch(n_classes, per_class, latent_dim=512, seed=None):
    """generate.py:85-91: `per_class` N(0,1) latents for every class label 0..n_classes-1 (host RNG, as the reference)."""
    rng = np.random.RandomState(seed) if seed is not None else np.random
    z = torch.as_tensor(rng.normal(0, 1, (n_classes * per_class, latent_dim)), dtype=torch.float32)
    labels = torch.as_tensor(np.array([c for _ in range(per_class) for c in range(n_classes)]), dtype=torch.long)
    return z, labels
