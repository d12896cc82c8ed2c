"""Inference-only generator pass, the compute of the reference's `generate.py` (generate.py:57-67 model setup in eval
mode, :85-93 latent / label construction and `generator(z, labels, trunc)`), restructured for B200:

  * no autograd graph (the reference builds one: it has no `no_grad`);
  * the whole forward pass - mapping network, 7 blocks, per-block noise draws - is captured ONCE into a CUDA graph
    fed from static (z, labels) buffers, so a call is two small copies and one graph launch;
  * W-space truncation (generate.py default `--trunc_mode w`) runs the 1000 mapping passes of generator.py:97-108 as one
    batched pass (models/generator.py: Generator.truncate);
  * eval-mode BatchNorm is folded into the convolution in front of it (Generator.fold_batchnorm), so no BatchNorm kernel
    runs.  Re-create the runner (or call `generator.fold_batchnorm()` again) after loading other weights.

`generate_dataset` / `main` are the reference script itself on top of that runner: same options
(generate.py:27-47), same sampling loop (:70-103), same output files in the same formats (:105-123), with the generated
batches accumulated in one preallocated host array instead of an O(n^2) `np.concatenate` per iteration.

    python generate.py --model runs/kinetic-gan/exp1/models/generator_10000.pth --n_classes 60 --gen_qtd 1000
"""
import argparse
import os
import pickle
from collections import Counter

import numpy as np
import torch

from . import ops


class GeneratorRunner:
    """`runner(z, labels) -> (N, C, T, V)` in eval mode.  `z`: (N, latent) float32, `labels`: (N,) int64; host (pinned)
    or device tensors.  The returned tensor is the graph's static output buffer: copy it (or `to_host`) before the next call."""

    def __init__(self, generator, batch, latent_dim=512, trunc=None, graphs=True, device=None, cache_mean=False):
        """`cache_mean` (W-space truncation only): estimate the mean latent ONCE here instead of from 1000 fresh host draws in
        every call as generator.py:97-108 does (statistically the same estimator; the per-call draw costs ~10 ms of host RNG
        and forces eager launches) - the truncated pass then replays as a CUDA graph like the plain one."""
        self.G = generator.eval()
        self.device = device or next(generator.parameters()).device
        # W-space truncation draws its 1000 latents from the HOST RNG on every call (generator.py:98): a captured graph would freeze
        # them (and a pageable H2D copy cannot be captured), so that mode launches eagerly
        self.batch, self.trunc, self.graphs = batch, trunc, graphs and (trunc is None or cache_mean)
        self.cache_mean = cache_mean
        self._version = None                # Generator.weights_version the derived state below was built from
        self.z = torch.zeros(batch, latent_dim, device=self.device)
        self.labels = torch.zeros(batch, dtype=torch.long, device=self.device)
        self.out = None
        self._graph = None
        self.launches_per_call = 0
        self._ring = None               # to_host(): two (device staging, pinned host, D2H-done event) slots + the copy stream
        self._ring_k = 0
        self.host_ready = None          # event of the last to_host() copy

    def _refresh(self):
        """(Re)builds what is derived from the generator's weights: BatchNorm folds, the cached W-space mean, and - by dropping
        it - the captured graph.  Runs at the first call and whenever Generator.weights_version moved (load_state_dict, .to(), ...)."""
        if self._version == getattr(self.G, "weights_version", 0) and self._version is not None:
            return
        self.G.eval()
        self.G.fold_batchnorm()             # eval-mode BatchNorm becomes part of the convolution weights: no BN launches
        self.G._w_mean = self.G.estimate_w_mean() if (self.trunc is not None and self.cache_mean) else None
        self._graph = None
        self._version = getattr(self.G, "weights_version", 0)

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
        self._refresh()
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
        """Asynchronous copy of the last result into a pinned host buffer (what generate.py's `.cpu()` at :95 does), overlapped
        with the NEXT call: the result is first duplicated on the device (the graph's output buffer is overwritten by the next
        replay), then a side stream moves the duplicate over PCIe.  Returns the pinned tensor; it is valid once
        `runner.host_ready.synchronize()` (or any later device synchronisation) returns, and until the second next to_host()."""
        if self._ring is None:
            mk = lambda: [torch.empty_like(self.out), torch.empty(self.out.shape, dtype=self.out.dtype).pin_memory(), None]
            self._ring = ([mk(), mk()], torch.cuda.Stream(device=self.device))
        slots, copy_stream = self._ring
        slot = slots[self._ring_k]
        self._ring_k ^= 1
        cur = torch.cuda.current_stream(self.device)
        if slot[2] is not None:
            cur.wait_event(slot[2])                  # the staging buffer's previous D2H copy has drained
        slot[0].copy_(self.out)
        staged = torch.cuda.Event()
        staged.record(cur)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(staged)
            slot[1].copy_(slot[0], non_blocking=True)
            slot[2] = torch.cuda.Event()
            slot[2].record(copy_stream)
        self.host_ready = slot[2]
        return slot[1]


def class_conditioned_batch(n_classes, per_class, latent_dim=512, seed=None):
    """generate.py:85-91: `per_class` N(0,1) latents for every class label 0..n_classes-1 (host RNG, as the reference)."""
    rng = np.random.RandomState(seed) if seed is not None else np.random
    z = torch.as_tensor(rng.normal(0, 1, (n_classes * per_class, latent_dim)), dtype=torch.float32)
    labels = torch.as_tensor(np.array([c for _ in range(per_class) for c in range(n_classes)]), dtype=torch.long)
    return z, labels


# ------------------------------------------------------------------------------------------------------------------
# generate.py as a function + CLI
# ------------------------------------------------------------------------------------------------------------------
def trunc_z(latent, mean_size, truncation):
    """generate.py:14-21, truncation trick on Z: pull every latent towards the mean of `mean_size` fresh N(0,1) draws
    (host RNG, as the reference); one vector expression instead of the per-row loop."""
    t = torch.as_tensor(np.random.normal(0, 1, (mean_size, *latent.shape[1:])), dtype=latent.dtype, device=latent.device)
    m = t.mean(0, keepdim=True)
    return m + truncation * (latent - m)


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--batch_size", type=int, default=10, help="samples per class generated in one round")
    p.add_argument("--latent_dim", type=int, default=512, help="length of the latent vector z")
    p.add_argument("--mlp_dim", type=int, default=4, help="number of Linear layers in the mapping network")
    p.add_argument("--n_classes", type=int, default=60, help="number of action classes")
    p.add_argument("--label", type=int, default=-1, help="generate this class only (-1: every class)")
    p.add_argument("--t_size", type=int, default=64, help="frames per sequence (T)")
    p.add_argument("--v_size", type=int, default=25, help="joints per frame (V); fixed by --dataset, kept for compatibility")
    p.add_argument("--channels", type=int, default=3, help="coordinates per joint (C)")
    p.add_argument("--dataset", type=str, default="ntu", help="skeleton layout: ntu or h36m")
    p.add_argument("--model", type=str, default="runs/kinetic-gan/exp1/models/generator_ntu_xsub_mlp4_1370000.pth", help="generator checkpoint (state_dict)")
    p.add_argument("--stochastic", action='store_true', help="repeat ONE latent point for the whole batch (only the per-block noise varies)")
    p.add_argument("--stochastic_file", type=str, default="-", help=".npy of latents to take that point from (- : draw a fresh one)")
    p.add_argument("--stochastic_index", type=int, default=0, help="row of --stochastic_file to use")
    p.add_argument("--gen_qtd", type=int, default=1000, help="samples wanted per class")
    p.add_argument("--trunc", type=float, default=0.95, help="truncation factor")
    p.add_argument("--trunc_mode", type=str, default='w', choices=['z', 'w', '-'], help="truncate in Z space, in W space, or not at all")
    p.add_argument("--mean_size", type=int, default=1000, help="latents used to estimate the truncation mean")
    # not in the reference
    p.add_argument("--precision", default="tf32", choices=["fp32", "fp32x3", "tf32"], help="libkgan arithmetic mode (DESIGN.md §4): tf32 tensor cores / fp32 FMA kernels / fp32x3 = fp32-accurate tensor-core split")
    p.add_argument("--out", type=str, default=None, help="output directory (default: actions/ of the latest run, generate.py:24-26)")
    return p


def output_stem(opt):
    """File-name stem of generate.py:113-123."""
    return (str(opt.n_classes if opt.label == -1 else opt.label) + '_' + str(opt.gen_qtd)
            + ('_trunc' + str(opt.trunc) if opt.trunc_mode != '-' else '') + ('_stochastic' if opt.stochastic else ''))


def generate_dataset(generator, opt, device=None):
    """The sampling loop generate.py:70-103.  Returns (data, z, labels) exactly as the reference writes them:
    data (n, C, T, V[, 1 for ntu]) float32, z (n, latent) float32 (after Z truncation, as the reference stores it),
    labels (2, n) int (the label row duplicated, generate.py:110)."""
    device = device or next(generator.parameters()).device
    generator.eval()
    classes = list(np.arange(opt.n_classes)) if opt.label == -1 else [opt.label]
    qtd = opt.batch_size
    rounds = -(-opt.gen_qtd // qtd)                         # every class gains `qtd` samples per round: all finish together
    batch = qtd * len(classes)
    total = rounds * batch
    runner = GeneratorRunner(generator, batch, opt.latent_dim, trunc=opt.trunc if opt.trunc_mode == 'w' else None,
                             graphs=device.type == "cuda", device=device)
    imgs = None
    labels_all = np.empty(total, dtype=np.int64)
    if opt.stochastic:                                      # one latent point repeated (generate.py:81-83)
        if opt.stochastic_file != '-':
            z0 = np.expand_dims(np.load(opt.stochastic_file)[opt.stochastic_index], 0)
        else:
            z0 = np.random.normal(0, 1, (1, opt.latent_dim))
        z = torch.as_tensor(z0, dtype=torch.float32).repeat(batch, 1)
    pending = None                                          # (round, pinned result, its D2H event): drained one round later

    def drain(p):
        nonlocal imgs
        rr, gen, ev = p
        if ev is not None:
            ev.synchronize()
        if imgs is None:
            imgs = np.empty((total,) + tuple(gen.shape[1:]), np.float32)
        imgs[rr * batch:(rr + 1) * batch] = gen.numpy()

    z_all = np.empty((total, opt.latent_dim), np.float32)
    for r in range(rounds):
        if not opt.stochastic:
            z = torch.as_tensor(np.random.normal(0, 1, (batch, opt.latent_dim)), dtype=torch.float32)
        if opt.trunc_mode == 'z':
            z = trunc_z(z, opt.mean_size, opt.trunc)
        labels_np = np.array([num for _ in range(qtd) for num in classes])
        sl = slice(r * batch, (r + 1) * batch)
        z_all[sl], labels_all[sl] = z.numpy(), labels_np
        runner(z, torch.as_tensor(labels_np, dtype=torch.long))
        if device.type == "cuda":
            cur = (r, runner.to_host(), runner.host_ready)  # this round's copy overlaps the next round's generator pass
            if pending is not None:
                drain(pending)
            pending = cur
        else:
            drain((r, runner.out.clone(), None))
    if pending is not None:
        drain(pending)
    counts = Counter(labels_all.tolist())
    assert all(counts[c] >= opt.gen_qtd for c in classes)
    if opt.dataset == 'ntu':
        imgs = np.expand_dims(imgs, axis=-1)
    labels2 = np.concatenate((np.expand_dims(labels_all, 0), np.expand_dims(labels_all, 0)), axis=0)
    return imgs, z_all, labels2


def write_outputs(actions_out, opt, imgs, z_all, labels2):
    stem = os.path.join(actions_out, output_stem(opt))
    with open(stem + '_gen_data.npy', 'wb') as f:
        np.save(f, imgs)
    with open(stem + '_gen_z.npy', 'wb') as f:
        np.save(f, z_all)
    with open(stem + '_gen_label.pkl', 'wb') as f:
        pickle.dump(labels2, f)
    return stem


def main(argv=None):
    from .models.generator import Generator
    from .train import check_runs

    opt = build_parser().parse_args(argv)
    print(opt)
    actions_out = opt.out
    if actions_out is None:
        actions_out = os.path.join(check_runs('kinetic-gan', id=-1), 'actions')
    os.makedirs(actions_out, exist_ok=True)
    with open(os.path.join(os.path.dirname(actions_out.rstrip('/')) or '.', "gen_config.txt"), "w") as f:
        f.write(os.path.basename(__file__) + '|' + str(opt))
    device = torch.device("cuda", 0)
    ops.set_precision(opt.precision)
    generator = Generator(opt.latent_dim, opt.channels, opt.n_classes, opt.t_size, mlp_dim=opt.mlp_dim, dataset=opt.dataset).to(device)
    generator.load_state_dict(torch.load(opt.model, map_location=device), strict=False)          # generate.py:66
    imgs, z_all, labels2 = generate_dataset(generator, opt, device)
    print(write_outputs(actions_out, opt, imgs, z_all, labels2))


if __name__ == "__main__":
    main()
