"""The caller of the hot path: `kinetic-gan.py` re-hosted on the B200 trainer.

    python kinetic-gan.py --data_path train_data.npy --label_path train_label.pkl [reference options ...]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 kinetic-gan.py ...            # batch-sharded DDP

Every option of the reference script (kinetic-gan.py:23-44) exists with the same name, meaning and default; the run
directory (`runs/kinetic-gan/expN/{models,actions}`, `config.txt`), the checkpoints (`generator_%d.pth`,
`discriminator_%d.pth`: plain state_dicts with the reference's keys, loadable by the reference's generate.py), the action
samples (`actions/%d.npy`, kinetic-gan.py:84-91) and the loss history (`plot_loss.mat`, utils/general.py:24-27) are written
in the reference's formats.  What differs is how the loop body runs (wgan_gp.WGANGPTrainer: CUDA graphs, fused Adam, no
per-iteration host sync - losses are read back every `--log_interval` iterations instead of every iteration,
kinetic-gan.py:176-182) and where the batches come from (feeder.BatchStream).  `batch_size` is per GPU.
Options that are not in the reference: --precision, --log_interval, --max_iters, --no_graphs, --seed, --out.
"""
import argparse
import os
import re

import numpy as np
import torch

from . import ops
from .ddp import Comm
from .feeder import BatchStream, Feeder
from .models.discriminator import Discriminator
from .models.generator import Generator
from .wgan_gp import WGANGPTrainer


def human_sorted(names):
    return sorted(names, key=lambda k: [int(s) if s.isdigit() else s.lower() for s in re.split('([0-9]+)', k)])


def check_runs(method, id=None, root="runs"):
    """utils/general.py:13-20: `runs/<method>/exp<k>`; a new directory when `id` is None, else the id-th existing one."""
    exps_path = os.path.join(root, method)
    os.makedirs(exps_path, exist_ok=True)
    exps = [e for e in human_sorted(os.listdir(exps_path)) if 'exp' in e]
    if id:
        return os.path.join(exps_path, exps[id])
    out = os.path.join(exps_path, 'exp' + str(len(exps) + 1))
    os.makedirs(out)
    return out


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--n_epochs", type=int, default=1200, help="passes over the dataset")
    p.add_argument("--batch_size", type=int, default=32, help="samples per batch and per GPU")
    p.add_argument("--lr", type=float, default=0.0002, help="Adam step size")
    p.add_argument("--b1", type=float, default=0.5, help="Adam beta1")
    p.add_argument("--b2", type=float, default=0.999, help="Adam beta2")
    p.add_argument("--n_cpu", type=int, default=8, help="kept for compatibility (batches are assembled by one gather thread)")
    p.add_argument("--latent_dim", type=int, default=512, help="length of the latent vector z")
    p.add_argument("--mlp_dim", type=int, default=4, help="number of Linear layers in the mapping network")
    p.add_argument("--n_classes", type=int, default=60, help="number of action classes")
    p.add_argument("--t_size", type=int, default=64, help="frames per sequence (T)")
    p.add_argument("--v_size", type=int, default=25, help="joints per frame (V); fixed by --dataset, kept for compatibility")
    p.add_argument("--channels", type=int, default=3, help="coordinates per joint (C)")
    p.add_argument("--n_critic", type=int, default=5, help="critic updates per generator update")
    p.add_argument("--lambda_gp", type=int, default=10, help="weight of the gradient penalty in the critic loss")
    p.add_argument("--sample_interval", type=int, default=5000, help="iterations between saved action samples")
    p.add_argument("--checkpoint_interval", type=int, default=10000, help="iterations between checkpoints (-1: never)")
    p.add_argument("--dataset", type=str, default="ntu", help="skeleton layout: ntu or h36m")
    p.add_argument("--data_path", type=str, required=True, help=".npy with the training sequences")
    p.add_argument("--label_path", type=str, required=True, help=".pkl with (names, labels)")
    # not in the reference
    p.add_argument("--precision", default="tf32", choices=["fp32", "fp32x3", "tf32"], help="libkgan arithmetic mode (DESIGN.md §4): tf32 tensor cores / fp32 FMA kernels / fp32x3 = fp32-accurate tensor-core split")
    p.add_argument("--log_interval", type=int, default=100, help="iterations between loss read-backs / prints")
    p.add_argument("--max_iters", type=int, default=-1, help="stop after this many iterations (-1: n_epochs decides)")
    p.add_argument("--no_graphs", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    p.add_argument("--seed", type=int, default=None, help="seed of torch / numpy RNGs (rank-offset for the per-rank draws)")
    p.add_argument("--out", type=str, default="runs", help="root of the run directories")
    return p


def sample_action(generator, n_row, latent_dim, path, device, keep_buffers=False):
    """kinetic-gan.py:84-91: 10 actions per class through the TRAINING-mode generator, saved as one (10*n_row, C, T, V) .npy.
    The pass updates the BatchNorm running statistics (as in the reference).  `keep_buffers`: restore them afterwards - used under
    DDP, where only rank 0 samples: its replica's buffers (and with them the checkpoint) would otherwise drift from the other
    ranks' by one extra momentum update per sample interval."""
    z = torch.as_tensor(np.random.normal(0, 1, (10 * n_row, latent_dim)), dtype=torch.float32, device=device)
    labels = torch.as_tensor(np.array([num for _ in range(10) for num in range(n_row)]), dtype=torch.long, device=device)
    saved = [(b, b.detach().clone()) for b in generator.buffers()] if keep_buffers else []
    with torch.no_grad():
        gen = generator(z, labels)
        for b, s in saved:
            b.copy_(s)
    with open(path, 'wb') as f:
        np.save(f, gen.cpu().numpy())


def save_losses(out, loss_d, loss_g):
    from scipy.io import savemat

    savemat(os.path.join(out, 'plot_loss.mat'), {'d_loss': np.asarray(loss_d, np.float32), 'g_loss': np.asarray(loss_g, np.float32)})


class HostDraws:
    """The host-RNG draws of one iteration - z ~ N(0,1) (kinetic-gan.py:140) and alpha ~ U[0,1) (:97), numpy's global RNG
    as in the reference - staged through a ring of pinned buffers: the asynchronous H2D copy of draw i may still be queued
    behind step i-1's kernels when draw i+1 is made, so a buffer is only refilled after its own copy has completed."""

    def __init__(self, batch, latent_dim, device, depth=3):
        self.device, self.cuda, self.k = device, device.type == "cuda", 0
        self.host = [(torch.empty(batch, latent_dim), torch.empty(batch, 1, 1, 1)) for _ in range(depth)]
        if self.cuda:
            self.host = [(a.pin_memory(), b.pin_memory()) for a, b in self.host]
            self.dev = [(torch.empty_like(a, device=device), torch.empty_like(b, device=device)) for a, b in self.host]
            self.done = [None] * depth

    def next(self):
        s = self.k % len(self.host)
        self.k += 1
        zh, ah = self.host[s]
        if self.cuda and self.done[s] is not None:
            self.done[s].synchronize()
        zh.copy_(torch.from_numpy(np.random.normal(0, 1, tuple(zh.shape))))
        ah.copy_(torch.from_numpy(np.random.random(tuple(ah.shape))))
        if not self.cuda:
            return zh.clone(), ah.clone()
        zd, ad = self.dev[s]
        zd.copy_(zh, non_blocking=True)
        ad.copy_(ah, non_blocking=True)
        self.done[s] = torch.cuda.Event()
        self.done[s].record(torch.cuda.current_stream(self.device))
        return zd, ad


def train(opt, comm=None):
    """Runs the loop of kinetic-gan.py:119-192.  Returns (run directory, d-loss history, g-loss history)."""
    comm = comm or Comm()
    cuda = torch.cuda.is_available()
    device = torch.device("cuda", comm.local_rank) if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(device)
        ops.set_precision(opt.precision)
    if opt.seed is not None:
        torch.manual_seed(opt.seed)
        np.random.seed(opt.seed + comm.rank)

    out = models_out = actions_out = None
    if comm.rank == 0:
        out = check_runs('kinetic-gan', root=opt.out)
        models_out, actions_out = os.path.join(out, 'models'), os.path.join(out, 'actions')
        os.makedirs(models_out, exist_ok=True)
        os.makedirs(actions_out, exist_ok=True)
        with open(os.path.join(out, "config.txt"), "w") as f:
            f.write(os.path.basename(__file__) + '|' + str(opt))

    generator = Generator(opt.latent_dim, opt.channels, opt.n_classes, opt.t_size, opt.mlp_dim, dataset=opt.dataset).to(device)
    discriminator = Discriminator(opt.channels, opt.n_classes, opt.t_size, opt.latent_dim, dataset=opt.dataset).to(device)
    trainer = WGANGPTrainer(generator, discriminator, opt.lr, opt.b1, opt.b2, opt.n_critic, opt.lambda_gp, comm=comm)

    feeder = Feeder(opt.data_path, opt.label_path, dataset=opt.dataset)
    stream = BatchStream(feeder, opt.batch_size, opt.t_size, device, comm.rank, comm.world_size, comm=comm)
    if len(stream) == 0:
        raise ValueError("dataset of %d samples is smaller than one global batch of %d" % (len(feeder), opt.batch_size * comm.world_size))
    if opt.seed is not None:                        # per-rank torch draws (the generator's per-block noise) after the shared model init
        torch.manual_seed(opt.seed + 7919 * (comm.rank + 1))

    draws = HostDraws(opt.batch_size, opt.latent_dim, device)

    loss_d, loss_g, pending = [], [], []
    g_last = None
    batches_done, stop = 0, False
    for epoch in range(opt.n_epochs):
        for i, (real, labels) in enumerate(stream):
            batches_done = epoch * len(stream) + i
            if cuda and not opt.no_graphs and trainer._graphs is None:
                trainer.capture_graphs(real, labels, torch.zeros(opt.batch_size, opt.latent_dim, device=device),
                                       torch.full((opt.batch_size, 1, 1, 1), 0.5, device=device))
            z, alpha = draws.next()
            d_loss, g_loss, _ = trainer.iteration(i, real, labels, z, alpha)
            if g_loss is not None:
                g_last = g_loss
            # losses stay on the device; one stacked read-back per log interval (the reference syncs every iteration, :176-182)
            pending.append(torch.stack((d_loss.reshape(()), g_last.reshape(()))).clone())
            if len(pending) >= opt.log_interval or (opt.max_iters > 0 and batches_done + 1 >= opt.max_iters):
                vals = torch.stack(pending).cpu().numpy()
                pending = []
                loss_d.extend(vals[:, 0].tolist())
                loss_g.extend(vals[:, 1].tolist())
                if comm.rank == 0:
                    print("[Epoch %d/%d] [Batch %d/%d] [D loss: %f] [G loss: %f]" % (epoch, opt.n_epochs, i, len(stream), vals[-1, 0], vals[-1, 1]),
                          flush=True)
            if batches_done % opt.sample_interval == 0 or (opt.checkpoint_interval != -1 and batches_done % opt.checkpoint_interval == 0):
                trainer.synchronize_updates()       # the optimizer updates run on a side stream: land them before the weights are read
            if comm.rank == 0 and batches_done % opt.sample_interval == 0:
                sample_action(generator, opt.n_classes, opt.latent_dim, os.path.join(actions_out, str(batches_done) + '.npy'), device,
                              keep_buffers=comm.world_size > 1)
                if pending:                         # the history written now includes this iteration (kinetic-gan.py:184-188)
                    vals = torch.stack(pending).cpu().numpy()
                    pending = []
                    loss_d.extend(vals[:, 0].tolist())
                    loss_g.extend(vals[:, 1].tolist())
                save_losses(out, loss_d, loss_g)
            if comm.rank == 0 and opt.checkpoint_interval != -1 and batches_done % opt.checkpoint_interval == 0:
                # parameters are views of one flat buffer (wgan_gp.FlatParams): clone so that each entry owns its storage
                torch.save({k: v.detach().clone() for k, v in generator.state_dict().items()}, os.path.join(models_out, "generator_%d.pth" % batches_done))
                torch.save({k: v.detach().clone() for k, v in discriminator.state_dict().items()},
                           os.path.join(models_out, "discriminator_%d.pth" % batches_done))
            if opt.max_iters > 0 and batches_done + 1 >= opt.max_iters:
                stop = True
                break
        if stop:
            break
    if pending:
        vals = torch.stack(pending).cpu().numpy()
        loss_d.extend(vals[:, 0].tolist())
        loss_g.extend(vals[:, 1].tolist())
    if comm.rank == 0:
        save_losses(out, loss_d, loss_g)
    return out, loss_d, loss_g


def main(argv=None):
    opt = build_parser().parse_args(argv)
    comm = Comm()
    if comm.rank == 0:
        print(opt)
    try:
        train(opt, comm)
    finally:
        comm.close()


if __name__ == "__main__":
    main()
