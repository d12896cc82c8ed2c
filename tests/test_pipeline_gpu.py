"""Feeder -> BatchStream -> WGANGPTrainer -> checkpoints -> generate, end to end on the device:
the batches that reach the GPU are bit-identical to the host-side reference collation, the CLI loop trains through CUDA
graphs, and its checkpoint drives generate.py's sampling loop."""
import os
from importlib import import_module

import numpy as np
import pytest
import torch

import kgan_b200 as kgan
from test_pipeline_cpu import RefFeeder, _train_opts, make_dataset

pytestmark = pytest.mark.gpu

feeder_mod = import_module("kinetic-gan_b200.feeder")
train_mod = import_module("kinetic-gan_b200.train")
gen_mod = import_module("kinetic-gan_b200.generate")


def test_batch_stream_cuda_ring_is_bit_exact(tmp_path):
    """More batches than ring slots, a consumer kernel reading every batch: values equal DataLoader(shuffle, drop_last) + crop."""
    dp, lp, data, labels = make_dataset(str(tmp_path), n=67)
    f = feeder_mod.Feeder(dp, lp)
    torch.manual_seed(11)
    loader = torch.utils.data.DataLoader(RefFeeder(data, labels, "ntu"), batch_size=4, shuffle=True, drop_last=True, num_workers=0)
    want = [(x[:, :, :16, :].float(), y.long()) for x, y in loader]
    torch.manual_seed(11)
    stream = feeder_mod.BatchStream(f, 4, 16, torch.device("cuda", 0), depth=2)
    sums, got = [], []
    for x, y in stream:
        assert x.is_cuda and y.is_cuda
        sums.append((x.double() * 3).sum())              # work enqueued on the consumer stream before the slot is recycled
        got.append((x.clone(), y.clone()))
    assert len(got) == len(want) == 16
    for (x, y), (xr, yr), s in zip(got, want, sums):
        assert torch.equal(x.cpu(), xr) and torch.equal(y.cpu(), yr)
        assert abs(s.item() - 3 * xr.double().sum().item()) < 1e-6


@pytest.mark.parametrize("graphs", [True, False])
def test_train_cli_and_generate_on_device(tmp_path, graphs):
    dp, lp, _, _ = make_dataset(str(tmp_path), n=41)
    out_root = os.path.join(str(tmp_path), "runs")
    opt = _train_opts(dp, lp, out_root, precision="fp32")
    opt.no_graphs = not graphs
    run, loss_d, loss_g = train_mod.train(opt)
    n_iter = 2 * (41 // 4)
    assert len(loss_d) == n_iter and np.isfinite(loss_d).all() and np.isfinite(loss_g).all()
    assert loss_g[1] == loss_g[0] and loss_g[5] != loss_g[0]           # generator step every n_critic iterations
    sd = torch.load(os.path.join(run, "models", "generator_8.pth"))
    G = kgan.Generator(512, 3, 6, 16, 2).cuda()
    G.load_state_dict(sd)
    assert all(torch.isfinite(v).all() for v in sd.values() if v.is_floating_point())
    a = np.load(os.path.join(run, "actions", "8.npy"))
    assert a.shape == (60, 3, 16, 25) and np.abs(a).max() <= 1.0
    gopt = gen_mod.build_parser().parse_args(["--n_classes", "6", "--t_size", "16", "--mlp_dim", "2", "--batch_size", "2", "--gen_qtd", "4",
                                             "--trunc_mode", "w", "--mean_size", "64"])
    kgan.set_precision("fp32")
    imgs, z, labels2 = gen_mod.generate_dataset(G, gopt, torch.device("cuda", 0))
    assert imgs.shape == (24, 3, 16, 25, 1) and np.isfinite(imgs).all() and np.abs(imgs).max() <= 1.0
    gopt.trunc_mode = "-"
    imgs2, _, _ = gen_mod.generate_dataset(G, gopt, torch.device("cuda", 0))          # CUDA-graph path of GeneratorRunner
    assert imgs2.shape == imgs.shape and np.isfinite(imgs2).all()


def test_same_seed_same_first_losses_graphs_vs_eager(tmp_path):
    """The CUDA-graph loop and the eager loop run the same arithmetic: identical seeds give the same loss history (fp32 path;
    atomics in the weight-gradient reductions and the CUDA RNG offsets consumed by the capture warm-up - the noise weights start at 0 - make the
    last digits run-dependent, hence a tolerance)."""
    dp, lp, _, _ = make_dataset(str(tmp_path), n=25)
    hist = []
    for graphs in (True, False):
        opt = _train_opts(dp, lp, os.path.join(str(tmp_path), "runs%d" % graphs), precision="fp32", max_iters=6)
        opt.no_graphs = not graphs
        _, loss_d, _ = train_mod.train(opt)
        hist.append(np.array(loss_d))
    assert hist[0].shape == hist[1].shape == (6,)
    assert np.allclose(hist[0], hist[1], rtol=1e-2, atol=1e-3)
