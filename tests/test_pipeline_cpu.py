"""The callers either side of the hot path: Feeder / BatchStream against the reference's
Dataset + DataLoader semantics (feeder/feeder.py:43-80, kinetic-gan.py:68-74,129-131), the training CLI's run directory,
checkpoints and samples (kinetic-gan.py:16-21,84-91,184-192), and generate.py's outputs (generate.py:70-123).
C-ABI primitives are emulated on CPU (tests/emu_backend.py)."""
import os
import pickle
from importlib import import_module

import numpy as np
import pytest
import torch

import kgan_b200 as kgan

feeder_mod = import_module("kinetic-gan_b200.feeder")
train_mod = import_module("kinetic-gan_b200.train")
gen_mod = import_module("kinetic-gan_b200.generate")


def make_dataset(tmp, dataset="ntu", n=23, c=3, t=20, v=25, n_classes=6, seed=0):
    rng = np.random.RandomState(seed)
    shape = (n, c, t, v, 2) if dataset == "ntu" else (n, c, t, v)
    data = (rng.rand(*shape) * 7 - 3).astype(np.float32)
    labels = rng.randint(0, n_classes, n).tolist()
    dp, lp = os.path.join(tmp, dataset + "_data.npy"), os.path.join(tmp, dataset + "_label.pkl")
    np.save(dp, data)
    with open(lp, "wb") as f:
        pickle.dump((["s%d" % i for i in range(n)], labels), f)
    return dp, lp, data, np.array(labels)


class RefFeeder(torch.utils.data.Dataset):
    """feeder/feeder.py:57-80 restated for the test: global min / max, person 0, per-sample normalisation."""

    def __init__(self, data, labels, dataset):
        self.data, self.label, self.dataset = data, labels, dataset
        self.max, self.min = data.max(), data.min()

    def __len__(self):
        return len(self.label)

    def __getitem__(self, i):
        a = np.array(self.data[i, :, :, :, 0]) if self.dataset == "ntu" else np.array(self.data[i])
        return 2 * ((a - self.min) / (self.max - self.min)) - 1, self.label[i]


@pytest.mark.parametrize("dataset", ["ntu", "h36m"])
def test_feeder_items_and_batches(tmp_path, dataset):
    v = 25 if dataset == "ntu" else 16
    dp, lp, data, labels = make_dataset(str(tmp_path), dataset, v=v)
    f = feeder_mod.Feeder(dp, lp, dataset=dataset)
    ref = RefFeeder(data, labels, dataset)
    assert len(f) == len(ref) and (f.N, f.C, f.T, f.V) == (23, 3, 20, v)
    for i in (0, 7, 22):
        a, la = f[i]
        b, lb = ref[i]
        assert np.array_equal(a, b) and la == lb and a.min() >= -1 and a.max() <= 1
    idx = [5, 2, 19, 2, 0]
    x, y = f.batch(idx, t_size=16)
    want = np.stack([ref[i][0][:, :16, :] for i in idx]).astype(np.float32)
    assert x.dtype == np.float32 and np.array_equal(x, want) and np.array_equal(y, labels[idx])


def test_feeder_class_subset(tmp_path):
    dp, lp, data, labels = make_dataset(str(tmp_path))
    f = feeder_mod.Feeder(dp, lp, classes=[4, 1])
    keep = np.isin(labels, [4, 1])
    assert len(f) == keep.sum()
    assert np.array_equal(f.label, np.where(labels[keep] == 4, 0, 1))          # re-indexed into `classes` (feeder.py:62)
    assert f.max == data.max() and f.min == data.min()                        # range of the WHOLE file (feeder.py:57 precedes the filter)


def test_batch_stream_matches_dataloader_order(tmp_path):
    """Same torch seed -> same shuffled batches as DataLoader(Feeder, shuffle=True, drop_last=True) + the crop / cast of
    kinetic-gan.py:129-131, for two epochs."""
    dp, lp, data, labels = make_dataset(str(tmp_path))
    f = feeder_mod.Feeder(dp, lp)
    torch.manual_seed(123)
    loader = torch.utils.data.DataLoader(RefFeeder(data, labels, "ntu"), batch_size=4, shuffle=True, drop_last=True, num_workers=0)
    want = [[(x[:, :, :16, :].type(torch.FloatTensor), y.type(torch.LongTensor)) for x, y in loader] for _ in range(2)]
    torch.manual_seed(123)
    stream = feeder_mod.BatchStream(f, 4, 16, "cpu")
    assert len(stream) == len(loader) == 5
    for epoch in range(2):
        got = list(stream)
        assert len(got) == 5
        for (x, y), (xr, yr) in zip(got, want[epoch]):
            assert x.dtype == torch.float32 and y.dtype == torch.int64
            assert torch.equal(x, xr) and torch.equal(y, yr)


def test_batch_stream_rank_sharding(tmp_path):
    """world=2: the two ranks' batches are the two halves of each global batch of the same permutation; nothing is
    visited twice in an epoch."""
    dp, lp, data, labels = make_dataset(str(tmp_path))
    f = feeder_mod.Feeder(dp, lp)
    per_rank = []
    for r in range(2):
        torch.manual_seed(9)
        per_rank.append(list(feeder_mod.BatchStream(f, 4, 16, "cpu", rank=r, world=2)))
    torch.manual_seed(9)
    whole = list(feeder_mod.BatchStream(f, 8, 16, "cpu"))
    assert len(per_rank[0]) == len(per_rank[1]) == len(whole) == 2
    for b in range(2):
        assert torch.equal(torch.cat((per_rank[0][b][0], per_rank[1][b][0])), whole[b][0])
        assert torch.equal(torch.cat((per_rank[0][b][1], per_rank[1][b][1])), whole[b][1])


def test_batch_stream_worker_error_surfaces(tmp_path):
    """A failure inside the gather thread is re-raised in the consumer instead of hanging the loop."""
    dp, lp, _, _ = make_dataset(str(tmp_path))
    f = feeder_mod.Feeder(dp, lp)

    def boom(*a, **k):
        raise RuntimeError("disk gone")

    f.batch = boom
    with pytest.raises(RuntimeError, match="disk gone"):
        list(feeder_mod.BatchStream(f, 4, 16, "cpu"))


def _train_opts(dp, lp, out, **kw):
    argv = ["--data_path", dp, "--label_path", lp, "--out", out, "--n_classes", "6", "--t_size", "16", "--mlp_dim", "2", "--batch_size", "4",
            "--n_epochs", "2", "--sample_interval", "4", "--checkpoint_interval", "4", "--log_interval", "3", "--seed", "5"]
    for k, v in kw.items():
        argv += ["--" + k, str(v)]
    return train_mod.build_parser().parse_args(argv)


def test_train_cli_defaults_match_reference():
    """Option names and defaults of kinetic-gan.py:23-44."""
    opt = train_mod.build_parser().parse_args(["--data_path", "d", "--label_path", "l"])
    ref = dict(n_epochs=1200, batch_size=32, lr=0.0002, b1=0.5, b2=0.999, n_cpu=8, latent_dim=512, mlp_dim=4, n_classes=60, t_size=64,
               v_size=25, channels=3, n_critic=5, lambda_gp=10, sample_interval=5000, checkpoint_interval=10000, dataset="ntu")
    for k, v in ref.items():
        assert getattr(opt, k) == v, k


def test_train_run_directory_checkpoints_and_generate(emu, tmp_path):
    from scipy.io import loadmat

    dp, lp, _, _ = make_dataset(str(tmp_path))
    out_root = os.path.join(str(tmp_path), "runs")
    opt = _train_opts(dp, lp, out_root)
    run, loss_d, loss_g = train_mod.train(opt)
    # run directory layout of kinetic-gan.py:16-21,46-49
    assert run == os.path.join(out_root, "kinetic-gan", "exp1")
    assert os.path.isdir(os.path.join(run, "models")) and os.path.isdir(os.path.join(run, "actions"))
    assert open(os.path.join(run, "config.txt")).read().startswith("train.py|Namespace(")
    n_iter = 2 * (23 // 4)
    assert len(loss_d) == len(loss_g) == n_iter and np.isfinite(loss_d).all() and np.isfinite(loss_g).all()
    # g_loss only changes on generator steps (i % n_critic == 0 within the epoch), kinetic-gan.py:160-176
    assert loss_g[1] == loss_g[0] and loss_g[4] == loss_g[0]
    mat = loadmat(os.path.join(run, "plot_loss.mat"))
    assert mat["d_loss"].size == n_iter and mat["g_loss"].size == n_iter
    # samples: 10 per class, (10 * n_classes, C, T, V) (kinetic-gan.py:84-91)
    for it in (0, 4, 8):
        a = np.load(os.path.join(run, "actions", "%d.npy" % it))
        assert a.shape == (60, 3, 16, 25) and a.dtype == np.float32 and np.abs(a).max() <= 1.0
    # checkpoints: plain state_dicts with the reference's keys, each entry owning its storage
    sd = torch.load(os.path.join(run, "models", "generator_8.pth"))
    G = kgan.Generator(512, 3, 6, 16, 2)
    assert set(sd) == set(G.state_dict())
    G.load_state_dict(sd)
    assert all(v.untyped_storage().nbytes() == v.numel() * v.element_size() for v in sd.values())
    sd_d = torch.load(os.path.join(run, "models", "discriminator_8.pth"))
    assert set(sd_d) == set(kgan.Discriminator(3, 6, 16, 512).state_dict())
    # a second run gets the next directory (utils/general.py:13-20)
    run2, _, _ = train_mod.train(_train_opts(dp, lp, out_root, max_iters=1))
    assert run2.endswith("exp2")

    # ---- generate.py on that checkpoint: file names, shapes, dtypes (generate.py:105-123)
    for mode, stem in (("w", "6_5_trunc0.95"), ("z", "6_5_trunc0.95"), ("-", "6_5")):
        gopt = gen_mod.build_parser().parse_args(["--model", os.path.join(run, "models", "generator_8.pth"), "--n_classes", "6", "--t_size", "16",
                                                 "--mlp_dim", "2", "--batch_size", "2", "--gen_qtd", "5", "--trunc_mode", mode, "--mean_size", "50"])
        assert gen_mod.output_stem(gopt) == stem
        G.load_state_dict(torch.load(gopt.model), strict=False)
        imgs, z, labels2 = gen_mod.generate_dataset(G, gopt, torch.device("cpu"))
        n = 3 * 2 * 6                                                      # ceil(5 / 2) rounds of 2 per class
        assert imgs.shape == (n, 3, 16, 25, 1) and imgs.dtype == np.float32
        assert z.shape == (n, 512) and labels2.shape == (2, n) and np.array_equal(labels2[0], labels2[1])
        assert np.array_equal(labels2[0][:12], np.tile(np.arange(6), 2))
        adir = os.path.join(str(tmp_path), "actions_" + mode.replace("-", "none"))
        os.makedirs(adir)
        written = gen_mod.write_outputs(adir, gopt, imgs, z, labels2)
        assert np.load(written + "_gen_data.npy").shape == imgs.shape
        assert np.load(written + "_gen_z.npy").shape == z.shape
        assert pickle.load(open(written + "_gen_label.pkl", "rb")).shape == (2, n)


def test_generate_stochastic_and_single_label(emu, tmp_path):
    """--stochastic repeats ONE latent (generate.py:81-83): outputs differ only through the per-block noise; --label k."""
    G = kgan.Generator(512, 3, 6, 16, 2)
    gopt = gen_mod.build_parser().parse_args(["--n_classes", "6", "--t_size", "16", "--mlp_dim", "2", "--batch_size", "3", "--gen_qtd", "3",
                                             "--trunc_mode", "-", "--stochastic", "--label", "4"])
    assert gen_mod.output_stem(gopt) == "4_3_stochastic"
    imgs, z, labels2 = gen_mod.generate_dataset(G, gopt, torch.device("cpu"))
    assert imgs.shape == (3, 3, 16, 25, 1) and (labels2 == 4).all()
    assert np.array_equal(z[0], z[1]) and np.array_equal(z[0], z[2])


def test_trunc_z_matches_reference_loop():
    """generate.py:14-21 row loop == the vector expression."""
    np.random.seed(3)
    lat = torch.randn(7, 12)
    got = gen_mod.trunc_z(lat.clone(), 40, 0.9)
    np.random.seed(3)
    t = torch.as_tensor(np.random.normal(0, 1, (40, 12)), dtype=torch.float32)
    m = t.mean(0, keepdim=True)
    want = lat.clone()
    for i, _ in enumerate(want):
        want[i] = m + 0.9 * (want[i] - m)
    assert torch.allclose(got, want, atol=1e-6)


def test_train_cli_two_ranks_gloo(tmp_path):
    """The CLI loop under torch.distributed (world_size 2, gloo): both ranks walk rank 0's permutation and take disjoint
    halves of every global batch; replicas stay identical (flat-gradient all-reduce) so the critic losses they log differ only
    through their different samples; only rank 0 writes a run directory."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dp, lp, _, _ = make_dataset(str(tmp_path))
    script = os.path.join(root, "tests", "ddp_train_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", OMP_NUM_THREADS="2")
    procs = [subprocess.Popen([sys.executable, script, dp, lp, str(tmp_path)], env=dict(env, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r)))
             for r in range(2)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    a = torch.load(os.path.join(str(tmp_path), "train_w2_r0.pt"), weights_only=False)
    b = torch.load(os.path.join(str(tmp_path), "train_w2_r1.pt"), weights_only=False)
    assert a["run"] is not None and b["run"] is None
    assert a["seen"].shape == b["seen"].shape == (3, 2)
    # rank 0's permutation (train.py re-seeds torch with seed + 7919 * (rank + 1) after the shared model init; rank 1's own
    # seed is different and must not matter): global batches of 4, rank r takes [2r, 2r + 2)
    torch.manual_seed(11 + 7919)
    perm = feeder_mod.epoch_permutation(23)
    for i in range(3):
        assert np.array_equal(a["seen"][i], perm[4 * i:4 * i + 2]) and np.array_equal(b["seen"][i], perm[4 * i + 2:4 * i + 4])
    assert len(a["loss_d"]) == len(b["loss_d"]) == 3 and np.isfinite(a["loss_d"]).all() and np.isfinite(b["loss_d"]).all()


def test_train_rejects_dataset_smaller_than_a_global_batch(emu, tmp_path):
    """drop_last semantics leave zero batches: the reference would silently run zero iterations per epoch; here it is an error."""
    dp, lp, _, _ = make_dataset(str(tmp_path), n=3)
    with pytest.raises(ValueError, match="smaller than one global batch"):
        train_mod.train(_train_opts(dp, lp, os.path.join(str(tmp_path), "runs")))


def test_feeder_without_normalisation_and_h36m_layout(tmp_path):
    """norm=False returns the stored values (how mmd-actions.py reads generated data); h36m files have no person axis."""
    dp, lp, data, labels = make_dataset(str(tmp_path), "h36m", n=9, c=2, t=12, v=16, n_classes=3)
    f = feeder_mod.Feeder(dp, lp, norm=False, dataset="h36m", mmap=False)
    a, la = f[4]
    assert np.array_equal(a, data[4]) and la == labels[4] and (f.N, f.C, f.T, f.V) == (9, 2, 12, 16)
    x, y = f.batch([8, 0], t_size=50)                      # t_size beyond the stored length: whole sequence
    assert x.shape == (2, 2, 12, 16) and np.array_equal(x, data[[8, 0]]) and y.tolist() == [labels[8], labels[0]]
    got = list(feeder_mod.BatchStream(f, 4, None, "cpu", shuffle=False))
    assert len(got) == 2 and torch.equal(got[1][0], torch.from_numpy(data[4:8]))
