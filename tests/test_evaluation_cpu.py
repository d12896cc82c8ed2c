"""MMD evaluation against a loop-by-loop restatement of evaluation/mmd-actions.py:26-113 (its numpy
path), including the reference's quirks: first action of every class only, joints as the sample axis, max over 14
bandwidths starting from 0, NaN never winning."""
import os
import pickle
from importlib import import_module

import numpy as np
import pytest
import torch

import kgan_b200  # noqa: F401

ev = import_module("kinetic-gan_b200.evaluation")
feeder_mod = import_module("kinetic-gan_b200.feeder")


def ref_rkhs_mmd(s1, s2, bw):                                   # mmd-actions.py:26-55, numpy branch
    m = s1.shape[0]

    def k(a, b):
        return np.exp(-np.sum((np.expand_dims(a, 1) - b) ** 2, axis=-1) / bw)

    h = k(s1, s1) + k(s2, s2) - 2 * k(s1, s2)
    with np.errstate(invalid="ignore"):
        return np.sqrt(np.sum(h - np.diag(np.diag(h))) / (m * (m - 1)))


def ref_sequence_mmd(q1, q2, bw, mode):                         # :57-76
    if mode == "avg":
        return sum(ref_rkhs_mmd(q1[:, f, :], q2[:, f, :], bw) / q1.shape[1] for f in range(q1.shape[1]))
    return ref_rkhs_mmd(q1.reshape(q1.shape[0], -1), q2.reshape(q2.shape[0], -1), bw)


def ref_calculate(gen, real, label, mode):                      # :79-113
    n_cls = label.shape[-1]
    gl, rl = [[] for _ in range(n_cls)], [[] for _ in range(n_cls)]
    for i in range(len(gen)):
        c = np.argmax(label[i])
        if len(gl[c]) < 2000:
            gl[c].append(gen[i])
            rl[c].append(real[i])
    res = []
    for c in range(n_cls):
        best = 0
        for j in range(-4, 10):
            v = ref_sequence_mmd(np.asarray(gl[c])[0], np.asarray(rl[c])[0], 10 ** j, mode)
            if v > best:
                best = v
        res.append(best)
    return np.mean(res), np.array(res)


@pytest.mark.parametrize("mode", ["avg", "joint"])
def test_calculate_mmd_matches_reference_loops(mode):
    rng = np.random.RandomState(0)
    n_cls, per, V, T, C = 4, 3, 7, 6, 3
    gen = rng.uniform(-1, 1, (n_cls * per, V, T, C)).astype(np.float32)
    real = (gen * 0.5 + rng.uniform(-1, 1, gen.shape) * 0.5).astype(np.float32)
    lab = np.zeros((n_cls * per, n_cls))
    lab[np.arange(n_cls * per), np.tile(np.arange(n_cls), per)] = 1
    want, want_pc = ref_calculate(gen.astype(np.float64), real.astype(np.float64), lab, mode)
    got, got_pc = ev.calculate_mmd(gen, real, lab, mode, device="cpu", return_per_class=True, dtype=torch.float64)
    assert np.allclose(got_pc, want_pc, rtol=1e-9, atol=1e-12) and abs(got - want) < 1e-10
    assert got > 0
    # float32 (the reference's tensor dtype): estimates that vanish in float64 carry ~sqrt(eps) of rounding noise
    got32, got32_pc = ev.calculate_mmd(gen, real, lab, mode, device="cpu", return_per_class=True)
    assert np.abs(got32_pc - want_pc).max() < 5e-4


def test_mmd_class_surface_and_identical_sets():
    rng = np.random.RandomState(1)
    a, b = rng.randn(9, 5).astype(np.float32), rng.randn(9, 5).astype(np.float32)
    m = ev.MMD("avg")
    assert abs(m.rkhs_mmd(a, b, 10.0) - ref_rkhs_mmd(a.astype(np.float64), b.astype(np.float64), 10.0)) < 1e-5
    assert m.rkhs_mmd(torch.as_tensor(a), torch.as_tensor(a), 1.0) == 0.0            # identical sets: h == 0 exactly
    q1, q2 = rng.randn(6, 4, 3).astype(np.float32), (rng.randn(6, 4, 3) + 3.0).astype(np.float32)     # well separated: no negative estimate
    for mode in ("avg", "joint"):
        m.reset(mode)
        want = ref_sequence_mmd(q1.astype(np.float64), q2.astype(np.float64), 100.0, mode)
        assert want > 0.1 and abs(m.compute_sequence_mmd(q1, q2, 100.0) - want) < 1e-5
    # a negative MMD^2 estimate is NaN after the square root, in the reference and here alike
    close = rng.randn(6, 1, 3).astype(np.float32)
    m.reset("joint")
    v = [m.compute_sequence_mmd(close, close[::-1].copy() + 1e-3, bw) for bw in (1e-2, 1.0, 1e2)]
    r = [ref_sequence_mmd(close.astype(np.float64), (close[::-1].copy() + 1e-3).astype(np.float64), bw, "joint") for bw in (1e-2, 1.0, 1e2)]
    assert [np.isnan(a) for a in v] == [np.isnan(b) for b in r]
    m.reset("other")
    with pytest.raises(Exception, match="undefined mode"):
        m.compute_sequence_mmd(q1, q2, 1.0)
    with pytest.raises(AssertionError):
        ev.MMD("joint").rkhs_mmd(a, b[:5], 1.0)


def test_first_per_class_and_script(tmp_path):
    """Selection loop (:136-164) and the end-to-end script on files in generate.py's output format."""
    rng = np.random.RandomState(2)
    n_cls, n, C, T, V = 10, 260, 2, 20, 16
    labels = (np.arange(n) * 7 % n_cls)
    real = rng.uniform(-3, 5, (n, C, T, V)).astype(np.float32)
    fake = rng.uniform(-1, 1, (n, C, T, V)).astype(np.float32)
    paths = {}
    for name, arr in (("real", real), ("fake", fake)):
        paths[name] = (os.path.join(str(tmp_path), name + ".npy"), os.path.join(str(tmp_path), name + ".pkl"))
        np.save(paths[name][0], arr)
        with open(paths[name][1], "wb") as f:
            pickle.dump((["s%d" % i for i in range(n)], labels.tolist()), f)
    ds = feeder_mod.Feeder(paths["fake"][0], paths["fake"][1], norm=False, dataset="h36m")
    acts, labs = ev.first_per_class(ds, np.arange(n_cls), per_class=5, t_size=16)
    assert acts.shape == (50, C, 16, V) and labs.tolist() == [c for c in range(n_cls) for _ in range(5)]
    first3 = np.nonzero(labels == 3)[0][:5]
    assert np.array_equal(acts[15:20], fake[first3][:, :, :16, :])
    with pytest.raises(IndexError):
        ev.first_per_class(ds, np.arange(n_cls), per_class=100, t_size=16)           # only 26 per class
    # script: needs 100 per class -> 1000 samples
    n = 1000
    labels = np.arange(n) % n_cls
    real = rng.uniform(-3, 5, (n, C, T, V)).astype(np.float32)
    fake = rng.uniform(-1, 1, (n, C, T, V)).astype(np.float32)
    for name, arr in (("real", real), ("fake", fake)):
        np.save(paths[name][0], arr)
        with open(paths[name][1], "wb") as f:
            pickle.dump((["s%d" % i for i in range(n)], labels.tolist()), f)
    res = ev.main(["--data_real", paths["real"][0], "--labels_real", paths["real"][1], "--data_fake", paths["fake"][0],
                   "--labels_fake", paths["fake"][1], "--t_size", "16", "--dataset", "h36m", "--out", os.path.join(str(tmp_path), "runs")])
    cfg = open(os.path.join(str(tmp_path), "runs", "mmd-actions", "exp1", "config.txt")).read()
    assert "MMD_avg: " in cfg and 0 < res < 2
    # same number from the reference loops on the same selection
    realn = 2 * ((real - real.min()) / (real.max() - real.min())) - 1
    sel = np.concatenate([np.nonzero(labels == c)[0][:100] for c in range(n_cls)])
    onehot = np.eye(n_cls)[labels[sel]]
    want, _ = ref_calculate(fake[sel][:, :, :16].transpose(0, 3, 2, 1).astype(np.float64), realn[sel][:, :, :16].transpose(0, 3, 2, 1).astype(np.float64), onehot, "avg")
    assert abs(res - want) < 5e-4


# ---- against the reference's OWN code (tests/golden/mmd_ref.npz: class MMD, calcualte_mmd and the selection loops of the unmodified
# evaluation/mmd-actions.py, exec'd by tests/golden/make_mmd_golden.py) -------------------------------------------------------------
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mmd_ref.npz")


@pytest.mark.parametrize("mode", ["avg", "joint"])
def test_mmd_class_matches_reference_class(mode):
    g = np.load(GOLD)
    m = ev.MMD(mode)
    mine = np.array([m.compute_sequence_mmd(g["seq_1"], g["seq_2"], bw) for bw in g["bandwidths"]])
    grid = m.sequence_mmd_grid(g["seq_1"], g["seq_2"], tuple(float(b) for b in g["bandwidths"])).numpy()
    for ref in (g["seq_mmd_torch_" + mode], g["seq_mmd_numpy_" + mode]):
        ok = np.isfinite(ref)                      # bandwidths where MMD^2 rounds below zero give NaN in the reference too
        assert np.array_equal(np.isfinite(mine), ok)
        # fp32 evaluation of sqrt(difference of near-equal sums): absolute agreement at the 1e-4 level for tiny values
        assert np.allclose(mine[ok], ref[ok], rtol=2e-4, atol=2e-4), (mine, ref)
        assert np.allclose(grid[ok], ref[ok], rtol=2e-4, atol=2e-4)
    if mode == "avg":
        r = np.array([ev.MMD("avg").rkhs_mmd(g["seq_1"][:, 0], g["seq_2"][:, 0], bw) for bw in g["bandwidths"][3:8]])
        assert np.allclose(r, g["rkhs_numpy"], rtol=2e-4, atol=2e-4, equal_nan=True)


@pytest.mark.parametrize("mode", ["avg", "joint"])
def test_calculate_mmd_matches_reference_function(mode):
    g = np.load(GOLD)
    mine = ev.calculate_mmd(g["calc_gen"], g["calc_real"], g["calc_label"], mode, device="cpu")
    assert abs(mine - float(g["calc_result_" + mode])) < 2e-4 * max(1.0, abs(float(g["calc_result_" + mode]))), (mine, g["calc_result_" + mode])


def test_selection_matches_reference_loops_item0_quirk():
    """Dataset item 0 belongs to class 3: the reference's scan of every class after the first starts at item 1, so item 0 is never
    selected; the ids and labels below come from the reference's own loops."""
    g = np.load(GOLD)
    labels = g["select_labels"]
    n = len(labels)
    data = np.zeros((n, 2, 6, 3), np.float32)
    data[:, 0, 0, 0] = np.arange(n)

    class DS:
        label = labels

        def __getitem__(self, i):
            return data[i], int(labels[i])

    acts, lab = ev.first_per_class(DS(), np.arange(10), 100, int(g["select_t_size"]))
    assert tuple(acts.shape) == tuple(g["select_shape"])
    assert np.array_equal(acts[:, 0, 0, 0].astype(int), g["select_ids_real"])
    assert np.array_equal(lab, g["select_label_batch"])
    assert 0 not in g["select_ids_real"] and labels[0] == 3


def test_frechet_distance_matches_scipy_statement():
    """fid-actions.py:106-157 (scipy.linalg.sqrtm of the product, real part) against the symmetric-eigenvalue evaluation, for
    full-rank and rank-deficient covariances (fewer samples than dimensions: the reference's eps fallback case)."""
    from scipy import linalg

    rng = np.random.RandomState(3)
    for n, d in ((400, 24), (16, 24)):
        f1 = rng.randn(n, d) @ rng.randn(d, d) * 0.3 + rng.randn(d)
        f2 = rng.randn(n, d) @ rng.randn(d, d) * 0.3
        mu1, s1 = np.mean(f1, 0), np.cov(f1, rowvar=False)
        mu2, s2 = np.mean(f2, 0), np.cov(f2, rowvar=False)
        covmean = linalg.sqrtm(s1.dot(s2))          # (reference: `sqrtm(..., disp=False)[0]`; scipy >= 1.16 dropped `disp`)
        if not np.isfinite(covmean).all():
            off = np.eye(d) * 1e-6
            covmean = linalg.sqrtm((s1 + off).dot(s2 + off))
        ref = (mu1 - mu2).dot(mu1 - mu2) + np.trace(s1) + np.trace(s2) - 2 * np.trace(np.real(covmean))
        m1, c1 = ev.activation_statistics(f1, device="cpu")
        m2, c2 = ev.activation_statistics(f2, device="cpu")
        assert np.allclose(c1.numpy(), s1) and np.allclose(m2.numpy(), mu2)
        mine = ev.frechet_distance(m1, c1, m2, c2)
        assert abs(mine - ref) < 1e-6 * max(1.0, abs(ref)) + (1e-3 if n < d else 0.0), (n, d, mine, ref)
        assert abs(ev.frechet_distance(mu1, s1, mu1, s1)) < 1e-8 * np.trace(s1)
