"""MMD between generated and real actions, the quality metric of evaluation/mmd-actions.py.

The reference evaluates, for every class, 14 bandwidths (10^-4 .. 10^9) and - in 'avg' mode - every frame, with one
`rkhs_mmd` call each: three pairwise-distance tensors, ~10 small kernels and a `.item()` host synchronisation per call,
i.e. ~54 000 calls for NTU-60 at 64 frames (mmd-actions.py:96-110).  Here the whole (class x bandwidth x frame) grid is one
batched tensor expression on the device and ONE read-back: the arithmetic of a single cell is unchanged (direct squared
distances in fp32 as at mmd-actions.py:36-38, no Gram-matrix shortcut, so small bandwidths do not lose digits).

What the reference computes, including its quirks, is kept (published numbers come from this code):
  * per class only the FIRST selected generated / real action enters (`new_gen[0]`, `new_real[0]`, mmd-actions.py:108): after
    the (0, 3, 2, 1) transpose at :186-187 that is a (V, T, C) array, read as "V samples of a T-frame sequence of dimension C";
  * mmd = sqrt( sum(off-diagonal of Kxx + Kyy - 2 Kxy) / (m (m - 1)) ), which needs as many generated as real samples;
  * the score of a class is the maximum over the bandwidths, the result the mean over classes (:105-113).

`MMD` keeps the reference's class surface (mode, rkhs_mmd, compute_sequence_mmd) on top of the same batched kernel;
`first_per_class` is the selection loop of mmd-actions.py:136-164; `main` the script (same options).
FID (evaluation/fid-actions.py): the feature extractor is the third-party pretrained Inception network of `pytorch_fid` (no
weights in this image, no network) and stays out of scope; what the reference itself computes on top of the features - mean /
covariance (:160-183) and the Frechet distance between the two Gaussians (:106-157) - is here as `activation_statistics` and
`frechet_distance`, on the device.
"""
import argparse
import os

import numpy as np
import torch

BANDWIDTHS = tuple(10.0 ** j for j in range(-4, 10))           # mmd-actions.py:107


def _mmd_grid(x, y, bandwidths):
    """x, y: (..., m, d) sample sets (same m).  -> (len(bandwidths), ...) values of
    sqrt(sum_offdiag(Kxx + Kyy - 2 Kxy) / (m (m - 1)))  with the RBF kernel exp(-|a - b|^2 / bandwidth)  (mmd-actions.py:26-55)."""
    m = x.shape[-2]
    assert y.shape[-2] == m, "the estimator pairs the two sample sets: both need the same number of samples"

    def sq(a, b):
        return (a.unsqueeze(-2) - b.unsqueeze(-3)).pow(2).sum(-1)                     # (..., m, m), direct differences

    dxx, dyy, dxy = sq(x, x), sq(y, y), sq(x, y)
    off = 1.0 - torch.eye(m, dtype=x.dtype, device=x.device)
    out = []
    for bw in bandwidths:                                                             # 14 fused elementwise passes over small tensors
        h = torch.exp(-dxx / bw) + torch.exp(-dyy / bw) - 2.0 * torch.exp(-dxy / bw)
        out.append(((h * off).sum((-1, -2)) / (m * (m - 1))).pow(0.5))
    return torch.stack(out)


class MMD:
    """Same surface as the reference class (mmd-actions.py:14-76).  mode: 'avg' = frame-averaged MMD, 'joint' = the whole
    sequence as one vector.  Inputs may be numpy arrays or tensors (any device); results are Python floats."""

    def __init__(self, mode, use_torch=True):
        self.mode, self.use_torch = mode, use_torch

    def reset(self, new_mode):
        self.mode = new_mode

    @staticmethod
    def _t(a):
        return a if torch.is_tensor(a) else torch.as_tensor(np.asarray(a))

    def rkhs_mmd(self, samples_1, samples_2, bandwidth):
        """Two sample groups of shape (N, dim)."""
        return float(_mmd_grid(self._t(samples_1), self._t(samples_2), (float(bandwidth),))[0])

    def sequence_mmd_grid(self, sequence_1, sequence_2, bandwidths=BANDWIDTHS):
        """(..., N, len, dim) x 2 -> (len(bandwidths), ...) sequence MMDs, all bandwidths and all leading batch entries at once."""
        s1, s2 = self._t(sequence_1), self._t(sequence_2)
        if self.mode == 'avg':                                                         # mean over frames of the per-frame MMD (:63-65)
            return _mmd_grid(s1.transpose(-3, -2), s2.transpose(-3, -2), bandwidths).mean(-1)
        if self.mode == 'joint':                                                       # the sequence as one (len * dim) vector (:66-74)
            return _mmd_grid(s1.flatten(-2), s2.flatten(-2), bandwidths)
        raise Exception('undefined mode')

    def compute_sequence_mmd(self, sequence_1, sequence_2, bandwidth):
        """Sequences of shape (N, len, dim)."""
        return float(self.sequence_mmd_grid(sequence_1, sequence_2, (float(bandwidth),))[0])


def calculate_mmd(gen, real, label, mode='avg', device=None, bandwidths=BANDWIDTHS, return_per_class=False, dtype=torch.float32):
    """mmd-actions.py:79-113.  gen, real: (n, V, T, C) (already transposed as at :186-187), label: (n, n_classes) one-hot.
    Per class the first (up to 2000 are collected, only index 0 is used - see the module docstring) generated / real action,
    the maximum over `bandwidths`, then the mean over classes.  `dtype`: the reference's tensors are float32 (Feeder output); an
    MMD^2 estimate that is zero up to rounding then shows up as ~1e-4 after the square root - float64 removes that noise."""
    label = np.asarray(label)
    cls = label.argmax(-1)
    n_classes = label.shape[-1]
    first = np.array([int(np.nonzero(cls == c)[0][0]) for c in range(n_classes)])      # raises like the reference if a class is empty
    dev = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    g = torch.as_tensor(np.asarray(gen)[first], dtype=dtype, device=dev)               # (classes, V, T, C): V "samples" per class
    r = torch.as_tensor(np.asarray(real)[first], dtype=dtype, device=dev)
    grid = MMD(mode).sequence_mmd_grid(g, r, bandwidths)                               # (bandwidths, classes)
    # `if new_new_r > new_r` starting from 0 (:106-110): NaN never wins, negative values cannot occur
    per_class = torch.nan_to_num(grid, nan=0.0).clamp_min(0.0).max(0).values.cpu().numpy()   # the one device->host read
    result = float(np.mean(per_class))
    return (result, per_class) if return_per_class else result


def first_per_class(dataset, classes, per_class=100, t_size=64):
    """The selection loops of mmd-actions.py:136-164: walking the dataset from the start, class after class, the first
    `per_class` items of each class, cropped to `t_size` frames.  -> (actions (n, C, t, V), labels (n,)).
    Quirk kept: when a class is complete the reference resets `i = 0` and then still executes the loop's `i += 1` (:147-149), so
    every class after the first is scanned from dataset item 1 - item 0 can only ever be selected for the first class."""
    labels = np.asarray(dataset.label)
    actions, out_labels = [], []
    for k, c in enumerate(classes):
        idx = np.nonzero(labels == c)[0]
        if k > 0:
            idx = idx[idx >= 1]
        idx = idx[:per_class]
        if len(idx) < per_class:
            raise IndexError("class %d has only %d of the %d samples asked for" % (c, len(idx), per_class))
        for i in idx:
            actions.append(dataset[int(i)][0][:, :t_size, :])
            out_labels.append(int(c))
    return np.asarray(actions), np.asarray(out_labels)


def activation_statistics(features, device=None):
    """fid-actions.py:160-183 after the feature extraction: mu = mean over samples, sigma = np.cov(features, rowvar=False)
    (unbiased), as float64 tensors on the device; the covariance is one GEMM."""
    dev = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    f = torch.as_tensor(np.asarray(features) if not torch.is_tensor(features) else features, dtype=torch.float64, device=dev)
    mu = f.mean(0)
    d = f - mu
    return mu, d.t() @ d / (f.shape[0] - 1)


def frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    """fid-actions.py:106-157: d^2 = |mu1 - mu2|^2 + Tr(S1) + Tr(S2) - 2 Tr(sqrt(S1 S2)), evaluated on the device in float64.
    The reference takes scipy's Schur-based `sqrtm` of the (non-symmetric) product on the host.  Here Tr(sqrt(S1 S2)) is the sum
    of the square roots of the eigenvalues of S1 S2, which are those of the SYMMETRIC matrix S1^(1/2) S2 S1^(1/2): two `eigh`
    calls and two GEMMs, no complex arithmetic, and rank-deficient covariances (fewer samples than the 2048 feature dimensions -
    the case in which the reference falls back to adding `eps` to the diagonals, :139-144) need no special path: eigenvalues that
    come out slightly negative from rounding are clamped to zero.  Accepts numpy arrays or tensors; returns a Python float."""
    dev = sigma1.device if torch.is_tensor(sigma1) else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    t = lambda a: torch.atleast_1d(torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a, dtype=torch.float64, device=dev))
    mu1, mu2 = t(mu1), t(mu2)
    s1, s2 = torch.atleast_2d(t(sigma1)), torch.atleast_2d(t(sigma2))
    assert mu1.shape == mu2.shape, 'Training and test mean vectors have different lengths'
    assert s1.shape == s2.shape, 'Training and test covariances have different dimensions'
    w, q = torch.linalg.eigh((s1 + s1.t()) / 2)
    root1 = (q * w.clamp_min(0).sqrt()) @ q.t()                       # S1^(1/2)
    m = root1 @ s2 @ root1
    tr_covmean = torch.linalg.eigvalsh((m + m.t()) / 2).clamp_min(0).sqrt().sum()
    diff = mu1 - mu2
    (_,) = (eps,)                                                      # kept in the signature for drop-in calls; not needed by this evaluation
    return float(diff.dot(diff) + torch.trace(s1) + torch.trace(s2) - 2 * tr_covmean)


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--data_real", type=str, required=True, help=".npy with the real sequences")
    p.add_argument("--labels_real", type=str, required=True, help=".pkl with the real (names, labels)")
    p.add_argument("--data_fake", type=str, required=True, help="*_gen_data.npy written by generate.py")
    p.add_argument("--labels_fake", type=str, required=True, help="*_gen_label.pkl written by generate.py")
    p.add_argument("--mmd_mode", type=str, default="avg", choices=['avg', 'joint'], help="avg: per-frame dynamics, joint: whole sequence")
    p.add_argument("--t_size", type=int, default=64, help="frames per sequence (T)")
    p.add_argument("--dataset", type=str, default="h36m", help="skeleton layout: ntu or h36m")
    p.add_argument("--out", type=str, default="runs", help="root of the run directories (not in the reference)")
    return p


def evaluate(opt):
    """mmd-actions.py:129-196 as a function: real data normalised to [-1, 1] by the Feeder, generated data taken as is."""
    from .feeder import Feeder

    real_set = Feeder(opt.data_real, opt.labels_real, norm=True, dataset=opt.dataset)
    fake_set = Feeder(opt.data_fake, opt.labels_fake, norm=False, dataset=opt.dataset)
    classes = np.arange(10 if opt.dataset == 'h36m' else 60)
    real, lab_r = first_per_class(real_set, classes, 100, opt.t_size)
    fake, lab_f = first_per_class(fake_set, classes, 100, opt.t_size)
    assert lab_f.tolist() == lab_r.tolist()
    onehot = np.zeros((lab_r.size, len(classes)))
    onehot[np.arange(lab_r.size), lab_r] = 1
    return calculate_mmd(fake.transpose(0, 3, 2, 1), real.transpose(0, 3, 2, 1), onehot, opt.mmd_mode)


def main(argv=None):
    from .train import check_runs

    opt = build_parser().parse_args(argv)
    print(opt)
    out = check_runs('mmd-actions', root=opt.out)
    result = evaluate(opt)
    with open(os.path.join(out, "config.txt"), "w") as f:
        f.write(os.path.basename(__file__) + '|' + str(opt) + '\n' + 'MMD_' + str(opt.mmd_mode) + ': ' + str(result))
    print(result)
    return result


if __name__ == "__main__":
    main()
