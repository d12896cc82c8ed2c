"""Generates tests/golden/mmd_ref.npz from the UNMODIFIED reference evaluation script (build container only: needs /root/reference).

evaluation/mmd-actions.py is a script (argparse + dataset paths at import, mmd-actions.py:11-12,117-131), so its pieces are
exec'd from the file by line range, source untouched:
    lines 14-76    class MMD                     (both its torch and its numpy branch)
    lines 79-115   calcualte_mmd(gen, real, label)
    lines 134-164  the two selection loops (real / fake), run on a synthetic in-memory dataset
`.cuda()` is the identity here (no GPU in the build container); `opt` is a namespace with the two options the code reads.

    python tests/golden/make_mmd_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/evaluation/mmd-actions.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mmd_ref.npz")


def ref_lines(a, b):
    with open(REF) as f:
        return "".join(f.readlines()[a - 1:b])


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self
    out = {}
    rng = np.random.RandomState(7)
    ns = {"np": np, "torch": torch, "opt": types.SimpleNamespace(mmd_mode="avg", t_size=4, dataset="h36m")}
    exec(ref_lines(14, 76), ns)
    exec(ref_lines(79, 115), ns)
    # --- class MMD: sequences (N, len, dim)
    s1 = rng.uniform(-1, 1, (9, 6, 3)).astype(np.float32)
    s2 = (0.6 * s1 + 0.4 * rng.uniform(-1, 1, s1.shape)).astype(np.float32)
    out["seq_1"], out["seq_2"] = s1, s2
    bws = [10.0 ** j for j in range(-4, 10)]
    out["bandwidths"] = np.array(bws)
    for mode in ("avg", "joint"):
        m_np, m_t = ns["MMD"](mode, 0), ns["MMD"](mode, 1)
        with np.errstate(invalid="ignore"):
            out["seq_mmd_numpy_" + mode] = np.array([m_np.compute_sequence_mmd(s1, s2, bw) for bw in bws], dtype=np.float64)
        out["seq_mmd_torch_" + mode] = np.array([m_t.compute_sequence_mmd(torch.tensor(s1), torch.tensor(s2), bw) for bw in bws], dtype=np.float64)
    out["rkhs_numpy"] = np.array([ns["MMD"]("avg", 0).rkhs_mmd(s1[:, 0], s2[:, 0], bw) for bw in bws[3:8]])
    # --- calcualte_mmd (torch branch: use_torch = 1 at :80)
    n_cls, per, V, T, C = 5, 3, 7, 6, 3
    gen = rng.uniform(-1, 1, (n_cls * per, V, T, C)).astype(np.float32)
    real = (gen * 0.5 + rng.uniform(-1, 1, gen.shape) * 0.5).astype(np.float32)
    lab = np.zeros((n_cls * per, n_cls))
    lab[np.arange(n_cls * per), rng.permutation(np.repeat(np.arange(n_cls), per))] = 1
    out["calc_gen"], out["calc_real"], out["calc_label"] = gen, real, lab
    for mode in ("avg", "joint"):
        ns["opt"].mmd_mode = mode
        out["calc_result_" + mode] = np.float64(ns["calcualte_mmd"](gen, real, lab))
    # --- selection loops :136-164 on a synthetic dataset whose item 0 belongs to a class other than the first
    n_items, n_classes = 1400, 10
    labels = rng.randint(0, n_classes, n_items)
    labels[0] = 3
    data = np.zeros((n_items, 2, 6, 3), np.float32)
    data[:, 0, 0, 0] = np.arange(n_items)                       # the item's id travels in its first element
    ds = [(data[i], int(labels[i])) for i in range(n_items)]
    sel = {"np": np, "opt": ns["opt"], "dataset_real": ds, "dataset_fake": ds}
    exec(ref_lines(134, 164), sel)
    out["select_labels"] = labels
    out["select_ids_real"] = np.array([int(a[0, 0, 0]) for a in sel["real_actions_batch"]])
    out["select_label_batch"] = np.array(sel["label_batch"])
    out["select_t_size"] = np.int64(ns["opt"].t_size)
    out["select_shape"] = np.array(np.array(sel["real_actions_batch"]).shape)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    sys.exit(main())
