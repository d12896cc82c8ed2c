"""MMD evaluation on the device: the batched (class x bandwidth x frame) expression gives the CPU result, for the NTU-60
problem size of evaluation/mmd-actions.py (60 classes, 25 joints, 64 frames, 3 coordinates)."""
from importlib import import_module

import numpy as np
import pytest
import torch

import kgan_b200  # noqa: F401

pytestmark = pytest.mark.gpu
ev = import_module("kinetic-gan_b200.evaluation")


@pytest.mark.parametrize("mode", ["avg", "joint"])
def test_mmd_device_matches_host(mode):
    rng = np.random.RandomState(3)
    n_cls, V, T, C = 60, 25, 64, 3
    gen = rng.uniform(-1, 1, (n_cls * 2, V, T, C)).astype(np.float32)
    real = (0.7 * gen + 0.3 * rng.uniform(-1, 1, gen.shape) + 0.2).astype(np.float32)
    lab = np.eye(n_cls)[np.tile(np.arange(n_cls), 2)]
    dev, dev_pc = ev.calculate_mmd(gen, real, lab, mode, device="cuda", return_per_class=True)
    host, host_pc = ev.calculate_mmd(gen, real, lab, mode, device="cpu", return_per_class=True)
    assert np.isfinite(dev) and dev > 0
    assert np.abs(dev_pc - host_pc).max() < 5e-4 and abs(dev - host) < 2e-4
