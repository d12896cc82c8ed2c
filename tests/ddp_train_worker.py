"""Worker of tests/test_pipeline_cpu.py::test_train_cli_two_ranks_gloo: the training CLI loop under torch.distributed
(CPU, gloo, C-ABI primitives emulated).  argv: data_path label_path out_dir."""
import os
import sys
from importlib import import_module

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
torch.set_num_threads(2)

import emu_backend  # noqa: E402
import kgan_b200  # noqa: E402,F401


class _MP:
    def setattr(self, obj, name, val):
        setattr(obj, name, val)


emu_backend.install(_MP())
train_mod = import_module("kinetic-gan_b200.train")
feeder_mod = import_module("kinetic-gan_b200.feeder")
ddp = import_module("kinetic-gan_b200.ddp")

seen = []
orig_batch = feeder_mod.Feeder.batch


def recording_batch(self, indices, *a, **k):
    seen.append(np.asarray(indices).copy())
    return orig_batch(self, indices, *a, **k)


feeder_mod.Feeder.batch = recording_batch
dp, lp, out = sys.argv[1:4]
comm = ddp.Comm(backend="gloo")
opt = train_mod.build_parser().parse_args(
    ["--data_path", dp, "--label_path", lp, "--out", os.path.join(out, "runs_w%d" % comm.world_size), "--n_classes", "6", "--t_size", "16",
     "--mlp_dim", "2", "--batch_size", "2", "--n_epochs", "1", "--max_iters", "3", "--sample_interval", "1000", "--checkpoint_interval", "-1",
     "--log_interval", "1", "--seed", str(11 + 100 * comm.rank)])          # different seeds per rank: the permutation must still be rank 0's
run, loss_d, loss_g = train_mod.train(opt, comm)
torch.save({"seen": np.stack(seen[:3]), "loss_d": loss_d, "run": run}, os.path.join(out, "train_w%d_r%d.pt" % (comm.world_size, comm.rank)))
comm.close()
