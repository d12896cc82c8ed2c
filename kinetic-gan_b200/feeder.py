"""Input pipeline on the data side of the hot path.

`Feeder` keeps the reference's dataset surface (feeder/feeder.py:21-80: same constructor, `__len__`, `__getitem__`
returning one normalised `(C, T, V)` array and its label), so existing DataLoader code keeps working.

`BatchStream` is what the B200 trainer consumes instead of `DataLoader(Feeder, shuffle=True, drop_last=True)`
(kinetic-gan.py:68-74) followed by the crop / cast / copy at kinetic-gan.py:129-131:

  * one worker thread assembles WHOLE batches (a fancy-index gather out of the memory-mapped `.npy`, person 0 only for
    NTU, frames cropped to `t_size` BEFORE normalising, so only the bytes that reach the GPU are touched) straight into
    pinned staging buffers, `depth` batches ahead; the reference collates N per-sample arrays in worker processes;
  * the host->device copies are asynchronous on a side stream; the consumer's stream waits on an event, never the host;
  * the epoch permutation follows torch's RandomSampler recipe (a seed drawn from the global torch RNG, then
    `torch.randperm`), so under the same `torch.manual_seed` the sample order equals the reference DataLoader's;
  * under data parallelism every rank walks the SAME permutation and takes its own `batch_size` slice of each global
    batch of `batch_size * world` samples (weak scaling; the permutation seed is rank 0's).

Values are identical to the reference's: `2 * ((x - min) / (max - min)) - 1` in the array's dtype with the GLOBAL
min / max of the file (feeder.py:57,76), evaluated in the same operation order.
"""
import pickle
import queue
import threading

import numpy as np
import torch


class Feeder(torch.utils.data.Dataset):
    """Feeder for skeleton-based action synthesis (feeder/feeder.py:21-80).
    data_path: '.npy' of shape (N, C, T, V, M) for NTU and (N, C, T, V) for h36m; label_path: pickle of (names, labels)."""

    def __init__(self, data_path, label_path, classes=None, norm=True, dataset='ntu', mmap=True):
        self.data_path, self.label_path = data_path, label_path
        self.classes, self.norm, self.dataset = classes, norm, dataset
        self.load_data(mmap)

    def load_data(self, mmap):
        with open(self.label_path, 'rb') as f:
            self.sample_name, self.label = pickle.load(f)
        self.label = np.array(self.label, dtype=int)
        self.data = np.load(self.data_path, mmap_mode='r') if mmap else np.load(self.data_path)
        self.max, self.min = self.data.max(), self.data.min()            # global range of the whole file (feeder.py:57)
        if self.classes is not None:                                       # feeder.py:59-62
            keep = np.where(np.isin(self.label, self.classes))
            tmp = self.label[keep]
            self.data = self.data[keep]
            self.label = np.nonzero(tmp[:, None] == self.classes)[1]
        if self.dataset == 'ntu':
            self.N, self.C, self.T, self.V, self.M = self.data.shape
        else:
            self.N, self.C, self.T, self.V = self.data.shape

    def __len__(self):
        return len(self.label)

    def normalise(self, a):
        return 2 * ((a - self.min) / (self.max - self.min)) - 1 if self.norm else a

    def __getitem__(self, index):
        a = np.array(self.data[index, :, :, :, 0]) if self.dataset == 'ntu' else np.array(self.data[index])
        return self.normalise(a), self.label[index]

    def batch(self, indices, t_size=None, out=None, labels_out=None):
        """`indices` -> ((B, C, t, V) float32, (B,) int64): the collated, cropped (kinetic-gan.py:129) and cast (:130)
        batch the reference builds from B `__getitem__` calls, assembled with one gather.  mmap reads want sorted indices;
        the rows are put back in the requested order."""
        idx = np.asarray(indices, dtype=np.int64)
        order = np.argsort(idx, kind="stable")
        t = self.T if t_size is None else min(t_size, self.T)
        if self.dataset == 'ntu':
            raw = self.data[idx[order], :, :t, :, 0]
        else:
            raw = self.data[idx[order], :, :t, :]
        vals = self.normalise(np.asarray(raw))
        if out is None:
            out = np.empty((len(idx), self.C, t, self.V), np.float32)
        out[order] = vals                                                  # cast to float32 here (kinetic-gan.py:130 `.type(Tensor)`)
        lab = self.label[idx]
        if labels_out is None:
            return out, lab.astype(np.int64)
        labels_out[:] = lab
        return out, labels_out


def epoch_permutation(n, seed_generator=None):
    """The order a DataLoader with a RandomSampler yields for one epoch: the loader iterator first draws its worker
    base seed from the global torch RNG (torch/utils/data/dataloader.py, `_BaseDataLoaderIter.__init__`; unused here but it
    advances the RNG), then the sampler seeds a fresh generator with another int64 draw and takes `torch.randperm(n)`
    (torch/utils/data/sampler.py)."""
    torch.empty((), dtype=torch.int64).random_(generator=seed_generator)
    seed = int(torch.empty((), dtype=torch.int64).random_(generator=seed_generator).item())
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g).numpy()


class BatchStream:
    """Iterable over the (real, labels) batches of ONE epoch per `iter()`, resident on `device`.

        stream = BatchStream(Feeder(...), batch_size, t_size, device, rank, world)
        for real, labels in stream: ...            # len(stream) batches, drop_last semantics

    With `device` a CUDA device the tensors live in a ring of `depth` device buffers filled by asynchronous copies from
    pinned memory; a yielded pair is valid for the work the consumer enqueues before asking for the next batch (the ring
    slot is overwritten, stream-ordered after that work, `depth` batches later)."""

    def __init__(self, feeder, batch_size, t_size=None, device="cpu", rank=0, world=1, depth=3, shuffle=True, comm=None):
        self.feeder, self.batch_size, self.rank, self.world = feeder, batch_size, rank, world
        self.t_size = feeder.T if t_size is None else min(t_size, feeder.T)
        self.device = torch.device(device)
        self.depth, self.shuffle, self.comm = max(2, depth), shuffle, comm
        self.cuda = self.device.type == "cuda"
        shape = (batch_size, feeder.C, self.t_size, feeder.V)
        self._host = [(torch.empty(shape, dtype=torch.float32), torch.empty(batch_size, dtype=torch.int64)) for _ in range(self.depth)]
        if self.cuda:
            self._host = [(a.pin_memory(), b.pin_memory()) for a, b in self._host]
            self._dev = [(torch.empty(shape, dtype=torch.float32, device=self.device), torch.empty(batch_size, dtype=torch.int64, device=self.device))
                         for _ in range(self.depth)]
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._copied = [None] * self.depth           # H2D-done events (consumer waits on them)
            self._released = [None] * self.depth         # consumer-side events: slot may be overwritten after this

    def __len__(self):
        return len(self.feeder) // (self.batch_size * self.world)         # drop_last over the GLOBAL batch

    def _permutation(self):
        n = len(self.feeder)
        if not self.shuffle:
            return np.arange(n)
        perm = torch.from_numpy(epoch_permutation(n))
        if self.comm is not None and self.world > 1:                       # every rank walks rank 0's permutation
            import torch.distributed as dist

            t = perm.to(self.device) if self.cuda else perm
            dist.broadcast(t, src=0)
            perm = t.cpu()
        return perm.numpy()

    def _worker(self, perm, q, free):
        try:
            gb = self.batch_size * self.world
            for b in range(len(self)):
                slot = free.get()
                if slot is None:
                    return
                lo = b * gb + self.rank * self.batch_size
                x, y = self._host[slot]
                self.feeder.batch(perm[lo:lo + self.batch_size], self.t_size, out=x.numpy(), labels_out=y.numpy())
                q.put(slot)
            q.put(None)
        except BaseException as e:                                         # surface worker failures in the consumer
            q.put(e)

    def __iter__(self):
        perm = self._permutation()
        q, free = queue.Queue(), queue.Queue()
        for s in range(self.depth):
            free.put(s)
        worker = threading.Thread(target=self._worker, args=(perm, q, free), daemon=True)
        worker.start()
        try:
            while True:
                slot = q.get()
                if slot is None:
                    break
                if isinstance(slot, BaseException):
                    raise slot
                if self.cuda:
                    cur = torch.cuda.current_stream(self.device)
                    with torch.cuda.stream(self._copy_stream):
                        if self._released[slot] is not None:               # the consumer's kernels that read this slot are done
                            self._copy_stream.wait_event(self._released[slot])
                        self._dev[slot][0].copy_(self._host[slot][0], non_blocking=True)
                        self._dev[slot][1].copy_(self._host[slot][1], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(self._copy_stream)
                    cur.wait_event(ev)
                    self._copied[slot] = ev
                    out = self._dev[slot]
                else:
                    out = (self._host[slot][0].clone(), self._host[slot][1].clone())
                yield out
                if self.cuda:
                    rel = torch.cuda.Event()
                    rel.record(torch.cuda.current_stream(self.device))
                    self._released[slot] = rel
                    # the pinned buffer may be refilled once its H2D copy has completed
                    self._copied[slot].synchronize()
                free.put(slot)
        finally:
            free.put(None)
