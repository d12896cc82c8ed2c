"""WGAN-GP training step with the reference's semantics (kinetic-gan.py:94-114 gradient penalty,
:137-174 loop body, Adam at :77-78), restructured for B200:

  * parameters, gradients and Adam moments of each network live in ONE flat fp32 buffer, so the optimizer is a
    single fused launch (kgan_adam_step) and the DDP gradient exchange is a single all-reduce (ddp.py);
  * the critic step runs G under no_grad and the generator step skips the critic's weight gradients: both are
    zeroed before any optimizer reads them in the reference (kinetic-gan.py:137,157), so results are identical
    (SURVEY.md §7 I6);
  * no per-iteration host synchronisation: losses stay on the device unless asked for.
"""
import contextlib

import numpy as np
import torch

from . import functional, ops


def compute_gradient_penalty(D, real_samples, fake_samples, labels, alpha=None, return_gradients=False):
    """Calculates the gradient penalty loss for WGAN GP (kinetic-gan.py:94-114).  `alpha` (N,1,1,1) defaults to the
    reference's host draw np.random.random at :97."""
    n = real_samples.size(0)
    if alpha is None:
        alpha = torch.as_tensor(np.random.random((n, 1, 1, 1)), dtype=real_samples.dtype, device=real_samples.device)
    interpolates = ops.interpolate(alpha.reshape(n).contiguous(), real_samples.contiguous(),
                                   fake_samples.contiguous()).requires_grad_(True)
    # the forward nodes of this pass receive no weight gradients from the step (they hang off the second-order graph only through LeakyReLU
    # slopes): its graph convs may run as single kernels with the adjacency product inside the GEMM (functional.no_weight_grads_expected)
    with functional.no_weight_grads_expected():
        d_interpolates = D(interpolates, labels)
    fake = torch.ones(n, 1, dtype=real_samples.dtype, device=real_samples.device)
    # only d/d(interpolates) is asked for: skip the weight / bias / adjacency gradients every node would otherwise compute
    # for the engine to drop (functional.data_grads_only); the graph built here still depends on all parameters
    with functional.data_grads_only():
        gradients = torch.autograd.grad(outputs=d_interpolates, inputs=interpolates, grad_outputs=fake, create_graph=True,
                                        retain_graph=True, only_inputs=True)[0]
    flat = gradients.reshape(n, -1)
    gradient_penalty = ((flat.norm(2, dim=1) - 1) ** 2).mean()
    return (gradient_penalty, gradients) if return_gradients else gradient_penalty


class FlatParams:
    """All parameters of a module re-seated as views of ONE flat fp32 buffer (plus flat grad / Adam moments).
    Offsets are 128-byte aligned so that vectorised and bulk (TMA) loads of any weight are legal."""

    ALIGN = 32

    def __init__(self, module):
        self.params = [p for p in module.parameters()]
        dev = self.params[0].device
        offs, off = [], 0
        for p in self.params:
            offs.append(off)
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.flat = torch.zeros(off, device=dev, dtype=torch.float32)
        self.grad = torch.zeros_like(self.flat)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.numel = sum(p.numel() for p in self.params)
        for p, o in zip(self.params, offs):
            n = p.numel()
            self.flat[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat[o:o + n].view_as(p)
            p.grad = self.grad[o:o + n].view_as(p)
        self.step = 0
        ops.register_flat(self.flat)             # packed tf32 weight images follow THIS buffer's optimizer steps

    def zero_grad(self):
        self.grad.zero_()
        for p in self.params:               # autograd accumulates in place into these views
            if p.grad is None or p.grad.data_ptr() < self.grad.data_ptr():
                raise RuntimeError("parameter gradient was re-seated outside the flat buffer")

    def adam(self, lr, b1, b2, eps=1e-8, grad_scale=1.0):
        self.step += 1
        ops.adam_step(self.flat, self.grad, self.m, self.v, lr, b1, b2, eps, self.step, grad_scale)
        ops.invalidate_packed_weights(self.flat)   # tf32 path: this network's packed weight images are stale now ...
        ops.repack_weights(self.flat)              # ... and those in use are refreshed by one batched launch


class WGANGPTrainer:
    """The loop body kinetic-gan.py:137-174 as two methods.  `iteration(i, ...)` = critic step every iteration,
    generator step when i % n_critic == 0 (:160).  Optional `comm` (ddp.Comm) sums the flat gradients across ranks."""

    def __init__(self, generator, discriminator, lr=0.0002, b1=0.5, b2=0.999, n_critic=5, lambda_gp=10, comm=None):
        self.G, self.D = generator, discriminator
        self.lr, self.b1, self.b2, self.n_critic, self.lambda_gp = lr, b1, b2, n_critic, lambda_gp
        self.comm = comm
        if comm is not None:
            comm.broadcast_module(self.G)
            comm.broadcast_module(self.D)
        self.fg, self.fd = FlatParams(self.G), FlatParams(self.D)
        self.world = 1 if comm is None else comm.world_size
        self._graphs = None
        self._copy = None                         # prefetch(): copy stream, two staging slots, the event of the copy in flight
        self.use_graphs = True      # set False to run eagerly even after capture_graphs() (per-kernel profiling)

    def _reduce(self, flat):
        if self.comm is not None and self.world > 1:
            self.comm.all_reduce_sum_(flat.grad)

    # ---- forward + backward halves (everything that can be captured into a CUDA graph) ----------------------------
    def _fake(self, labels, z, noises=None):
        """kinetic-gan.py:143 for the critic update: G's autograd graph is never used there (its grads are zeroed at :157)."""
        with torch.no_grad():
            return self.G(z, labels, noises=noises)

    def _d_grads(self, real, labels, z, alpha=None, noises=None, fake=None):
        """kinetic-gan.py:137-154: fills the critic's flat gradient buffer, returns (d_loss, gp).  `fake`: the generator's output
        for (z, labels) when it has been computed already (the CUDA-graph path runs that pass as a graph of its own)."""
        self.fd.zero_grad()
        if fake is None:
            fake = self._fake(labels, z, noises)
        # the critic has no BatchNorm: every sample is processed independently, so the real and the fake pass
        # (kinetic-gan.py:146,148) run as ONE pass over the concatenated batch - same values, half the launches
        n = real.size(0)
        share = self.D.share_adjacency() if hasattr(self.D, "share_adjacency") else contextlib.nullcontext()
        with share:                                  # both critic calls use the same A * edge_importance tensors (one backward sweep)
            validity = self.D(torch.cat((real, fake), 0), torch.cat((labels, labels), 0))
            real_validity, fake_validity = validity[:n], validity[n:]
            gp = compute_gradient_penalty(self.D, real, fake, labels, alpha)
        d_loss = -torch.mean(real_validity) + torch.mean(fake_validity) + self.lambda_gp * gp
        with functional.param_grads_in_place():      # weight-gradient kernels add straight into the flat gradient views
            d_loss.backward()
        return d_loss.detach(), gp.detach()

    def _g_grads(self, labels, z, noises=None):
        """kinetic-gan.py:167-173; the critic's weight gradients are not needed (zeroed at :137 before use)."""
        self.fg.zero_grad()
        for p in self.fd.params:
            p.requires_grad_(False)
        try:
            fake = self.G(z, labels, noises=noises)
            g_loss = -torch.mean(self.D(fake, labels))
            with functional.param_grads_in_place():
                g_loss.backward()
        finally:
            for p in self.fd.params:
                p.requires_grad_(True)
        return g_loss.detach()

    # ---- CUDA graphs: at the reference batch size the step is launch-bound ------------------
    def capture_graphs(self, real, labels, z, alpha):
        """Captures the step into THREE CUDA graphs fed from static input buffers (shapes taken from the arguments): the
        generator pass of the critic update (kinetic-gan.py:143), the critic's forward + backward (:146-154), and the generator
        update's forward + backward (:167-173).  The gradient all-reduce, the fused Adam step and the refresh of the packed weight
        images run on a SIDE stream after the graph that filled the gradients; the next critic update's generator pass - which
        depends on none of them - is already running on the main stream meanwhile (`d_step`), so the collective is off the
        critical path (overlap the exchange; round 1 ran it serially between graph replay and Adam).
        Afterwards `iteration()` copies its inputs into the static buffers and replays."""
        if self._graphs is not None:
            return
        st = {k: torch.empty_like(v) for k, v in dict(real=real, labels=labels, z=z, alpha=alpha).items()}
        for k, v in dict(real=real, labels=labels, z=z, alpha=alpha).items():
            st[k].copy_(v)
        # the warm-up and capture passes run G in training mode: keep its BatchNorm running statistics as they were
        buffers = [(b, b.detach().clone()) for m in (self.G, self.D) for b in m.buffers()]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):       # warm-up on a side stream: builds every cached table / descriptor / workspace
            for _ in range(2):
                self._d_grads(st["real"], st["labels"], st["z"], st["alpha"])
                self._g_grads(st["labels"], st["z"])
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # parameters: their packed tf32 images are persistent buffers refreshed after every optimizer step (ops.repack_weights),
        # so the graphs only read them.  Images of temporaries (second-order terms) must be packed INSIDE the graphs: drop
        # whatever the eager warm-up cached so that no capture-time lookup hits a buffer no replay would refresh.
        ops.clear_temporary_packs()
        gf, gd, gg = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        l0 = ops.launches
        with torch.cuda.graph(gf):
            fake = self._fake(st["labels"], st["z"])                 # stays allocated in the shared pool: the critic graph reads it
        lf = ops.launches
        ops.clear_temporary_packs()
        with torch.cuda.graph(gd, pool=gf.pool()):
            d_out = self._d_grads(st["real"], st["labels"], st["z"], st["alpha"], fake=fake)
        l1 = ops.launches
        ops.clear_temporary_packs()
        with torch.cuda.graph(gg, pool=gf.pool()):
            g_out = self._g_grads(st["labels"], st["z"])
        l2 = ops.launches
        ops.clear_temporary_packs()
        with torch.no_grad():
            for b, saved in buffers:
                b.copy_(saved)
        self._graphs = dict(static=st, f=gf, d=gd, g=gg, fake=fake, d_out=d_out, g_out=g_out, f_launches=lf - l0, d_launches=l1 - lf,
                            g_launches=l2 - l1, side=torch.cuda.Stream(), d_done=None, g_done=None)

    def _update_on_side_stream(self, flat, key):
        """all-reduce + fused Adam + packed-image refresh of one network on the side stream, after everything queued on the main
        stream so far; `key` ('d_done' / 'g_done') names the event later consumers of those weights wait for."""
        g = self._graphs
        main, side = torch.cuda.current_stream(), g["side"]
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self._reduce(flat)
            flat.adam(self.lr, self.b1, self.b2, grad_scale=1.0 / self.world)
            ev = torch.cuda.Event()
            ev.record(side)
        g[key] = ev

    def synchronize_updates(self):
        """Makes the main stream wait for the optimizer updates still running on the side stream (call before reading parameters
        outside `iteration()`: checkpoints, evaluation, eager passes) and for a prefetch() still in flight."""
        g = self._graphs
        if g is not None:
            for key in ("d_done", "g_done"):
                if g[key] is not None:
                    torch.cuda.current_stream().wait_event(g[key])
        if self._copy is not None:
            for ev in self._copy["events"]:
                if ev is not None:
                    torch.cuda.current_stream().wait_event(ev)

    def prefetch(self, real, labels, z, alpha=None):
        """Starts the host -> device copy of the NEXT iteration's inputs (pinned host tensors) on a copy stream, so that it overlaps the
        iteration in flight, and returns the device tensors to hand to `iteration()` / `d_step()` (which wait for that copy).  Two staging
        slots, each with its own completion event: a slot is rewritten only after the iteration that read it has been enqueued."""
        c = self._copy
        if c is None:
            c = self._copy = {"stream": torch.cuda.Stream(), "slots": [None, None], "events": [None, None], "ids": [(), ()], "i": 0}
        c["i"] ^= 1
        i = c["i"]
        src = {"real": real, "labels": labels, "z": z, "alpha": alpha}
        dev = self.fd.flat.device
        slot = c["slots"][i]
        if slot is None or any(v is not None and (k not in slot or slot[k].shape != v.shape or slot[k].dtype != v.dtype) for k, v in src.items()):
            slot = c["slots"][i] = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in src.items() if v is not None}
        cs = c["stream"]
        cs.wait_stream(torch.cuda.current_stream())        # the slot's previous reader (two iterations back) is on the main stream
        with torch.cuda.stream(cs):
            for k, v in src.items():
                if v is not None:
                    slot[k].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        c["events"][i], c["ids"][i] = ev, tuple(id(t) for t in slot.values())
        return slot["real"], slot["labels"], slot["z"], slot.get("alpha")

    def _await_prefetch(self, *tensors):
        """Main stream waits for the copy that filled the staging slot these tensors belong to (no-op for other tensors)."""
        c = self._copy
        if c is None:
            return
        mine = {id(t) for t in tensors if t is not None}
        for ev, ids in zip(c["events"], c["ids"]):
            if ev is not None and mine.intersection(ids):
                torch.cuda.current_stream().wait_event(ev)

    def d_step(self, real, labels, z, alpha=None, noises=None):
        """kinetic-gan.py:137-155."""
        self._await_prefetch(real, labels, z, alpha)
        g = self._graphs
        if g is not None and self.use_graphs and noises is None:
            st = g["static"]
            if alpha is None:
                alpha = torch.as_tensor(np.random.random(tuple(st["alpha"].shape)), dtype=torch.float32)
            main = torch.cuda.current_stream()
            for k, v in (("real", real), ("labels", labels), ("z", z), ("alpha", alpha)):
                if v is not st[k]:
                    st[k].copy_(v, non_blocking=True)
            if g["g_done"] is not None:                # the generator pass reads G's weights: its last update must have landed
                main.wait_event(g["g_done"])
            g["f"].replay()                            # overlaps the critic's all-reduce / Adam of the previous iteration
            if g["d_done"] is not None:
                main.wait_event(g["d_done"])
            g["d"].replay()
            ops.launches += g["f_launches"] + g["d_launches"]
            d_loss, gp = g["d_out"]
            self._update_on_side_stream(self.fd, "d_done")
            return d_loss, gp
        if g is not None:
            self.synchronize_updates()
        d_loss, gp = self._d_grads(real, labels, z, alpha, noises)
        self._reduce(self.fd)
        self.fd.adam(self.lr, self.b1, self.b2, grad_scale=1.0 / self.world)
        return d_loss, gp

    def g_step(self, labels, z, noises=None):
        """kinetic-gan.py:167-174.  With CUDA graphs the (labels, z) of the preceding d_step are reused, as in the reference loop."""
        self._await_prefetch(labels, z)
        g = self._graphs
        if g is not None and self.use_graphs and noises is None:
            st = g["static"]
            for k, v in (("labels", labels), ("z", z)):
                if v is not st[k]:
                    st[k].copy_(v, non_blocking=True)
            self.synchronize_updates()                 # the generator update differentiates through the UPDATED critic (kinetic-gan.py:155 -> :170)
            g["g"].replay()
            ops.launches += g["g_launches"]
            # the graphs share one memory pool: the critic graph's temporaries may occupy the bytes of this output, so
            # hand out a copy that survives the next d_step replay
            g_loss = g["g_out"].clone()
            self._update_on_side_stream(self.fg, "g_done")
            return g_loss
        if g is not None:
            self.synchronize_updates()
        g_loss = self._g_grads(labels, z, noises)
        self._reduce(self.fg)
        self.fg.adam(self.lr, self.b1, self.b2, grad_scale=1.0 / self.world)
        return g_loss

    def iteration(self, i, real, labels, z, alpha=None, noises_d=None, noises_g=None):
        d_loss, gp = self.d_step(real, labels, z, alpha, noises_d)
        g_loss = self.g_step(labels, z, noises_g) if i % self.n_critic == 0 else None
        return d_loss, g_loss, gp
