"""Parity of the BENCHED path with the oracle: precision 'tf32' (tcgen05 kind::tf32 kernels), BASELINE.json configs[2]
(kinetic-gan-mlp8, NTU-120 shape 25 x 64 x 3, 120 classes), batches large enough that the tensor-core plans - not the
exact SIMT kernels that serve GEMMs below 256 rows - run, eagerly and through `WGANGPTrainer.capture_graphs()` replay.
Ground truth: the fp64 CPU oracle (oracle/networks.py, pinned to the unmodified reference by tests/golden).

Stated tolerances (rel-L2 against fp64), and why:
  * every LAYER (block) output and the network outputs: <= 1e-3 (north_star's bound for a TF32 path).  One tf32 GEMM with
    round-to-nearest operands and fp32 accumulation carries ~3e-4 (two operands at 2^-11/sqrt(3) relative rms each);
    blocks chain 2-3 GEMMs on top of an input that already carries the error of the blocks before it, so the CUMULATIVE
    error at block i grows like 3e-4 * sqrt(#GEMMs so far).  The critic (17 GEMMs deep) stays under 1e-3 through all 6 blocks
    and at its output.  The generator is 8 mapping layers + 7 blocks = 25 GEMMs deep with three training-mode BatchNorms
    (which divide by a batch standard deviation that carries the error too): its block-0 output (10 GEMMs) is held to 1e-3,
    the later blocks and the result to 3e-3 = 1e-3 * sqrt(depth / 3) (measured 1.3e-3 .. 2.1e-3, printed below).  The
    per-LAYER figure on identical inputs (tests/test_tf32_gpu.py) is <= 1e-3 everywhere;
  * gradients (first order, gradient penalty double backward, trainer steps) WITH THE ACTIVATION PATTERN PINNED: <= 3e-3.
    The critic is piecewise linear (LeakyReLU 0.2).  A parameter gradient of the first blocks has been through the forward
    chain, the data-gradient chain and (for the penalty) the double-backward chain: 25-40 tf32 GEMMs in series,
    3e-4 * sqrt(40) ~ 1.9e-3 in the worst case.  That is the arithmetic of the kernels, and it is compared with the fp64 oracle
    evaluated on the SAME activation pattern (oracle `masks=`: the slopes of the CUDA forward pass);
  * the same gradients against the oracle's OWN activation pattern: <= 6e-2, reported as "raw".  Every pre-activation that lies
    within the forward error of zero has a different slope (1 vs 0.2) in the two evaluations; a fraction f of flipped slopes
    costs 0.8 * sqrt(f) in rel-L2 per activation layer whatever the arithmetic.  With a forward error of ~5e-4 and unit-scale
    pre-activations f ~ 5e-4..1e-3: 2-3e-2 per layer, ~4e-2 after 6 blocks (measured below).  The same law governs the fp32 path
    (forward 1e-6 -> gradients 5e-4, tests/test_parity_gpu.py) and any TF32 evaluation of the reference itself
    (bench.py `gpu_reference.tf32_grad_rel_l2` measures it on this GPU); it is the conditioning of LeakyReLU at zero, not an error
    of the kernels;
  * tf32-representable inputs (all-ones cotangents, 0/1 masks, small integers): EXACT (test_tf32_exact_on_representable_data) -
    the operands are rounded to nearest by the producing kernels, never truncated-and-rescaled.
Run with `-s` to see the achieved figures; they are also written to gpurun_out/parity_tf32.txt when that directory exists."""
import os
from importlib import import_module

import numpy as np
import pytest
import torch

import emu_backend as emu
import kgan_b200 as kgan
from oracle import networks as onet
from oracle.graph import SkeletonTables
from helpers import draw_noises, inputs, rel_l2

pytestmark = pytest.mark.gpu
ops = kgan.ops
CFG = onet.Config(dataset="ntu", n_classes=120, t_size=64, mlp_dim=8, channels=3)        # BASELINE.json configs[2]
TOL_OUT, TOL_G, TOL_GRAD, TOL_RAW = 1e-3, 3e-3, 3e-3, 6e-2
_LOG = []


_FAIL = []


def report(name, value, tol):
    """Prints / logs the achieved figure; a figure above `tol` is recorded and fails the test at its end (check()), so one run
    shows every tensor's number."""
    line = "%-58s rel-L2 %.3e  (tol %.0e)%s" % (name, value, tol, "" if value < tol else "   <-- ABOVE TOLERANCE")
    if not value < tol:
        _FAIL.append(line)
    _LOG.append(line)
    print(line)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_tf32.txt"), "a") as f:
            f.write(line + "\n")
    return value


def check():
    bad = list(_FAIL)
    _FAIL.clear()
    assert not bad, "\n".join(bad)


@pytest.fixture(autouse=True)
def tf32_path():
    _FAIL.clear()
    kgan.set_precision("tf32")
    yield
    kgan.set_precision("fp32")
    ops._persist.clear()
    ops._batches.clear()
    ops.clear_temporary_packs()


def build():
    tables = SkeletonTables("ntu")
    pg = onet.synth_params(onet.g_param_shapes(CFG, tables), 1)
    pd = onet.synth_params(onet.d_param_shapes(CFG, tables), 2)
    G = kgan.Generator(CFG.latent_dim, CFG.channels, CFG.n_classes, CFG.t_size, CFG.mlp_dim)
    D = kgan.Discriminator(CFG.channels, CFG.n_classes, CFG.t_size, CFG.latent_dim)
    G.load_state_dict(pg)
    D.load_state_dict(pd)
    return G.cuda(), D.cuda(), pg, pd, tables


def f64(p, grad=False):
    return {k: (v.double().requires_grad_(True) if (grad and v.is_floating_point() and onet.is_trainable(k)) else
                (v.double() if v.is_floating_point() else v)) for k, v in p.items()}


def tensor_core_share(prof):
    fam = ops.profile_stop(prof)
    fam.pop("_sites")
    tc = sum(v["n"] for k, v in fam.items() if k.endswith("_tf32"))
    simt = sum(v["n"] for k, v in fam.items() if k in ("tapconv_fwd", "tapconv_wgrad"))
    return tc, simt


def test_tf32_exact_on_representable_data():
    """Operands that ARE tf32 numbers (small integers, 0/1 masks, all-ones cotangents - the `grad_outputs` of the gradient
    penalty, kinetic-gan.py:101-105) go through the tensor-core kernels exactly: forward, data gradient and weight gradient
    equal the fp64 statement bit for bit (all partial sums are integers below 2^24).  Round 1's truncate-and-rescale epilogue
    (x 1.000353 / x 1.000706) put a deterministic +3.5e-4 / +7e-4 on exactly these inputs."""
    geo = kgan.geometry
    lib = import_module("kinetic-gan_b200._lib").lib()
    gen = torch.Generator().manual_seed(3)
    cases = {"gcn_tma": (dict(c_in=32, c_out=64, t_in=64, v_in=12, K=3), 8), "tcn_tma": (dict(c_in=64, c_out=64, t_in=64, v_in=12, kt=3, pad=1), 8),
             "tcn_gather": (dict(c_in=64, c_out=64, t_in=64, v_in=11, kt=3, pad=1), 8), "linear_kmajor": (dict(c_in=632, c_out=632, t_in=1, v_in=1), 300),
             "small_plane": (dict(c_in=256, c_out=512, t_in=16, v_in=5, K=3), 15)}
    for name, (kw, n) in cases.items():
        g = geo.TapConvGeom(**kw)
        assert lib.kgan_tapconv_tf32_workspace(g.fwd.cstruct(n, 0, 1)) > 0, name          # tensor-core eligible
        x = torch.randint(-3, 4, (n, g.K * g.c_in, g.t_in, g.v_in), generator=gen).float()
        w = torch.randint(-2, 3, (g.K * g.c_out, g.c_in, g.kt, 1), generator=gen).float()
        go = torch.ones(n, g.c_out, g.t_out, g.v_out)                                        # the all-ones cotangent
        mask = (torch.rand(n, g.c_out, g.t_out, g.v_out, generator=gen) > 0.5).float()       # a 0/1 mask
        y = ops.tapconv_fwd(x.cuda(), w.cuda(), g.fwd)
        assert torch.equal(y.cpu().double(), emu.tapconv_fwd(x.double(), w.double(), g.fwd)), name
        gx = ops.tapconv_fwd(go.cuda(), w.cuda(), g.dgrad)
        assert torch.equal(gx.cpu().double(), emu.tapconv_fwd(go.double(), w.double(), g.dgrad)), name
        gx = ops.tapconv_fwd(mask.cuda(), w.cuda(), g.dgrad)
        assert torch.equal(gx.cpu().double(), emu.tapconv_fwd(mask.double(), w.double(), g.dgrad)), name
        for cot in (go, mask):
            dw = ops.tapconv_wgrad(x.cuda(), cot.cuda(), g.fwd, tuple(w.shape))
            assert torch.equal(dw.cpu().double(), emu.tapconv_wgrad(x.double(), cot.double(), g.fwd, tuple(w.shape))), name


def test_rounded_activations_are_read_exactly():
    """What the rounding contract buys: on operands that went through kgan_round_tf32 (what every libkgan kernel stores in tf32
    mode) the tensor-core result equals the fp64 product of THOSE operands up to fp32 accumulation (~1e-6), i.e. the only
    tf32 error of a layer is the rounding of its operands - no truncation, no bias."""
    geo = kgan.geometry
    gen = torch.Generator().manual_seed(5)
    for kw, n in ((dict(c_in=64, c_out=128, t_in=64, v_in=12, K=3), 16), (dict(c_in=128, c_out=128, t_in=32, v_in=12, kt=3, pad=1), 16)):
        g = geo.TapConvGeom(**kw)
        x = ops.round_tf32(torch.randn(n, g.K * g.c_in, g.t_in, g.v_in, generator=gen).cuda())
        w = ops.round_tf32((torch.randn(g.K * g.c_out, g.c_in, g.kt, 1, generator=gen) / (g.c_in * g.kt * g.K) ** 0.5).cuda())
        go = ops.round_tf32(torch.randn(n, g.c_out, g.t_out, g.v_out, generator=gen).cuda())
        assert (x.view(torch.int32) & 0x1FFF).abs().max().item() == 0                       # low 13 mantissa bits are zero
        y = ops.tapconv_fwd(x, w, g.fwd)
        dw = ops.tapconv_wgrad(x, go, g.fwd, tuple(w.shape))
        e_y = rel_l2(y, emu.tapconv_fwd(x.cpu().double(), w.cpu().double(), g.fwd))
        e_w = rel_l2(dw, emu.tapconv_wgrad(x.cpu().double(), go.cpu().double(), g.fwd, tuple(w.shape)))
        report("rounded operands: forward (ck=%d)" % g.c_in, e_y, 3e-4)
        report("rounded operands: weight gradient (ck=%d)" % g.c_in, e_w, 3e-6)
        # the forward output is stored tf32-rounded (2^-11 / sqrt(3) ~ 2.4e-4 relative rms), the weight gradient is not
    check()


def test_generator_bench_config_vs_oracle():
    """G (training mode, BatchNorm batch statistics) at batch 256: every block output and the result against fp64."""
    n = 256
    G, _, pg, _, tables = build()
    x = inputs(CFG, n, 31)
    noises = draw_noises(CFG, n, 32)
    G.train()
    blocks = []
    hooks = [m.register_forward_hook(lambda m, i, o: blocks.append(o[0].detach())) for m in G.st_gcn_networks]
    prof = ops.profile_start()
    fake = G(x["z"].cuda(), x["labels"].cuda(), noises=[t.cuda() for t in noises])
    tc, simt = tensor_core_share(prof)
    for h in hooks:
        h.remove()
    ref_blocks = []
    with torch.no_grad():
        ref = onet.generator_forward(f64(pg), x["z"].double(), x["labels"], CFG, tables, [t.double() for t in noises], True, {},
                                     collect=ref_blocks)
    print("generator forward: %d tensor-core launches, %d exact-SIMT tap convolutions" % (tc, simt))
    assert tc >= 8 + 2 * 5                  # the mapping network and the blocks down to 32 channels run on the tensor cores
    for i, (b, r) in enumerate(zip(blocks, ref_blocks)):
        report("G block %d output (N=%d)" % (i, n), rel_l2(b, r), TOL_OUT if i == 0 else TOL_G)
    report("G output (N=%d)" % n, rel_l2(fake, ref), TOL_G)
    G.load_state_dict(pg)                   # the training-mode pass above updated the BatchNorm running statistics
    G.eval()
    ev_noises = draw_noises(CFG, n, 33)
    with torch.no_grad():
        ev = G(x["z"].cuda(), x["labels"].cuda(), noises=[t.cuda() for t in ev_noises])
        ref = onet.generator_forward(f64(pg), x["z"].double(), x["labels"], CFG, tables, [t.double() for t in ev_noises], False, {})
    report("G output, eval mode (N=%d)" % n, rel_l2(ev, ref), TOL_G)
    check()


def slopes(blocks, refs):
    """LeakyReLU slopes of the CUDA forward pass (block outputs keep the sign of their pre-activation), cut to the oracle's joints
    (the critic carries dummy joints to keep planes 16-byte aligned)."""
    return [torch.where(b[..., :r.shape[-1]].detach().cpu() > 0, 1.0, 0.2).double() for b, r in zip(blocks, refs)]


def test_critic_bench_config_vs_oracle():
    """D at batch 64 (every GEMM of every block >= 256 rows): block outputs, output, first-order gradients, gradient penalty and
    its double-backward parameter gradients against fp64."""
    n = 64
    wg = import_module("kinetic-gan_b200.wgan_gp")
    _, D, _, pd, tables = build()
    x = inputs(CFG, n, 41)
    pd64 = f64(pd, grad=True)
    keys = list(pd64)
    xr = x["real"].cuda().requires_grad_(True)
    blocks = []
    hooks = [m.register_forward_hook(lambda m, i, o: blocks.append(o[0].detach())) for m in D.st_gcn_networks]
    prof = ops.profile_start()
    dv = D(xr, x["labels"].cuda())
    tc, simt = tensor_core_share(prof)
    print("critic forward: %d tensor-core launches, %d exact-SIMT tap convolutions" % (tc, simt))
    assert tc >= 10                          # D1..D5: gcn + tcn (the residual convs ride in the tcn launches) on the tensor cores; D0 / head: SIMT
    ref_blocks = []
    xr64 = x["real"].double().requires_grad_(True)
    dv_ref = onet.discriminator_forward(pd64, xr64, x["labels"], CFG, tables, collect=ref_blocks)
    for i, (b, r) in enumerate(zip(blocks, ref_blocks)):
        r = r.detach()
        report("D block %d output (N=%d)" % (i, n), rel_l2(b[..., :r.shape[-1]], r), TOL_OUT)
    report("D output (N=%d)" % n, rel_l2(dv, dv_ref.detach()), TOL_OUT)
    # first-order gradients of a random-cotangent loss: same activation pattern (arithmetic), then the oracle's own pattern (raw)
    (dv * x["cot_d"].cuda()).sum().backward()
    dv_pin = onet.discriminator_forward(pd64, xr64, x["labels"], CFG, tables, masks=slopes(blocks, ref_blocks))
    for tag, out, tol in (("", dv_pin, TOL_GRAD), (" [raw]", dv_ref, TOL_RAW)):
        gref = torch.autograd.grad((out * x["cot_d"].double()).sum(), [xr64] + [pd64[k] for k in keys], allow_unused=True)
        report("D grad wrt input" + tag, rel_l2(xr.grad, gref[0]), tol)
        gpar = dict(zip(keys, gref[1:]))
        for k, p in D.named_parameters():
            report("D first-order grad%s %s" % (tag, k), rel_l2(p.grad, gpar[k]), tol)
    # gradient penalty (kinetic-gan.py:94-114): value, gradients w.r.t. the interpolates, double-backward parameter gradients
    D.zero_grad(set_to_none=True)
    blocks.clear()
    fake = torch.tanh(torch.randn(n, 3, 64, 25, generator=torch.Generator().manual_seed(43)))
    gp, grads = wg.compute_gradient_penalty(D, x["real"].cuda(), fake.cuda(), x["labels"].cuda(), alpha=x["alpha"].cuda(), return_gradients=True)
    gp.backward()
    for h in hooks:
        h.remove()
    for tag, masks, tol_v, tol_g in (("", slopes(blocks, ref_blocks), TOL_OUT, TOL_GRAD), (" [raw]", None, 2e-3, TOL_RAW)):
        pd64 = f64(pd, grad=True)
        gp_ref, grads_ref = onet.gradient_penalty(pd64, x["real"].double(), fake.double(), x["labels"], x["alpha"].double(), CFG, tables,
                                                  return_grad=True, masks=masks)
        report("GP value" + tag, abs(gp.item() - gp_ref.item()) / abs(gp_ref.item()), tol_v)
        report("GP gradients wrt interpolates" + tag, rel_l2(grads, grads_ref.detach()), tol_g)
        gref = dict(zip(keys, torch.autograd.grad(gp_ref, [pd64[k] for k in keys], allow_unused=True)))
        for k, p in D.named_parameters():
            if gref[k] is None or gref[k].abs().max().item() == 0:
                assert p.grad is None or p.grad.abs().max().item() == 0, k                    # exact zeros: biases, label_emb, dead partitions
                continue
            report("GP double-backward grad%s %s" % (tag, k), rel_l2(p.grad, gref[k]), tol_g)
    check()


def test_trainer_graph_replay_bench_config_vs_oracle():
    """Three iterations (i = 1, 2: critic updates; i = 5: critic + generator update) of kinetic-gan.py:137-174 through
    WGANGPTrainer.capture_graphs() REPLAY at batch 64, tf32.  Per step: losses against the fp64 oracle evaluated at the same
    parameters (<= 1e-3); the flat gradient buffer (a) against the SAME step launched eagerly - identical arithmetic and activation
    pattern, only the order of fp32 atomics differs: <= 1e-5, which ties the replayed path to the eager one that
    test_critic_bench_config_vs_oracle checks tensor by tensor - and (b) raw against the oracle (own activation pattern, see the
    module docstring: <= 6e-2); the fused Adam update against an fp64 Adam on the same gradients.
    The generator's noise weights are zero (their init, generator.py:16), so the device-drawn noise does not enter the values."""
    n = 64
    wg = import_module("kinetic-gan_b200.wgan_gp")
    G, D, pg, pd, tables = build()
    for k in pg:
        if k.endswith("noise.weight"):
            pg[k] = torch.zeros_like(pg[k])
    G.load_state_dict(pg)
    G.train()
    tr = wg.WGANGPTrainer(G, D, CFG.lr, CFG.b1, CFG.b2, CFG.n_critic, CFG.lambda_gp)
    x0 = {k: v.cuda() for k, v in inputs(CFG, n, 50).items()}
    tr.capture_graphs(x0["real"], x0["labels"], x0["z"], x0["alpha"])
    zeros = [torch.zeros(*s, dtype=torch.float64) for s in onet.noise_shapes(CFG, n, tables)]
    m_d, v_d = torch.zeros_like(tr.fd.flat, dtype=torch.float64), torch.zeros_like(tr.fd.flat, dtype=torch.float64)

    def flat_rel(a, b):
        return ((a.double() - b.double()).norm() / b.double().norm()).item()

    def flat_vs_oracle(module, gref):
        num = den = 0.0
        for k, p in module.named_parameters():
            if k not in gref:
                continue
            g = torch.zeros_like(p, dtype=torch.float64).cpu() if gref[k] is None else gref[k]
            num += (p.grad.detach().cpu().double() - g).pow(2).sum().item()
            den += g.pow(2).sum().item()
        return (num / den) ** 0.5

    for step, i in enumerate((1, 2, 5), start=1):
        xi = inputs(CFG, n, 50 + i)
        xc = {k: v.cuda() for k, v in xi.items()}
        # oracle at the product's CURRENT parameters (BatchNorm running statistics are irrelevant in training mode)
        pg64 = f64({k: v.detach().cpu() for k, v in G.state_dict().items()}, grad=True)
        pd64 = f64({k: v.detach().cpu() for k, v in D.state_dict().items()}, grad=True)
        before = tr.fd.flat.double().clone()
        tr._d_grads(xc["real"], xc["labels"], xc["z"], xc["alpha"])                 # the same step, launched eagerly
        eager = tr.fd.grad.clone()
        d_loss, g_loss, gp = tr.iteration(i, xc["real"], xc["labels"], xc["z"], xc["alpha"])
        torch.cuda.synchronize()
        report("iter %d: critic flat gradient, graph replay vs eager" % i, flat_rel(tr.fd.grad, eager), 1e-5)
        d_ref, gp_ref, _ = onet.d_loss_fn(pg64, pd64, xi["real"].double(), xi["labels"], xi["z"].double(), xi["alpha"].double(), zeros, CFG, tables, {})
        report("iter %d (graph replay): d_loss" % i, abs(d_loss.item() - d_ref.item()) / max(1.0, abs(d_ref.item())), TOL_OUT)
        report("iter %d (graph replay): gp" % i, abs(gp.item() - gp_ref.item()) / max(1.0, abs(gp_ref.item())), TOL_OUT)
        keys = [k for k, _ in D.named_parameters()]
        gref = dict(zip(keys, torch.autograd.grad(d_ref, [pd64[k] for k in keys], allow_unused=True)))
        report("iter %d (graph replay): critic flat gradient [raw]" % i, flat_vs_oracle(D, gref), TOL_RAW)
        # fused Adam (kgan_adam_step) on the flat buffer vs fp64 Adam on the same gradient
        g = tr.fd.grad.double()
        m_d = CFG.b1 * m_d + (1 - CFG.b1) * g
        v_d = CFG.b2 * v_d + (1 - CFG.b2) * g * g
        want = before - CFG.lr / (1 - CFG.b1 ** step) * m_d / ((v_d / (1 - CFG.b2 ** step)).sqrt() + 1e-8)
        assert (tr.fd.flat.double() - want).abs().max().item() < 1e-6
        if i % CFG.n_critic == 0:
            # the generator update ran AFTER the critic's Adam step: oracle with the updated critic
            pd64 = f64({k: v.detach().cpu() for k, v in D.state_dict().items()})
            g_ref = onet.g_loss_fn(pg64, pd64, xi["labels"], xi["z"].double(), zeros, CFG, tables, {})
            report("iter %d (graph replay): g_loss" % i, abs(g_loss.item() - g_ref.item()) / max(1.0, abs(g_ref.item())), TOL_OUT)
            kg = [k for k, _ in G.named_parameters() if not k.endswith("noise.weight")]
            gref = dict(zip(kg, torch.autograd.grad(g_ref, [pg64[k] for k in kg], allow_unused=True)))
            report("iter %d (graph replay): generator flat gradient [raw]" % i, flat_vs_oracle(G, gref), TOL_RAW)
        else:
            assert g_loss is None
    check()
