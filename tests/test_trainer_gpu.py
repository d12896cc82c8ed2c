"""WGAN-GP trainer on the GPU: two iterations of kinetic-gan.py:137-174 through the CUDA kernels against the golden fixtures
of the unmodified reference (fp32 path), CUDA-graph replay against eager launches, and the batched refresh of the packed
tf32 weight images after the fused Adam step."""
from importlib import import_module

import numpy as np
import pytest
import torch

import kgan_b200 as kgan
from oracle import networks as onet
from helpers import CASES, draw_noises, inputs, load_golden, sub

pytestmark = pytest.mark.gpu
ops = kgan.ops


def build(cfg):
    G = kgan.Generator(cfg.latent_dim, cfg.channels, cfg.n_classes, cfg.t_size, cfg.mlp_dim, dataset=cfg.dataset)
    D = kgan.Discriminator(cfg.channels, cfg.n_classes, cfg.t_size, cfg.latent_dim, dataset=cfg.dataset)
    G.load_state_dict(onet.synth_params(onet.g_param_shapes(cfg), 1))
    D.load_state_dict(onet.synth_params(onet.d_param_shapes(cfg), 2))
    return G.cuda().train(), D.cuda()


def loss_ok(mine, gold, key):
    """Within 2e-4 of the reference's fp64 value, or - after an optimizer step, where Adam turns fp32 rounding noise of
    near-zero gradients (BatchNorm over 2-3 samples) into steps of size lr - within twice the reference's OWN fp32-vs-fp64
    distance (ntu_small d_loss1: reference fp32 2.7211 vs fp64 2.6866)."""
    r64, r32 = float(gold["f64/" + key]), float(gold["f32/" + key])
    return abs(mine - r64) < max(2e-4 * max(1.0, abs(r64)), 2.0 * abs(r32 - r64))


@pytest.mark.parametrize("case", list(CASES))
def test_two_iterations_vs_golden_fp32(case):
    """Same check as tests/test_trainer_cpu.py, but every operator is a libkgan.so kernel (fp32 SIMT path): losses and
    the parameters after a critic+generator update and a second critic update."""
    gold, cfg, n = load_golden(case), CASES[case]["cfg"], CASES[case]["n"]
    wg = import_module("kinetic-gan_b200.wgan_gp")
    kgan.set_precision("fp32")
    G, D = build(cfg)
    tr = wg.WGANGPTrainer(G, D, cfg.lr, cfg.b1, cfg.b2, cfg.n_critic, cfg.lambda_gp)
    dev = lambda ts: [t.cuda() for t in ts]
    for i in range(2):
        xi = {k: v.cuda() for k, v in inputs(cfg, n, 10 + i, torch.float32).items()}
        d_loss, g_loss, _ = tr.iteration(i, xi["real"], xi["labels"], xi["z"], xi["alpha"],
                                         dev(draw_noises(cfg, n, 100 + 2 * i)), dev(draw_noises(cfg, n, 101 + 2 * i)))
        assert loss_ok(d_loss.item(), gold, "train/d_loss%d" % i), (i, d_loss.item())
        if i == 0:
            assert loss_ok(g_loss.item(), gold, "train/g_loss0"), g_loss.item()
    # Adam normalises every gradient to a step of ~lr: compare the parameter DELTAS' bulk, not single elements
    for net, m in (("g", G), ("d", D)):
        for k, v in m.state_dict().items():
            ref = gold["f64/train/%s_after/%s" % (net, k)]
            floor = np.abs(gold["f32/train/%s_after/%s" % (net, k)] - ref).mean()     # the reference's own fp32-vs-fp64 distance
            mine = sub(v) if v.numel() > 4096 else v.detach().double().cpu().numpy()
            assert np.abs(mine - ref).max() < 3 * cfg.lr + 1e-6, k          # never further than a few optimizer steps
            # and on average much closer - except where the gradient is pure rounding noise (a conv bias in front of a
            # BatchNorm: analytically zero), which Adam turns into +-lr steps in the reference's fp32 run as well
            if mine.size >= 16:             # a statistical statement: not meaningful for a handful of elements
                assert np.abs(mine - ref).mean() < 0.25 * cfg.lr + 1e-7 + 2.0 * floor, k


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_graph_replay_matches_eager(precision):
    """capture_graphs() + replay == eager launches for the critic update (iterations 1..3: no generator update, and the
    generator's noise weights are zeroed, so the step is deterministic up to the order of fp32 atomics)."""
    cfg, n = CASES["ntu_small"]["cfg"], 8
    wg = import_module("kinetic-gan_b200.wgan_gp")
    kgan.set_precision(precision)
    try:
        flats = []
        for graphs in (False, True):
            G, D = build(cfg)
            with torch.no_grad():               # synth_params draws non-zero noise weights: zero them so that the device-drawn
                for blk in G.st_gcn_networks:   # noise (different random numbers eagerly and under graph capture) has no effect
                    blk.noise.weight.zero_()
            tr = wg.WGANGPTrainer(G, D, cfg.lr, cfg.b1, cfg.b2, cfg.n_critic, cfg.lambda_gp)
            x0 = {k: v.cuda() for k, v in inputs(cfg, n, 20, torch.float32).items()}
            if graphs:
                tr.capture_graphs(x0["real"], x0["labels"], x0["z"], x0["alpha"])
            for i in range(1, 4):
                xi = {k: v.cuda() for k, v in inputs(cfg, n, 20 + i, torch.float32).items()}
                tr.iteration(i, xi["real"], xi["labels"], xi["z"], xi["alpha"])
            torch.cuda.synchronize()
            flats.append(tr.fd.flat.clone())
            if precision == "tf32":
                # every persistent packed image equals a fresh pack of the CURRENT weights
                lib = import_module("kinetic-gan_b200._lib").lib()
                assert len(ops._persist) > 0
                for e in ops._persist.values():
                    fresh = torch.empty_like(e.wp)
                    assert lib.kgan_tapconv_pack_tf32(e.cs, e.w.data_ptr(), fresh.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
                    torch.cuda.synchronize()
                    assert torch.equal(fresh, e.wp)
        a, b = flats
        # three Adam steps of size lr each: identical up to atomics-order noise amplified by Adam's normalisation
        assert (a - b).abs().max().item() < 3 * cfg.lr
        assert (a - b).abs().mean().item() < (2e-2 if precision == "fp32" else 1e-1) * cfg.lr
    finally:
        kgan.set_precision("fp32")
        ops._persist.clear()
        ops._batches.clear()


def test_prefetch_overlapped_inputs_match_direct_inputs():
    """WGANGPTrainer.prefetch (the next iteration's host -> device copy on a copy stream, two staging slots) feeds the captured graphs
    the same numbers as handing pinned host tensors to iteration() directly: first critic loss bit-equal (the forward pass has no
    atomics), parameters after three critic updates equal up to atomics-order noise."""
    cfg, n = CASES["ntu_small"]["cfg"], 8
    wg = import_module("kinetic-gan_b200.wgan_gp")
    kgan.set_precision("fp32")
    try:
        res = []
        for use_prefetch in (False, True):
            G, D = build(cfg)
            with torch.no_grad():
                for blk in G.st_gcn_networks:
                    blk.noise.weight.zero_()
            tr = wg.WGANGPTrainer(G, D, cfg.lr, cfg.b1, cfg.b2, cfg.n_critic, cfg.lambda_gp)
            x0 = {k: v.cuda() for k, v in inputs(cfg, n, 20, torch.float32).items()}
            tr.capture_graphs(x0["real"], x0["labels"], x0["z"], x0["alpha"])
            host = [{k: v.pin_memory() for k, v in inputs(cfg, n, 20 + i, torch.float32).items() if k in ("real", "labels", "z", "alpha")}
                    for i in range(1, 5)]
            losses = []
            nxt = tr.prefetch(**host[0]) if use_prefetch else None
            for i in range(1, 4):
                if use_prefetch:
                    x, nxt = nxt, tr.prefetch(**host[i])          # the copy of iteration i + 1 overlaps iteration i
                    d_loss, _, _ = tr.iteration(i, *x)
                else:
                    h = host[i - 1]
                    d_loss, _, _ = tr.iteration(i, h["real"], h["labels"], h["z"], h["alpha"])
                losses.append(d_loss.item())
            tr.synchronize_updates()
            torch.cuda.synchronize()
            res.append((losses, tr.fd.flat.clone()))
        (l0, f0), (l1, f1) = res
        assert l0[0] == l1[0], (l0, l1)
        assert max(abs(a - b) for a, b in zip(l0, l1)) < 1e-3 * max(1.0, abs(l0[-1]))
        assert (f0 - f1).abs().max().item() < 3 * cfg.lr
    finally:
        ops._persist.clear()
        ops._batches.clear()
