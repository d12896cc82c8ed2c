#!/usr/bin/env python
"""Headline benchmark: WGAN-GP training throughput (samples/s) of Kinetic-GAN at the NTU 25x64x3 shape.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl kgan|reference] [--batch B] [--precision tf32|fp32|fp32_fma]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one iteration of the training loop body kinetic-gan.py:137-174 on one synthetic batch per GPU: the critic
update (G forward, 3 critic forwards, gradient penalty with double backward, Adam) and, every n_critic=5 iterations,
the generator update.  Workload (BASELINE.json configs[2]/[3]): kinetic-gan-mlp8, NTU-120 shape (25 joints x 64 frames
x 3, 120 classes), random-init weights, synthetic data; weak scaling (fixed per-GPU batch).

One JSON line is printed by rank 0; see DESIGN.md "Measurement" for every key.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic FLOPs per train sample, NTU-120 mlp8: 1.6 F_G + 10.4 F_D (SURVEY.md §8d / BASELINE.md §3)
# BASELINE.json configs: forward FLOPs per sample of G and D from SURVEY.md §6/§8d
SHAPES = {
    "ntu120": dict(dataset="ntu", n_classes=120, t_size=64, mlp_dim=8, channels=3, joints=25, f_g=32.24e6, f_d=550.62e6,
                   name="kinetic-gan-mlp8 NTU-120 xsub shape (25x64x3, 120 classes)"),
    "ntu60": dict(dataset="ntu", n_classes=60, t_size=64, mlp_dim=4, channels=3, joints=25, f_g=28.28e6, f_d=532.19e6,
                  name="kinetic-gan-mlp4 NTU-60 xsub shape (25x64x3, 60 classes)"),
    "h36m": dict(dataset="h36m", n_classes=10, t_size=32, mlp_dim=4, channels=2, joints=16, f_g=11.95e6, f_d=129.29e6,
                 name="kinetic-gan-mlp4 Human3.6M shape (16x32x2, 10 classes)"),
}
SHAPE = SHAPES["ntu120"]          # the headline configuration; --shape selects another (set_shape)
F_G, F_D = SHAPE["f_g"], SHAPE["f_d"]
FLOP_PER_SAMPLE = 1.6 * F_G + 10.4 * F_D
WORKLOAD = SHAPE["name"] + ", WGAN-GP training, n_critic=5"
GEN_WORKLOAD = "generate.py generator pass, " + SHAPE["name"] + ", eval mode, no_grad"
TRAFFIC_BATCH, TRAFFIC_SHAPE = None, None      # configuration of THIS run (set in __main__): a capture of another one is not attached
TRAFFIC_FILE = "r2_traffic_b4096.json"        # ncu DRAM-traffic capture of the training step (generate: r2_traffic_generate.json)
NCU_RANGE = os.environ.get("KGAN_NCU_RANGE") == "1"   # bracket the eager roofline pass with cudaProfilerStart/Stop (ncu --profile-from-start off)


_JSON_FD = None


def quiet_stdout():
    """stdout carries exactly ONE line, the JSON result: everything else a library may print there (NCCL's version banner
    under NCCL_DEBUG=VERSION, warnings) is routed to stderr for the duration of the run."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, (line + "\n").encode())


def set_shape(key):
    global SHAPE, F_G, F_D, FLOP_PER_SAMPLE, WORKLOAD, GEN_WORKLOAD
    SHAPE = SHAPES[key]
    F_G, F_D = SHAPE["f_g"], SHAPE["f_d"]
    FLOP_PER_SAMPLE = 1.6 * F_G + 10.4 * F_D
    WORKLOAD = SHAPE["name"] + ", WGAN-GP training, n_critic=5"
    GEN_WORKLOAD = "generate.py generator pass, " + SHAPE["name"] + ", eval mode, no_grad"


def oracle_cfg():
    from oracle import networks as onet

    return onet.Config(dataset=SHAPE["dataset"], n_classes=SHAPE["n_classes"], t_size=SHAPE["t_size"], mlp_dim=SHAPE["mlp_dim"],
                       channels=SHAPE["channels"])


LIB_MODE = {"tf32": "tf32", "fp32": "fp32x3", "fp32_fma": "fp32"}                     # --precision -> kgan.set_precision
DTYPE_NAME = {"tf32": "tf32", "fp32": "fp32 (3xTF32 split on the tensor cores + FMA kernels)", "fp32_fma": "fp32"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="kgan", choices=["kgan", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "generate", "tf32-deviation"],
                    help="train: WGAN-GP training samples/s (headline); generate: inference-only generated sequences/s (BASELINE.json configs[4])")
    ap.add_argument("--shape", default="ntu120", choices=list(SHAPES), help="network / data shape (BASELINE.json configs); headline: ntu120")
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: 4096 for train and for generate)")
    ap.add_argument("--trunc", type=float, default=None, help="generate: W-space truncation factor (generate.py --trunc_mode w); default off")
    ap.add_argument("--trunc-cached", action="store_true", help="generate: estimate the W-space mean once (GeneratorRunner cache_mean) instead of per call")
    ap.add_argument("--precision", default=os.environ.get("KGAN_PRECISION", "tf32"), choices=["fp32", "fp32_fma", "tf32"],
                    help="tf32: headline mode; fp32: fp32-accurate mode - 3xTF32 operand split on the tensor cores wherever a TMA-fed plan exists, "
                         "FMA kernels elsewhere (library mode 'fp32x3'); fp32_fma: FMA kernels only (library mode 'fp32')")
    ap.add_argument("--cpu-batch", type=int, default=32, help="batch of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-secondary", action="store_true", help="train workload: skip the generate-workload measurement attached as `secondary`")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the `gpu_reference` leg (unmodified reference, torch eager, same GPU)")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"], help="--impl reference: where the unmodified reference runs (the arm the "
                    "driver launches is the default, cpu; cuda is what the `gpu_reference` leg calls)")
    ap.add_argument("--ref-tf32", type=int, default=0, help="--impl reference --ref-device cuda: torch.backends.*.allow_tf32")
    ap.add_argument("--first-index", type=int, default=0, help="--impl reference: batch index of the first timed iteration (n_critic phase)")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["bf16_tflops_sustained"], p["hbm_gbs"], "measured"
    except Exception:
        return 1400.0, 6650.0, "fallback"      # B200_PROFILING.md fallback (sustained)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 7 and f[0] == str(self.index):
                self.rows.append(f)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def physical_index(local_rank):
    """nvidia-smi index of the GPU this rank drives."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    ids = [v for v in vis.split(",") if v.strip().isdigit()]
    return int(ids[local_rank]) if local_rank < len(ids) else local_rank


def oracle_trainer(batch, per_sample_loop=True):
    """The CPU restatement of the reference step (oracle/networks.py) at NTU-120 mlp8 - `kind: port`."""
    import torch

    from oracle import networks as onet
    from oracle.graph import SkeletonTables

    cfg = oracle_cfg()
    tables = SkeletonTables(SHAPE["dataset"])
    pg = onet.synth_params(onet.g_param_shapes(cfg, tables), 1, reference_init=True)
    pd = onet.synth_params(onet.d_param_shapes(cfg, tables), 2)
    tr = onet.Trainer(cfg, pg, pd, tables, per_sample_loop=per_sample_loop)
    g = torch.Generator().manual_seed(0)

    def batch_fn():
        real = torch.rand(batch, SHAPE["channels"], SHAPE["t_size"], SHAPE["joints"], generator=g) * 2 - 1
        labels = torch.randint(0, SHAPE["n_classes"], (batch,), generator=g)
        z = torch.randn(batch, 512, generator=g)
        alpha = torch.rand(batch, 1, 1, 1, generator=g)
        nz = [torch.randn(*s, generator=g) for s in onet.noise_shapes(cfg, batch, tables)]
        nz2 = [torch.randn(*s, generator=g) for s in onet.noise_shapes(cfg, batch, tables)]
        return real, labels, z, alpha, nz, nz2

    return tr, batch_fn


def time_oracle(batch, steps, warmup, first_index):
    import torch

    torch.set_num_threads(os.cpu_count() or 1)
    tr, batch_fn = oracle_trainer(batch)
    i = first_index
    for _ in range(warmup):
        tr.iteration(i, *batch_fn())
        i += 1
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.iteration(i, *batch_fn())
        i += 1
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def reference_kind():
    from oracle import ref_runner

    return "reference" if ref_runner.available() else "port"


def run_reference(args):
    """`--impl reference`: the reference's own implementation of the path on the box's host cores - the UNMODIFIED reference
    modules from baseline/_ref (oracle/ref_runner.py: `kind: "reference"`), all host threads, each step one iteration of the
    loop body kinetic-gan.py:137-174 on a bounded batch (`--cpu-batch`, the reference's default 32).  Falls back to the oracle
    port (`kind: "port"`) only if the install is absent.  `--ref-device cuda` runs the same unmodified code in torch eager on the
    GPU (the `gpu_reference` leg of the kgan arm)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_runner

    kind = reference_kind()
    on_gpu = args.ref_device == "cuda"
    b = args.batch if (on_gpu and args.batch) else args.cpu_batch
    if kind == "reference":
        sps, ms, who = ref_runner.time_training(SHAPE, b, args.steps, args.warmup, device=args.ref_device, tf32=bool(args.ref_tf32),
                                                first_index=args.first_index)
    else:
        assert not on_gpu, "the GPU leg needs the installed reference (baseline/_ref)"
        sps, ms, who = time_oracle(b, args.steps, args.warmup, first_index=args.first_index)
    what = ("unmodified reference modules (baseline/_ref), torch %s" % ("eager on " + str(who) if on_gpu else "CPU operators") if kind == "reference"
            else "CPU port of the reference step (oracle/networks.py)")
    line = {
        "impl": "reference", "metric": "wgan_gp_train_samples_per_s", "value": sps, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": ("tf32" if args.ref_tf32 else "fp32") if on_gpu else "fp32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "per_gpu_batch": b, "device": args.ref_device, "note": what + (", host cores only" if not on_gpu else "")},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": who if not on_gpu else 0, "kind": kind,
                         "sample": "%d iterations of batch %d (%s) after %d warm-up" % (args.steps, b, SHAPE["name"], args.warmup)},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(json.dumps(line))


def sub_bench(extra, timeout=900):
    """Runs `bench.py <extra>` in a fresh process (own CUDA context / torch thread pool, global monkey-patches of the import shim
    stay out of this process) and returns its JSON line, or {"error": ...}."""
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "KGAN_NCU_RANGE", "KGAN_SITES_OUT"):
        env.pop(k, None)
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + extra, capture_output=True, text=True, timeout=timeout, env=env)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"error": (r.stderr.strip().splitlines() or ["rc=%d" % r.returncode])[-1][:300]}
        return json.loads(lines[-1])
    except Exception as e:          # noqa: BLE001
        return {"error": repr(e)[:300]}


def cpu_baseline_leg(workload, shape_key, cpu_batch, trunc=None):
    """`cpu_baseline`: the reference arm on the host cores, bounded sample (training: one n_critic cycle i = 5..9 after i = 4 at
    batch 32; generate: 2 calls of batch 256)."""
    if workload == "train":
        j = sub_bench(["--impl", "reference", "--shape", shape_key, "--steps", "5", "--warmup", "1", "--first-index", "4", "--cpu-batch", str(cpu_batch)])
    else:
        j = sub_bench(["--impl", "reference", "--workload", "generate", "--shape", shape_key, "--steps", "2", "--warmup", "1", "--cpu-batch", str(cpu_batch)]
                      + (["--trunc", str(trunc)] if trunc is not None else []))
    return j.get("cpu_baseline", j)


def gpu_reference_leg(workload, shape_key, batch):
    """`gpu_reference`: the UNMODIFIED reference in torch eager (cuDNN / cuBLAS) on the same B200, same per-GPU batch, exact fp32 and
    with TF32 tensor cores allowed; for training also what TF32 does to the reference's own critic gradients."""
    out = {"what": "unmodified reference modules (baseline/_ref) driven by oracle/ref_runner.py, torch eager, 1 GPU, batch %d" % batch}
    for name, tf32 in (("fp32", 0), ("tf32", 1)):
        extra = ["--impl", "reference", "--ref-device", "cuda", "--ref-tf32", str(tf32), "--shape", shape_key, "--batch", str(batch),
                 "--workload", workload, "--steps", "5" if workload == "train" else "3", "--warmup", "1", "--first-index", "4"]
        j = sub_bench(extra)
        out[name] = {k: j[k] for k in ("value", "unit", "ms_per_step", "error") if k in j}
    if workload == "train":
        j = sub_bench(["--impl", "reference", "--ref-device", "cuda", "--workload", "tf32-deviation", "--shape", shape_key])
        out["tf32_deviation"] = j
    return out


def make_roofline(fam, sites, passes, step_tflops):
    """Roofline of the dominant kernel family of the step (CUDA events around every launch of `passes` eager steps).
    achieved = algorithmic bytes (every operand tensor once: inputs + outputs, ops.py `_io`) / summed launch durations for an
    HBM-bound family, algorithmic FLOPs / durations for a tensor-bound one; the family's arithmetic intensity against the
    measured ridge point decides which.  `traffic` = measured DRAM bytes per launch of the family's kernel from the committed
    `ncu --set full` capture (profiles/r1_traffic.json), when present."""
    tensor_peak, hbm_peak, peak_kind = peaks()
    name, st = max(fam.items(), key=lambda kv: kv[1]["ms"])
    sec = st["ms"] * 1e-3
    tf = st["flops"] / sec / 1e12
    gbs = st["bytes"] / sec / 1e9
    ai = st["flops"] / max(st["bytes"], 1.0)
    ridge = tensor_peak * 1e12 / (hbm_peak * 1e9)
    total_ms = sum(v["ms"] for v in fam.values())
    traffic = traffic_src = None
    try:                                        # per-family DRAM bytes per launch, tools/ncu_traffic_summary.py over the same eager pass
        with open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)) as f:
            t = json.load(f)
        if t.get("per_gpu_batch") != TRAFFIC_BATCH or t.get("shape", "ntu120") != TRAFFIC_SHAPE:
            raise KeyError("capture is for another configuration")
        traffic = t["families"][name]["dram_bytes_per_launch"]
        traffic_src = "profiles/%s (ncu dram__bytes_read.sum + dram__bytes_write.sum, per launch, per-GPU batch %s)" % (TRAFFIC_FILE, t.get("per_gpu_batch"))
    except Exception:
        pass
    r = {"kernel": name, "arithmetic_intensity_flop_per_byte": ai, "ridge_flop_per_byte": ridge,
         "algorithmic_bytes_per_launch": st["bytes"] / st["n"], "traffic": traffic, "traffic_source": traffic_src, "peak_kind": peak_kind,
         "tensor": {"achieved": tf, "peak": tensor_peak, "unit": "TFLOP/s", "frac": tf / tensor_peak, "peak_kind": peak_kind + " bf16 sustained"},
         "hbm": {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "peak_kind": peak_kind + " copy bandwidth"},
         "launches_per_step": st["n"] / passes, "avg_launch_ms": st["ms"] / st["n"], "share_of_kgan_kernel_time": st["ms"] / total_ms,
         "step_algorithmic_tflops": step_tflops,
         "families": {k: {"ms_per_step": v["ms"] / passes, "launches_per_step": v["n"] / passes,
                          "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["flops"] else None,
                          "gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["bytes"] else None} for k, v in fam.items()},
         "top_sites": [{"site": k, "ms_per_step": round(v["ms"] / passes, 3), "launches_per_step": v["n"] / passes,
                        "us_per_launch": round(v["ms"] / v["n"] * 1e3, 1),
                        "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["bytes"] else None}
                       for k, v in sorted(sites.items(), key=lambda kv: -kv[1]["ms"])[:40]]}
    if os.environ.get("KGAN_SITES_OUT"):          # full per-site table (diagnostics; not part of the JSON line)
        with open(os.environ["KGAN_SITES_OUT"] + (".generate" if "generate" in TRAFFIC_FILE else ".train"), "w") as f:
            json.dump({k: {"ms_per_step": v["ms"] / passes, "n_per_step": v["n"] / passes, "us": v["ms"] / v["n"] * 1e3,
                           "gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["bytes"] else None,
                           "tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["flops"] else None} for k, v in sites.items()}, f, indent=0)
    if ai < ridge:
        r.update(bound="hbm", achieved=gbs, peak=hbm_peak, unit="GB/s", frac=gbs / hbm_peak)
    else:
        r.update(bound="tensor", achieved=tf, peak=tensor_peak, unit="TFLOP/s", frac=tf / tensor_peak)
    return r


def run_kgan(args):
    import torch

    import kgan_b200 as kgan
    from importlib import import_module

    wg = import_module("kinetic-gan_b200.wgan_gp")
    ddp = import_module("kinetic-gan_b200.ddp")
    ops = kgan.ops
    comm = ddp.Comm()
    dev = torch.device("cuda", comm.local_rank)
    torch.cuda.set_device(dev)
    kgan.set_precision(LIB_MODE[args.precision])
    if os.environ.get("KGAN_STAGED_POLICY"):          # A/B switch of this harness (the library itself reads no environment)
        kgan.geometry.STAGED_POLICY = os.environ["KGAN_STAGED_POLICY"]
    B, K, W = args.batch, args.steps, args.warmup

    torch.manual_seed(0)
    G = kgan.Generator(512, SHAPE["channels"], SHAPE["n_classes"], SHAPE["t_size"], mlp_dim=SHAPE["mlp_dim"], dataset=SHAPE["dataset"]).to(dev)
    D = kgan.Discriminator(SHAPE["channels"], SHAPE["n_classes"], SHAPE["t_size"], 512, dataset=SHAPE["dataset"]).to(dev)
    G.train()
    tr = wg.WGANGPTrainer(G, D, comm=comm)

    # synthetic inputs: a small pool of distinct batches, per-rank RNG streams
    gcpu = torch.Generator().manual_seed(1234 + comm.rank)
    POOL = 4
    host = []
    for _ in range(POOL):
        host.append(dict(real=(torch.rand(B, SHAPE["channels"], SHAPE["t_size"], SHAPE["joints"], generator=gcpu) * 2 - 1).pin_memory(),
                         labels=torch.randint(0, SHAPE["n_classes"], (B,), generator=gcpu).pin_memory(),
                         z=torch.randn(B, 512, generator=gcpu).pin_memory(),
                         alpha=torch.rand(B, 1, 1, 1, generator=gcpu).pin_memory()))
    resident = [{k: v.to(dev) for k, v in h.items()} for h in host]
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())

    def step_resident(i):
        x = resident[i % POOL]
        return tr.iteration(i, x["real"], x["labels"], x["z"], x["alpha"])

    nxt = [None]

    def step_e2e(i):
        # every step: ONE host -> device copy of a batch from pinned memory - the NEXT step's, started on the trainer's copy stream so that
        # it overlaps this step (WGANGPTrainer.prefetch, what train.py's BatchStream does) - and the read-back of this step's loss
        x = nxt[0] if nxt[0] is not None else tr.prefetch(**host[i % POOL])
        nxt[0] = tr.prefetch(**host[(i + 1) % POOL])
        d_loss, _, _ = tr.iteration(i, *x)
        return d_loss.item()                      # device -> host read of the step's result

    def timed(fn, first):
        comm.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(first, first + K):
            fn(i)
        tr.synchronize_updates()                  # the last iteration's all-reduce / Adam (side stream) belong to the timed region
        e1.record()
        comm.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        comm.all_reduce_max_(ms)
        return ms.item()

    graphs = not args.no_graphs
    if graphs:
        r0 = resident[0]
        tr.capture_graphs(r0["real"], r0["labels"], r0["z"], r0["alpha"])
    it = 0
    for _ in range(max(W, 3)):
        step_resident(it)
        it += 1
    torch.cuda.synchronize()
    l0 = ops.launches
    sampler = ClockSampler(physical_index(comm.local_rank))
    sampler.start()
    ms_total = timed(step_resident, it)
    clocks = sampler.stop()
    launches = ops.launches - l0
    it += K
    value = B * comm.world_size * K / (ms_total * 1e-3)

    e2e = None
    if not args.no_e2e:
        step_e2e(it)
        it += 1
        # keep the n_critic phase of the timed window identical to the resident run
        while it % 5 != (max(W, 3)) % 5:
            step_e2e(it)
            it += 1
        ms_e2e = timed(step_e2e, it)
        it += K
        e2e = {"value": B * comm.world_size * K / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4}

    # roofline of the dominant kernel family, measured live with CUDA events around every launch of one more pass
    tr.use_graphs = False                       # eager launches: CUDA events around every libkgan kernel
    ncu_steps = int(os.environ.get("KGAN_NCU_STEPS", "5"))      # steps of the pass inside the cudaProfilerStart/Stop range
    if NCU_RANGE:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    prof = ops.profile_start()
    for i in range(it, it + 5):
        step_resident(i)
        if NCU_RANGE and i - it + 1 == ncu_steps:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
    torch.cuda.synchronize()
    fam = ops.profile_stop(prof)
    sites = fam.pop("_sites")
    tr.use_graphs = graphs
    it += 5
    roofline = make_roofline(fam, sites, 5, FLOP_PER_SAMPLE * value / 1e12 / comm.world_size)

    hbm_peak_gb = torch.cuda.max_memory_allocated(dev) / 1e9      # activations of one step + graphs' pools + parameters
    # release the training state before the other legs (the generate measurement and the reference on the same GPU need the memory)
    del tr, G, D, resident
    kgan.ops._persist.clear()
    kgan.ops._batches.clear()
    kgan.ops.clear_temporary_packs()
    torch.cuda.empty_cache()

    secondary = None
    if not args.no_secondary:                   # the metric's second half (BASELINE.json configs[4]): generated sequences / s, batch 4096 per GPU
        ga = argparse.Namespace(**vars(args))
        ga.batch, ga.trunc, ga.trunc_cached = 4096, None, False
        secondary = measure_generate(ga, comm, dev)
        torch.cuda.empty_cache()

    cpu = gpu_ref = None
    if comm.rank == 0 and comm.world_size == 1:
        if not args.no_cpu_baseline:
            cpu = cpu_baseline_leg("train", args.shape, args.cpu_batch)
        if not args.no_gpu_reference and reference_kind() == "reference":
            gpu_ref = gpu_reference_leg("train", args.shape, min(B, 1024))     # (its per-sample mapping loop is O(N^2): 1024 keeps the leg short)

    if comm.rank == 0:
        line = {
            "metric": "wgan_gp_train_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": comm.world_size, "steps": K,
            "warmup": max(W, 3), "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE_NAME[args.precision], "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": B, "global_batch": B * comm.world_size,
                       "parallelism": "dp%d" % comm.world_size, "n_critic": 5, "flop_per_sample": FLOP_PER_SAMPLE,
                       "cuda_graphs": graphs,
                       "batch_note": "throughput vs per-GPU batch on this GPU (profiles/r2_batch_sweep.txt): 1024: 55.9k, 2048: 61.4k, 4096: 65.7k, "
                                     "6144: 66.6k samples/s; 38 GB of HBM at 4096",
                       "l2_policy": "per-step working set (activations of 4 critic passes at batch %d, >1 GB) exceeds the 126 MB L2; "
                                    "inputs rotate over a pool of %d batches" % (B, POOL)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "gpu_reference": gpu_ref, "secondary": secondary, "hbm_peak_gb": hbm_peak_gb,
        }
        emit(json.dumps(line))
    comm.close()




def time_oracle_generate(batch, steps, warmup, trunc=None):
    """The reference's generate.py:93 call on the host cores: eval-mode generator with its per-sample mapping loop
    (generator.py:84-85), restated in oracle/networks.py."""
    import numpy as np
    import torch

    from oracle import networks as onet
    from oracle.graph import SkeletonTables

    torch.set_num_threads(os.cpu_count() or 1)
    cfg = oracle_cfg()
    tables = SkeletonTables(SHAPE["dataset"])
    pg = onet.synth_params(onet.g_param_shapes(cfg, tables), 1, reference_init=True)
    g = torch.Generator().manual_seed(0)

    def once():
        z = torch.randn(batch, 512, generator=g)
        labels = torch.randint(0, SHAPE["n_classes"], (batch,), generator=g)
        nz = [torch.randn(*s, generator=g) for s in onet.noise_shapes(cfg, batch, tables)]
        tl = torch.as_tensor(np.random.normal(0, 1, (1000, cfg.latent_dim + cfg.n_classes)), dtype=torch.float32) if trunc is not None else None
        with torch.no_grad():
            return onet.generator_forward(pg, z, labels, cfg, tables, nz, training=False, trunc=trunc, trunc_latents=tl, per_sample_loop=True)

    for _ in range(warmup):
        once()
    t0 = time.perf_counter()
    for _ in range(steps):
        once()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def run_reference_generate(args):
    """`--impl reference --workload generate`: generate.py:93 `generator(z, labels)` of the UNMODIFIED reference (eval mode, its
    per-sample mapping loop, no no_grad - as generate.py runs it), host cores by default."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import ref_runner

    kind = reference_kind()
    on_gpu = args.ref_device == "cuda"
    b = args.batch if on_gpu else args.cpu_batch * 8
    if kind == "reference" and args.trunc is None:
        sps, ms, who = ref_runner.time_generate(SHAPE, b, args.steps, args.warmup, device=args.ref_device, tf32=bool(args.ref_tf32))
    else:                                       # W-space truncation draws on torch.cuda.FloatTensor in the reference (generator.py:98): port
        kind = "port"
        sps, ms, who = time_oracle_generate(b, args.steps, args.warmup, args.trunc)
    emit(json.dumps({
        "impl": "reference", "metric": "generated_sequences_per_s", "value": sps, "unit": "seq/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": ("tf32" if args.ref_tf32 else "fp32") if on_gpu else "fp32",
        "data": "synthetic", "config": {"workload": GEN_WORKLOAD, "per_gpu_batch": b, "trunc": args.trunc, "device": args.ref_device,
                                        "note": ("unmodified reference generator (baseline/_ref)" if kind == "reference" else
                                                 "CPU port of the reference generator call (oracle/networks.py)") + ("" if on_gpu else ", host cores only")},
        "cpu_baseline": {"value": sps, "unit": "seq/s", "cores": who if not on_gpu else 0, "kind": kind,
                         "sample": "%d generator calls of batch %d after %d warm-up" % (args.steps, b, args.warmup)},
        "e2e": {"value": sps, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_tf32_deviation(args):
    from oracle import ref_runner

    emit(json.dumps(ref_runner.tf32_gradient_deviation(SHAPE)))


def run_generate(args):
    import torch

    import kgan_b200 as kgan
    from importlib import import_module

    ddp = import_module("kinetic-gan_b200.ddp")
    comm = ddp.Comm()
    dev = torch.device("cuda", comm.local_rank)
    torch.cuda.set_device(dev)
    kgan.set_precision(LIB_MODE[args.precision])
    line = measure_generate(args, comm, dev)
    if comm.rank == 0 and comm.world_size == 1:
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg("generate", args.shape, args.cpu_batch, args.trunc)
        if not args.no_gpu_reference and reference_kind() == "reference" and args.trunc is None:
            line["gpu_reference"] = gpu_reference_leg("generate", args.shape, args.batch)
    if comm.rank == 0:
        emit(json.dumps(line))
    comm.close()


def measure_generate(args, comm, dev):
    """BASELINE.json configs[4]: inference-only generator throughput, NTU 25x64x3, batch 4096 per GPU; independent replicas
    (no collective).  -> the JSON line as a dict (every rank measures; the times are max-reduced)."""
    import torch

    import kgan_b200 as kgan
    from importlib import import_module

    gen = import_module("kinetic-gan_b200.generate")
    ops = kgan.ops
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    torch.manual_seed(0)
    G = kgan.Generator(512, SHAPE["channels"], SHAPE["n_classes"], SHAPE["t_size"], mlp_dim=SHAPE["mlp_dim"], dataset=SHAPE["dataset"]).to(dev)
    with torch.no_grad():                       # a trained generator has non-zero noise weights; exercise that path
        for blk in G.st_gcn_networks:
            blk.noise.weight.normal_(0, 0.1)
    runner = gen.GeneratorRunner(G, B, 512, trunc=args.trunc, graphs=not args.no_graphs, device=dev, cache_mean=args.trunc_cached)
    gcpu = torch.Generator().manual_seed(4321 + comm.rank)
    POOL = 4
    host = [dict(z=torch.randn(B, 512, generator=gcpu).pin_memory(), labels=torch.randint(0, SHAPE["n_classes"], (B,), generator=gcpu).pin_memory())
            for _ in range(POOL)]
    resident = [{k: v.to(dev) for k, v in h.items()} for h in host]
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())

    def step_resident(i):
        x = resident[i % POOL]
        return runner(x["z"], x["labels"])

    def step_e2e(i):
        h = host[i % POOL]
        runner(h["z"], h["labels"])
        return runner.to_host()

    def timed(fn, first):
        comm.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(first, first + K):
            fn(i)
        e1.record()
        comm.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        comm.all_reduce_max_(ms)
        return ms.item()

    it = 0
    for _ in range(W):
        step_resident(it)
        it += 1
    torch.cuda.synchronize()
    l0 = ops.launches
    sampler = ClockSampler(physical_index(comm.local_rank))
    sampler.start()
    ms_total = timed(step_resident, it)
    clocks = sampler.stop()
    launches = ops.launches - l0
    it += K
    value = B * comm.world_size * K / (ms_total * 1e-3)
    e2e = None
    if not args.no_e2e:
        out = step_e2e(it)
        it += 1
        ms_e2e = timed(step_e2e, it)
        it += K
        e2e = {"value": B * comm.world_size * K / (ms_e2e * 1e-3), "unit": "seq/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": out.numel() * out.element_size()}
    global TRAFFIC_FILE
    TRAFFIC_FILE = "r2_traffic_generate.json"
    runner.graphs = False
    if NCU_RANGE:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    prof = ops.profile_start()
    for i in range(it, it + 3):
        step_resident(i)
    torch.cuda.synchronize()
    if NCU_RANGE:
        torch.cuda.profiler.stop()
    fam = ops.profile_stop(prof)
    sites = fam.pop("_sites")
    runner.graphs = not args.no_graphs and (args.trunc is None or args.trunc_cached)
    roofline = make_roofline(fam, sites, 3, F_G * value / 1e12 / comm.world_size)
    # lower bound of the whole pass: z + output (+ noise) = ~21 KB of compulsory HBM traffic per sequence
    roofline["pass_hbm_floor_frac"] = (value / comm.world_size) * 21.2e3 / (peaks()[1] * 1e9)

    del runner, G
    ops._persist.clear()
    ops._batches.clear()
    ops.clear_temporary_packs()
    if True:
        return ({
            "metric": "generated_sequences_per_s", "value": value, "unit": "seq/s", "n_gpus": comm.world_size, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_NAME[args.precision], "data": "synthetic",
            "config": {"workload": GEN_WORKLOAD, "per_gpu_batch": B, "global_batch": B * comm.world_size, "parallelism": "replicas%d" % comm.world_size,
                       "trunc": args.trunc, "trunc_mean_cached": bool(args.trunc_cached), "flop_per_sequence": F_G, "cuda_graphs": not args.no_graphs,
                       "l2_policy": "inputs rotate over a pool of %d batches; the activations of one pass at batch %d exceed the 126 MB L2" % (POOL, B)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": None})


if __name__ == "__main__":
    a = parse()
    quiet_stdout()
    set_shape(a.shape)
    if a.batch is None:
        a.batch = 4096
    TRAFFIC_BATCH, TRAFFIC_SHAPE = a.batch, a.shape
    if a.workload == "tf32-deviation":
        run_tf32_deviation(a)
    elif a.workload == "generate":
        run_reference_generate(a) if a.impl == "reference" else run_generate(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_kgan(a)
