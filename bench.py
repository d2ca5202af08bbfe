#!/usr/bin/env python
"""LC-GAN G+D training throughput on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--res R] [--batch B] [--impl ours|reference]
                  [--mode train|inference] [--freeze-d] [--config1]

A *step* is one reference iteration (loader.py:44-54): generator step + EMA update + discriminator
step on a global batch of B images, with the reference's loss schedule (aux losses on even
iterations, R1 on iteration % 8 == 1, sparsity on even G steps).  The timed region starts at an
iteration index that is a multiple of 8, so K = 8*n covers whole schedule cycles.

  value    img/s, inputs resident in HBM, timed on the device with CUDA events (max over ranks)
  e2e      img/s through the same public modules with the step's real images / views / latents
           copied from pinned host memory and both losses read back (.item()) every step
  roofline the dominant kernel (by summed device time in an event-instrumented pass): algorithmic
           FLOPs or bytes / event-measured duration, against MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the reference's own modules (baseline/_ref: cnn.py, custom_layers.py, loss.py,
           ema.py, unmodified) stepped by the restated worker.py schedule on the host cores (oracle/reference_arm.py),
           on a bounded sample: half-iterations (alternating G and D steps) at batch 1 and the bench resolution.
           --config1 runs BASELINE config 1 instead (256x256, batch 4, one warm-up iteration + one 8-iteration cycle).
  --mode inference: BASELINE config 5 - generator_ema forward sweep over batch 1..64 (lcgan_b200.inference).
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

METRIC = "G+D train img/s"
FLOP_PER_IMG_ITER = {256: 2.03e12, 512: 2.65e12, 1024: 3.27e12}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--res", type=int, default=int(os.environ.get("LCGAN_BENCH_RES", "1024")),
                    help="image resolution; 1024 = BASELINE.json's headline config (fits one B200 at batch 32)")
    ap.add_argument("--batch", type=int, default=32, help="global batch (reference recipes: 32)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--freeze-d", action="store_true", help="post-freezeD schedule (worker.py:127-131)")
    ap.add_argument("--mode", default="train", choices=["train", "inference"])
    ap.add_argument("--config1", action="store_true",
                    help="reference arm only: BASELINE config 1 (256x256 batch 4 on CPU, warm-up + one 8-iteration cycle)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tensor_burst": d["bf16_tflops"],
                "tensor_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor_sustained": 1400.0, "src": "fallback"}


def workload_name(res, batch):
    hp = {256: "ffhq_256", 512: "afhq_v2_512 (freezeD_layer 4)", 1024: "ffhq_1024 (freezeD_layer 5)"}.get(res, "custom")
    return f"LC-GAN {res}x{res} training batch {batch} ({hp} hyperparams), synthetic data"


# ---------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ---------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own modules on the host cores (oracle/reference_arm.py)
# ---------------------------------------------------------------------------------------------
def _lr(res):
    return 1e-3 if res == 1024 else 2e-3                     # README.md:29/45/49


def cpu_reference_sample(res, half_steps=4, warmup=0, batch=1):
    """Time `half_steps` half-iterations of the reference schedule (k even: G step + EMA of iteration k/2, k odd: D
    step; 4 half-steps = iterations 0 and 1 = all four loss terms) on the host cores, fp32, all threads, at
    `batch` and the bench resolution.  Returns (img_per_s, cores, kind, sample description, seconds per half-step)."""
    from oracle import reference_arm as RA
    cores = os.cpu_count() or 1
    tr, kind, what = RA.make_trainer(res, _lr(res), batch, cores)
    if warmup:
        RA.time_half_steps(tr, 0, warmup)
    ts = RA.time_half_steps(tr, warmup, half_steps)
    import torch
    imgs = batch * half_steps / 2.0
    desc = (f"{what}; torch CPU fp32, {torch.get_num_threads()} threads; half-iterations {warmup}..{warmup + half_steps - 1} "
            f"of the reference schedule (even = G step + EMA, odd = D step; R1 at iteration 1) at batch {batch}, {res}x{res}: "
            f"{', '.join(f'{t:.1f}s' for t in ts)}; img/s = batch * half_steps / 2 / time")
    return imgs / sum(ts), cores, kind, desc, ts


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config1:
        return run_config1(args)
    K, W = args.steps, args.warmup
    v, cores, kind, desc, ts = cpu_reference_sample(args.res, half_steps=K, warmup=W)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "img/s", "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": 1000.0 * sum(ts) / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.res, args.batch), "resolution": args.res,
                       "sample_batch": 1, "step": "one half-iteration (G step + EMA, or D step) of the reference "
                       "schedule at batch 1 - a bounded sample of the batch-32 iteration",
                       "note": "CPU arm runs on host cores only"},
            "cpu_baseline": {"value": v, "unit": "img/s", "cores": cores, "kind": kind, "sample": desc},
            "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


def run_config1(args):
    """BASELINE.json configs[0] / BASELINE.md section 4: 256x256, batch 4, CPU: one warm-up iteration, then one
    8-iteration cycle (all loss variants at the reference's frequencies), per-variant step times."""
    from oracle import reference_arm as RA
    res, batch = 256, 4
    cores = os.cpu_count() or 1
    tr, kind, what = RA.make_trainer(res, _lr(res), batch, cores)
    RA.time_half_steps(tr, 0, 2)                               # iteration 0 as warm-up
    ts = RA.time_half_steps(tr, 16, 16)                        # iterations 8..15: one whole cycle
    names = {}
    for k, t in zip(range(16, 32), ts):
        it = k // 2
        v = ("G even" if it % 2 == 0 else "G odd") if k % 2 == 0 else \
            ("D even" if it % 2 == 0 else ("D odd + R1" if it % 8 == 1 else "D odd"))
        names.setdefault(v, []).append(round(t, 2))
    total = sum(ts)
    value = batch * 8 / total
    import torch
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": 0, "steps": 8, "warmup": 1,
            "ms_per_step": 1000.0 * total / 8, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "LC-GAN 256x256 G+D training (adv + aux + R1 + l_s at reference frequencies) batch 4 on CPU",
                       "resolution": res, "global_batch": batch, "schedule": "iterations 8..15 (one 8-iteration cycle)",
                       "step_seconds": names},
            "cpu_baseline": {"value": value, "unit": "img/s", "cores": cores, "kind": kind,
                             "sample": f"{what}; {torch.get_num_threads()} threads; BASELINE config 1, full cycle"},
            "e2e": {"value": value, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    _emit(line)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def _emit(line):
    """Print the one JSON line on the process's real stdout (libraries such as NCCL write their own
    banners to fd 1; everything else is routed to stderr)."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.mode == "inference":
        return run_inference(args)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    from lcgan_b200 import _lib, cnn, ops, train_step as T
    from lcgan_b200.config import Config, Hyper

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    if rank == 0:
        _lib.build()          # quiet: stdout carries exactly one JSON line
    if world > 1:
        dist.barrier()
    ops.set_precision(args.precision)
    assert args.batch % world == 0
    b = args.batch // world                                   # worker.py:35
    res = args.res
    cfg = Config(img_resolution=res)
    hp = Hyper(lr=1e-3 if res == 1024 else 2e-3)             # README.md:29/45/49
    torch.manual_seed(0)
    G, D = cnn.Generator(cfg.namespace()).to(dev), cnn.Discriminator(cfg.namespace()).to(dev)
    use_graphs = not args.no_graphs
    if world > 1 and not use_graphs:
        from torch.nn.parallel import DistributedDataParallel as DDP     # worker.py:88-96
        G = DDP(G, device_ids=[local_rank], broadcast_buffers=False, find_unused_parameters=True)
        D = DDP(D, device_ids=[local_rank], broadcast_buffers=False, find_unused_parameters=True)
    fl = {256: 3, 512: 4, 1024: 5}.get(res, 3)
    kw = dict(freeze_d_start=0 if args.freeze_d else 10 ** 9, freeze_d_layer=fl)
    # graphs: data parallel with the gradient all-reduce captured in the graphs; eager: torch DDP as
    # the reference wraps it
    tr = T.GraphedTrainer(G, D, hp, b, dev, world=world, **kw) if use_graphs else T.Trainer(G, D, hp, **kw)

    gcpu = torch.Generator().manual_seed(1000 + rank)
    n_pool = 2
    host = [{k: (torch.rand(b, 3, res, res, generator=gcpu) * 2 - 1).pin_memory()
             for k in ("image", "geometry_change", "appearance_change")} for _ in range(n_pool)]
    host_z = [{k: torch.randn(b, 64, generator=gcpu).pin_memory()
               for k in ("rand1", "rand2", "resample1", "resample2", "drand1", "drand2")} for _ in range(n_pool)]
    resident = [{k: v.to(dev) for k, v in h.items()} for h in host]

    def latents():
        z = {k: torch.randn(b, 64, device=dev) for k in ("rand1", "rand2", "resample1", "resample2")}
        zd = {k: torch.randn(b, 64, device=dev) for k in ("rand1", "rand2")}
        return z, zd

    def step_eager(it):
        z, zd = latents()                                     # device RNG, like worker.py:145-146,182-185
        gl = tr.g_step(it, z)
        tr.ema.update(it)
        dl = tr.d_step(it, zd, resident[it % n_pool])
        return gl, dl

    def step_graph(it):
        # fresh latents (device RNG) and the step's images into the graphs' static input buffers
        for t in list(tr.z.values()) + list(tr.zd.values()):
            t.normal_()
        for k, v in resident[it % n_pool].items():
            tr.data[k].copy_(v)
        graph_launches[0] += tr.iteration_graphed(it)

    def step_e2e(it):
        h, hz = host[it % n_pool], host_z[it % n_pool]
        if use_graphs:
            main = torch.cuda.current_stream()
            for k in ("rand1", "rand2", "resample1", "resample2"):
                tr.z[k].copy_(hz[k], non_blocking=True)      # pinned host -> static device buffers
            tr.zd["rand1"].copy_(hz["drand1"], non_blocking=True)
            tr.zd["rand2"].copy_(hz["drand2"], non_blocking=True)
            # this iteration's images are only read by the D step: their 1.2 GB host->device copy runs
            # on a side stream underneath the G step (ordered after the previous D replay, which reads
            # the same static buffers, and before this iteration's D replay)
            copy_stream.wait_stream(main)
            with torch.cuda.stream(copy_stream):
                for k, v in h.items():
                    tr.data[k].copy_(v, non_blocking=True)
            tr.replay_g(it)
            gl = tr.g_loss.item()                             # the reference reads both losses back every
            main.wait_stream(copy_stream)                     # iteration (worker.py:177,214)
            tr.replay_d(it)
            return gl, tr.d_loss.item()
        data = {k: v.to(dev, non_blocking=True) for k, v in h.items()}
        zs = {k: v.to(dev, non_blocking=True) for k, v in hz.items()}
        z = {k: zs[k] for k in ("rand1", "rand2", "resample1", "resample2")}
        zd = {"rand1": zs["drand1"], "rand2": zs["drand2"]}
        return tr.iteration(it, z, zd, data)

    def step_eager_for_profile(it):
        z, zd = latents()
        tr.g_step(it, z)
        tr.ema.update(it)
        tr.d_step(it, zd, resident[it % n_pool])

    graph_launches = [0]
    copy_stream = torch.cuda.Stream(device=dev)
    if use_graphs:
        tr.capture(warmup=2)
    step_resident = step_graph if use_graphs else step_eager

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, first_it, k):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0, gl0 = _lib.launches, graph_launches[0]
        e0.record()
        for i in range(k):
            fn(first_it + i)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), (_lib.launches - launches0) + (graph_launches[0] - gl0)

    K, W = args.steps, max(args.warmup, 3)
    it0 = 0
    for i in range(W):
        step_resident(it0 + i)
    it0 = ((it0 + W + 7) // 8) * 8                           # timed region starts on a cycle boundary
    clocks = Clocks(local_rank) if rank == 0 else None
    ms, launches = timed(step_resident, it0, K)
    clk = clocks.stop() if clocks else None
    value = args.batch * K / (ms / 1000.0)
    it0 += ((K + 7) // 8) * 8

    e2e = None
    if not args.no_e2e:
        step_e2e(it0); it0 += 8                               # warm the pinned-copy path
        ms_e, _ = timed(step_e2e, it0, K)
        it0 += ((K + 7) // 8) * 8
        h2d = sum(v.numel() * v.element_size() for v in host[0].values()) + \
            sum(v.numel() * v.element_size() for v in host_z[0].values())
        e2e = {"value": args.batch * K / (ms_e / 1000.0), "unit": "img/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": 8 * world}

    roof, kernels = None, None
    if use_graphs:                                           # release the graphs' private memory pool
        import gc
        tr.graphs.clear()
        gc.collect()
        torch.cuda.empty_cache()
    if not args.no_roofline:
        # every rank runs the instrumented cycle (the steps contain DDP collectives); rank 0 reports
        roof, kernels = roofline_pass(step_eager_for_profile, it0, _lib, peaks())
    if world > 1:
        dist.barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        del tr, G, D
        torch.cuda.empty_cache()
        v, cores, kind, desc, _ = cpu_reference_sample(res, half_steps=4)
        cpu = {"value": v, "unit": "img/s", "cores": cores, "kind": kind, "sample": desc}

    if rank == 0:
        flop = FLOP_PER_IMG_ITER.get(res)
        pk = peaks()
        line = {
            "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": workload_name(res, args.batch), "global_batch": args.batch, "local_batch": b,
                       "resolution": res, "parallelism": f"dp{world}", "schedule": "timed region = iterations "
                       f"{it0 - 2 * ((K + 7) // 8) * 8 - (8 if not args.no_e2e else 0)}..+{K} (cycle-aligned)",
                       "freezeD": bool(args.freeze_d), "cuda_graphs": bool(use_graphs), "l2": "activations per layer exceed the 126 MB L2; no flush needed",
                       "model_tflop_per_img_iter": flop / 1e12 if flop else None,
                       "model_tflops_achieved": value * flop / 1e12 if flop else None,
                       "frac_of_bf16_peak_sustained": (value * flop / 1e12) / (world * pk["tensor_sustained"]) if flop else None,
                       "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
                       "peaks": pk["src"]},
            "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
            "kernels": kernels,
        }
        _emit(line)
    if world > 1:
        # leave without NCCL's teardown: destroy_process_group() can wait forever on communicators that captured
        # CUDA graphs once used (seen on the B200 box), and the line above is already out
        dist.barrier()
        torch.cuda.synchronize()
        sys.stderr.flush()
        os._exit(0)


def roofline_pass(step_fn, it0, _lib, pk):
    """One extra (untimed-for-the-headline) cycle with a CUDA event pair around every launch of our
    library: per-kernel device time, algorithmic FLOPs (tap convs) and bytes, tap convs tagged by shape.
    The dominant kernel by summed time is the one reported, at its dominant shape."""
    import torch
    from lcgan_b200 import ops
    ops.set_profile_shapes(True)
    _lib.profile_begin()
    for i in range(8):
        step_fn(it0 + i)
    torch.cuda.synchronize()
    stats = _lib.profile_end()
    ops.set_profile_shapes(False)
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    return summarise_kernels(stats, pk, traffic)


def summarise_kernels(stats, pk, traffic_table):
    """stats: {"kernel" or "kernel|shape": {n, ms, flops, bytes}} from the instrumented cycle -> (roofline object,
    the 12 heaviest kernels).  The roofline object describes the dominant kernel AT ITS DOMINANT SHAPE (most summed
    time): algorithmic FLOPs and bytes per launch over the event-measured launch time; a shape whose arithmetic
    intensity is below the measured ridge is held against the HBM peak, otherwise against the sustained bf16 peak.
    `traffic` is the ncu DRAM byte count of that same shape when profiles/ncu_traffic.json has it.  `all_shapes`
    keeps the kernel's aggregate, `hbm_top` the heaviest pure memory kernel."""
    per_kernel, shapes = {}, {}
    for name, st in stats.items():
        k = name.split("|")[0]
        agg = per_kernel.setdefault(k, {"n": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        for f in agg:
            agg[f] += st[f]
        if "|" in name:
            shapes.setdefault(k, {})[name.split("|", 1)[1]] = st
    rows = []
    for name, st in per_kernel.items():
        rows.append({"kernel": name, "launches": st["n"], "ms": st["ms"],
                     "tflops": (st["flops"] / (st["ms"] / 1e3) / 1e12) if st["flops"] and st["ms"] > 0 else None,
                     "gbs": (st["bytes"] / (st["ms"] / 1e3) / 1e9) if st["bytes"] and st["ms"] > 0 else None})
    if not rows:
        return None, []
    rows.sort(key=lambda r: -r["ms"])
    total = sum(r["ms"] for r in rows) or 1.0
    for r in rows:
        r["share"] = r["ms"] / total
    ridge = pk["tensor_sustained"] * 1e12 / (pk["hbm"] * 1e9)       # FLOP per byte

    def obj(kernel, st, share, shape=None):
        n = max(st["n"], 1)
        t = st["ms"] / 1e3
        ai = st["flops"] / st["bytes"] if st["flops"] and st["bytes"] else 0.0
        key = kernel if shape is None else f"{kernel}|{shape}"
        o = {"kernel": kernel, "shape": shape, "launches": st["n"], "avg_launch_ms": st["ms"] / n, "share_of_step": share,
             "flop_per_launch": st["flops"] / n, "bytes_per_launch": st["bytes"] / n,
             "traffic": traffic_table.get(key), "timing": "CUDA event pair around every launch, eager pass"}
        if st["flops"] and ai >= ridge:
            a = st["flops"] / t / 1e12
            o.update({"bound": "tensor", "achieved": a, "peak": pk["tensor_sustained"], "unit": "TFLOP/s",
                      "frac": a / pk["tensor_sustained"],
                      "peak_source": pk["src"] + " (sustained: kernel timed inside a long step)"})
        else:
            a = st["bytes"] / t / 1e9 if t > 0 else 0.0
            o.update({"bound": "hbm", "achieved": a, "peak": pk["hbm"], "unit": "GB/s", "frac": a / pk["hbm"],
                      "peak_source": pk["src"]})
            if st["flops"]:
                o["tflops"] = st["flops"] / t / 1e12
        return o

    top = rows[0]
    sh = shapes.get(top["kernel"])
    if sh:
        name, st = max(sh.items(), key=lambda kv: kv[1]["ms"])
        roof = obj(top["kernel"], st, st["ms"] / total, name)
        roof["all_shapes"] = {"launches": top["launches"], "share_of_step": top["share"], "tflops": top["tflops"],
                              "gbs": top["gbs"], "frac_of_tensor_peak": (top["tflops"] or 0) / pk["tensor_sustained"]}
        roof["shapes"] = [dict(shape=k, launches=v["n"], ms=v["ms"],
                               tflops=v["flops"] / (v["ms"] / 1e3) / 1e12 if v["flops"] and v["ms"] > 0 else None,
                               gbs=v["bytes"] / (v["ms"] / 1e3) / 1e9 if v["bytes"] and v["ms"] > 0 else None)
                          for k, v in sorted(sh.items(), key=lambda kv: -kv[1]["ms"])[:10]]
    else:
        roof = obj(top["kernel"], per_kernel[top["kernel"]], top["share"])
    mem = [r for r in rows if r["tflops"] is None and r["gbs"] is not None]
    if mem and mem[0]["kernel"] != top["kernel"]:
        roof["hbm_top"] = obj(mem[0]["kernel"], per_kernel[mem[0]["kernel"]], mem[0]["share"])
    # the heaviest (kernel, shape) pairs of the whole step, whatever kernel they belong to
    flat = [(k, sh_, v) for k, d_ in shapes.items() for sh_, v in d_.items()]
    flat.sort(key=lambda e: -e[2]["ms"])
    roof["top_shapes"] = [dict(kernel=k, shape=sh_, launches=v["n"], ms=v["ms"], share=v["ms"] / total,
                               tflops=v["flops"] / (v["ms"] / 1e3) / 1e12 if v["flops"] and v["ms"] > 0 else None,
                               gbs=v["bytes"] / (v["ms"] / 1e3) / 1e9 if v["bytes"] and v["ms"] > 0 else None)
                          for k, sh_, v in flat[:24]]
    return roof, rows[:12]


# ---------------------------------------------------------------------------------------------
# BASELINE config 5: generator-only inference sweep (worker.py:427-441, 447-485)
# ---------------------------------------------------------------------------------------------
G_GMAC_PER_IMG = {256: 56.3, 512: 71.5, 1024: 86.9}          # SURVEY section 8a (forward MACs per image)


def run_inference(args):
    import torch
    import copy
    from lcgan_b200 import _lib, cnn, ops
    from lcgan_b200.config import Config
    from lcgan_b200.inference import GeneratorRunner
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    _lib.build()
    ops.set_precision(args.precision)
    res = args.res
    torch.manual_seed(0)
    G = cnn.Generator(Config(img_resolution=res).namespace()).to(dev)
    with torch.no_grad():                                     # a few training-mode forwards give avg_latent a value
        for _ in range(3):
            G(torch.randn(8, 64, device=dev), torch.randn(8, 64, device=dev))
    runner = GeneratorRunner(copy.deepcopy(G), w_psi=0.7)
    K, W = args.steps, max(args.warmup, 3)
    sweep, clocks, pk = [], None, peaks()
    flop_img = 2 * G_GMAC_PER_IMG.get(res, 0) * 1e9
    for b in runner.batch_sizes:
        zg, za = runner.static_inputs(b)
        hz = [torch.randn(b, 64).pin_memory(), torch.randn(b, 64).pin_memory()]
        hout = torch.empty(b, 3, res, res, dtype=torch.uint8).pin_memory()
        for _ in range(W):
            zg.normal_(); za.normal_(); runner.replay(b)
        torch.cuda.synchronize()
        if b == runner.batch_sizes[-1]:
            clocks = Clocks(dev.index or 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            zg.normal_(); za.normal_()
            runner.replay(b)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        # end to end: latents from pinned host memory, uint8 images back to pinned host memory, every step
        e0.record()
        for _ in range(K):
            zg.copy_(hz[0], non_blocking=True); za.copy_(hz[1], non_blocking=True)
            _, _, u8 = runner.replay(b)
            hout.copy_(u8, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms_e = e0.elapsed_time(e1) / K
        sweep.append({"batch": b, "ms": ms, "img_s": b / ms * 1e3, "e2e_img_s": b / ms_e * 1e3,
                      "tflops": b * flop_img / (ms / 1e3) / 1e12 if flop_img else None,
                      "frac_of_bf16_peak_sustained": b * flop_img / (ms / 1e3) / 1e12 / pk["tensor_sustained"] if flop_img else None,
                      "launches_per_forward": runner.launches[b]})
    clk = clocks.stop() if clocks else None
    best = max(sweep, key=lambda r: r["img_s"])
    bl = runner.batch_sizes[-1]
    line = {"metric": "generator-only inference img/s", "value": best["img_s"], "unit": "img/s", "n_gpus": 1, "steps": K,
            "warmup": W, "ms_per_step": best["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"LC-GAN {res}x{res} generator-only inference sweep (generator_ema, w_psi 0.7, batch 1-64), "
                                   "random-init weights", "resolution": res, "best_batch": best["batch"],
                       "cuda_graphs": True, "l2": "activations per layer exceed the 126 MB L2 from batch 2 up",
                       "g_forward_gmac_per_img": G_GMAC_PER_IMG.get(res), "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9},
            "clocks": clk,
            "e2e": {"value": best["e2e_img_s"], "unit": "img/s", "h2d_bytes_per_step": best["batch"] * 2 * 64 * 4,
                    "d2h_bytes_per_step": best["batch"] * 3 * res * res},
            "gpu_launches": sum(runner.launches[b] for b in runner.batch_sizes) * (2 * K + W), "sweep": sweep,
            "roofline": {"kernel": "generator forward (all kernels)", "bound": "tensor", "achieved": best["tflops"],
                         "peak": pk["tensor_sustained"], "unit": "TFLOP/s", "frac": best["frac_of_bf16_peak_sustained"],
                         "traffic": None, "peak_source": pk["src"]},
            "cpu_baseline": None}
    _emit(line)


if __name__ == "__main__":
    main()
