#!/usr/bin/env python
"""bench.py -- points/sec of the SO(3) equivariant backbone, fwd+bwd, BASELINE.json config 2
(classic SPConvNets backbone, N=1024 points, A=60 anchors, 8 clouds per GPU, fp32, train mode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); clouds are sharded over ranks,
the only collective is the gradient all-reduce (NCCL).  Rank 0 prints ONE JSON line.

  value     whole-job points/sec, inputs resident in HBM, K steps between barrier+sync, CUDA events
  e2e       same metric with the input clouds in pinned HOST memory (H2D inside the timed region)
            and the loss read back (D2H) every step
  roofline  the dominant kernel of the step (largest summed CUDA-event time); the per-call CUDA events are recorded
            in a repeat of the same K steps right after the headline region (they cost ~0.9 ms of a 19 ms step)
  cpu_baseline  the oracle port of the reference on the host cores, bounded sample (rank 0, N=1 only)

--impl reference times the reference's algorithm on the host CPU (oracle/so3.py + oracle_ops.c,
the restatement pinned on the reference's own outputs; the reference Python itself cannot travel
to the GPU box) with all host threads on the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "points/sec fwd+bwd SPConv backbone N=1024 A=60"
N_POINTS, N_ANCHORS, CLOUDS_PER_GPU = 1024, 60, 8
WORKLOAD = "configs[1]: SPConvNets classic equivariant backbone fwd+bwd, synthetic sphere-shell clouds N=1024 A=60 batch=8/GPU"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sync-bn", action="store_true", help="N>1: keep BatchNorm statistics per rank (default: SyncBatchNorm, as the reference trainer converts its model)")
    ap.add_argument("--gemm-mode", type=int, default=None, help="contraction arithmetic: 0 fp32 FFMA, 1 tcgen05 3xTF32, 2 tcgen05 1xTF32, 3 tcgen05 bf16x3 (default)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace('.', '').isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------- CPU oracle leg
def oracle_step_fn(n_points, clouds, seed=0):
    """fwd+bwd(+Adam) of the classic backbone with the oracle restatement on the host CPU."""
    import torch
    from oracle import so3 as O
    from equi_articulated_pose_b200 import so3_constants as C
    params = O.backbone_params(input_num=n_points)
    sd = O.init_backbone_state(params, seed=seed)
    for k in sd:
        sd[k].requires_grad_(True)
    for bi, blk in enumerate(params):
        for li, layer in enumerate(blk):
            co = layer['args']['dim_out']
            for pre in (f'backbone.{bi}.blocks.{li}.inter_conv.norm.', f'backbone.{bi}.blocks.{li}.norm.'):
                sd[pre + 'running_mean'], sd[pre + 'running_var'] = torch.zeros(co), torch.ones(co)
    leaves = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.Adam(leaves, lr=1e-3)
    anchors, intra = torch.from_numpy(C.anchors_all()), torch.from_numpy(C.intra_idx())
    pts = O.synthetic_cloud(clouds, n_points, 2000)
    xyz = pts.permute(0, 2, 1).contiguous()
    feats = torch.ones(clouds, 1, n_points, N_ANCHORS)

    def step():
        opt.zero_grad(set_to_none=True)
        _, of = O.backbone_forward(sd, params, xyz, feats, anchors, intra, C.kernel_points_base(), training=True)
        loss = of.square().mean()
        loss.backward()
        opt.step()
        return float(loss)
    return step


def time_oracle(steps, warmup, clouds=1, budget_s=240.0):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = oracle_step_fn(N_POINTS, clouds)
    t0 = time.perf_counter()
    step()
    first = time.perf_counter() - t0
    done_warm = 1
    while done_warm < warmup and (done_warm + steps) * first < budget_s:
        step()
        done_warm += 1
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"points_per_s": clouds * N_POINTS * steps / dt, "ms_per_step": 1e3 * dt / steps, "cores": cores,
            "warmup_done": done_warm,
            "sample": f"{clouds} cloud(s) of N={N_POINTS} per step, fwd+bwd+Adam, fp32, torch CPU {cores} threads, "
                      f"{steps} timed step(s)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = time_oracle(args.steps, args.warmup)
    line = {"metric": METRIC, "value": r["points_per_s"], "unit": "points/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": r["warmup_done"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "reference algorithm on host CPU (oracle port), bounded sample"},
            "cpu_baseline": {"value": r["points_per_s"], "unit": "points/s", "cores": r["cores"], "kind": "port",
                             "sample": r["sample"]},
            "e2e": {"value": r["points_per_s"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------- roofline bookkeeping
def algorithmic_work(name, a):
    """(kind, amount) per call: 'flop' for the contractions, 'byte' for the HBM-bound kernels
    (DESIGN.md section 4 states the per-unit figures)."""
    if name == "vgtkb_gemm_nt":
        M, N, K = a[0], a[1], a[2]
        return "flop", 2.0 * M * N * K
    if name == "vgtkb_gemm_tn":
        M, N, R = a[0], a[1], a[2]
        return "flop", 2.0 * M * N * R
    if name == "vgtkb_gather_gemm_nt":      # points, anchors, kk, c, n
        return "flop", 2.0 * a[0] * a[1] * a[2] * a[3] * a[4]
    if name == "vgtkb_gather_gemm_tn":      # points, anchors, kk, c, m
        return "flop", 2.0 * a[0] * a[1] * a[2] * a[3] * a[4]
    if name in ("vgtkb_inter_group_forward", "vgtkb_inter_group_backward"):
        b, n, p, nn, an, k, ci = a[:7]
        return "byte", 4.0 * b * an * ci * (n + p * k) + 12.0 * b * n + 4.0 * b * p * nn
    if name in ("vgtkb_intra_group_forward", "vgtkb_intra_group_backward"):
        rows, an, kk, c = a[:4]
        return "byte", 4.0 * rows * an * c * (1 + kk)
    if name in ("vgtkb_norm_act_forward",):
        g, rows, c = a[:3]
        return "byte", 8.0 * g * rows * c + (4.0 * g * rows * c if a[8] else 0.0)
    if name in ("vgtkb_norm_stats",):
        g, rows, c = a[:3]
        return "byte", 4.0 * g * rows * c
    if name in ("vgtkb_norm_act_backward",):
        g, rows, c = a[:3]
        return "byte", 5 * 4.0 * g * rows * c
    if name in ("vgtkb_inter_conv_forward",):
        b, n, p, nn, an, k, ci, co = a[:8]
        return "flop", 2.0 * b * p * an * k * ci * (nn + co)
    return "byte", 0.0


def summarize_profile(records, steps, peaks):
    agg = {}
    for name, a, e0, e1 in records:
        ms = e0.elapsed_time(e1)
        kind, amt = algorithmic_work(name, a)
        d = agg.setdefault(name, {"ms": 0.0, "calls": 0, "flop": 0.0, "byte": 0.0})
        d["ms"] += ms
        d["calls"] += 1
        d[kind] += amt
    shapes = {}
    for name, a, e0, e1 in records:
        if name in ("vgtkb_gemm_nt", "vgtkb_gemm_tn", "vgtkb_gather_gemm_nt", "vgtkb_gather_gemm_tn",
                    "vgtkb_inter_group_forward", "vgtkb_inter_group_backward"):
            key = name[6:] + str(tuple(int(v) for v in (a[:5] if "gather" in name else (a[:3] if "gemm" in name else a[:7]))))
            d = shapes.setdefault(key, [0.0, 0])
            d[0] += e0.elapsed_time(e1)
            d[1] += 1
    shape_table = {k: {"ms_per_step": v[0] / steps, "calls_per_step": v[1] / steps}
                   for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][0])[:24]}
    total = sum(d["ms"] for d in agg.values()) or 1.0
    table = {k: {"ms_per_step": d["ms"] / steps, "calls_per_step": d["calls"] / steps, "share": d["ms"] / total}
             for k, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
    traffic_all = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic_all = json.load(open(tpath))

    def roof_of(name):
        d = agg[name]
        if d["flop"] > 0:
            ach = d["flop"] / (d["ms"] * 1e-3) / 1e12
            peak = peaks.get("bf16_tflops_sustained", 1400.0)
            return {"bound": "tensor", "kernel": name, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": traffic_all.get(name),
                    "peak_source": peaks.get("_source", "fallback") + " (bf16 sustained: kernel timed inside a long step; "
                                   "the fp32-parity contraction issues 3 bf16 passes per product, so frac <= 1/3 by construction)",
                    "avg_launch_ms": d["ms"] / d["calls"], "algorithmic_per_launch": d["flop"] / d["calls"]}
        ach = d["byte"] / (d["ms"] * 1e-3) / 1e9
        peak = peaks.get("hbm_gbs", 6650.0)
        return {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic_all.get(name), "peak_source": peaks.get("_source", "fallback"),
                "avg_launch_ms": d["ms"] / d["calls"], "algorithmic_per_launch": d["byte"] / d["calls"]}

    top = max(agg, key=lambda k: agg[k]["ms"])
    roof = roof_of(top)
    # the north star asks for both: HBM fraction on the gather/grouping path, tensor fraction on the contraction
    others = {k: roof_of(k) for k in agg if k != top and (agg[k]["flop"] > 0 or agg[k]["byte"] > 0)}
    return roof, table, shape_table, others


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured (MEASURED_PEAKS.json)"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_source": "fallback (B200_PROFILING.md)"}


# ---------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from equi_articulated_pose_b200 import lib, ops, blocks, synthetic, dataparallel as dp   # nothing under oracle/ on this arm

    lib.load()                                      # fail loudly if the CUDA library is missing
    rank, local_rank, world = dp.init_from_env()
    assert torch.cuda.is_available(), "bench.py --impl ours needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if args.gemm_mode is not None:
        ops.set_gemm_mode(args.gemm_mode)

    params = blocks.backbone_params(input_num=N_POINTS)
    net = blocks.SO3Backbone(params)
    net.load_state_dict(synthetic.init_backbone_state(params, seed=0), strict=False)
    net = net.to(dev).train()
    sync_bn = world > 1 and not args.no_sync_bn
    if sync_bn:
        blocks.convert_sync_batchnorm(net)         # trainer_unsup_arti_align.py:430
    bucket = dp.FlatGradBucket(net.parameters())
    opt = torch.optim.Adam(bucket.params, lr=1e-3, fused=True)

    total_clouds = CLOUDS_PER_GPU * world          # weak scaling: 8 clouds per GPU
    lo, hi = dp.shard_range(total_clouds, rank, world)
    clouds_host = synthetic.synthetic_cloud(total_clouds, N_POINTS, 2000)[lo:hi].contiguous().pin_memory()
    clouds_dev = clouds_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def step(pts):
        bucket.zero_()
        out = net(pts)
        loss = out.feats.square().mean()
        loss.backward()
        bucket.all_reduce_mean()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(clouds_dev)
        flush.zero_()
    barrier()

    # ---- timed region 1: inputs resident in HBM ------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    k0, c0 = lib.COUNTERS["kernels"], lib.COUNTERS["launch_calls"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(clouds_dev)
        flush.zero_()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.COUNTERS["kernels"] - k0
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region 1b: the same K steps with a CUDA-event pair around every C-ABI call (roofline bookkeeping).
    # ~600 event records per step cost ~0.9 ms of a 19 ms step (measured), so they stay out of the headline region;
    # the per-kernel durations they deliver are the ones `roofline` / `kernel_table` report.
    lib.PROFILE = []
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    p0.record()
    for _ in range(args.steps):
        step(clouds_dev)
        flush.zero_()
    p1.record()
    barrier()
    ms_profiled = p0.elapsed_time(p1)
    records, lib.PROFILE = lib.PROFILE, None

    # ---- timed region 2: end to end (pinned host input -> device, loss -> host, every step) -----
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    # the training loop a user writes: every step uploads its clouds from pinned host memory and downloads its loss; the
    # loss of step i is READ on the host (event wait + float()) while step i+1 is already enqueued, so the host never
    # stalls the device (two pinned loss slots; the last one is read before the region closes)
    loss_slots = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event() for _ in range(2)]
    losses_read = []
    for i in range(args.steps):
        pts = clouds_host.to(dev, non_blocking=True)
        loss = step(pts)
        loss_slots[i & 1].copy_(loss.detach(), non_blocking=True)
        loss_ready[i & 1].record()
        flush.zero_()
        if i > 0:
            loss_ready[(i - 1) & 1].synchronize()
            losses_read.append(float(loss_slots[(i - 1) & 1]))
    loss_ready[(args.steps - 1) & 1].synchronize()
    losses_read.append(float(loss_slots[(args.steps - 1) & 1]))
    loss_host.copy_(loss_slots[(args.steps - 1) & 1])
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)
    assert len(losses_read) == args.steps and all(v == v for v in losses_read), "every step's loss is read back"

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        peaks = load_peaks()
        roof, table, shape_table, other_roofs = summarize_profile(records, args.steps, peaks)
        pts_per_step = total_clouds * N_POINTS
        line = {"metric": METRIC, "value": pts_per_step * args.steps / (ms * 1e-3), "unit": "points/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "clouds_per_gpu": CLOUDS_PER_GPU, "n_points": N_POINTS, "anchors": N_ANCHORS,
                           "step": "fwd + bwd + gradient all-reduce (N>1) + fused Adam", "parallelism": f"dp{world}", "sync_batchnorm": bool(sync_bn),
                           "gemm_mode": ops.get_gemm_mode(),
                           "l2": "256 MiB buffer written between steps (L2 flush); per-step activations >> 126 MB L2"},
                "e2e": {"value": pts_per_step * args.steps / (ms_e2e * 1e-3), "unit": "points/s",
                        "h2d_bytes_per_step": clouds_host.numel() * 4 * world, "d2h_bytes_per_step": 4 * world,
                        "readback": "every step's loss is copied to pinned host memory and read there; the read of step i "
                                    "overlaps step i+1 (two slots)"},
                "gpu_launches": launches, "clocks": clocks,
                "roofline_region": {"ms_per_step": ms_profiled / args.steps,
                                    "note": "roofline / kernel_table: the same K steps repeated with a CUDA-event pair around every "
                                            "C-ABI call on the launching stream; `value` is timed without those events"},
                "roofline": roof, "rooflines_other": other_roofs, "kernel_table": table, "shape_table": shape_table,
                "loss": float(loss_host)}
        if world == 1 and not args.no_cpu_baseline:
            r = time_oracle(1, 1)
            line["cpu_baseline"] = {"value": r["points_per_s"], "unit": "points/s", "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
