#!/usr/bin/env python
"""bench.py -- points/sec of the SO(3) equivariant backbone, fwd+bwd (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4]

  --config 2 (default, the headline: BASELINE.json configs[1])  classic SPConvNets backbone, N=1024 points, A=60 anchors,
             8 clouds per GPU, fp32 storage / fp32-parity tensor-core arithmetic, train mode
  --config 4 (configs[3])  dense scan: N=4096, 32 clouds per GPU, loss = feats^2 + chamfer_dist (HBM-bound gather stress)
  --config 3 (configs[2])  model-38 backbone shape (3 stride-1 separable blocks 64/128/512, 64 neighbours), synthetic 'oven'
             clouds N=512, 8 clouds per GPU, bf16 fast arithmetic (single-pass bf16 tensor-core contraction)

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); clouds are sharded over ranks, the exchange steps are the
gradient all-reduce (NCCL) and the SyncBatchNorm statistics.  Rank 0 prints ONE JSON line.

  value     whole-job points/sec, inputs resident in HBM, K steps between barrier+sync, CUDA events; the step (forward +
            backward + gradient exchange + Adam) is captured once in a CUDA graph and replayed (--no-graph: eager launches)
  e2e       same metric through the public API with the input clouds in pinned HOST memory (H2D inside the timed region)
            and the loss read back (D2H) every step
  roofline  the dominant entry point of the step (largest summed CUDA-event time), measured in an eager repeat of the same K
            steps with a CUDA-event pair around every C-ABI call
  cpu_baseline      the oracle port of the reference on the host cores, bounded sample (rank 0, N=1 only)
  ref_gpu_baseline  (config 2, N=1) the reference's own CUDA kernels recompiled for sm_100a + eager torch fp32 on this GPU, in a
                    separate process (tests/tools/ref_gpu_baseline.py), with its own clock record
  dp_parity (N>1)   before the timed region: 2 clouds per rank with SyncBatchNorm + the flat gradient bucket must reproduce
                    the single-process result computed on rank 0

--impl reference times the reference's algorithm on the host CPU (oracle/so3.py + oracle_ops.c, the restatement pinned on
the reference's own outputs; the reference Python itself cannot travel to the GPU box) with all host threads on the same
workload.
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
N_ANCHORS = 60
CONFIGS = {
    2: {"n_points": 1024, "clouds": 8, "kind": "classic", "cloud": "shell", "chamfer": False, "gemm_mode": 3, "dtype": "f32",
        "workload": "configs[1]: SPConvNets classic equivariant backbone fwd+bwd, synthetic sphere-shell clouds N=1024 A=60 batch=8/GPU"},
    4: {"n_points": 4096, "clouds": 32, "kind": "classic", "cloud": "shell", "chamfer": True, "gemm_mode": 3, "dtype": "f32",
        "workload": "configs[3]: dense scan, classic backbone fwd+bwd N=4096 A=60 batch=32/GPU + chamfer_dist loss (n=1024 vs m=4096)"},
    3: {"n_points": 512, "clouds": 8, "kind": "model38", "cloud": "oven", "chamfer": False, "gemm_mode": 4, "dtype": "bf16",
        "workload": "configs[2]: model-38 (unsup_arti_align) backbone shape fwd+bwd, synthetic 'oven' clouds N=512 A=60 batch=8/GPU, "
                    "bf16 fast arithmetic"},
}
# step-0 loss of config 2 (seed-0 init, synthetic_cloud(8, 1024, 2000), loss = feats.square().mean()) evaluated by the
# fp64 oracle: tests/golden/make_bench_golden.py -> tests/golden/bench_config2_b8.npz (loss64_square; fp32 oracle: 1.31455362)
ORACLE_STEP0_LOSS = {2: 1.3145534467305648}
FP32_ALU_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # B200: 148 SMs x 128 FMA lanes x 2 flop x max SM clock (non-tensor)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-GPU baseline leg (config 2, N=1)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying a CUDA graph")
    ap.add_argument("--no-sync-bn", action="store_true", help="N>1: keep BatchNorm statistics per rank (default: SyncBatchNorm, as the reference trainer converts its model)")
    ap.add_argument("--gemm-mode", type=int, default=None, help="contraction arithmetic: 0 fp32 FFMA, 1 tcgen05 3xTF32, 2 tcgen05 1xTF32, 3 tcgen05 bf16x3 (default), 4 single-pass bf16 (config 3)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------- clocks
_NVML_SAMPLER = r"""
import sys, time
import pynvml as N
N.nvmlInit()
arg = sys.argv[1]
h = N.nvmlDeviceGetHandleByUUID(arg) if arg.startswith("GPU-") else N.nvmlDeviceGetHandleByIndex(int(arg))
mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
get = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
bits = [N.nvmlClocksEventReasonHwSlowdown, N.nvmlClocksEventReasonHwThermalSlowdown,
        N.nvmlClocksEventReasonSwThermalSlowdown, N.nvmlClocksEventReasonSwPowerCap]
period = float(sys.argv[2])
while True:
    sm, r = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM), get(h)
    print("%.6f,%d,%d,,%s" % (time.time(), sm, mx, ",".join("Active" if r & b else "Not Active" for b in bits)), flush=True)
    time.sleep(period)
"""


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe: nvidia-smi's clocks line).
    Default: a helper PROCESS reads the same NVML counters every 10 ms and timestamps them; the rows between mark_begin() and
    mark_end() count (the 10-step timed region lasts ~150 ms, in which `nvidia-smi -lms 100` delivers one or two samples).
    The helper is started, and has delivered its first row, BEFORE the region: NVML initialisation in another process while the
    steps run costs time -- `nvidia-smi -lms 100` launched at the region's start measured 523-533 k points/s where the helper
    measures 538-540 k on the same box (the end-to-end region, which runs with no sampler at all, reads 519-536 k either way).
    VGTKB_CLOCK_SAMPLER=smi (or no pynvml): the `nvidia-smi -lms 100` process."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.source = index, [], None, None
        self.t_begin, self.t_end = None, None

    def _spawn(self, cmd, stamped):
        self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

        def read():
            for line in self.proc.stdout:
                cols = [c.strip() for c in line.split(",")]
                if stamped:
                    try:
                        self.rows.append((float(cols[0]), cols[1:]))
                    except ValueError:
                        pass
                else:
                    self.rows.append((time.time(), cols))
        self.t = threading.Thread(target=read, daemon=True)
        self.t.start()

    def start(self):
        if os.environ.get("VGTKB_CLOCK_SAMPLER", "nvml") == "nvml":
            try:
                import pynvml  # noqa: F401  (only to know the helper can import it)
                try:
                    import torch
                    dev = "GPU-" + str(torch.cuda.get_device_properties(self.index).uuid)
                except Exception:  # noqa: BLE001  (older torch without .uuid: plain index)
                    dev = str(self.index)
                self._spawn([sys.executable, "-c", _NVML_SAMPLER, dev, "0.01"], True)
                t0 = time.time()
                while not self.rows and time.time() - t0 < 5.0 and self.proc.poll() is None:
                    time.sleep(0.01)                      # the helper needs ~0.3 s to start: wait for its first row HERE
                if self.rows:
                    self.source = "nvml helper process, 10 ms period"
                    return self
                self.proc.terminate()
            except Exception:  # noqa: BLE001
                pass
            self.proc, self.rows = None, []
        try:
            self._spawn(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                         "-lms", "100"], False)
            self.source = "nvidia-smi -lms 100"
        except OSError:
            self.proc = None
        return self

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        rows = [r for ts, r in self.rows if (self.t_begin is None or ts >= self.t_begin) and (self.t_end is None or ts <= self.t_end)]
        if not rows:                                       # region shorter than one period: the nearest samples
            rows = [r for _, r in self.rows[-2:]]
        sm = sorted(int(float(r[0])) for r in rows if r and r[0].replace('.', '').isdigit())
        mx = [int(float(r[1])) for r in rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "sm_mhz_min": sm[0] if sm else None, "source": self.source}


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


def time_oracle(steps, warmup, n_points, clouds=1, budget_s=240.0):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = oracle_step_fn(n_points, clouds)
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
    return {"points_per_s": clouds * n_points * steps / dt, "ms_per_step": 1e3 * dt / steps, "cores": cores,
            "warmup_done": done_warm,
            "sample": f"{clouds} cloud(s) of N={n_points} per step, classic backbone fwd+bwd+Adam, fp32, torch CPU {cores} threads, "
                      f"{steps} timed step(s)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    r = time_oracle(args.steps, args.warmup, 1024 if cfg["kind"] == "model38" else min(cfg["n_points"], 1024))
    line = {"metric": METRIC, "value": r["points_per_s"], "unit": "points/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": r["warmup_done"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "note": "reference algorithm on host CPU (oracle port, kind: port), bounded sample"},
            "cpu_baseline": {"value": r["points_per_s"], "unit": "points/s", "cores": r["cores"], "kind": "port",
                             "sample": r["sample"]},
            "e2e": {"value": r["points_per_s"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------- roofline bookkeeping
def algorithmic_work(name, a):
    """(kind, amount) per call: 'flop' for the contractions, 'byte' for the HBM-bound kernels
    (DESIGN.md section 4 states the per-unit figures)."""
    if name in ("vgtkb_gemm_nt", "vgtkb_gemm_nt_presplit"):
        M, N, K = a[0], a[1], a[2]
        return "flop", 2.0 * M * N * K
    if name in ("vgtkb_gemm_tn", "vgtkb_gemm_tn_presplit", "vgtkb_gemm_tn_planes"):
        M, N, R = a[0], a[1], a[2]
        return "flop", 2.0 * M * N * R
    if name in ("vgtkb_gather_gemm_nt", "vgtkb_gather_gemm_nt_planes", "vgtkb_gather_gemm_tn", "vgtkb_gather_gemm_tn_planes"):
        return "flop", 2.0 * a[0] * a[1] * a[2] * a[3] * a[4]      # points, anchors, kk, c, n|m
    if name in ("vgtkb_inter_group_forward", "vgtkb_inter_group_backward"):
        b, n, p, nn, an, k, ci = a[:7]
        return "byte", 4.0 * b * an * ci * (n + p * k) + 12.0 * b * n + 4.0 * b * p * nn
    if name in ("vgtkb_intra_group_forward", "vgtkb_intra_group_backward"):
        rows, an, kk, c = a[:4]
        return "byte", 4.0 * rows * an * c * (1 + kk)
    if name in ("vgtkb_norm_act_forward", "vgtkb_norm_act_forward_planes"):
        g, rows, c = a[:3]
        return "byte", 8.0 * g * rows * c + (4.0 * g * rows * c if a[8] else 0.0) + (4.0 * g * rows * c if len(a) > 10 and a[10] else 0.0)
    if name in ("vgtkb_norm_stats", "vgtkb_norm_sums"):
        g, rows, c = a[:3]
        return "byte", 4.0 * g * rows * c
    if name in ("vgtkb_norm_act_backward", "vgtkb_norm_act_backward_planes"):
        g, rows, c = a[:3]
        return "byte", (5 + (1 if len(a) > 13 and a[13] else 0)) * 4.0 * g * rows * c
    if name == "vgtkb_inter_conv_forward":            # grouping + contraction (2 B P A K Ci (nn + Co))
        b, n, p, nn, an, k, ci, co = a[:8]
        return "flop", 2.0 * b * p * an * k * ci * (nn + co)
    if name == "vgtkb_inter_conv_backward":           # dW + dG contractions + the scatter product
        b, n, p, nn, an, k, ci, co = a[:8]
        return "flop", 2.0 * b * p * an * k * ci * (nn + 2 * co)
    if name == "vgtkb_chamfer_forward":               # 8 n m flop per cloud pair and direction
        b, n, m = a[0], a[1], a[3]
        return "alu", 2 * 8.0 * b * n * m
    return "byte", 0.0


def summarize_profile(records, steps, peaks):
    agg = {}
    for name, a, e0, e1 in records:
        ms = e0.elapsed_time(e1)
        kind, amt = algorithmic_work(name, a)
        d = agg.setdefault(name, {"ms": 0.0, "calls": 0, "flop": 0.0, "byte": 0.0, "alu": 0.0})
        d["ms"] += ms
        d["calls"] += 1
        d[kind] += amt
    shapes = {}
    for name, a, e0, e1 in records:
        if "gemm" in name or "inter_group" in name or "inter_conv" in name:
            nargs = 5 if "gather" in name else (3 if "gemm" in name else (8 if "inter_conv" in name else 7))
            key = name[6:] + str(tuple(int(v) for v in a[:nargs]))
            d = shapes.setdefault(key, [0.0, 0])
            d[0] += e0.elapsed_time(e1)
            d[1] += 1
    shape_table = {k: {"ms_per_step": v[0] / steps, "calls_per_step": v[1] / steps}
                   for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][0])[:28]}
    total = sum(d["ms"] for d in agg.values()) or 1.0
    table = {k: {"ms_per_step": d["ms"] / steps, "calls_per_step": d["calls"] / steps, "share": d["ms"] / total}
             for k, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
    traffic_all = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic_all = json.load(open(tpath))

    def roof_of(name):
        d = agg[name]
        tr = traffic_all.get(name)
        if d["flop"] > 0:
            ach = d["flop"] / (d["ms"] * 1e-3) / 1e12
            peak = peaks.get("bf16_tflops_sustained", 1400.0)
            return {"bound": "tensor", "kernel": name, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": tr,
                    "peak_source": peaks.get("_source", "fallback") + " (bf16 sustained: kernel timed inside a long step; "
                                   "the fp32-parity contraction issues 3 bf16 passes per product, so frac <= 1/3 by construction)",
                    "avg_launch_ms": d["ms"] / d["calls"], "algorithmic_per_launch": d["flop"] / d["calls"]}
        if d["alu"] > 0:
            ach = d["alu"] / (d["ms"] * 1e-3) / 1e12
            return {"bound": "fp32-alu", "kernel": name, "achieved": ach, "peak": FP32_ALU_PEAK_TFLOPS, "unit": "TFLOP/s",
                    "frac": ach / FP32_ALU_PEAK_TFLOPS, "traffic": tr, "peak_source": "nominal: 148 SMs x 128 FMA x 2 x 1.965 GHz",
                    "avg_launch_ms": d["ms"] / d["calls"], "algorithmic_per_launch": d["alu"] / d["calls"]}
        ach = d["byte"] / (d["ms"] * 1e-3) / 1e9
        peak = peaks.get("hbm_gbs", 6650.0)
        return {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": tr, "peak_source": peaks.get("_source", "fallback"),
                "avg_launch_ms": d["ms"] / d["calls"], "algorithmic_per_launch": d["byte"] / d["calls"]}

    top = max(agg, key=lambda k: agg[k]["ms"])
    roof = roof_of(top)
    # the north star asks for both: HBM fraction on the gather/grouping path, tensor fraction on the contraction
    others = {k: roof_of(k) for k in agg if k != top and (agg[k]["flop"] > 0 or agg[k]["byte"] > 0 or agg[k]["alu"] > 0)}
    return roof, table, shape_table, others


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured (MEASURED_PEAKS.json)"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_source": "fallback (B200_PROFILING.md)"}


def ref_gpu_baseline(local_rank):
    """The reference-style GPU path (its own kernels recompiled + eager torch) in a separate process on the same GPU."""
    tool = os.path.join(ROOT, "tests", "tools", "ref_gpu_baseline.py")
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "vgtk_ref_grouping.so")):
        return {"unavailable": "oracle/_ref not built (the reference kernels are compiled where /root/reference is mounted)"}
    sampler = ClockSampler(local_rank).start()
    try:
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(local_rank)))
        out = subprocess.run([sys.executable, tool, "--steps", "5", "--warmup", "2"], capture_output=True, text=True, timeout=600, env=env)
        line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
        if out.returncode != 0 or not line:
            return {"unavailable": (out.stderr.strip().splitlines() or ["failed"])[-1][:200]}
        d = json.loads(line[-1])
    except Exception as e:  # noqa: BLE001  (a baseline leg must never take the product line down)
        return {"unavailable": str(e)[:200]}
    finally:
        clocks = sampler.stop()
    return {"value": d["points_per_s"], "unit": "points/s", "ms_per_step": d["ms_per_step"], "peak_mem_gb": d["peak_mem_gb"],
            "what": d["what"], "allow_tf32": d["allow_tf32"], "clocks": clocks,
            "note": "same workload (8 clouds x 1024 points, fwd+bwd+Adam); run after the product's timed regions"}


# ---------------------------------------------------------------------------------- our arm
def build_workload(cfg, dev, world, rank):
    import torch
    from equi_articulated_pose_b200 import blocks, synthetic, dataparallel as dp
    n = cfg["n_points"]
    if cfg["kind"] == "classic":
        params = blocks.backbone_params(input_num=n)
    else:
        params = blocks.model38_backbone_params(input_num=n)
    net = blocks.SO3Backbone(params)
    net.load_state_dict(synthetic.init_backbone_state(params, seed=0), strict=False)
    net = net.to(dev).train()
    total = cfg["clouds"] * world                   # weak scaling: fixed clouds per GPU
    lo, hi = dp.shard_range(total, rank, world)
    if cfg["cloud"] == "shell":
        clouds = synthetic.synthetic_cloud(total, n, 2000)[lo:hi]
    else:
        clouds = synthetic.articulated_cloud(cfg["cloud"], total, n, 3000)[lo:hi]
    return net, params, clouds.contiguous(), total


def dp_parity_check(dev, rank, world, sync_bn):
    """2 clouds per rank through SyncBatchNorm + the flat gradient bucket == one process with all 2*world clouds
    (what tests/test_gpu_syncbn.py asserts on 2 GPUs), checked on rank 0 inside the benchmark run."""
    import torch
    import torch.distributed as dist
    from equi_articulated_pose_b200 import blocks, synthetic, dataparallel as dp
    n = 256
    params = blocks.backbone_params(input_num=n)

    def build():
        net = blocks.SO3Backbone(params)
        net.load_state_dict(synthetic.init_backbone_state(params, seed=0), strict=False)
        return net.to(dev).train()
    clouds = synthetic.synthetic_cloud(2 * world, n, 4321).to(dev)
    lo, hi = dp.shard_range(2 * world, rank, world)
    net = build()
    if sync_bn:
        blocks.convert_sync_batchnorm(net)
    bucket = dp.FlatGradBucket(net.parameters())
    bucket.zero_()
    out = net(clouds[lo:hi])
    loss = out.feats.square().mean()
    loss.backward()
    bucket.all_reduce_mean()
    feats = [torch.empty_like(out.feats) for _ in range(world)]
    dist.all_gather(feats, out.feats.detach().contiguous())
    lsum = loss.detach().clone()
    dist.all_reduce(lsum)
    res = {"checked": bool(sync_bn)}
    if rank == 0 and sync_bn:
        ref = build()
        rb = dp.FlatGradBucket(ref.parameters())
        rb.zero_()
        rout = ref(clouds)
        rloss = rout.feats.square().mean()
        rloss.backward()
        rb.collect()
        f = torch.cat(feats, 0)
        rf = rout.feats.detach()
        e_f = float((f - rf).abs().max() / rf.abs().max())
        e_l = abs(float(lsum) / world - float(rloss.detach())) / abs(float(rloss.detach()))
        e_g = float((bucket.flat - rb.flat).abs().max() / rb.flat.abs().max())
        res.update(feats_rel_err=e_f, loss_rel_err=e_l, grad_rel_err=e_g, ok=bool(e_f < 1e-4 and e_l < 1e-4 and e_g < 3e-2),
                   bars="features 1e-4, loss 1e-4, gradients 3e-2 of the bucket maximum (fp32 gradient noise floor, DESIGN 4)")
        del ref, rb, rout, rloss
    del net, bucket, out, loss
    flag = torch.tensor([1 if res.get("ok", True) else 0], device=dev)
    dist.broadcast(flag, 0)
    res["ok"] = bool(int(flag.item()))
    return res


def log(msg):
    if os.environ.get("VGTKB_BENCH_LOG", "0") == "1":
        print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from equi_articulated_pose_b200 import lib, ops, blocks, dataparallel as dp, graph as G   # nothing under oracle/ on this arm

    lib.load()                                      # fail loudly if the CUDA library is missing
    rank, local_rank, world = dp.init_from_env()
    assert torch.cuda.is_available(), "bench.py --impl ours needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cfg = CONFIGS[args.config]
    ops.set_gemm_mode(cfg["gemm_mode"] if args.gemm_mode is None else args.gemm_mode)
    n_points = cfg["n_points"]

    sync_bn = world > 1 and not args.no_sync_bn
    log(f"rank {rank}: dp_parity check")
    parity = dp_parity_check(dev, rank, world, sync_bn) if world > 1 else None
    log(f"rank {rank}: dp_parity {parity}")

    net, params, clouds_host, total_clouds = build_workload(cfg, dev, world, rank)
    if sync_bn:
        blocks.convert_sync_batchnorm(net)         # trainer_unsup_arti_align.py:430
    bucket = dp.FlatGradBucket(net.parameters())
    opt = torch.optim.Adam(bucket.params, lr=1e-3, fused=True, capturable=True)
    clouds_host = clouds_host.pin_memory()
    clouds_dev = clouds_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    if cfg["chamfer"]:
        import equi_articulated_pose_b200 as eap
        eap.install()
        from extensions.chamfer_dist import ChamferDistance
        chamfer = ChamferDistance()
        noise = 0.01 * torch.randn(clouds_dev.shape[0], 1024, 3, generator=torch.Generator().manual_seed(7)).to(dev)

    def step(pts):
        bucket.zero_()
        out = net(pts)
        loss = out.feats.square().mean()
        if cfg["chamfer"]:                          # Y = X[:, :1024] + noise requires grad, n=1024 vs m=4096
            y = (pts[:, :1024] + noise).requires_grad_(True)
            loss = loss + chamfer(y, pts)
        loss.backward()
        bucket.all_reduce_mean()
        opt.step()
        return loss.detach()                        # (a live loss would keep the autograd graph alive across the capture)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # step 0 in eager mode: its loss is checked against the oracle's value for this exact workload
    log(f"rank {rank}: eager steps")
    loss0 = float(step(clouds_dev))
    loss_check = None
    want = ORACLE_STEP0_LOSS.get(args.config)
    if want is not None and world == 1 and ops.get_gemm_mode() == cfg["gemm_mode"]:
        rel = abs(loss0 - want) / abs(want)
        loss_check = {"step0_loss": loss0, "oracle_fp64": want, "rel_err": rel, "bar": 1e-4, "ok": bool(rel < 1e-4)}
        assert rel < 1e-4, f"step-0 loss {loss0} differs from the oracle's {want} (rel {rel:.2e})"
    for _ in range(max(args.warmup, 3) - 1):
        step(clouds_dev)
        flush.zero_()
    barrier()

    # ---- roofline bookkeeping region (eager, a CUDA-event pair around every C-ABI call on the launching stream): the
    # per-kernel durations `roofline` / `kernel_table` report.  Runs BEFORE the graph is captured, on the same K steps.
    lib.PROFILE = []
    k0 = lib.COUNTERS["kernels"]
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
    launches_per_step = (lib.COUNTERS["kernels"] - k0) // args.steps

    # ---- capture the step once; timed regions replay it.  N = 1: one graph (forward + backward + Adam).  N > 1: the NCCL
    # gradient all-reduce stays an eager call between two graphs (forward + backward + bucket | Adam) -- the SyncBatchNorm
    # exchanges inside the first graph are plain kernels over peer memory with a device-side sequence counter, so they
    # replay; an NCCL-based SyncBatchNorm (no peer mailboxes) keeps the step eager.
    use_graph = not args.no_graph and (world == 1 or not sync_bn or ops.peer_mailbox_for(None) is not None)
    graph_note = "eager launches" + (" (--no-graph)" if args.no_graph else "")
    run_step = step
    log(f"rank {rank}: capturing the step (graph={use_graph})")
    if use_graph and world == 1:
        captured = G.CapturedStep(step, [clouds_dev], warmup=2)
        run_step = captured
        graph_note = "one CUDA graph per step (forward + backward + Adam), replayed"
    elif use_graph:
        def step_a(pts):
            bucket.zero_()
            out = net(pts)
            loss = out.feats.square().mean()
            loss.backward()
            bucket.collect()
            return loss.detach()
        captured = G.CapturedStep(step_a, [clouds_dev], warmup=2)
        captured_b = G.CapturedStep(lambda: opt.step(), [], warmup=1)

        def run_step(pts):
            loss = captured(pts)
            bucket.all_reduce_mean()
            captured_b()
            return loss
        graph_note = "two CUDA graphs per step (forward + backward + SyncBatchNorm peer exchanges | Adam) around one eager NCCL all-reduce"
    log(f"rank {rank}: capture done")
    dev_in = captured.static_in[0] if use_graph else clouds_dev      # graph: the static input buffer (holds the same clouds)
    for _ in range(2):
        run_step(dev_in)
        flush.zero_()
    barrier()

    # ---- timed region 1: inputs resident in HBM ------------------------------------------------
    log(f"rank {rank}: timed regions")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record()
    for _ in range(args.steps):
        run_step(dev_in)
        flush.zero_()
    e1.record()
    barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

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
    staging = dev_in if use_graph else torch.empty_like(clouds_dev)
    for i in range(args.steps):
        staging.copy_(clouds_host, non_blocking=True)          # H2D of this step's input (graph: into its static input buffer)
        loss = run_step(staging)
        loss_slots[i & 1].copy_(loss, non_blocking=True)       # D2H of this step's result
        loss_ready[i & 1].record()
        flush.zero_()
        if i > 0:
            loss_ready[(i - 1) & 1].synchronize()
            losses_read.append(float(loss_slots[(i - 1) & 1]))
    loss_ready[(args.steps - 1) & 1].synchronize()
    losses_read.append(float(loss_slots[(args.steps - 1) & 1]))
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
        pts_per_step = total_clouds * n_points
        line = {"metric": METRIC, "value": pts_per_step * args.steps / (ms * 1e-3), "unit": "points/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic",
                "config": {"workload": cfg["workload"], "clouds_per_gpu": cfg["clouds"], "n_points": n_points, "anchors": N_ANCHORS,
                           "step": "fwd + bwd + gradient all-reduce (N>1) + fused Adam", "parallelism": f"dp{world}",
                           "sync_batchnorm": bool(sync_bn), "gemm_mode": ops.get_gemm_mode(), "launch": graph_note,
                           "l2": "256 MiB buffer written between steps (L2 flush); per-step activations >> 126 MB L2"},
                "e2e": {"value": pts_per_step * args.steps / (ms_e2e * 1e-3), "unit": "points/s",
                        "h2d_bytes_per_step": clouds_host.numel() * 4 * world, "d2h_bytes_per_step": 4 * world,
                        "readback": "every step's loss is copied to pinned host memory and read there; the read of step i "
                                    "overlaps step i+1 (two slots)"},
                "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
                "roofline_region": {"ms_per_step": ms_profiled / args.steps,
                                    "note": "roofline / kernel_table: the same K steps in eager mode with a CUDA-event pair around "
                                            "every C-ABI call on the launching stream; `value` is timed without those events"},
                "roofline": roof, "rooflines_other": other_roofs, "kernel_table": table, "shape_table": shape_table,
                "loss": losses_read[-1], "loss_check": loss_check}
        if parity is not None:
            line["dp_parity"] = bool(parity.get("ok")) if parity.get("checked") else None
            line["dp_parity_detail"] = parity
        if args.config == 4:
            o = dict(other_roofs, **{roof["kernel"]: roof})
            line["config4"] = {k: o.get(k) for k in ("vgtkb_inter_group_forward", "vgtkb_chamfer_forward", "vgtkb_furthest_point_sampling",
                                                     "vgtkb_ball_query") if k in o or k in table}
            for k in ("vgtkb_furthest_point_sampling", "vgtkb_ball_query"):
                if k in table:
                    line["config4"][k] = table[k]
        if world == 1 and not args.no_cpu_baseline:
            r = time_oracle(1, 1, min(n_points, 1024))
            line["cpu_baseline"] = {"value": r["points_per_s"], "unit": "points/s", "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"]}
        if world == 1 and args.config == 2 and not args.no_ref_gpu:
            del flush
            torch.cuda.empty_cache()
            line["ref_gpu_baseline"] = ref_gpu_baseline(local_rank)
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
