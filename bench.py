#!/usr/bin/env python
"""bench.py - multi-view hand poses/sec of the HandMvNet forward path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

Own arm.  A "step" is one forward of the release HO3D configuration (5 views, 256x256 crops, ResNet-50-paper
backbone -> cross-attention fusion -> Chebyshev GCN head) over a synthetic batch of B=64 samples per GPU
(BASELINE.json configs[1]; weak scaling: every rank gets its own 64 samples, no data-path collective, only an
NCCL all-gather of the [B,21,3] poses).  `value` = samples all ranks processed / max-over-ranks device time with
the inputs resident in HBM; `e2e` = the same through `HandMvNet.forward_host_async` (hmv_forward_host_u8_async) with
pinned HOST buffers holding the uint8 image crops the reference's data pipeline produces in front of ToTensor +
Normalize (datasets/ho3d.py:35-40), host->device copies of the step's inputs and device->host copies of the poses
inside the timed region (`e2e.fp32_input` = the same loop fed host-normalised fp32 tensors, 4x the bytes).
`roofline` describes the dominant kernel class of the step (by device time), timed per launch with CUDA events on its
stream in a second pass over the same steps, against the roofline that bounds it; `roofline.classes` lists every
tcgen05 kernel class the same way and `roofline.all_tc` their aggregate.  `gpu_eager_baseline` (N = 1) is the oracle's
full forward run as eager PyTorch on the same GPU (cudnn.benchmark, fp32 with TF32 convolutions and bf16 autocast) -
the north star's "x the reference's eager forward at B=64"; `latency_b1` is the p50 / p99 of one B=1 forward
(reference protocol src/eval_fps.py:68-106) for the 5- and 8-view configurations; `cpu_baseline` is the CPU oracle (a
port of the reference's torch forward) timed on this box's host cores on the same bounded sample the reference arm uses.

Reference arm (`--impl reference`): the reference's algorithm on the host CPU cores (oracle port, all threads),
same metric / config, each step a bounded sample of the workload.
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

METRIC = "multi-view hand poses/sec (HandMvNet forward, HO3D release config, 5 views)"
METRIC_FMT = "multi-view hand poses/sec (HandMvNet forward, release config, {v} views)"
UNIT = "poses/s"
FLOP_PER_SAMPLE_V5 = 108.515e9        # SURVEY.md §8d (FlopCounterMode over the reference forward)
FLOP_PER_SAMPLE = {5: 108.515e9, 8: 173.58e9}
FLOP_PER_SAMPLE_HRNET = {5: 153.26e9}   # FlopCounterMode over the oracle's HO3D_HandMvNet_HR forward (HRNet-w40, 52.8 M parameters)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the run through NVML (in-process: an `nvidia-smi -lms`
    child was measured to stall the GPU for tens of ms per query on this pool, polluting the timed region)."""

    def __init__(self, gpu_index, period_s=0.02):
        self.rows, self.idx, self.period = [], gpu_index, period_s
        self._stop = threading.Event()
        self._active = threading.Event()       # queries run only while set (see resume / pause)
        self._thread = None
        self.error = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v for v in vis.split(",") if v.strip() != ""]
            try:
                return int(ids[self.idx])
            except (ValueError, IndexError):
                return self.idx
        return self.idx

    def _run(self):
        try:
            import pynvml as nv
            if os.environ.get("HMV_BENCH_SAMPLER_LATE", "0") == "1":
                self._active.wait()                # (experiment) NVML is not even initialised before the timed steps are enqueued
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            mode = os.environ.get("HMV_BENCH_SAMPLER", "full")
            # first calls outside the timed region (they run during the clock ramp): the first query of each kind is the slow one
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            if mode != "clock":
                nv.nvmlDeviceGetPowerUsage(h)
                get_reasons(h)
            while not self._stop.is_set():
                if not self._active.wait(timeout=0.05):
                    continue
                t = time.perf_counter()
                clk = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0 if mode != "clock" else 0.0
                rs = int(get_reasons(h)) if mode != "clock" else 0
                self.rows.append((t, clk, pw, rs, time.perf_counter() - t))
                time.sleep(self.period)
            nv.nvmlShutdown()
        except Exception as e:  # noqa: BLE001
            self.error = repr(e)

    def start(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    # The sampler queries only between resume() - called once all K timed steps are enqueued, i.e. while the GPU is
    # still working through them - and pause() at the end of the region, so that it never competes with the thread
    # that enqueues CUDA work.  (The 50-140 ms stalls once blamed on it were cudaGraphInstantiate inside the second
    # timed step: see the warm-up comment in run_own.)
    def resume(self):
        self._active.set()

    def pause(self):
        self._active.clear()

    def stop(self, t_begin=None, t_end=None):
        """Summary over the samples taken in [t_begin, t_end] (perf_counter seconds): the timed region."""
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2.0)
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                "hw_power_brake_slowdown": 0x80}
        rows = [r for r in self.rows if (t_begin is None or r[0] >= t_begin) and (t_end is None or r[0] <= t_end)]
        sm = sorted(r[1] for r in rows)
        reasons = sorted(name for name, bit in bits.items() if any(r[3] & bit for r in rows))
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": getattr(self, "max_sm", None), "reasons": reasons,
                "power_w_max": max((r[2] for r in rows), default=None), "samples": len(rows),
                "samples_total": len(self.rows), "query_ms_max": max((r[4] for r in self.rows), default=0.0) * 1e3,
                "query_ms_max_in_region": max((r[4] for r in rows), default=0.0) * 1e3,
                "source": f"NVML, {self.period * 1e3:.0f} ms period, during the timed region (from the moment its last step is enqueued)", "error": self.error}


def cpu_oracle_throughput(batch, iters, warmup, views=5):
    """The CPU oracle (port of the reference's torch forward) on all host threads."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import handmvnet_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.release_config(views, True)
    sd = O.make_state_dict(cfg, seed=0, randomize_norm=False)
    x, bbox, intr = O.make_inputs(batch, views, seed=1234)
    times = []
    for i in range(warmup + iters):
        t0 = time.perf_counter()
        O.forward(sd, cfg, x, bbox, intr)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return batch * len(times) / total, total / len(times) * 1e3, torch.get_num_threads()


CPU_SAMPLE_B = 2       # samples per CPU step: the ONE bounded sample both `--impl reference` and `cpu_baseline` time


def run_reference(args, lines):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_b = CPU_SAMPLE_B
    ps, ms, cores = cpu_oracle_throughput(sample_b, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": ps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "HO3D_HandMvNet release config, 5 views, B=64/GPU (reference CPU forward timed on a "
                               f"bounded sample of B={sample_b} per step)", "views": 5, "image": 256},
        "cpu_baseline": {"value": ps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"oracle/handmvnet_oracle.py forward, B={sample_b} x {args.steps} steps, fp32, all host threads"},
        "e2e": {"value": ps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    lines.append(json.dumps(line))


def run_own(args, lines):
    import torch
    import torch.distributed as dist
    from handmvnet_b200 import HandMvNet
    from handmvnet_b200.config import release_config

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    views, B = args.views, args.batch
    cfg = release_config(views, True, backbone=args.backbone)
    torch.manual_seed(0)
    model = HandMvNet(cfg["train"], cfg["model"], cfg["data"], precision="bf16", micro_batch=args.micro_batch)
    model.to(dev).eval()
    model.freeze()
    model.prepare(dev)

    g = torch.Generator().manual_seed(1234 + rank)
    c = torch.rand(B, views, 2, generator=g) * torch.tensor([320.0, 240.0]) + torch.tensor([160.0, 120.0])
    side = 100 + 150 * torch.rand(B, views, 1, generator=g)
    bbox_host = torch.cat([c - side / 2, c + side / 2], dim=-1).pin_memory()
    f = 500 + 200 * torch.rand(B, views, 1, generator=g)
    intr_host = torch.cat([f, f, torch.full((B, views, 1), 320.0), torch.full((B, views, 1), 240.0)], dim=-1).pin_memory()
    if args.no_e2e:                                    # large-batch sweeps: inputs are generated on the device
        gd = torch.Generator(device=dev).manual_seed(1234 + rank)
        x_host = None
        x = torch.randn(B, views, 3, 256, 256, device=dev, generator=gd)
    else:
        x_host = torch.randn(B, views, 3, 256, 256, generator=g).pin_memory()        # 251 MB at B=64 (> 126 MB L2)
        x = x_host.to(dev)
    bbox, intr = bbox_host.to(dev), intr_host.to(dev)
    cam = {"intrinsic": intr}
    gathered = [torch.empty(B, 21, 3, device=dev) for _ in range(world)] if world > 1 else None

    def step():
        out = model(x, bbox, cam)
        if world > 1:
            dist.all_gather(gathered, out["joints_cam"])       # optional NCCL gather of the poses (252 B/sample)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank, period_s=float(os.environ.get("HMV_BENCH_SAMPLER_PERIOD", "0.02")))
    if rank == 0 and not args.no_clocks:
        sampler.start()
    # clock ramp: a GPU coming out of idle needs ~1 s of load before it holds its boost clocks; this pre-warm is
    # not counted as one of the W warm-up steps
    # (time-bounded, so the number of iterations differs between ranks: no collective inside this loop)
    t_ramp = time.perf_counter()
    # The results are held the way the timed loop holds them (`out = step()` keeps step k's tensors alive while step
    # k+1 allocates its own), so the two alternating sets of output buffers - and the library's pointer-keyed CUDA
    # graphs for them, captured the second time a set is seen - exist before the timed region: in round 2 the second
    # timed step was measured to spend 1-165 ms of host time in cudaGraphInstantiate, starving the GPU.
    out = None
    n_pre = 0
    while time.perf_counter() - t_ramp < args.ramp_seconds or n_pre < 4:
        out = model(x, bbox, cam)
        torch.cuda.synchronize(dev)
        n_pre += 1
    for _ in range(args.warmup):
        out = step()
    barrier()
    launches0 = model.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    import gc
    gc.collect()
    gc.disable()              # a generation-2 collection inside the timed loop stalls the enqueueing thread for tens of ms
    barrier()
    t_region0 = time.perf_counter()
    e0.record()
    marks[0].record()
    t_enq = time.perf_counter()
    enq_t = [t_enq]
    for i in range(args.steps):
        out = step()
        marks[i + 1].record()
        enq_t.append(time.perf_counter())
    enqueue_ms = (time.perf_counter() - t_enq) * 1e3 / args.steps
    enq_steps = sorted(((enq_t[i + 1] - enq_t[i]) * 1e3, i) for i in range(args.steps))[-3:]
    e1.record()
    sampler.resume()          # the steps are enqueued; the GPU executes them while the clocks are sampled
    barrier()
    t_region1 = time.perf_counter()
    sampler.pause()
    gc.enable()
    dev_ms = e0.elapsed_time(e1)
    step_ms = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps))
    launches = model.launch_count() - launches0
    assert torch.isfinite(out["joints_cam"]).all()

    # ---- end to end through the host-buffer API (pinned host inputs, H2D + D2H inside the timed region) ----
    # Streaming loop of a caller that feeds batch after batch (eval_fps.py:79-92): step k+1 is enqueued before the
    # results of step k are awaited, so its host->device copy overlaps step k's compute.  Every step still copies its
    # own inputs from pinned host memory and reads its poses back, all inside the timed region.
    def host_loop(x_h):
        for _ in range(2):
            model.forward_host_async(x_h, bbox_host, cam_host, want_heatmap=False).result(recycle=True)
        barrier()
        t0 = time.perf_counter()
        prev, ho = None, None
        for _ in range(args.steps):
            tk = model.forward_host_async(x_h, bbox_host, cam_host, want_heatmap=False)
            if prev is not None:
                ho = prev.result(recycle=True)
            prev = tk
        ho = prev.result(recycle=True)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        barrier()
        assert torch.isfinite(ho["joints_cam"]).all()
        return dt, ho["joints_cam"].numel() * 4 + ho["joints_crop_img"].numel() * 4

    small = bbox_host.numel() * 4 + intr_host.numel() * 4
    e2e_u8_s = e2e_f32_s = float("nan")
    d2h = B * 21 * 3 * 4 + B * views * 21 * 2 * 4
    if not args.no_e2e:
        cam_host = {"intrinsic": intr_host}
        # headline: uint8 crops, what the reference's data pipeline holds in front of ToTensor + Normalize
        xu_host = torch.randint(0, 256, (B, views, 3, 256, 256), generator=g, dtype=torch.uint8).pin_memory()
        e2e_u8_s, d2h = host_loop(xu_host)
        # the same loop fed host-normalised fp32 tensors (4x the bytes: host-memory bound when 8 ranks share one socket)
        e2e_f32_s, _ = host_loop(x_host)
        # latency form of the same call (one batch at a time, wait before the next)
        t0 = time.perf_counter()
        for _ in range(max(3, args.steps // 4)):
            model.forward_host(xu_host, bbox_host, cam_host, want_heatmap=False)
        e2e_sync_ms = (time.perf_counter() - t0) * 1e3 / max(3, args.steps // 4)

    # ---- per-launch timing of the tcgen05 kernels: same steps again with CUDA events around every launch ----
    model.profile(True)
    for _ in range(args.steps):
        model(x, bbox, cam)
    import tempfile
    csv_path = os.path.join(tempfile.gettempdir(), f"hmv_tc_launches_{os.getpid()}.csv")
    tc_ms, tc_flops, tc_n = model.profile_read(csv_path)
    phases = {k: v / args.steps for k, v in model.profile_phases().items()}
    model.profile(False)
    clocks = sampler.stop(t_region0, t_region1) if rank == 0 and not args.no_clocks else None
    if rank == 0 and os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        import shutil
        shutil.copyfile(csv_path, os.path.join(ROOT, "gpurun_out", "tc_launches.csv"))

    times = torch.tensor([dev_ms, 0.0 if args.no_e2e else e2e_u8_s * 1e3, 0.0 if args.no_e2e else e2e_f32_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_u8_ms, e2e_f32_ms = float(times[0]), float(times[1]), float(times[2])
    line = None
    if rank == 0:
        peaks = load_peaks()
        value = B * world * args.steps / (dev_ms * 1e-3)
        step_ms_mean = dev_ms / args.steps
        line = {
            "metric": (METRIC if views == 5 else METRIC_FMT.format(v=views)) + (" [HRNet-w40 backbone]" if args.backbone == "hrnet" else ""), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms_mean,
            "step_ms": {"min": step_ms[0], "median": step_ms[len(step_ms) // 2], "max": step_ms[-1], "cpu_enqueue": enqueue_ms,
                        "cpu_enqueue_slowest": [{"step": i, "ms": round(ms, 3)} for ms, i in reversed(enq_steps)]},
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{'HO3D' if views == 5 else 'DexYCB-style'}_HandMvNet{'_HR' if args.backbone == 'hrnet' else ''} release config, synthetic {views}-view B={B} per GPU, bf16, random-init weights",
                       "backbone": "hrnet-w40" if args.backbone == "hrnet" else "resnet50_paper",
                       "views": views, "image": 256, "batch_per_gpu": B, "micro_batch": args.micro_batch,
                       "parallelism": f"batch-sharded x{world} (replicated weights, NCCL all-gather of poses)",
                       "l2": f"inputs {x.numel() * 4 / 2**20:.0f} MiB + {0.445 * B:.1f} GB of activations per step exceed the 126 MB L2"},
            "clocks": clocks,
            "e2e": None if args.no_e2e else {
                "value": B * world * args.steps / (e2e_u8_ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": xu_host.numel() + small, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_u8_ms / args.steps,
                "sync_call_ms": e2e_sync_ms, "input": "uint8 [B,V,3,256,256] image crops in pinned host memory",
                "fp32_input": {"value": B * world * args.steps / (e2e_f32_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_f32_ms / args.steps,
                               "h2d_bytes_per_step": x.numel() * 4 + small,
                               "note": "same loop fed host-normalised fp32 [B,V,3,256,256] tensors (the reference module's own input type)"},
                "api": "HandMvNet.forward_host_async -> hmv_forward_host_u8_async / hmv_host_wait (pinned host buffers; ToTensor + Normalize of "
                       "datasets/ho3d.py:35-40 run inside the stem kernel; poses copied back; at most 2 steps in flight); "
                       "sync_call_ms = blocking HandMvNet.forward_host per call"},
            "gpu_launches": launches,
            "roofline": roofline_report(csv_path, args.steps, step_ms_mean, peaks, views if args.backbone == "resnet" else 0, B, args.micro_batch, tc_ms, tc_flops, tc_n,
                                        (FLOP_PER_SAMPLE_HRNET.get(views, 30.42e9 * views + 1.2e9) if args.backbone == "hrnet"
                                         else FLOP_PER_SAMPLE.get(views, 21.40e9 * views + 1.5e9)) * value / world * 1e-12, phases,
                                        sm_mhz=(clocks or {}).get("sm_mhz")),
        }
        if world == 1 and not args.no_cpu_baseline:
            ps, ms, cores = cpu_oracle_throughput(CPU_SAMPLE_B, 10, 3)
            line["cpu_baseline"] = {"value": ps, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"oracle/handmvnet_oracle.py forward, B={CPU_SAMPLE_B} x 10 steps (+3 warm-up), fp32, all host threads "
                                              "(the same bounded sample `--impl reference` times)",
                                    "ms_per_forward": ms}
    # the two legs below allocate: release the model's workspace first
    del model
    if world == 1 and rank == 0:
        if not args.no_eager:
            line["gpu_eager_baseline"] = gpu_eager_baseline(dev, views, B, step_ms_mean, args.backbone)
        if not args.no_latency and args.backbone == "resnet":
            line["latency_b1"] = latency_b1(dev)
    if rank == 0:
        lines.append(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def kernel_class(name):
    """layer3.2.conv2 -> layer3.x.conv2 ; joints_late_fusion.attn_fusion.0.qkv -> fusion.x.qkv"""
    import re
    name = re.sub(r"layer(\d)\.\d+\.", r"layer\1.x.", name)
    name = re.sub(r"joints_late_fusion\.attn_fusion\.\d+\.", "fusion.x.", name)
    return name


def roofline_report(csv_path, steps, step_ms, peaks, views, batch, micro_batch, tc_ms, tc_flops, tc_n, model_tflops, phases, sm_mhz=None, num_sms=148):
    """Per kernel class (launches of one class share a geometry): CUDA-event time per launch, algorithmic FLOPs and HBM
    bytes (SURVEY.md §8d / DESIGN.md §4: every input, residual, output element and every weight once), the roofline
    that bounds the class (the larger of FLOPs / sustained tensor peak and bytes / measured HBM peak) and the achieved
    fraction of it.  The headline object describes the class with the largest share of the step."""
    import csv
    classes = {}
    with open(csv_path) as f:
        for row in csv.DictReader(f):
            c = classes.setdefault(kernel_class(row["layer"]), {"n": 0, "ms": 0.0, "gflop": 0.0, "mbytes": 0.0, "mmas": 0.0, "mma_cycles": 0.0})
            c["n"] += 1
            c["ms"] += float(row["ms"])
            c["gflop"] += float(row["gflop"])
            c["mbytes"] += float(row["mbytes"])
            c["mmas"] += float(row.get("mmas") or 0.0)
            c["mma_cycles"] += float(row.get("mma_cycles") or 0.0)
    tpeak, hpeak = peaks["bf16_tflops_sustained"], peaks["hbm_gbs"]
    traffic = load_traffic(views, batch, micro_batch)
    out = []
    for name, c in classes.items():
        ms_l = c["ms"] / c["n"]
        t_tensor = c["gflop"] / c["n"] / tpeak            # GFLOP / (TFLOP/s) = ms
        t_hbm = c["mbytes"] / c["n"] / hpeak              # MB / (GB/s) = ms
        bound = "tensor" if t_tensor >= t_hbm else "hbm"
        rec = {"kernel": name, "launches_per_step": c["n"] / steps, "ms_per_launch": ms_l, "ms_per_step": c["ms"] / steps,
               "share_of_step": c["ms"] / steps / step_ms, "bound": bound,
               "achieved": c["gflop"] / c["ms"] if bound == "tensor" else c["mbytes"] / c["ms"],
               "peak": tpeak if bound == "tensor" else hpeak, "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
               "frac": max(t_tensor, t_hbm) / ms_l, "tflops": c["gflop"] / c["ms"], "gbs": c["mbytes"] / c["ms"],
               "algorithmic_mbytes_per_launch": c["mbytes"] / c["n"], "gflop_per_launch": c["gflop"] / c["n"]}
        if c["mma_cycles"] > 0:
            # tensor-pipe floor at the shapes the kernel issues: an MMA over 128 rows x N columns x K = 16 takes max(N / 2, 48)
            # cycles (tools/mma_issue_bench.cu), so layers with N = 64 / 128 cannot reach the FLOP roofline
            mhz = sm_mhz or 1550.0
            rec["mma_per_launch"] = c["mmas"] / c["n"]
            rec["tensor_shape_floor_ms"] = c["mma_cycles"] / c["n"] / num_sms / (mhz * 1e3)
            rec["frac_tensor_shape"] = rec["tensor_shape_floor_ms"] / ms_l
        out.append(rec)
    out.sort(key=lambda r: -r["ms_per_step"])
    top = dict(out[0]) if out else {}
    tr = (traffic or {}).get("classes", {}).get(top.get("kernel"))
    if traffic is None:
        top["traffic"], top["traffic_note"] = None, "no ncu capture committed for this workload"
    elif tr is None or abs(tr["launches_per_step"] - top["launches_per_step"]) > 1e-6 or traffic["tc_launches_per_step"] != round(tc_n / steps):
        top["traffic"] = None
        top["traffic_note"] = (f"stale ncu capture: {traffic['source']} saw {traffic['tc_launches_per_step']} tcgen05 launches per step, "
                               f"this run timed {tc_n / steps:.0f}")
    else:
        top["traffic"] = tr["dram_bytes_per_launch"]
        top["traffic_unit"] = "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu)"
        top["traffic_over_algorithmic"] = tr["dram_bytes_per_launch"] / (top["algorithmic_mbytes_per_launch"] * 1e6)
        top["traffic_per_step_bytes_all_tc"] = traffic["tc_dram_bytes_per_step"]
        top["traffic_source"] = traffic["source"]
    top["peak_source"] = peaks["source"] + (" (sustained cuBLAS bf16: kernel timed inside a long step)" if top.get("bound") == "tensor"
                                             else " (device copy bandwidth)")
    top["timing"] = "second pass over the same steps with CUDA events around every launch"
    agg = tc_flops / (tc_ms * 1e-3) * 1e-12 if tc_ms > 0 else 0.0
    top["all_tc"] = {"kernel": "every tcgen05 launch of the step (implicit-GEMM convs / linears, fused tails, fused seams)", "bound": "tensor",
                     "achieved": agg, "peak": tpeak, "unit": "TFLOP/s", "frac": agg / tpeak, "launches_per_step": tc_n / steps,
                     "kernel_ms_per_step": tc_ms / steps, "kernel_share_of_step": tc_ms / steps / step_ms,
                     "sum_of_class_bounds_ms": sum(r["frac"] * r["ms_per_step"] for r in out),
                     "end_to_end_model_tflops": model_tflops}
    top["phase_ms_per_step"] = phases
    top["classes"] = [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()
                       if k in ("kernel", "launches_per_step", "ms_per_launch", "ms_per_step", "bound", "frac", "tflops", "gbs", "tensor_shape_floor_ms", "frac_tensor_shape")} for r in out]
    top["tensor_shape_floor_note"] = ("tensor_shape_floor_ms = sum over the launch's tcgen05.mma of max(N / 2, 48) cycles, / SMs, at the SM clock "
                                      "sampled during the timed region: what the tensor pipe needs at the N the layer allows (measured: "
                                      "tools/mma_issue_bench.cu); `frac` stays the FLOP / HBM roofline fraction")
    return top


def load_traffic(views, batch, micro_batch):
    """DRAM bytes per launch per kernel class from the committed ncu launch list of one B=64 / 5-view step
    (profiles/r02/tc_traffic.json, written by tools/ncu_step_traffic.py); None for any other workload."""
    path = os.path.join(ROOT, "profiles", "r02", "tc_traffic.json")
    if views != 5 or batch != 64 or micro_batch != 64 or not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f)


def gpu_eager_baseline(dev, views, B, own_ms, backbone="resnet"):
    """The reference's forward as eager PyTorch on THIS GPU (north star: ">= 20x the reference's eager PyTorch forward on
    1 B200 at B=64"): the oracle's full forward (every stage, constants resident on the device) with the protocol of
    src/eval_fps.py:17,79-94 - cudnn.benchmark on, no_grad, fp32 (cuDNN convolutions use TF32 by default, as in the
    reference) and bf16 autocast - timed with CUDA events."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import handmvnet_oracle as O
    torch.cuda.empty_cache()
    cfg = O.release_config(views, True, backbone)
    sd = {k: v.to(dev) for k, v in O.make_state_dict(cfg, seed=0, randomize_norm=backbone == "hrnet").items()}
    x, bbox, intr = O.make_inputs(B, views, seed=1234)
    x, bbox, intr = x.to(dev), bbox.to(dev), intr.to(dev)
    prev = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True

    def timed(iters=8, warm=3):
        for _ in range(warm):
            O.forward(sd, cfg, x, bbox, intr)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out = O.forward(sd, cfg, x, bbox, intr)
        e1.record()
        torch.cuda.synchronize(dev)
        assert torch.isfinite(out["joints_cam"]).all()
        return e0.elapsed_time(e1) / iters

    ms32 = timed()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ms16 = timed()
    torch.backends.cudnn.benchmark = prev
    del sd, x
    torch.cuda.empty_cache()
    return {"what": "oracle/handmvnet_oracle.py full forward as eager PyTorch on the same GPU, cudnn.benchmark=True, CUDA-event timed "
                    "(restates the loop of src/eval_fps.py:79-94 without the MANO post-processing)",
            "batch": B, "views": views,
            "fp32_tf32conv": {"ms_per_step": ms32, "value": B / ms32 * 1e3, "unit": UNIT},
            "bf16_autocast": {"ms_per_step": ms16, "value": B / ms16 * 1e3, "unit": UNIT},
            "speedup_vs_eager_fp32": ms32 / own_ms, "speedup_vs_eager_bf16": ms16 / own_ms,
            "target": 20.0}


def latency_b1(dev, iters=300, warm=50):
    """p50 / p99 of ONE B=1 forward on device-resident inputs (BASELINE.json metric; reference protocol
    src/eval_fps.py:68-106: one sample per call, timed per call), 5-view HO3D and 8-view DexYCB configurations."""
    import torch
    from handmvnet_b200 import HandMvNet
    from handmvnet_b200.config import release_config
    out = {}
    for v in (5, 8):
        cfg = release_config(v, True)
        torch.manual_seed(0)
        m = HandMvNet(cfg["train"], cfg["model"], cfg["data"], precision="bf16", micro_batch=1)
        m.to(dev).eval()
        m.freeze()
        m.prepare(dev)
        x = torch.randn(1, v, 3, 256, 256, device=dev)
        bbox = torch.tensor([220.0, 140.0, 420.0, 340.0], device=dev).expand(1, v, 4).contiguous()
        cam = {"intrinsic": torch.tensor([600.0, 600.0, 320.0, 240.0], device=dev).expand(1, v, 4).contiguous()}
        for _ in range(warm):
            m(x, bbox, cam)
        torch.cuda.synchronize(dev)
        n0 = m.launch_count()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            o = m(x, bbox, cam)
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        assert torch.isfinite(o["joints_cam"]).all()
        out[f"views{v}"] = {"p50_ms": ts[len(ts) // 2], "p99_ms": ts[int(len(ts) * 0.99)], "min_ms": ts[0], "iters": iters,
                            "kernels_per_forward": (m.launch_count() - n0) // iters}
        del m
        torch.cuda.empty_cache()
    out["protocol"] = "B=1 per call, device-resident fp32 inputs, CUDA events around each call (CUDA-graph replay inside the library)"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64, help="samples per GPU per step")
    ap.add_argument("--micro-batch", type=int, default=64, help="samples per internal pass")
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--views", type=int, default=5, help="camera views per sample (5 = HO3D release config, 8 = DexYCB)")
    ap.add_argument("--backbone", default="resnet", choices=["resnet", "hrnet"], help="resnet = *_HandMvNet.yaml (north star), hrnet = *_HandMvNet_HR.yaml (HRNet-w40)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (large-batch sweeps)")
    ap.add_argument("--no-clocks", action="store_true", help="do not run the NVML clock sampler")
    ap.add_argument("--no-eager", action="store_true", help="skip the eager-PyTorch-on-GPU baseline leg (N=1 only)")
    ap.add_argument("--no-latency", action="store_true", help="skip the B=1 latency leg (N=1 only)")
    ap.add_argument("--ramp-seconds", type=float, default=1.5, help="untimed load before the warm-up steps (clock ramp)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if os.environ.get("HMV_BENCH_WATCHDOG"):          # debugging aid: dump every thread's stack and exit after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["HMV_BENCH_WATCHDOG"]), exit=True)
    # Only the JSON line may reach stdout: libraries (NCCL prints its version banner there) are pointed at stderr
    # for the duration of the run and the real stdout is restored for the final print.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    lines = []
    try:
        if args.impl == "reference":
            run_reference(args, lines)
        else:
            run_own(args, lines)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for ln in lines:
        print(ln, flush=True)


if __name__ == "__main__":
    main()
