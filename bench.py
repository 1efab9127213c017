#!/usr/bin/env python3
"""Headline benchmark: kerr_schild at 3840x2160 (BASELINE.json configs[2]) through the C ABI.

A "step" is one frame of the hot path: camera/tetrad -> prepass -> init_rays -> integrate -> render_data ->
adaptive sampling -> integrate -> render_data -> render, in the reference's default operating mode (prepass as the
metric's JSON asks, adaptive sampling on at threshold 32 - main.cpp:1123-1158, graphics_settings.hpp:39-40).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|3|4|5] [--impl reference]

* value   Mray-steps/s: trial Verlet steps of the frame (SURVEY.md 8(d) unit of work) / device frame time; every input
          (background pyramid, camera, parameters) and the delivered frame stay in HBM; CUDA events on the frames' own
          streams; frames are queued back to back with `frames_in_flight` frame slots busy.
* e2e     the same metric with the camera uploaded (H2D) and the assembled frame delivered into pinned HOST memory on
          rank 0 (D2H) for every frame, through grt_comm_render, host clock.
* N > 1   the frame is tiled in interleaved 54-row stripes over N ranks (one process per GPU, torchrun). Every rank traces its
          stripes (+ halo, + the redundant 1/16 prepass) and its render kernel stores the pixels straight into rank 0's frame over
          NVLink (libgrt's grt_comm_*: peer-mapped frame, stream memory operations for the hand-shake, no collective kernel);
          torch.distributed (NCCL) only carries the 64-byte IPC handles, the barriers and the max-over-ranks reductions.
          Strong scaling: the work of ONE frame is fixed, value = its single-GPU step count / max-over-ranks time.
* --impl reference   the reference's own integration kernel (oracle/_ref: /root/reference/cl.cl compiled as C++ for the host,
          DESIGN.md) on all host cores (thread count set explicitly), on a bounded pixel subsample of the same frame.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

# one hardware work queue per frame slot's stream (the default 8 serialises the slots beyond the eighth); has to be in the environment
# before CUDA is initialised in this process (libgrt sets it too, for hosts that do not load PyTorch first)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import geodesic_raytracing_b200 as grt  # noqa: E402

STRIPE_ROWS = int(os.environ.get("GRT_BENCH_STRIPE_ROWS", "54"))
CAM_POS, CAM_QUAT = (0.0, 0.0, -4.0, 0.0), (-0.70710678, 0.0, 0.0, 0.70710678)
# BASELINE.json `configs` (index = position in that list + 1; config 1 is the reference's CPU-only plumbing case)
CONFIGS = {
    2: dict(metric="schwarzschild", size=(1920, 1080), features={}, note="fixed camera"),
    3: dict(metric="kerr_schild", size=(3840, 2160), features={}, note="headline config"),
    4: dict(metric="alcubierre", size=(1920, 1080), features={"redshift": 1}, geodesic=dict(speed=0.3, tau=2.0), note="redshift + anisotropic filter on, camera attached to a geodesic"),
    5: dict(metric="double_unequal_kerr", size=(3840, 2160), features={}, note="row-tiled over the GPUs of one box"),
}
_WORLD = int(os.environ.get("WORLD_SIZE", "1"))
# measured optima (tools/gpu_sweep_*.sh; profiles/round1_scaling.md, round2_*): frame slots in flight : integration blocks per SM and slot
_DEFAULT_KNOBS = {1: (4, 3)}.get(_WORLD, (12, 1))
FRAMES_IN_FLIGHT = int(os.environ.get("GRT_BENCH_FRAMES_IN_FLIGHT", str(_DEFAULT_KNOBS[0])))
BLOCKS_PER_SM = int(os.environ.get("GRT_BENCH_BLOCKS_PER_SM", str(_DEFAULT_KNOBS[1])))
# 1: stripes dealt by measured cost (grt_balanced_stripes).  Off by default: at 8 GPUs the heaviest rank ends where it was (one-GPU
# emulation of all eight ranks, profiles/round2_scaling.md) - the predicted loads are equal to 3 %, the real ones stay 12 % apart
BALANCE = os.environ.get("GRT_BENCH_BALANCE", "0") != "0"
# how a frame is shared out for N > 1: "interleaved" 54-row stripes, "balanced" (the same stripes dealt by measured cost) or "contiguous"
# (one block of rows per rank with boundaries at equal measured cost: one halo per rank instead of one per stripe)
PARTITION = os.environ.get("GRT_BENCH_PARTITION", "balanced" if BALANCE else "contiguous")      # measured at 8 GPUs: 3.42 ms per frame against 3.86 interleaved
REFINE_ROUNDS = int(os.environ.get("GRT_BENCH_REFINE_ROUNDS", "3"))
REFINE_BY = os.environ.get("GRT_BENCH_REFINE_BY", "ms")
COMM_FRAMES = 16             # frame buffers the owner rotates through (>= the largest frames-in-flight setting)
BYTES_PER_RAY = 96           # 56 B read + 40 B written per ray by the integration stage, 0 B per step
ISSUE_SLOTS_PER_CLOCK = 148 * 4   # one warp instruction per scheduler and clock


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def ncu_facts():
    """Per-launch DRAM bytes, warp instructions per frame and per-step SASS instruction counts of k_integrate from the committed
    `ncu --set full` capture of the headline command (newest profiles/round*_ncu_traffic.json)."""
    best = {}
    for name in sorted(os.listdir(os.path.join(ROOT, "profiles"))):
        if name.endswith("_ncu_traffic.json"):
            best = json.load(open(os.path.join(ROOT, "profiles", name)))
            best["file"] = "profiles/" + name
    return best


def workload_config(cfg_id, n_gpus):
    cfg = CONFIGS[cfg_id]
    w, h = cfg["size"]
    return {
        "workload": f"BASELINE config {cfg_id}: {cfg['metric']} {w}x{h} ({cfg['note']}); bundled script geodesic_raytracing_b200/content (a re-statement pinned by value against the reference's scripts/{cfg['metric']}.js, tests/golden/reference_metric_vectors.json); camera (0,0,-4,0) quat (-sin(pi/4),0,0,cos(pi/4)), fov 90, universe 20, default $cfg parameters",
        "mode": "reference defaults: prepass as the metric's JSON asks, adaptive sampling on (threshold 32), anisotropy 8" + (", redshift on" if cfg["features"].get("redshift") else ", redshift off")
                + (f", camera on a geodesic (speed {cfg['geodesic']['speed']}, proper time {cfg['geodesic']['tau']})" if "geodesic" in cfg else ""),
        "background": "synthetic 4096x2048 RGBA8 (the reference's nasa.png is not redistributable)",
        "out_format": "rgba8 (sRGB, the reference's screenshot encoding)",
        "parallelism": "1 GPU" if n_gpus == 1 else f"{n_gpus} GPUs, " + {
            "interleaved": f"interleaved {STRIPE_ROWS}-row stripes + halo",
            "balanced": f"{STRIPE_ROWS}-row stripes + halo dealt to the ranks by measured cost, longest first (grt_balanced_stripes; profile measured during set-up)",
            "contiguous": f"one contiguous block of rows per rank, boundaries at equal measured cost (grt_contiguous_bounds from a stripe-cost profile + {REFINE_ROUNDS} rounds of grt_refine_bounds feedback from the ranks' measured {'times' if REFINE_BY == 'ms' else 'ray-steps'}, all during set-up, untimed)"}[PARTITION]
                       + ", partial prepass per rank, peer stores into rank 0's frame (grt_comm_*), no collective kernel",
        "l2": "inputs larger than L2: every frame reads and writes its slot's ray planes, work lists and shading records (> 260 MB per frame at 4K, 126 MB L2) and the frames in flight use different slots; no explicit flush",
        "pipelining": f"frames are queued back to back, {FRAMES_IN_FLIGHT} in flight on the context's frame slots (the reference keeps two render_state objects in flight, main.cpp:1463-1469); fps_at_2_in_flight and fps_latency (one frame at a time) are reported next to it",
        "integrate_blocks_per_sm_per_slot": BLOCKS_PER_SM,
        "fast_math": True,
    }


class ClockSampler:
    """nvidia-smi clocks + event reasons during the warm-up frames and the timed region, every 50 ms (B200_PROFILING.md recipe)."""

    FIELDS = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(5)
        except Exception:
            self.proc.kill()
        sm, reasons = [], set()
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                out["sm_max_mhz"] = float(p[2])
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def host_threads():
    """Cores this process may use (the cgroup / affinity mask, not the machine's)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline_sample(cfg_id, stride):
    """The reference's own integration kernel (oracle/_ref: cl.cl's do_generic_rays compiled as C++ for the host, one OpenMP
    iteration per work-item) on a bounded subsample of the frame's rays: every `stride`-th pixel in x and y of the pixel grid,
    integration stage only (the stage that is > 90 % of a frame).  The OpenMP thread count is set explicitly to the cores of this
    process - a launcher such as torchrun exports OMP_NUM_THREADS=1.  Falls back to the C++ restatement (kind "port") when no
    oracle/_ref library was prebuilt."""
    import oracle

    cfg = CONFIGS[cfg_id]
    content = grt.Content(grt.CONTENT_DIR)
    metric = content.load_by_name(cfg["metric"])
    threads = host_threads()
    try:
        orc = oracle.reference_for_metric(metric, native=True, threads=threads)     # oracle/_ref: the reference's own cl.cl kernels, compiled for the host
        orc.kind = "reference"
    except FileNotFoundError:
        try:
            orc = oracle.reference_for_metric(metric, threads=threads)              # the generic-ISA build of the same kernels
            orc.kind = "reference"
        except FileNotFoundError:
            orc = oracle.for_metric(metric, native=True, threads=threads)           # the C++ restatement (bit-exact against oracle/_ref, tests/test_reference_kernels.py)
            orc.kind = "port"
    assert orc.info["threads"] == threads, (orc.info["threads"], threads)
    feat = orc.features(max_acceleration_change=metric.info.max_acceleration_change, **cfg["features"])
    w, h = cfg["size"]
    return orc, feat, w // stride, h // stride


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return                                   # one CPU arm per job: the other ranks of a torchrun launch have nothing to add
    cfg = CONFIGS[args.config]
    w, h = cfg["size"]
    total = args.steps + args.warmup
    stride = 4 if total <= 4 else (8 if total <= 30 else 16)
    orc, feat, cw, ch = cpu_baseline_sample(args.config, stride)
    times, steps = [], 0
    for i in range(total):
        r = orc.integrate_sample(w, h, feat, 0, 0, cw, ch, stride)
        if i >= args.warmup:
            times.append(r["seconds"])
            steps += r["trial_steps"]
    t = sum(times)
    value = steps / t / 1e6
    cores = orc.info["threads"]
    sample = f"every {stride}th pixel in x and y of the {w}x{h} grid ({cw}x{ch} = {cw * ch} rays, {steps // max(1, args.steps)} ray-steps per step), integration stage, no prepass cull"
    line = {
        "impl": "reference", "metric": "ray_step_throughput", "value": value, "unit": "Mray-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / max(1, args.steps), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.config, args.gpus),
        "cpu_baseline": {"value": value, "unit": "Mray-steps/s", "cores": cores, "per_core": value / max(1, cores), "kind": orc.kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mray-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": (f"oracle/_ref = the reference's cl.cl compiled as C++ for the host cores (g++ -O3 -march=x86-64-v3, OpenMP over work-items, {cores} threads set explicitly); no OpenCL runtime exists in this image. "
                 "This arm is the integration stage alone on a pixel subsample without the prepass cull; the GPU arm's value is the whole frame (all stages) - the comparison is conservative for the GPU"
                 if orc.kind == "reference" else "C++ restatement of the reference's OpenCL kernels (no prebuilt oracle/_ref library found)"),
    }
    print(json.dumps(line), flush=True)


class Job:
    """One rank of the benchmark: context, communicator, frame submission and timing."""

    def __init__(self, args):
        import torch

        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world != 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.device)
            self.dist = dist
        self.content = grt.Content(grt.CONTENT_DIR)
        self.bg = grt.synthetic_background(4096, 2048)
        self.seq = 0
        self.ctx = None
        self.submit_s, self.submits = 0.0, 0

    def bind(self, cfg_id, partition=None):
        """A fresh context for one BASELINE config: metric, features, camera, background, communicator."""
        partition = partition or PARTITION
        torch = self.torch
        cfg = CONFIGS[cfg_id]
        if self.ctx is not None:
            self.release()
        self.cfg_id, self.cfg = cfg_id, cfg
        self.w, self.h = cfg["size"]
        self.metric = self.content.load_by_name(cfg["metric"])
        ctx = grt.Context(self.local_rank)           # raises when libgrt.so or a B200 is missing: no fallback
        ctx.set_fast_math(True)
        ctx.set_frames_in_flight(FRAMES_IN_FLIGHT)
        ctx.set_tuning(BLOCKS_PER_SM, 0)
        ctx.set_metric(self.metric)
        ctx.set_features(max_acceleration_change=self.metric.info.max_acceleration_change, **cfg["features"])
        bg_dev = torch.from_numpy(self.bg).to(self.device)
        torch.cuda.synchronize()
        ctx.set_background(bg_dev, 0, 4096, 2048)
        ctx.set_camera(CAM_POS, CAM_QUAT)
        if "geodesic" in cfg:
            ctx.snapshot_geodesic(cfg["geodesic"]["speed"], True)
            ctx.set_camera_on_geodesic(cfg["geodesic"]["tau"])
        self.ctx = ctx
        # How the frame is shared out for N > 1 (GRT_BENCH_PARTITION): by measured cost - sort-first rendering with feedback.  The cost
        # profile (trial steps of every half stripe of this view, rendered one by one) is a deterministic function of program, features
        # and camera, so every rank measures the same profile itself; the feedback rounds all-gather what every rank's block cost.  All
        # of it happens during set-up, untimed - what an interactive host refreshes now and then from frames it renders anyway.
        emu_world = int(os.environ.get("GRT_BENCH_EMULATE_WORLD", "0")) if self.world == 1 else 0
        part_world, part_rank = (emu_world, int(os.environ.get("GRT_BENCH_EMULATE_RANK", "0"))) if emu_world else (self.world, self.rank)
        # (emulation, diagnostic only: one GPU renders the stripes one rank of an N-GPU job would own - the per-rank pipeline without the delivery)
        self.stripe_costs = None
        self.partition = "interleaved"
        if part_world > 1 and partition == "contiguous":
            # one block of rows per rank, boundaries at equal measured cost, then REFINE_ROUNDS rounds of feedback from what the blocks really cost
            self.partition = "contiguous"
            self.stripe_costs = ctx.measure_stripe_costs(self.w, self.h, STRIPE_ROWS)
            bounds = grt.contiguous_bounds(self.h, part_world, self.stripe_costs, STRIPE_ROWS)
            self.block_costs = []
            tried = []
            for _ in range(REFINE_ROUNDS + 1):
                # the feedback signal: ray-steps of every block (REFINE_BY=steps), or what matters in the end - the time of every rank's own
                # pipeline over a dozen frames without delivery (REFINE_BY=ms: equal ray-steps are not equal times, the ranks that hold the
                # shadow's edge carry the long rays)
                measured = self.block_ms_of(bounds, part_world, part_rank, bool(emu_world)) if REFINE_BY == "ms" else [m / 1e6 for m in self.block_costs_of(bounds, part_world, part_rank, bool(emu_world))]
                self.block_costs.append([round(m, 2) for m in measured])
                tried.append((max(measured), list(bounds)))
                if len(self.block_costs) <= REFINE_ROUNDS:
                    bounds = grt.refine_bounds(self.h, bounds, measured, self.stripe_costs, STRIPE_ROWS)
            bounds = min(tried, key=lambda t: t[0])[1]          # every candidate was measured (the same numbers on every rank): keep the one whose slowest rank is fastest
            self.bounds = bounds
            self.stripes = [(bounds[part_rank], bounds[part_rank + 1])]
        elif part_world > 1 and partition == "balanced":
            self.partition = "balanced"
            self.stripe_costs = ctx.measure_stripe_costs(self.w, self.h, STRIPE_ROWS)
            self.stripes = grt.balanced_stripes(self.h, part_rank, part_world, self.stripe_costs, STRIPE_ROWS)
        else:
            self.stripes = grt.interleaved_stripes(self.h, part_rank, part_world, STRIPE_ROWS)
        # canonical work count of the frame (single-GPU trace of the whole frame), untimed
        self.frame_steps = 0
        if self.rank == 0:
            full = torch.empty((self.h, self.w, 4), dtype=torch.uint8, device=self.device)
            _, st_full = ctx.render(self.w, self.h, grt.GRT_FORMAT_RGBA8, out=full)
            self.frame_steps = int(st_full.trial_steps)
            del full
        # the communicator: rank 0 owns COMM_FRAMES frame buffers, every rank stores its stripes into them
        handle = ctx.comm_init(self.rank, self.world, self.w, self.h, grt.GRT_FORMAT_RGBA8, COMM_FRAMES)
        ctx.comm_connect(exchange_handles(handle, self.rank, self.world, self.device, self.dist))
        self.seq = 0
        self.host_out = [torch.empty((self.h, self.w, 4), dtype=torch.uint8).pin_memory() for _ in range(COMM_FRAMES)] if self.rank == 0 else None

    def block_costs_of(self, bounds, world, rank, emulated):
        """Trial ray-steps of every rank's block [bounds[r], bounds[r + 1]) rendered with its halo: each rank renders its own and the
        numbers are all-gathered (emulation on one GPU: this process renders all of them)."""
        torch = self.torch
        def cost(r):
            rows = bounds[r + 1] - bounds[r]
            scratch = torch.empty((rows, self.w, 4), dtype=torch.uint8, device=self.device)
            _, st = self.ctx.render(self.w, self.h, grt.GRT_FORMAT_RGBA8, out=scratch, stats=True, stripes=[(bounds[r], bounds[r + 1])])
            return float(st.trial_steps)
        if emulated or self.world == 1:
            return [cost(r) for r in range(world)]
        mine = torch.tensor([cost(rank)], dtype=torch.float64, device=self.device)
        parts = [torch.empty_like(mine) for _ in range(world)]
        self.dist.all_gather(parts, mine)
        return [float(p.item()) for p in parts]

    def block_ms_of(self, bounds, world, rank, emulated, frames=12):
        """ms per frame of every rank's own pipeline on its block [bounds[r], bounds[r + 1]): `frames` frames queued back to back into
        local buffers (no delivery, so no rank waits for another), CUDA events; each rank times its own and the numbers are all-gathered
        (emulation on one GPU: this process times all of them)."""
        torch = self.torch
        def ms(r):
            stripes = [(bounds[r], bounds[r + 1])]
            outs = [torch.empty((bounds[r + 1] - bounds[r], self.w, 4), dtype=torch.uint8, device=self.device) for _ in range(FRAMES_IN_FLIGHT)]
            for k in range(FRAMES_IN_FLIGHT):                       # every slot has its buffers for this geometry
                self.ctx.render(self.w, self.h, grt.GRT_FORMAT_RGBA8, out=outs[k], stats=False, stripes=stripes)
            self.ctx.synchronize(); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e0.record()
            ends = []
            for k in range(frames):
                self.ctx.render(self.w, self.h, grt.GRT_FORMAT_RGBA8, out=outs[k % FRAMES_IN_FLIGHT], stats=False, stripes=stripes)
                if k >= frames - FRAMES_IN_FLIGHT:
                    e = torch.cuda.Event(enable_timing=True); e.record(torch.cuda.ExternalStream(self.ctx.stream, device=self.device)); ends.append(e)
            self.ctx.synchronize(); torch.cuda.synchronize()
            return max(e0.elapsed_time(e) for e in ends) / frames
        if emulated or self.world == 1:
            return [ms(r) for r in range(world)]
        mine = torch.tensor([ms(rank)], dtype=torch.float64, device=self.device)
        parts = [torch.empty_like(mine) for _ in range(world)]
        self.dist.all_gather(parts, mine)
        return [float(p.item()) for p in parts]

    def release(self):
        """Every rank unmaps the others' regions, then - behind a barrier - frees its own."""
        self.barrier()
        self.ctx.comm_disconnect()
        self.barrier_host()
        self.ctx.close()
        self.ctx = None

    def barrier_host(self):
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def barrier(self):
        self.ctx.synchronize()
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def submit(self, to_host=False):
        """Queue the next frame on this rank (asynchronous): stripes rendered straight into the owner's frame buffer; with `to_host`
        the camera is uploaded again (H2D) and rank 0 copies the assembled frame into pinned host memory (D2H) behind it."""
        if to_host and "geodesic" not in self.cfg:
            self.ctx.set_camera(CAM_POS, CAM_QUAT)          # per-frame camera upload on the frame's stream, as main.cpp:2244-2262
        host = self.host_out[self.seq % COMM_FRAMES] if (to_host and self.rank == 0) else None
        t0 = time.perf_counter()
        self.ctx.comm_render(self.stripes, self.seq, host)
        self.submit_s += time.perf_counter() - t0
        self.submits += 1
        self.seq += 1
        return self.torch.cuda.ExternalStream(self.ctx.stream, device=self.device)

    def timed_batch(self, n_frames, in_flight):
        """ms for n_frames queued back to back with `in_flight` frame slots, CUDA events: the start event is recorded on an idle GPU,
        one end event behind the last frame of each slot (on rank 0 that is behind the arrival of every rank's stripes). Max over ranks."""
        torch = self.torch
        self.barrier()
        start = torch.cuda.Event(enable_timing=True)
        start.record(torch.cuda.current_stream(self.device))
        ends = []
        for k in range(n_frames):
            st = self.submit()
            if k >= n_frames - in_flight:
                e = torch.cuda.Event(enable_timing=True)
                e.record(st)
                ends.append(e)
        for e in ends:
            e.synchronize()
        self.barrier()
        ms = torch.tensor([max(start.elapsed_time(e) for e in ends)], dtype=torch.float64, device=self.device)
        self.last_rank_ms = [float(ms.item())]
        if self.world > 1:
            parts = [torch.empty_like(ms) for _ in range(self.world)]
            self.dist.all_gather(parts, ms)
            self.last_rank_ms = [float(p.item()) for p in parts]       # every rank's own clock over the same frames
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())

    def measure(self, steps, warmup, in_flight):
        self.barrier()
        self.ctx.set_frames_in_flight(in_flight)
        self.timed_batch(max(3, warmup, in_flight + 1), in_flight)      # every frame slot allocates its buffers on first use
        return self.timed_batch(steps, in_flight) / steps

    def e2e(self, steps, in_flight):
        torch = self.torch
        self.barrier()
        self.ctx.set_frames_in_flight(in_flight)
        for _ in range(in_flight):
            self.submit(to_host=True)
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.submit(to_host=True)
        self.barrier()
        s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(s, op=self.dist.ReduceOp.MAX)
        return float(s.item())


def exchange_handles(handle, rank, world, device, dist):
    """All ranks' communicator handles in rank order (64 bytes each): an all_gather over the job's process group."""
    if world == 1:
        return handle
    import torch
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return b"".join(bytes(p.cpu().tolist()) for p in parts)


def run_ours(args):
    job = Job(args)
    torch, rank, world = job.torch, job.rank, job.world
    job.bind(args.config)
    ctx = job.ctx

    sampler = ClockSampler(job.local_rank)
    job.barrier()
    ctx.set_frames_in_flight(FRAMES_IN_FLIGHT)
    warmup = max(3, args.warmup, FRAMES_IN_FLIGHT + 1)       # every frame slot allocates its buffers on first use: warm all of them
    if rank == 0:
        sampler.start()       # from the warm-up frames on (the same load): at 8 GPUs the timed region alone is shorter than one sampling period
    job.timed_batch(warmup, FRAMES_IN_FLIGHT)
    total_ms = job.timed_batch(args.steps, FRAMES_IN_FLIGHT)
    rank_ms = list(job.last_rank_ms)
    clocks = sampler.stop() if rank == 0 else None
    launch_ms = ctx.integrate_launch_ms()                  # integration launches of the last frames of the timed region, while frames overlap
    ms_per_step = total_ms / args.steps

    # the reference's depth (two render_state objects in flight) and one frame at a time
    ms_2 = job.measure(args.steps, 1, 2) if FRAMES_IN_FLIGHT != 2 else ms_per_step
    ctx.set_tuning(3, 0)                                   # one frame at a time: three blocks per SM whatever the pipelined setting is
    ms_1 = job.measure(max(4, args.steps // 2), 1, 1)
    ctx.set_tuning(BLOCKS_PER_SM, 0)

    # per-stage times of one frame alone, statistics read back (this rank's stripes)
    job.barrier()
    ctx.set_frames_in_flight(FRAMES_IN_FLIGHT)
    scratch = torch.empty((sum(b - a for a, b in job.stripes), job.w, 4), dtype=torch.uint8, device=job.device)
    lat, last_stats = [], None
    for _ in range(3):
        job.barrier()
        _, last_stats = ctx.render(job.w, job.h, grt.GRT_FORMAT_RGBA8, out=scratch, stats=True, stripes=job.stripes)
        lat.append(last_stats.ms_total)
    stage_log = ctx.launch_log()
    steps_all = torch.tensor([float(last_stats.trial_steps)], dtype=torch.float64, device=job.device)
    if world > 1:
        job.dist.all_reduce(steps_all, op=job.dist.ReduceOp.SUM)

    e2e_s = job.e2e(args.steps, FRAMES_IN_FLIGHT)
    job.barrier()

    line = None
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        frame_steps = job.frame_steps
        value = frame_steps * args.steps / (total_ms * 1e-3) / 1e6
        st = last_stats
        n_launch = max(1, st.integrate_launches)
        rays_per_launch = int(st.rays_integrated) / n_launch
        algo_bytes_per_launch = rays_per_launch * BYTES_PER_RAY
        mean_launch_ms = float(np.mean(launch_ms)) if launch_ms else float("nan")
        achieved_gbs = algo_bytes_per_launch / (mean_launch_ms * 1e-3) / 1e9
        info = job.metric.info
        facts = ncu_facts() if (args.config == 3 and world == 1) else {}      # the capture is of the headline command on one GPU
        clk_hz = (clocks["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)) * 1e6
        warp_instr = facts.get("warp_instructions_per_frame")
        line = {
            "metric": "ray_step_throughput", "value": value, "unit": "Mray-steps/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "fps": 1e3 / ms_per_step, "frames_in_flight": FRAMES_IN_FLIGHT,
            "fps_at_2_in_flight": 1e3 / ms_2, "fps_latency": 1e3 / ms_1, "frame_latency_ms": ms_1,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.config, world),
            "frame": {"ray_steps": frame_steps, "rays_integrated_rank0": int(st.rays_integrated), "extra_rays_rank0": int(st.extra_rays),
                      "rank0_stage_ms_one_frame_alone": stage_log, "rank0_frame_alone_ms": float(np.median(lat)), "sum_ranks_ray_steps": float(steps_all.item()),
                      "timed_region_ms_by_rank": rank_ms, "stripes_rank0": job.stripes, "partition": job.partition, "bounds": getattr(job, "bounds", None), "block_costs_by_round": getattr(job, "block_costs", None), "block_cost_unit": ("ms per frame of the rank's own pipeline" if REFINE_BY == "ms" else "M ray-steps"),
                      "stripe_costs_mray_steps": [round(c / 1e6, 2) for c in job.stripe_costs] if job.stripe_costs else None, "host_ms_per_submit_rank0": 1e3 * job.submit_s / max(1, job.submits)},
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved_gbs / peaks["hbm_gbs"],
                         "traffic": facts.get("mean_dram_bytes_per_launch"), "peak_kind": peak_kind,
                         "kernel": "k_integrate", "algorithmic_bytes_per_launch": algo_bytes_per_launch, "rays_per_launch": rays_per_launch,
                         "launch_ms": mean_launch_ms, "launches_timed": len(launch_ms),
                         "launch_ms_basis": f"CUDA events around every k_integrate launch of the last {FRAMES_IN_FLIGHT} frames of the timed region (rank 0), frames overlapping as in `value`",
                         "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean of the integrate launches of one frame; committed capture of this command, null when the run is not that command)",
                         "note": "register-resident loop: 96 B per ray and 0 B per ray-step, so the HBM fraction is tiny by construction; the kernel is FP32-issue bound - see roofline_fp32"},
            "roofline_fp32": {"bound": "issue", "unit": "G warp-instructions/s",
                              "achieved": (warp_instr * args.steps / (total_ms * 1e-3) / 1e9) if warp_instr else None,
                              "peak": ISSUE_SLOTS_PER_CLOCK * clk_hz / 1e9,
                              "frac": (warp_instr * args.steps / (total_ms * 1e-3)) / (ISSUE_SLOTS_PER_CLOCK * clk_hz) if warp_instr else None,
                              "warp_instructions_per_frame": warp_instr, "sass_instr_per_step": facts.get("sass_instr_per_step"), "sass_fp_instr_per_step": facts.get("sass_fp_instr_per_step"),
                              "ops_per_step_emitted": info.flops_accel + 110, "ops_per_step_reference_formulation": info.flops_expr + 2090,
                              "gsteps_per_s": frame_steps * args.steps / (total_ms * 1e-3) / 1e9,
                              "source": (facts.get("file") or "no capture for this command") + ": smsp__inst_executed.sum over the frame's kernels (the frame is deterministic); clock = median clocks.sm sampled during the timed region; peak = 148 SMs x 4 schedulers x clock"},
            "e2e": {"value": frame_steps * args.steps / e2e_s / 1e6, "unit": "Mray-steps/s", "fps": args.steps / e2e_s,
                    "h2d_bytes_per_step": 48 if "geodesic" not in job.cfg else 0, "d2h_bytes_per_step": job.h * job.w * 4,
                    "note": "whole frame through grt_comm_render into pinned host memory on rank 0; the reference arm it is compared with is the integration stage alone on a strided pixel sample of the same grid without prepass cull (conservative for the GPU)"},
            "gpu_launches": int(st.total_launches) * args.steps,
            "clocks": clocks,
            "kernel_info": {k: v for k, v in ctx.kernel_info().items() if k in ("k_integrate", "integrate_blocks_per_sm")},
        }

    if args.config == 3 and not args.no_side_measurements:
        # side measurements after the headline numbers; never allowed to break the line
        side = {}
        try:
            if world == 1:
                # SURVEY.md 8(d) asks for both operating modes: the same frame with adaptive sampling off (every pixel traced)
                ctx.set_features(max_acceleration_change=job.metric.info.max_acceleration_change, adaptive_sampling=0)
                full = torch.empty((job.h, job.w, 4), dtype=torch.uint8, device=job.device)
                _, st_off = ctx.render(job.w, job.h, grt.GRT_FORMAT_RGBA8, out=full)
                del full
                ms_off = job.measure(6, 1, FRAMES_IN_FLIGHT)
                side["full_sampling"] = {"mode": "adaptive sampling off: every pixel traced", "ray_steps": int(st_off.trial_steps), "rays_integrated": int(st_off.rays_integrated),
                                         "ms_per_frame": ms_off, "fps": 1e3 / ms_off, "value": int(st_off.trial_steps) / ms_off / 1e3, "unit": "Mray-steps/s",
                                         "frame_alone_ms": float(st_off.ms_total)}
        except Exception as exc:   # noqa: BLE001
            side["full_sampling"] = {"error": str(exc)}
        try:
            # BASELINE config 5 on the same GPUs: double_unequal_kerr 3840x2160, row-tiled (register-bound kernel: 167 registers, 3 blocks/SM)
            job.bind(5, partition="interleaved")      # no horizon, flat cost profile: the plain stripes (and no set-up collectives inside a side measurement)
            ms5 = job.measure(8, 2, FRAMES_IN_FLIGHT)
            ms5_1 = job.measure(4, 1, 1)
            if rank == 0:
                side["config5"] = {"workload": workload_config(5, world)["workload"], "ray_steps": job.frame_steps, "ms_per_step": ms5, "fps": 1e3 / ms5,
                                   "value": job.frame_steps / ms5 / 1e3, "unit": "Mray-steps/s", "fps_latency": 1e3 / ms5_1, "n_gpus": world,
                                   "kernel_info": {k: v for k, v in job.ctx.kernel_info().items() if k in ("k_integrate", "integrate_blocks_per_sm")}}
        except Exception as exc:   # noqa: BLE001
            side["config5"] = {"error": str(exc)}
        if rank == 0:
            line.update(side)

    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            stride = 4       # 518 400 rays of the 4K grid: 5 - 15 s on the box's host cores (a bounded sample, about 100 core-seconds)
            orc, feat, cw, ch = cpu_baseline_sample(args.config, stride)
            w, h = CONFIGS[args.config]["size"]
            r = orc.integrate_sample(w, h, feat, 0, 0, cw, ch, stride)
            line["cpu_baseline"] = {
                "value": r["trial_steps"] / r["seconds"] / 1e6, "unit": "Mray-steps/s", "cores": orc.info["threads"],
                "per_core": r["trial_steps"] / r["seconds"] / 1e6 / max(1, orc.info["threads"]), "kind": orc.kind,
                "sample": f"every {stride}th pixel in x and y of the {w}x{h} grid ({cw * ch} rays, {r['trial_steps']} ray-steps, {r['seconds']:.1f} s), integration stage, no prepass cull"}
        print(json.dumps(line), flush=True)
    job.release()
    if world > 1:
        job.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS), help="BASELINE.json config to time (default 3: the headline kerr_schild 4K frame)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-measurements", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
