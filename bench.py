#!/usr/bin/env python3
"""Headline benchmark: kerr_schild at 3840x2160 (BASELINE.json configs[2]) through the C ABI.

A "step" is one frame of the hot path: camera/tetrad -> prepass -> init_rays -> integrate -> render_data ->
adaptive sampling -> integrate -> render_data -> render, in the reference's default operating mode (prepass as the
metric's JSON asks, adaptive sampling on at threshold 32 - main.cpp:1123-1158, graphics_settings.hpp:39-40).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

* value   Mray-steps/s: trial Verlet steps of the frame (SURVEY.md 8(d) unit of work) / device frame time, output and
          all inputs resident in HBM, CUDA events on the launching stream, L2 flushed between frames.
* e2e     the same metric with the frame delivered into pinned HOST memory through grt_render_stripes
          (camera upload H2D, RGBA8 frame D2H, and for N > 1 the NCCL gather to rank 0, all inside the timed region).
* N > 1   the frame is tiled in interleaved 54-row stripes over N ranks (one process per GPU, torchrun), every rank
          traces its stripes (+ halo, + the redundant 1/16 prepass) and the finished tiles are gathered on rank 0 over
          NCCL/NVLink; strong scaling: the work of ONE frame is fixed, value = its single-GPU step count / max-over-ranks time.
* --impl reference   the reference's own integration kernel (oracle/_ref: /root/reference/cl.cl compiled as C++ for the host,
          DESIGN.md) on all host cores, on a bounded pixel subsample of the same frame.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import geodesic_raytracing_b200 as grt  # noqa: E402

WIDTH, HEIGHT = 3840, 2160
METRIC_NAME = "kerr_schild"
STRIPE_ROWS = 54
# Frame slots the benchmark keeps busy (grt_set_frames_in_flight): enough frames to cover the latency-bound critical path of a
# frame (prepass -> first pass -> refinement, each as long as its slowest ray) with bulk work. One frame alone takes ~45 ms on 1 GPU
# for ~30 ms of bulk work, and ~24 ms on 8 GPUs for ~4 ms of bulk work per GPU, so the count grows with the number of GPUs.
_WORLD = int(os.environ.get("WORLD_SIZE", "1"))
# measured optima (tools/gpu_sweep_bench.sh, tools/gpu_sweep_scaling.sh; profiles/round1_scaling.md): slots in flight : blocks per SM and slot
_DEFAULT_KNOBS = {1: (4, 3), 2: (6, 3), 4: (8, 2), 8: (8, 1)}.get(_WORLD, (min(8, 2 * _WORLD + 2), 2))
FRAMES_IN_FLIGHT = int(os.environ.get("GRT_BENCH_FRAMES_IN_FLIGHT", str(_DEFAULT_KNOBS[0])))
BLOCKS_PER_SM = int(os.environ.get("GRT_BENCH_BLOCKS_PER_SM", str(_DEFAULT_KNOBS[1])))
BYTES_PER_RAY = 96          # 56 B read + 40 B written per ray by the integration stage, 0 B per step
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # derived: SMs x lanes x FMA x clocks.max.sm (not in MEASURED_PEAKS.json)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def ncu_facts():
    """Per-launch DRAM bytes and per-step SASS instruction counts of k_integrate from the committed `ncu --set full`
    capture of this same command (profiles/round1_ncu_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "round1_ncu_traffic.json")
    return json.load(open(path)) if os.path.exists(path) else {}


def fp32_microbench():
    """FP32 pipe rate this GPU sustains on independent FFMA/FMUL chains (tools/microbench/fp32_pipes.cu, < 1 s): the
    practical ceiling for an FP32-issue-bound kernel; MEASURED_PEAKS.json has no FP32 figure."""
    exe = os.path.join(ROOT, "tools", "microbench", "fp32_pipes")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
    except Exception:
        return None
    best = {}
    for line in out.splitlines():
        p = line.split()
        if len(p) >= 9 and p[1] == "blocks/SM":
            best[p[0]] = max(best.get(p[0], 0.0), float(p[6]) / 1e3)     # T lane-ops/s
    return best or None


def workload_config(n_gpus):
    return {
        "workload": f"{METRIC_NAME}.js {WIDTH}x{HEIGHT}, camera (0,0,-4,0) quat (-sin(pi/4),0,0,cos(pi/4)), fov 90, universe 20, a=-0.5 rs=1",
        "mode": "reference defaults: prepass on (kerr_schild.json), adaptive sampling on (threshold 32), redshift off, anisotropy 8",
        "background": "synthetic 4096x2048 RGBA8 (the reference's nasa.png is not redistributable)",
        "out_format": "rgba8 (sRGB, the reference's screenshot encoding)",
        "parallelism": "1 GPU" if n_gpus == 1 else f"{n_gpus} GPUs, interleaved {STRIPE_ROWS}-row stripes + halo, redundant prepass, NCCL gather to rank 0",
        "l2": "256 MiB flush write queued before every frame; the per-frame working set (ray planes + shading records, > 260 MB) also exceeds the 126 MB L2",
        "pipelining": f"frames are queued back to back, {FRAMES_IN_FLIGHT} in flight on the context's frame slots (the reference keeps two render_state objects in flight, main.cpp:1463-1469); frame_latency_ms is the time of one frame alone",
        "integrate_blocks_per_sm_per_slot": BLOCKS_PER_SM,
        "fast_math": True,
    }


def gather_frame(local, height, width, rank, world, stripe_rows, device, buffers=None):
    """Gather every rank's stripes on rank 0 and reassemble the frame (rows, width, C). Returns the frame on rank 0, None elsewhere.
    `buffers` (a dict the caller keeps per frame slot) makes the receive bucket and the frame persistent."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return local
    rows = [sum(b - a for a, b in grt.interleaved_stripes(height, r, world, stripe_rows)) for r in range(world)]
    max_rows = max(rows)
    if local.shape[0] < max_rows:
        pad = torch.zeros((max_rows - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=device)
        local = torch.cat([local, pad], 0)
    buffers = buffers if buffers is not None else {}
    if rank == 0 and "bucket" not in buffers:
        buffers["bucket"] = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=device)
        buffers["frame"] = torch.empty((height, width) + tuple(local.shape[2:]), dtype=local.dtype, device=device)
    bucket = list(buffers["bucket"].unbind(0)) if rank == 0 else None
    dist.gather(local, bucket, dst=0)
    if rank != 0:
        return None
    frame = buffers["frame"]
    n_stripes = -(-height // stripe_rows)
    if height % stripe_rows == 0 and n_stripes % world == 0:
        # regular interleave: stripe s = group * world + rank  ->  one strided copy
        groups = n_stripes // world
        src = buffers["bucket"].view(world, groups, stripe_rows, width, -1).permute(1, 0, 2, 3, 4)
        frame.view(groups, world, stripe_rows, width, -1).copy_(src)
    else:
        for r in range(world):
            row = 0
            for (y0, y1) in grt.interleaved_stripes(height, r, world, stripe_rows):
                frame[y0:y1] = bucket[r][row:row + (y1 - y0)]
                row += y1 - y0
    return frame


class ClockSampler:
    """nvidia-smi clocks + event reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
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


def cpu_baseline_sample(stride, threads_note=True):
    """The reference's own integration kernel (oracle/_ref: cl.cl's do_generic_rays compiled as C++ for the host, one OpenMP
    iteration per work-item, all host cores) on a bounded subsample of the frame's rays: every `stride`-th pixel in x and y
    of the 4K pixel grid, integration stage only (the stage that is > 90 % of a frame).  Falls back to the C++ restatement
    (kind "port") when no oracle/_ref library was prebuilt."""
    import oracle

    content = grt.Content(grt.CONTENT_DIR)
    metric = content.load_by_name(METRIC_NAME)
    try:
        orc = oracle.reference_for_metric(metric, native=True)     # oracle/_ref: the reference's own cl.cl kernels, compiled for the host
        orc.kind = "reference"
    except FileNotFoundError:
        orc = oracle.for_metric(metric, native=True)               # the C++ restatement (bit-exact against oracle/_ref, tests/test_reference_kernels.py)
        orc.kind = "port"
    feat = orc.features(max_acceleration_change=metric.info.max_acceleration_change)
    cw, ch = WIDTH // stride, HEIGHT // stride
    return orc, feat, cw, ch


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.steps + args.warmup
    stride = 4 if total <= 4 else (8 if total <= 30 else 16)
    orc, feat, cw, ch = cpu_baseline_sample(stride)
    times, steps = [], 0
    for i in range(total):
        r = orc.integrate_sample(WIDTH, HEIGHT, feat, 0, 0, cw, ch, stride)
        if i >= args.warmup:
            times.append(r["seconds"])
            steps += r["trial_steps"]
    t = sum(times)
    value = steps / t / 1e6
    sample = f"every {stride}th pixel in x and y of the {WIDTH}x{HEIGHT} grid ({cw}x{ch} = {cw * ch} rays, {steps // max(1, args.steps)} ray-steps per step), integration stage, no prepass cull"
    line = {
        "impl": "reference", "metric": "ray_step_throughput", "value": value, "unit": "Mray-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / max(1, args.steps), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "Mray-steps/s", "cores": orc.info["threads"], "per_core": value / max(1, orc.info["threads"]), "kind": orc.kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mray-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": ("oracle/_ref = the reference's cl.cl compiled as C++ for the host cores (g++ -O3 -march=x86-64-v3, OpenMP over work-items); no OpenCL runtime exists in this image"
                 if orc.kind == "reference" else "C++ restatement of the reference's OpenCL kernels (no prebuilt oracle/_ref library found)"),
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world != 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)

    content = grt.Content(grt.CONTENT_DIR)
    metric = content.load_by_name(METRIC_NAME)
    ctx = grt.Context(local_rank)           # raises when libgrt.so or a B200 is missing: no fallback
    ctx.set_fast_math(True)
    ctx.set_frames_in_flight(FRAMES_IN_FLIGHT)
    ctx.set_tuning(BLOCKS_PER_SM, 0)
    ctx.set_metric(metric)
    bg = grt.synthetic_background(4096, 2048)
    bg_dev = torch.from_numpy(bg).to(device)
    torch.cuda.synchronize()
    ctx.set_background(bg_dev, 0, 4096, 2048)
    cam_pos, cam_quat = (0.0, 0.0, -4.0, 0.0), (-0.70710678, 0.0, 0.0, 0.70710678)
    ctx.set_camera(cam_pos, cam_quat)
    stripes = grt.interleaved_stripes(HEIGHT, rank, world, STRIPE_ROWS)
    if world == 1 and os.environ.get("GRT_BENCH_EMULATE_WORLD"):
        # diagnostic only (tools/gpu_probe_pipeline.sh): one GPU renders the stripes rank 0 of an N-GPU job would own, no gather -
        # separates the per-rank pipeline from the collective when studying scaling
        stripes = grt.interleaved_stripes(HEIGHT, 0, int(os.environ["GRT_BENCH_EMULATE_WORLD"]), STRIPE_ROWS)
    rows = sum(b - a for a, b in stripes)
    fmt = grt.GRT_FORMAT_RGBA8
    local = torch.empty((rows, WIDTH, 4), dtype=torch.uint8, device=device)
    stream = torch.cuda.ExternalStream(ctx.stream, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    # canonical work count of the frame (single-GPU trace of the whole frame), untimed
    if rank == 0:
        full = torch.empty((HEIGHT, WIDTH, 4), dtype=torch.uint8, device=device)
        _, st_full = ctx.render(WIDTH, HEIGHT, fmt, out=full)
        frame_steps = int(st_full.trial_steps)
        del full
    else:
        frame_steps = 0

    # two output buffers: frames alternate between the context's two frame slots
    locals_ = [local] + [torch.empty_like(local) for _ in range(FRAMES_IN_FLIGHT - 1)]
    gather_bufs = [dict() for _ in range(FRAMES_IN_FLIGHT)]

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def submit(k):
        """Queue frame k (device-resident output; with N > 1 the NCCL gather to rank 0 is queued behind it). Returns its stream."""
        flush.fill_(k & 1)                                        # evict L2 (256 MiB write) - see config.l2
        ctx.render(WIDTH, HEIGHT, fmt, out=locals_[k % FRAMES_IN_FLIGHT], stats=False, stripes=stripes)
        st = torch.cuda.ExternalStream(ctx.stream, device=device)
        if world > 1:
            with torch.cuda.stream(st):
                gather_frame(locals_[k % FRAMES_IN_FLIGHT], HEIGHT, WIDTH, rank, world, STRIPE_ROWS, device, gather_bufs[k % FRAMES_IN_FLIGHT])
        return st

    def timed_batch(n_frames):
        """ms for n_frames queued back to back (two in flight), measured with CUDA events: the start event is recorded on an idle
        GPU, one end event behind the last frame of each frame slot."""
        barrier()
        start = torch.cuda.Event(enable_timing=True)
        start.record(torch.cuda.current_stream(device))
        ends = []
        for k in range(n_frames):
            st = submit(k)
            if k >= n_frames - FRAMES_IN_FLIGHT:
                e = torch.cuda.Event(enable_timing=True)
                e.record(st)
                ends.append(e)
        for e in ends:
            e.synchronize()
        barrier()
        return max(start.elapsed_time(e) for e in ends)

    timed_batch(max(3, args.warmup))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    batch_ms = timed_batch(args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # single-frame latency and per-stage times: one frame at a time, statistics read back
    lat, last_stats = [], None
    for _ in range(3):
        barrier()
        flush.fill_(1)
        torch.cuda.synchronize()
        _, last_stats = ctx.render(WIDTH, HEIGHT, fmt, out=locals_[0], stats=True, stripes=stripes)
        lat.append(last_stats.ms_total)
    launches = last_stats.total_launches * args.steps
    total_ms = torch.tensor([batch_ms], dtype=torch.float64, device=device)
    steps_all = torch.tensor([float(last_stats.trial_steps)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(steps_all, op=dist.ReduceOp.SUM)
    total_ms = float(total_ms.item())
    ms_per_step = total_ms / args.steps

    # end to end: every frame delivered into pinned HOST memory on rank 0 (two buffers, two frames in flight), host clock
    host_out = [torch.empty((HEIGHT if world == 1 else rows, WIDTH, 4), dtype=torch.uint8).pin_memory() for _ in range(FRAMES_IN_FLIGHT)]
    host_frame = [torch.empty((HEIGHT, WIDTH, 4), dtype=torch.uint8).pin_memory() for _ in range(FRAMES_IN_FLIGHT)] if (world > 1 and rank == 0) else None

    def submit_e2e(k):
        ctx.set_camera(cam_pos, cam_quat)                      # per-frame camera upload (48 B H2D on the frame's stream), as main.cpp:2244-2262
        if world == 1:
            ctx.render(WIDTH, HEIGHT, fmt, out=host_out[k % FRAMES_IN_FLIGHT], stats=False)      # D2H queued behind the frame
        else:
            ctx.render(WIDTH, HEIGHT, fmt, out=locals_[k % FRAMES_IN_FLIGHT], stats=False, stripes=stripes)
            st = torch.cuda.ExternalStream(ctx.stream, device=device)
            with torch.cuda.stream(st):
                frame = gather_frame(locals_[k % FRAMES_IN_FLIGHT], HEIGHT, WIDTH, rank, world, STRIPE_ROWS, device, gather_bufs[k % FRAMES_IN_FLIGHT])
                if rank == 0:
                    host_frame[k % FRAMES_IN_FLIGHT].copy_(frame, non_blocking=True)

    for k in range(FRAMES_IN_FLIGHT):
        submit_e2e(k)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        submit_e2e(k)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        value = frame_steps * args.steps / (total_ms * 1e-3) / 1e6
        st = last_stats
        rays_per_launch = (int(st.rays_integrated)) / max(1, st.integrate_launches)
        integrate_ms = st.ms_integrate + st.ms_prepass          # prepass time is its integrate launch to within microseconds
        algo_bytes = int(st.rays_integrated) * BYTES_PER_RAY            # all integrate launches of the frame
        achieved_gbs = algo_bytes / (integrate_ms * 1e-3) / 1e9        # = per-launch bytes / mean launch duration
        info = metric.info
        flops_ref_step = info.flops_expr + 2090          # reference-formulation count, 16-component metric
        flops_own_step = info.flops_accel + 110          # emitted lowered acceleration + Verlet update + step control + radii
        gsteps = int(st.trial_steps) / (integrate_ms * 1e-3)
        facts = ncu_facts()
        micro = fp32_microbench() if (world == 1 and not args.no_cpu_baseline) else None   # side measurement, like the CPU baseline
        fp_instr = facts.get("sass_fp_instr_per_step")
        line = {
            "metric": "ray_step_throughput", "value": value, "unit": "Mray-steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "fps": 1e3 / ms_per_step, "frame_latency_ms": float(np.median(lat)), "frames_in_flight": FRAMES_IN_FLIGHT, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "frame": {"ray_steps": frame_steps, "rays_integrated_rank0": int(st.rays_integrated), "extra_rays_rank0": int(st.extra_rays),
                      "rank0_stage_ms": ctx.launch_log(), "sum_ranks_ray_steps": float(steps_all.item())},
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved_gbs / peaks["hbm_gbs"], "traffic": facts.get("mean_dram_bytes_per_launch"),
                         "peak_kind": peak_kind, "kernel": "k_integrate (all integrate launches of one frame on rank 0)",
                         "algorithmic_bytes_per_launch": algo_bytes / max(1, st.integrate_launches), "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean of the 3 integrate launches of a frame)", "launch_ms": integrate_ms / max(1, st.integrate_launches), "rays_per_launch": rays_per_launch,
                         "note": "register-resident loop: 96 B per ray and 0 B per ray-step, so the HBM fraction is tiny by construction; the kernel is FP32-issue bound - see fp32"},
            "fp32": {"gsteps_per_s_in_kernel": gsteps / 1e9, "flops_per_step_emitted": flops_own_step, "flops_per_step_reference_formulation": flops_ref_step,
                     "achieved_tflops_emitted": gsteps * flops_own_step / 1e12, "achieved_tflops_reference_formulation": gsteps * flops_ref_step / 1e12,
                     "peak_tflops_derived": FP32_PEAK_TFLOPS, "frac_emitted": gsteps * flops_own_step / 1e12 / FP32_PEAK_TFLOPS,
                     "sass_fp_instr_per_step": fp_instr, "sass_instr_per_step": facts.get("sass_instr_per_step"),
                     "sustained_gsteps_per_s": frame_steps * args.steps / (total_ms * 1e-3) / 1e9,
                     "sustained_fp_tlaneops": (frame_steps * args.steps / (total_ms * 1e-3)) * fp_instr / 1e12 if fp_instr else None,
                     "microbench_tlaneops": micro,
                     "frac_of_microbench_ffma_fmul_mix": ((frame_steps * args.steps / (total_ms * 1e-3)) * fp_instr / 1e12 / micro["mix_fma_mul"]) if (fp_instr and micro and micro.get("mix_fma_mul")) else None,
                     "note": "FP lane-op rate = ray-steps/s x FFMA+FMUL+FADD instructions per step (ncu); the microbenchmark is what independent FFMA/FMUL chains sustain on this GPU"},
            "e2e": {"value": frame_steps * args.steps / e2e_s / 1e6, "unit": "Mray-steps/s", "fps": args.steps / e2e_s, "h2d_bytes_per_step": 48,
                    "d2h_bytes_per_step": HEIGHT * WIDTH * 4},
            "gpu_launches": launches,
            "clocks": clocks,
            "kernel_info": {k: v for k, v in ctx.kernel_info().items() if k in ("k_integrate", "integrate_blocks_per_sm")},
        }
        if world == 1:
            # SURVEY.md 8(d) asks for both operating modes: the same frame with adaptive sampling off (all 8.29 M rays traced, the
            # clean ray-step rate), a side measurement after the headline numbers; never allowed to break the line
            try:
                ctx.set_features(adaptive_sampling=0)
                full = torch.empty((HEIGHT, WIDTH, 4), dtype=torch.uint8, device=device)
                _, st_off = ctx.render(WIDTH, HEIGHT, fmt, out=full)                      # one frame alone, with statistics
                outs_off = [full] + [torch.empty_like(full) for _ in range(min(FRAMES_IN_FLIGHT, 3) - 1)]
                n_off = 4
                ctx.synchronize(); torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e0.record(torch.cuda.current_stream(device))
                ends_off = []
                for k in range(n_off):
                    ctx.render(WIDTH, HEIGHT, fmt, out=outs_off[k % len(outs_off)], stats=False)
                    e = torch.cuda.Event(enable_timing=True); e.record(torch.cuda.ExternalStream(ctx.stream, device=device)); ends_off.append(e)
                ctx.synchronize(); torch.cuda.synchronize()
                ms_off = max(e0.elapsed_time(e) for e in ends_off) / n_off
                line["full_sampling"] = {"mode": "adaptive sampling off: every pixel traced", "ray_steps": int(st_off.trial_steps), "rays_integrated": int(st_off.rays_integrated),
                                         "ms_per_frame": ms_off, "fps": 1e3 / ms_off, "value": int(st_off.trial_steps) / ms_off / 1e3, "unit": "Mray-steps/s",
                                         "frame_latency_ms": float(st_off.ms_total), "frames_timed": n_off}
                del outs_off, full
            except Exception as exc:   # noqa: BLE001
                line["full_sampling"] = {"error": str(exc)}
            finally:
                ctx.set_features(adaptive_sampling=1)
        if world == 1 and not args.no_cpu_baseline:
            stride = 8
            orc, feat, cw, ch = cpu_baseline_sample(stride)
            r = orc.integrate_sample(WIDTH, HEIGHT, feat, 0, 0, cw, ch, stride)
            line["cpu_baseline"] = {
                "value": r["trial_steps"] / r["seconds"] / 1e6, "unit": "Mray-steps/s", "cores": orc.info["threads"],
                "per_core": r["trial_steps"] / r["seconds"] / 1e6 / max(1, orc.info["threads"]), "kind": orc.kind,
                "sample": f"every {stride}th pixel in x and y of the {WIDTH}x{HEIGHT} grid ({cw * ch} rays, {r['trial_steps']} ray-steps, {r['seconds']:.1f} s), integration stage, no prepass cull"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
