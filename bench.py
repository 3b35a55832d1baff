#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: Mrays/s and frame ms of `Camera::render`
(lib/src/camera.rs:76-91), beside the CPU render of the same frame.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl ours|reference]

A *step* is one frame of the workload (default c3 = BASELINE.json configs[2]: the soft-shadow scene at
3840x2160 with a 4x4 = 16-cell jittered area light, reflections, depth 5 — the frame the north-star target is
quoted on).  A *ray* is one World::intersect query: primary + reflect/refract + shadow.

    value        Mrays/s with the scene resident in HBM and the frame left on the device; per-step time = the
                 CUDA-event duration of the render kernel, max over ranks
    e2e          the same frame through Camera::render_b200 (flatten + commit + render + D2H of the f32 and 8-bit
                 canvases into pinned host memory), host wall clock
    roofline     algorithmic FP32 flops of the frame (detailed pass) / kernel time, against
                 the FP32 peak measured live by the K5 FMA micro-benchmark
    cpu_baseline the CPU oracle (a C++ port of the reference; Rust is not installed) on a bounded row sample

Multi-GPU: one process per GPU (torchrun), the frame's 8-row bands interleaved over the ranks (strong scaling of
one frame, no collective on the data path; every rank writes its bands into one shared host canvas).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scene builder name, kwargs, depth, description)
    "c1": ("soft_shadows", dict(width=1000, height=400, u_steps=10, v_steps=10), 5,
           "soft_shadows demo as shipped: 1000x400, 10x10-cell area light, depth 5"),
    "c2": ("reflect_refract", dict(width=1920, height=1080), 5,
           "chapter-11 reflect/refract scene at 1920x1080, point light, depth 5"),
    "c3": ("soft_shadows", dict(width=3840, height=2160, u_steps=4, v_steps=4), 5,
           "soft-shadow + reflection scene at 3840x2160, 4x4 = 16 jittered light cells, depth 5"),
    "c4": ("dragon_element", dict(width=1920, height=1080, n_u=320, n_v=160), 5,
           "synthetic 102k-triangle OBJ mesh in divided groups (here_be_dragons element) at 1920x1080"),
    "c5": ("stress", dict(width=3840, height=2160, n_spheres=100_000), 5,
           "100k-sphere field + cylinders/cones/cubes + CSG + checker plane at 3840x2160"),
}


def sample_clocks(stop: threading.Event, out: list, gpu_index: int) -> None:
    """One streaming nvidia-smi (-lms 50) for the whole timed region; every line is one sample."""
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    try:
        proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(gpu_index),
                                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return

    def reader():
        for line in proc.stdout:
            parts = [p.strip() for p in line.strip().split(",")]
            if len(parts) >= 6:
                out.append(parts)

    t = threading.Thread(target=reader, daemon=True)
    t.start()
    stop.wait()
    proc.terminate()
    t.join(timeout=2)


def clocks_summary(samples: list) -> dict:
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(samples[0][1]) if samples[0][1].isdigit() else None,
            "reasons": reasons, "samples": len(samples)}


def build_scene(api, workload: str):
    from ray_tracer_challenge_b200 import scenes

    name, kw, depth, desc = WORKLOADS[workload]
    cam, world = getattr(scenes, name)(api, **kw)
    return cam, world, depth, desc


def cpu_baseline(workload: str, threads: int, target_seconds: float = 12.0) -> dict:
    """The oracle (C++ port of the reference) on a bounded sample: every `ystep`-th row of the same frame."""
    from tests.oracle_binding import load_oracle

    oracle = load_oracle()
    cam, world, depth, _ = build_scene(oracle, workload)
    h = cam.height_pixels
    oracle.probe.set_threads(threads)
    # calibrate on a very sparse sample, then choose the row stride for ~target_seconds of work
    probe_step = max(1, h // 16)
    t0 = time.perf_counter()
    _, _, st = oracle.probe.render_rows(cam, world, depth, probe_step // 2, h, probe_step)
    dt = max(time.perf_counter() - t0, 1e-4)
    rows_probe = max(1, len(range(probe_step // 2, h - 1, probe_step)))
    per_row = dt / rows_probe
    rows_target = int(max(8, min(h - 1, target_seconds / per_row)))
    ystep = max(1, (h - 1) // rows_target)
    t0 = time.perf_counter()
    _, _, st = oracle.probe.render_rows(cam, world, depth, ystep // 2, h, ystep)
    dt = time.perf_counter() - t0
    rows = len(range(ystep // 2, h - 1, ystep))
    mrays = st.rays / dt / 1e6
    return {"value": round(mrays, 4), "unit": "Mrays/s", "cores": threads, "kind": "port",
            "sample": f"rows {ystep // 2}::{ystep} of the {cam.width_pixels}x{h} frame ({rows} rows, {st.rays} rays, {dt:.2f} s)",
            "est_frame_ms": round(dt * (h - 1) / rows * 1e3, 1),
            # the reference algorithm's own work for the frame (every ray against every object, Appendix E units)
            "reference_flops_per_frame": float(st.flops) * (h - 1) / rows}


def run_reference(args) -> None:
    """--impl reference: the reference's CPU algorithm (oracle port; Rust is not installed) with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tests.oracle_binding import load_oracle

    oracle = load_oracle()
    threads = oracle.probe.max_threads()
    oracle.probe.set_threads(threads)
    cam, world, depth, desc = build_scene(oracle, args.workload)
    h = cam.height_pixels
    # one step = every `ystep`-th row; size the stride so warmup+steps finish within a few minutes
    t0 = time.perf_counter()
    probe_step = max(1, h // 16)
    oracle.probe.render_rows(cam, world, depth, 0, h, probe_step)
    per_row = (time.perf_counter() - t0) / max(1, len(range(0, h - 1, probe_step)))
    budget = 150.0 / max(1, args.steps + args.warmup)
    ystep = max(1, int(np.ceil((h - 1) * per_row / budget)))
    rays = 0
    for _ in range(args.warmup):
        oracle.probe.render_rows(cam, world, depth, 0, h, ystep)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, _, st = oracle.probe.render_rows(cam, world, depth, 0, h, ystep)
        rays += st.rays
    dt = time.perf_counter() - t0
    value = rays / dt / 1e6
    rows = len(range(0, h - 1, ystep))
    line = {
        "impl": "reference", "metric": "Mrays/s (Camera::render: primary + reflect/refract + shadow rays per second)",
        "value": round(value, 4), "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": round(value, 4), "unit": "Mrays/s", "cores": threads, "kind": "port",
                         "sample": f"each step renders rows 0::{ystep} ({rows} of {h - 1} rows) of the frame"},
        "e2e": {"value": round(value, 4), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--fma", action="store_true", help="time the FMA-contracted kernel build instead of the IEEE one")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the renderer has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import ray_tracer_challenge_b200 as rt

    api = rt.new_session()
    api.set_render_options(device_ids=[local_rank], fma=args.fma)
    cam, world, depth, desc = build_scene(api, args.workload)
    w, h = cam.width_pixels, cam.height_pixels
    prepared = cam.prepare(world)
    n_shards = world_size if world_size > 1 else 0

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # ---- rays / flops of the frame (deterministic): one detailed pass, untimed
    detail = None
    if world_size == 1:
        prepared.render(depth, want_rgb=False, want_u8=False, detailed=True, fma=args.fma)
        detail = prepared.last_stats.as_dict()

    # ---- device-resident throughput
    for _ in range(args.warmup):
        prepared.render(depth, want_rgb=False, want_u8=False, shard=rank, n_shards=n_shards, fma=args.fma)
    clock_samples: list = []
    stop = threading.Event()
    sampler = threading.Thread(target=sample_clocks, args=(stop, clock_samples, local_rank), daemon=True)
    sampler.start()
    # keep the GPU under the same load while the sampler spins up (nvidia-smi needs ~100 ms for its first line)
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < 0.4:
        prepared.render(depth, want_rgb=False, want_u8=False, shard=rank, n_shards=n_shards, fma=args.fma)
    barrier()
    kernel_ms, rays, launches = 0.0, 0, 0
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        prepared.flush_l2()  # evict the scene and the previous frame between timed iterations
        prepared.render(depth, want_rgb=False, want_u8=False, shard=rank, n_shards=n_shards, fma=args.fma)
        kernel_ms += prepared.last_stats.kernel_ms
        rays += prepared.last_stats.rays
        launches += prepared.last_stats.launches
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    stop.set()
    sampler.join(timeout=2)

    from ray_tracer_challenge_b200 import multi

    kernel_ms_max = multi.reduce_scalar(dist, kernel_ms, "max", "cuda")
    rays_total = multi.reduce_scalar(dist, rays, "sum", "cuda")
    value = rays_total / (kernel_ms_max * 1e-3) / 1e6

    # ---- end to end through Camera::render_b200 into pinned host canvases
    lib = rt.device_library()
    import ctypes as C

    lib.rtc_host_alloc.restype = C.c_void_p
    lib.rtc_host_alloc.argtypes = [C.c_size_t]
    lib.rtc_host_free.argtypes = [C.c_void_p]
    e2e = None
    d2h = w * h * 3 * 4 + w * h * 3
    if world_size == 1:
        p_rgb, p_u8 = lib.rtc_host_alloc(w * h * 12), lib.rtc_host_alloc(w * h * 3)
        rgb = np.ctypeslib.as_array(C.cast(p_rgb, C.POINTER(C.c_float)), shape=(h, w, 3))
        u8 = np.ctypeslib.as_array(C.cast(p_u8, C.POINTER(C.c_uint8)), shape=(h, w, 3))
        stats = rt.SgStats()
        from ray_tracer_challenge_b200.api import U8P, fptr

        def one_shot():
            api.check(api.lib.sg_camera_render(api.ctx, cam.handle, world.handle, depth, fptr(rgb), u8.ctypes.data_as(U8P),
                                               C.byref(stats)))
            return stats.rays

        for _ in range(2):
            one_shot()
        e_rays, t0 = 0, time.perf_counter()
        e_steps = max(3, min(args.steps, 10))
        for _ in range(e_steps):
            e_rays += one_shot()
        e_dt = time.perf_counter() - t0
        prims, nodes, refs, _, counts = api.flatten(world)
        h2d = counts[0] * C.sizeof(rt.RtcPrim) + counts[1] * C.sizeof(rt.RtcNode) + counts[2] * 4 + counts[3] * 44
        e2e = {"value": round(e_rays / e_dt / 1e6, 3), "unit": "Mrays/s", "ms_per_frame": round(e_dt / e_steps * 1e3, 3),
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "path": "Camera::render_b200: flatten + commit + render + D2H (f32 + u8 canvases, pinned)"}
        # the same with the scene kept resident (animation-style repeated renders)
        for _ in range(2):
            prepared.render(depth, out_rgb=rgb, out_u8=u8, fma=args.fma)
        t0 = time.perf_counter()
        for _ in range(e_steps):
            prepared.render(depth, out_rgb=rgb, out_u8=u8, fma=args.fma)
        e2e["resident_scene_ms_per_frame"] = round((time.perf_counter() - t0) / e_steps * 1e3, 3)
        # the Canvas the reference returns is the f32 plane alone (canvas.rs:6-10; 8-bit values are made by to_ppm)
        for _ in range(2):
            prepared.render(depth, out_rgb=rgb, want_u8=False, fma=args.fma)
        t0 = time.perf_counter()
        for _ in range(e_steps):
            prepared.render(depth, out_rgb=rgb, want_u8=False, fma=args.fma)
        e2e["resident_scene_f32_only_ms_per_frame"] = round((time.perf_counter() - t0) / e_steps * 1e3, 3)
        lib.rtc_host_free(p_rgb)
        lib.rtc_host_free(p_u8)
    else:
        # every rank copies its bands straight into one shared host canvas (POSIX shared memory, page-locked by
        # each rank): the "simple host gather" of SURVEY.md §8e with no extra copy
        canvas = multi.open_shared_canvas(dist, rank, f"rtc_bench_{os.environ.get('MASTER_PORT', '0')}", w, h)
        rgb, u8 = canvas.rgb, canvas.u8
        lib.rtc_host_register.argtypes = [C.c_void_p, C.c_size_t]
        lib.rtc_host_unregister.argtypes = [C.c_void_p]
        registered = lib.rtc_host_register(canvas.address, canvas.nbytes) == 0
        for _ in range(2):
            prepared.render(depth, out_rgb=rgb, out_u8=u8, shard=rank, n_shards=n_shards, fma=args.fma)
        barrier()
        e_steps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(e_steps):
            prepared.render(depth, out_rgb=rgb, out_u8=u8, shard=rank, n_shards=n_shards, fma=args.fma)
        barrier()
        e_dt = multi.reduce_scalar(dist, time.perf_counter() - t0, "max", "cuda")
        e2e = {"value": round(rays_total / args.steps * e_steps / e_dt / 1e6, 3), "unit": "Mrays/s",
               "ms_per_frame": round(e_dt / e_steps * 1e3, 3), "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": int(d2h),
               "path": "resident scene, each rank renders its bands and copies them into one shared pinned host canvas"
                       + ("" if registered else " (cudaHostRegister failed: pageable copy)")}
        if registered:
            lib.rtc_host_unregister(canvas.address)
        del rgb, u8
        dist.barrier()
        canvas.close()

    # ---- roofline + CPU baseline (rank 0, N = 1 only)
    roofline, cpu = None, None
    if rank == 0:
        tflops = C.c_double()
        mhz = C.c_double()
        lib.rtc_measure_fp32_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.rtc_measure_fp32_peak(local_rank, C.byref(tflops), C.byref(mhz))
        props = torch.cuda.get_device_properties(local_rank)
        nominal = props.multi_processor_count * 128 * 2 * mhz.value * 1e6 / 1e12
        if detail is not None:
            ms_frame = kernel_ms_max / args.steps
            achieved = detail["flops"] / (ms_frame * 1e-3) / 1e12
            prof = {}
            try:  # the committed ncu capture of this workload's kernel (tools/ncu_profile_json.py)
                with open(os.path.join(ROOT, "profiles", f"r01_{args.workload}_profile.json")) as fh:
                    prof = json.load(fh)
            except Exception:
                pass
            roofline = {"bound": "fp32", "achieved": round(achieved, 3), "peak": round(tflops.value, 2), "unit": "TFLOP/s",
                        "frac": round(achieved / tflops.value, 4), "traffic": prof.get("dram_bytes"),
                        "traffic_source": prof.get("source"),
                        "issue_slot_utilisation_pct": prof.get("issue_active_pct"),
                        "fma_pipe_utilisation_pct": prof.get("pipe_fma_pct"),
                        # what the hardware executed (ncu SASS counts of the committed capture: FADD + FMUL + 2 FFMA + FMNMX +
                        # MUFU, predicated-on threads) over this run's kernel time
                        "executed_fp32_tflops": round(prof["executed_fp32_flops"] / (ms_frame * 1e-3) / 1e12, 3) if prof.get("executed_fp32_flops") else None,
                        "what": "achieved = algorithmic FP32 flops of the units the kernel EXECUTED (SURVEY.md Appendix E table x the "
                                "detailed pass's counters: tests skipped by the shadow filter's bundle reject are not counted) / kernel time",
                        "peak_kind": "measured live: K5 FMA micro-benchmark (MEASURED_PEAKS.json has no FP32 figure)",
                        "nominal_peak": round(nominal, 2), "frac_of_nominal": round(achieved / nominal, 4),
                        "flops_per_frame": detail["flops"], "rays_per_frame": detail["rays"],
                        "flops_per_ray": round(detail["flops"] / max(detail["rays"], 1), 1),
                        "node_visits_per_ray": round(detail["node_visits"] / max(detail["rays"], 1), 2),
                        "prim_tests_per_ray": round(sum(detail["prim_tests"]) / max(detail["rays"], 1), 2)}
        if world_size == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(args.workload, 1)
            if roofline and cpu.get("reference_flops_per_frame"):
                ref_flops = cpu["reference_flops_per_frame"]
                ref_tf = ref_flops / (kernel_ms_max / args.steps * 1e-3) / 1e12
                roofline["reference_algorithm"] = {
                    "flops_per_frame": ref_flops, "achieved": round(ref_tf, 3), "frac": round(ref_tf / tflops.value, 4),
                    "what": "the same frame's flops as the REFERENCE algorithm spends them (the oracle's counters: every ray "
                            "against every object) / our kernel time: work-equivalent throughput, not hardware utilisation"}
            try:
                from tests.oracle_binding import load_oracle

                n_threads = load_oracle().probe.max_threads()
                cpu["all_cores"] = cpu_baseline(args.workload, n_threads, 6.0)
            except Exception as e:  # pragma: no cover
                cpu["all_cores"] = {"error": str(e)}

    if rank == 0:
        line = {
            "metric": "Mrays/s (Camera::render: primary + reflect/refract + shadow rays per second)",
            "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(kernel_ms_max / args.steps, 4), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "resolution": [w, h], "depth": depth,
                       "kernel_build": "fma-contracted" if args.fma else "ieee (no contraction, bit-exact vs the oracle)",
                       "l2": "flushed between timed iterations (256 MiB memset)",
                       "sharding": f"{world_size} rank(s), interleaved 8-row bands"},
            "rays_per_frame": rays_total / args.steps, "wall_ms_per_step_incl_flush": round(wall_ms / args.steps, 3),
            "e2e": e2e, "gpu_launches": int(multi.reduce_scalar(dist, launches, "sum", "cuda")) if False else launches * world_size, "clocks": clocks_summary(clock_samples),
        }
        if roofline:
            line["roofline"] = roofline
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    prepared.release()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
