#!/usr/bin/env python
"""bench.py — Mrays/s of the per-pixel render path on N B200s (one process per GPU).

Workload (BASELINE.json configs[1]): RTIOW book-1 cover scene (~484 analytic
spheres, mixed materials, thin lens), 1200x800, 500 spp, depth 50. One "step" is
one full frame: camera in, packed uint32 framebuffer out.

  value : device-timed (CUDA events on the render stream) whole-job Mrays/s with
          the scene and frame resident in HBM; rays = ray segments the kernel counted.
  e2e   : the same frame through the reference-facing C-ABI call rt3_render() with a
          HOST frame buffer (N=1), or device render + NCCL gather + D2H into pinned
          host memory on rank 0 (N>1); host<->device copies inside the timed region.
  --impl reference : the CPU implementation of the same path (the oracle port of the
          RTIOW semantics -- the reference itself has no bounce loop) on all host threads.

N>1: the frame's row tiles are dealt round-robin to the ranks (no data-path collective
while rendering); the packed framebuffer is gathered with NCCL at frame end. The frame
is fixed, so scaling is strong.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, SPP, DEPTH, SEED, TILE_ROWS = 1200, 800, 500, 50, 1, 2
FLOP_PER_SPHERE_TEST, FLOP_PER_FACE_TEST = 17, 45
CPU_SAMPLE = dict(width=480, height=320, spp=48)    # bounded sample of the same workload for the CPU legs (~20 M ray segments)
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12
NCU_DRAM_BYTES_PER_LAUNCH = 23199488 + 0            # profiles/r01_v3_ncu_summary.txt (N = 1, the BASELINE config)


def hbm_line(kernel_ms):
    """Achieved DRAM bandwidth of the dominant kernel (ncu traffic / event time) against the measured copy bandwidth."""
    peak, src = 6650.0, "fallback of /opt/skills/guides/B200_PROFILING.md"
    try:
        peak, src = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        pass
    achieved = NCU_DRAM_BYTES_PER_LAUNCH / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    return {"achieved_gbs": achieved, "peak_gbs": peak, "frac": achieved / peak, "peak_source": src}


def workload_config():
    which = "configs[1]" if (W, H, SPP) == (1200, 800, 500) else "configs[3]: not the configuration the metric is quoted on"
    return {"workload": f"RTIOW book-1 cover scene, {W}x{H}, {SPP} spp, depth {DEPTH} (BASELINE {which})",
            "width": W, "height": H, "spp": SPP, "max_depth": DEPTH, "tile_rows": TILE_ROWS,
            "l2": "256 MiB buffer written between timed steps (L2 flush); scene+accumulators re-read from HBM"}


class ClockSampler:
    """Samples SM clocks / throttle reasons during the timed region (nvidia-smi, 200 ms)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 7 and r[0].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(int(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i] == "Active" for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(rows[0][1]), "reasons": reasons, "samples": len(rows),
                "power_w_max": max(float(r[2]) for r in rows)}


def usable_cores():
    """Host cores this process may actually use: the affinity mask, capped by the cgroup CPU quota when there is one."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            fields = open(path).read().split()
            if path.endswith("cpu.max"):
                quota, period = fields[0], float(fields[1])
            else:
                quota, period = fields[0], float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if quota not in ("max", "-1") and period > 0:
                n = max(1, min(n, int(float(quota) / period + 0.5)))
            break
        except (OSError, ValueError, IndexError):
            continue
    return n


def cpu_port_rate(n_threads, steps=1, warmup=0):
    """Mrays/s of the oracle port on a bounded sample of the workload (all ray segments counted)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import rt3_b200  # noqa: F401
    from rt3_b200 import abi, scenes
    import oraclelib
    w, h, spp = CPU_SAMPLE["width"], CPU_SAMPLE["height"], CPU_SAMPLE["spp"]
    scene, cam = scenes.rtiow_cover(w, h)
    params = abi.make_params(w, h, mode=abi.MODE_PATHTRACE, spp=spp, max_depth=DEPTH, seed=SEED)
    for _ in range(warmup):
        oraclelib.oracle_pathtrace(scene, cam, params, n_threads=n_threads)
    t0 = time.perf_counter()
    rays = 0
    for _ in range(steps):
        _, _, r = oraclelib.oracle_pathtrace(scene, cam, params, n_threads=n_threads)
        rays += r
    dt = time.perf_counter() - t0
    sample = f"same scene and camera at {w}x{h}, {spp} spp, depth {DEPTH}: {rays // max(steps, 1)} ray segments per step"
    return rays / dt / 1e6, dt / max(steps, 1), sample


def reference_native_rate():
    """The compiled reference's own render loop (oracle/_ref) on its default-scene shape: 1 core, 1 spp, depth 1."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import oraclelib
        if not oraclelib.have_ref() or not os.path.exists(oraclelib.REF_TEDDY):
            return None
        rs = oraclelib.RefScene()
        rs.add_object(oraclelib.REF_TEDDY, (0, 0, -3), 1.0 / 17.0, (1, 0, 0))
        rs.add_sphere((-2, 0, -5), 1.0, 8, 8, (0, 0, 1))
        rs.prerender()
        w, h = 200, 113
        _, sec = rs.render(w, h)
        return {"kind": "reference", "what": f"reference SequentialRenderer::render, Main.cpp default scene (3288 faces), {w}x{h}, 1 spp, depth 1",
                "cores": 1, "value": w * (h - 1) / sec / 1e6, "unit": "Mrays/s", "seconds": round(sec, 3)}
    except Exception as e:  # the reference leg is informative only
        return {"kind": "reference", "error": str(e)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_threads = usable_cores()
    rate, sec_per_step, sample = cpu_port_rate(n_threads, steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": rate, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(),
        "cpu_baseline": {"value": rate, "unit": "Mrays/s", "cores": n_threads, "kind": "port", "sample": sample,
                         "note": "the reference has no bounce loop/materials (raytracer_v4.glsl:279); this is the CPU restatement of the same path"},
        "e2e": {"value": rate, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_native": reference_native_rate(),
    }
    print(json.dumps(line), flush=True)


def main():
    global W, H, SPP
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="rt3", choices=["rt3", "reference"])
    ap.add_argument("--spp", type=int, default=SPP, help="override samples per pixel (non-default values are not the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="c2", choices=["c2", "c4"],
                    help="c2 (default, the metric's configuration) or c4: the same scene at 3840x2160, 1024 spp (BASELINE configs[3], the multi-GPU one)")
    ap.add_argument("--gather", choices=("peer", "nccl"), default="peer",
                    help="N>1 frame end: 'peer' = every rank's kernels store their rows into rank 0's frame over NVLink (CUDA IPC mapping) "
                         "and the frame ends with a one-element all-reduce; 'nccl' = pack + NCCL gather + de-interleave")
    ap.add_argument("--bvh", action="store_true", help="time the hierarchy path (RT3_FLAG_BVH) instead of the brute-force sweep: not the headline configuration")
    args = ap.parse_args()
    if args.config == "c4":
        W, H, SPP = 3840, 2160, 1024
        if args.spp == 500:
            args.spp = SPP
    if args.impl == "reference":
        run_reference_arm(args)
        return
    # stdout carries exactly one JSON line: whatever libraries print there while we run (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import rt3_b200  # noqa: F401
    from rt3_b200 import abi, distributed, scenes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render core has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    spp = args.spp
    scene, cam = scenes.rtiow_cover(W, H)
    ctx = abi.Context(local_rank)
    ctx.upload(scene)
    params = abi.make_params(W, H, mode=abi.MODE_PATHTRACE, spp=spp, max_depth=DEPTH, seed=SEED, flags=abi.FLAG_BVH if args.bvh else 0,
                             tile_rows=TILE_ROWS, part_index=rank, part_count=world)
    lib = ctx.lib
    stream = torch.cuda.Stream(device=dev)  # a real (non-NULL) stream: kernels, NCCL and the timing events all go here
    torch.cuda.set_stream(stream)
    peer = world > 1 and args.gather == "peer"
    shared = distributed.SharedFrame(ctx, dist, H * W, rank, world, dev) if peer else None
    if shared and not shared.ok:   # agreed on by all ranks: no mapping somewhere -> plain NCCL gather everywhere
        print(f"[bench] rank {rank}: shared frame unavailable ({shared.error}); gathering with NCCL", file=sys.stderr)
        shared.close()
        shared, peer = None, False
    frame = shared.tensor if (peer and rank == 0) else torch.zeros(H * W, dtype=torch.int32, device=dev)
    frame_ptr = shared.ptr if peer else frame.data_ptr()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    my_rows = lib.rt3_partition_rows(H, TILE_ROWS, rank, world)
    max_rows = max(lib.rt3_partition_rows(H, TILE_ROWS, r, world) for r in range(world))
    slab = torch.zeros(max_rows * W, dtype=torch.int32, device=dev)
    host_frame_t = torch.zeros(H * W, dtype=torch.int32).pin_memory()
    host_frame = host_frame_t.numpy().view(np.uint32).reshape(H, W)
    launches_per_step = 3  # clear_accum + pathtrace + resolve

    def device_step():
        """Render this rank's rows. N>1, peer: the rows land in rank 0's frame as they are resolved, the frame ends with a
        stream-ordered one-element all-reduce. N>1, nccl: pack, gather onto rank 0 (NCCL), de-interleave."""
        ctx.render_device(cam, params, frame_ptr, stream.cuda_stream)
        if peer:
            shared.finish()
        elif world > 1:
            ctx.pack_partition(frame.data_ptr(), slab.data_ptr(), W, H, TILE_ROWS, rank, world, stream.cuda_stream)
            gathered = distributed.gather_slabs(dist, slab, rank, world)  # NCCL, frame end only
            if rank == 0:
                for r in range(1, world):
                    ctx.unpack_partition(gathered[r].data_ptr(), frame.data_ptr(), W, H, TILE_ROWS, r, world, stream.cuda_stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(max(args.warmup, 0)):
        flush.fill_(1)
        device_step()
    sync_all()
    rays_step = torch.tensor([ctx.stats().rays], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(rays_step)
    rays_per_step = int(rays_step.item())  # deterministic: identical every step

    # ---- timed: device-resident ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    for s in range(args.steps):
        flush.fill_(s)  # L2 flush, outside the per-step events
        ev[s][0].record(stream)
        device_step()
        ev[s][1].record(stream)
    sync_all()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    st = ctx.stats()  # last step's dominant-kernel time and counters (this rank)
    clocks = sampler.stop() if sampler else None

    # ---- timed: end to end with host buffers ----
    e2e_steps = args.steps
    if world == 1:
        ctx.render(cam, params, out=host_frame)  # untimed: first call allocates the context's own frame buffer
    sync_all()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        if world == 1:
            ctx.render(cam, params, out=host_frame)     # rt3_render: blocking, D2H into the pinned host frame
        else:
            device_step()
            if rank == 0:
                host_frame_t.copy_(frame, non_blocking=True)
            if peer:
                shared.finish()  # the other ranks' next frame may only overwrite the shared frame once rank 0 has read this one
            torch.cuda.synchronize()
    sync_all()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    st_e2e = ctx.stats()

    # ---- N>1: the assembled frame must be the one-GPU frame, bit for bit (untimed; rank 0 renders every row itself) ----
    frame_matches = None
    if world > 1:
        device_step()
        sync_all()
        if rank == 0:
            whole = torch.zeros(H * W, dtype=torch.int32, device=dev)
            one = abi.make_params(W, H, mode=abi.MODE_PATHTRACE, spp=spp, max_depth=DEPTH, seed=SEED, flags=abi.FLAG_BVH if args.bvh else 0, tile_rows=TILE_ROWS)
            ctx.render_device(cam, one, whole.data_ptr(), stream.cuda_stream)
            torch.cuda.synchronize()
            frame_matches = bool(torch.equal(whole, frame))
            if not frame_matches:
                raise SystemExit(f"the frame assembled from {world} partitions differs from the one-GPU frame")
        sync_all()

    # ---- informational: the same frame through the hierarchy (RT3_FLAG_BVH), device-timed, not part of `value` ----
    hierarchy = None
    if world == 1 and not args.bvh:
        bvh_params = abi.make_params(W, H, mode=abi.MODE_PATHTRACE, spp=spp, max_depth=DEPTH, seed=SEED, flags=abi.FLAG_BVH, tile_rows=TILE_ROWS)
        ctx.render_device(cam, bvh_params, frame.data_ptr(), stream.cuda_stream)   # builds the hierarchy, warms up
        torch.cuda.synchronize()
        flush.fill_(7)
        ctx.render_device(cam, bvh_params, frame.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        sb = ctx.stats()
        hierarchy = {"value": sb.rays / (sb.device_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": sb.device_ms, "build_ms": sb.accel_build_ms,
                     "node_visits_per_ray": sb.accel_node_visits / max(sb.rays, 1), "prim_tests_per_ray": sb.accel_prim_tests / max(sb.rays, 1),
                     "note": "same workload through the device-built BVH (identical frame); informational, the metric is the brute-force sweep"}

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = rays_per_step * args.steps / (total_ms * 1e-3) / 1e6
        e2e_value = rays_per_step * e2e_steps / e2e_s / 1e6
        # roofline of the dominant kernel (pathtrace_kernel), this rank's launch
        flops = FLOP_PER_SPHERE_TEST * st.sphere_tests + FLOP_PER_FACE_TEST * st.face_tests
        achieved = flops / (st.trace_kernel_ms * 1e-3) / 1e12 if st.trace_kernel_ms > 0 else 0.0
        peak = ctx.measure_fma_peak()
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(),
            "rays_per_step": rays_per_step, "spheres": scene.n_spheres,
            "clocks": clocks, "gpu_launches": launches_per_step * args.steps + (0 if (world == 1 or peer) else (1 + (world - 1)) * args.steps),
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": ctypes.sizeof(abi.Camera) + ctypes.sizeof(abi.Params),
                    "d2h_bytes_per_step": W * H * 4, "ms_per_step": e2e_s / e2e_steps * 1e3,
                    "last_step_device_ms": st_e2e.device_ms, "last_step_d2h_ms": st_e2e.d2h_ms,
                    "path": "rt3_render (C ABI), pinned host frame" if world == 1 else
                            ("rt3_render_device into rank 0's frame over NVLink (rt3_frame_import) + all-reduce barrier + D2H on rank 0" if peer
                             else "rt3_render_device + NCCL gather + D2H on rank 0")},
            "roofline": {"bound": "fp32-issue", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH if (world == 1 and spp == SPP and args.config == "c2" and not args.bvh) else None,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one pathtrace_kernel launch of this workload, ncu --set full "
                                           "(profiles/r01_v3_ncu_summary.txt); the accumulators, the scene is 13 KB",
                         "kernel": "pathtrace_kernel", "kernel_ms": st.trace_kernel_ms,
                         "algorithmic": f"{FLOP_PER_SPHERE_TEST} FLOP x {st.sphere_tests} ray-sphere tests (rank 0 launch)",
                         "peak_source": "FFMA micro-kernel measured in this run (MEASURED_PEAKS.json has no FP32 figure)",
                         "note": "the path is neither HBM- nor tensor-bound (DRAM 0.002 % busy): achieved = 17 algorithmic FLOP per ray-sphere test / kernel time; "
                                 "the conservative prefilter executes 3 FMA per test, so the fraction can exceed 1 (DESIGN.md 3.1)",
                         "peak_nominal": NOMINAL_FP32_TFLOPS, "frac_of_nominal": achieved / NOMINAL_FP32_TFLOPS,
                         "hbm": hbm_line(st.trace_kernel_ms) if (world == 1 and spp == SPP and args.config == "c2" and not args.bvh) else None},
        }
        if hierarchy:
            line["hierarchy"] = hierarchy
        if world > 1:
            line["frame_matches_single_gpu"] = frame_matches
            line["config"]["frame_end"] = ("each rank's kernels store its rows into rank 0's frame over NVLink (CUDA IPC mapping), one-element all-reduce as barrier"
                                           if peer else "pack + NCCL gather onto rank 0 + de-interleave")
        if args.bvh:
            line["config"]["workload"] += " [--bvh: hierarchy traversal instead of the brute-force sweep; roofline figures do not apply]"
            line["accel"] = {"node_visits": st.accel_node_visits, "prim_tests": st.accel_prim_tests, "build_ms": st.accel_build_ms}
        if spp != SPP:
            line["config"]["workload"] += f" [OVERRIDE spp={spp}: not the BASELINE config]"
            line["config"]["spp"] = spp
        if world == 1 and not args.no_cpu_baseline:
            n_threads = usable_cores()
            rate, _, sample = cpu_port_rate(n_threads)
            line["cpu_baseline"] = {"value": rate, "unit": "Mrays/s", "cores": n_threads, "kind": "port", "sample": sample}
            line["reference_native"] = reference_native_rate()
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if shared:
        torch.cuda.synchronize()
        frame = None
        shared.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
