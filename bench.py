#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native path-tracing backend (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the CPU arm (oracle port) on the box's host cores

Workload (BASELINE.json configs[2], the config the north_star target is quoted on): synthetic fractal terrain,
9,999,394 triangles (2236^2 heightfield + a 2-triangle light), 1920x1080; the incoherent-ray set is 2^26 rays per
GPU (origins uniform in the scene AABB inflated 10 %, directions uniform on the sphere).
A "step" = one closest-hit pass (stage B2) over the whole resident ray set.  `value` = incoherent Mrays/s with the rays
resident in HBM; `e2e` = the same metric through foundation_pt_trace_closest with pinned HOST buffers (H2D of the
rays and D2H of the hits inside the timed region).  Also reported: spp/s @1080p (full wavefront loop, 8 bounces,
tile-sharded + NCCL reduce when N > 1), hit-ID mismatches against the CPU oracle (must be 0), the HBM roofline of the
traversal kernel, and the CPU baseline (the oracle on the host cores — substitutes for the unavailable lavapipe arm,
see BASELINE.md §2).  Weak scaling (default): every rank traces its own 2^26 rays; `--scaling strong` splits ONE 2^26-ray set
contiguously over the ranks; no collective on the ray-set path either way.  Extra keys say what the kernel does on rays
that work: `hit_fraction`, V/T split for hits and misses, `hits_only` (the hitting subset alone) and `secondary` (surface-started bounce
rays), each with its own roofline.  For N > 1 rank 0 re-renders the same samples unpartitioned and prints `n_gpu_vs_1_gpu_max_abs_diff`.
Between timed iterations the inputs (2 GiB of rays, 0.59 GB of BVH) exceed the 126 MB L2, so no explicit flush is needed.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "incoherent_mrays_per_s"
UNIT = "Mrays/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--terrain-n", type=int, default=2236)
    ap.add_argument("--log2-rays", type=int, default=26)
    ap.add_argument("--spp", type=int, default=64, help="samples per pixel per render step of the spp/s measurement (0 = skip)")
    ap.add_argument("--bounces", type=int, default=8)
    ap.add_argument("--cpu-log2-rays", type=int, default=26, help="CPU-baseline / parity sample (default: the whole 2^26 ray set, ~5 s on 16 host threads)")
    ap.add_argument("--ref-log2-rays", type=int, default=23, help="rays per step of the --impl reference arm (bounded sample of the same set)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: 2^log2-rays rays per GPU; strong: one 2^log2-rays set split contiguously over the ranks")
    ap.add_argument("--no-extra-sets", action="store_true", help="skip the hits-only and secondary ray sets")
    ap.add_argument("--gather", default="both", choices=["nccl", "direct", "both"], help="N > 1: gather mode(s) of the spp/s measurement")
    return ap.parse_args()


def workload_config(a, world):
    per = "rays/GPU" if a.scaling == "weak" else f"rays total, split contiguously over {world} ranks"
    return {"workload": f"fractal_terrain n={a.terrain_n} ({2 * a.terrain_n ** 2 + 2} tris) 1920x1080; incoherent ray set 2^{a.log2_rays} {per}, "
                        f"closest-hit (stage B2)", "rays_per_gpu": (1 << a.log2_rays) if a.scaling == "weak" else (1 << a.log2_rays) // world, "triangles": 2 * a.terrain_n ** 2 + 2,
            "resolution": "1920x1080", "spp_step": a.spp, "max_bounces": a.bounces,
            "parallelism": f"ray-set sharded x{world} (no collective); frame tile-sharded, owned tiles gathered inside the C ABI (packed ncclSend/Recv, or direct NVLink peer stores)",
            "l2_policy": "inputs (2 GiB rays + 0.59 GB BVH) larger than the 126 MB L2; no flush needed"}


class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop, self._t = [], set(), None, threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8), "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True); self._t.start()

    def stop(self):
        if self._t:
            self._stop.set(); self._t.join()
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch of the traversal kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "trace_kernel_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------------
def run_reference(a, rank, world):
    """CPU arm: the oracle port of the path (the reference has no ray tracer and cannot be built here — DESIGN.md),
    all host threads, bounded sample of the same workload per step.  Rank 0 only."""
    if rank != 0:
        return
    from foundation_b200 import scenes
    from oracle.pt_oracle import OracleScene, hw_threads

    sc = scenes.fractal_terrain(n=a.terrain_n)
    orc = OracleScene(sc)
    lo, hi = scenes.scene_bounds(sc)
    n = 1 << min(a.ref_log2_rays, a.log2_rays)
    rays = scenes.incoherent_rays(lo, hi, n, seed=4)
    cores = hw_threads()
    for _ in range(a.warmup):
        orc.trace_closest(rays[: n // 4])
    t0 = time.perf_counter()
    for _ in range(a.steps):
        orc.trace_closest(rays)
    dt = time.perf_counter() - t0
    v = n * a.steps / dt / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(a, world), reference_arm_rays_per_step=n,
                           reference_arm_note=f"bounded sample: the CPU arm traces the first 2^{int(np.log2(n))} rays of the same set per step (a rate, comparable per ray)"),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"first 2^{int(np.log2(n))} rays of the seed-4 incoherent set per step, CPU oracle BVH traversal, {cores} threads"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def main():
    a = parse()
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        return run_reference(a, rank, world)

    import torch
    from foundation_b200 import distributed as fdist
    from foundation_b200 import pt, scenes

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this backend has no CPU fallback")
    fdist.init_process_group()
    torch.cuda.set_device(local)
    dist = torch.distributed if world > 1 else None

    def barrier():
        if dist:
            dist.barrier(device_ids=[local])      # explicit device: no "using the device under current context" warning after the JSON line
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if not dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if not dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peak, peak_src = measured_peaks()

    # ---- scene (replicated per GPU: every rank runs the same deterministic device build) ----
    sc = scenes.fractal_terrain(n=a.terrain_n)
    with pt.PathTracer(sc.width, sc.height, device=local, seed=1, background=sc.background) as cold:   # first build in the process: module load, fresh
        build_first_ms = float(cold.load(sc).build_ms)                                                # device memory mapped into the scratch pool
    tr = pt.PathTracer(sc.width, sc.height, device=local, seed=1, background=sc.background)
    bs = tr.load(sc)                                                                                  # steady state: what a rebuild costs
    build_steady_ms, sort_steady_ms = float(bs.build_ms), float(bs.sort_ms)
    lo, hi = np.asarray(bs.scene_lo[:], np.float64), np.asarray(bs.scene_hi[:], np.float64)
    nset = 1 << a.log2_rays
    if a.scaling == "weak":
        rays = scenes.incoherent_rays(lo, hi, nset, seed=4 + rank)       # weak scaling: one full set per rank
    else:
        r0, r1 = fdist.ray_slice(nset, rank, world)                      # strong scaling: contiguous 1/world slice of ONE seed-4 set
        rays = scenes.incoherent_rays(lo, hi, nset, seed=4)[r0:r1].copy()
    nrays = len(rays)
    total_rays_per_step = sum_over_ranks(float(nrays))
    pin_rays = torch.empty(nrays * 8, dtype=torch.float32, pin_memory=True)
    pin_rays.numpy()[:] = rays.view(np.float32).reshape(-1)
    pin_hits = torch.empty(nrays * 4, dtype=torch.float32, pin_memory=True)
    tr.rays_upload(rays)

    def time_resident(steps, warmup):
        """W warm-up passes, then `steps` timed passes over the resident ray set: (device ms summed over the steps, launches)."""
        for _ in range(warmup):
            tr.rays_trace_closest()
        ms, ln = 0.0, 0
        for _ in range(steps):
            tr.rays_trace_closest()
            st = tr.stats(); ms += st.last_ms; ln += st.kernel_launches
        return ms, ln

    # ---- device-resident metric: W warm-up steps, then exactly K timed steps ----
    a.warmup = max(a.warmup, 3)                                           # timing rule: at least 3 warm-up steps
    for _ in range(a.warmup):
        tr.rays_trace_closest()
    sampler = ClockSampler(local); sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, launches = time_resident(a.steps, 0)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    dev_ms_max = max_over_ranks(dev_ms)
    wall_ms_max = max_over_ranks(wall_ms)
    value = total_rays_per_step * a.steps / (dev_ms_max * 1e-3) / 1e6
    launches_total = int(sum_over_ranks(launches))
    gh_resident = tr.rays_download_hits(0, nrays)[0] if rank == 0 else None        # the TIMED path's hits, before anything overwrites the buffer

    # ---- e2e: pinned host rays in, host hits out, through the C ABI call a user makes ----
    e2e_steps = max(1, min(a.e2e_steps, a.steps))
    tr.trace_closest_raw(pin_rays.data_ptr(), nrays, pin_hits.data_ptr())          # warm-up (allocations, page touching)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        tr.trace_closest_raw(pin_rays.data_ptr(), nrays, pin_hits.data_ptr())
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e_value = total_rays_per_step * e2e_steps / (e2e_ms * 1e-3) / 1e6
    e2e_hits = pin_hits.numpy().view(scenes.HIT_DTYPE).copy() if rank == 0 else None

    # ---- spp/s @1080p: full wavefront loop, tile-sharded; the owned tiles are gathered into rank 0's frame inside the C ABI ----
    render = {}
    spp_per_s, render_ms, frame_mean, render_rays, rays_per_path, ngpu_diff = None, None, None, 0.0, None, None
    if a.spp > 0:
        modes = ["single"] if world == 1 else (["nccl", "direct"] if a.gather == "both" else [a.gather])
        frames = {}
        for mode in modes:
            t_ctx = tr if mode in ("single", modes[0]) else pt.PathTracer(sc.width, sc.height, device=local, seed=1, background=sc.background)
            if t_ctx is not tr:
                t_ctx.load(sc)
            try:
                dr = fdist.DistributedRenderer(t_ctx, rank, world, direct=(mode == "direct"))
            except pt.FoundationPtError as e:                              # e.g. peer memory not mappable on this box: say so, keep the other mode
                render[mode] = {"error": str(e)}
                continue
            dr.render(0, 1, a.bounces)                                     # warm-up
            barrier()
            t0 = time.perf_counter()
            dr.render(1, a.spp, a.bounces)
            barrier()
            ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
            rst = t_ctx.stats()
            rr = sum_over_ranks(float(rst.rays_extend + rst.rays_shadow))
            paths = sc.width * sc.height * a.spp
            render[mode] = {"spp_per_s": a.spp / (ms * 1e-3), "render_ms": ms, "gather_ms_rank0": float(rst.gather_ms), "mrays_per_s": rr / (ms * 1e-3) / 1e6,
                            "rays_per_path": rr / paths}
            if rank == 0:
                frames[mode] = t_ctx.read_accum().copy()
            if t_ctx is not tr:
                t_ctx.close()
        best = max((m for m in render if "spp_per_s" in render[m]), key=lambda m: render[m]["spp_per_s"], default=None)
        if best:
            spp_per_s, render_ms, rays_per_path = render[best]["spp_per_s"], render[best]["render_ms"], render[best]["rays_per_path"]
            render_rays = render[best]["mrays_per_s"] * render_ms * 1e3
            if rank == 0:
                frame_mean = float(frames[best][..., :3].mean() / (a.spp + 1))
        if world > 1 and rank == 0 and frames:
            # correctness gate of every multi-GPU measurement: the same samples rendered unpartitioned on ONE GPU
            with pt.PathTracer(sc.width, sc.height, device=local, seed=1, background=sc.background) as one:
                one.load(sc)
                one.render(0, 1, a.bounces); one.render(1, a.spp, a.bounces)
                ref = one.read_accum()
            ngpu_diff = {m: float(np.abs(f.astype(np.float64) - ref.astype(np.float64)).max()) for m, f in frames.items()}
        barrier()

    # ---- what the kernel does on rays that work: the hitting subset alone, and surface-started bounce rays (rank 0's GPU) ----
    extra = {}
    if rank == 0 and not a.no_extra_sets:
        hit_mask = gh_resident["prim"] != 0xFFFFFFFF
        extra["hit_fraction"] = float(hit_mask.mean())
        sets = {"hits_only": rays[hit_mask]}
        cam = scenes.camera_rays(sc, 1 << max(16, min(23, a.log2_rays - 3)), 3)
        ch, _ = tr.trace_closest(cam)
        extra["primary_hit_fraction"] = float((ch["prim"] != 0xFFFFFFFF).mean())
        sets["secondary"] = scenes.secondary_rays(sc, cam, ch)
        del cam, ch
        for name, rs in sets.items():
            tr.rays_upload(rs)
            ms, _ = time_resident(5, 2)
            extra[name] = {"rays": int(len(rs)), "mrays_per_s": len(rs) * 5 / (ms * 1e-3) / 1e6, "kernel_ms": ms / 5, "_rays": rs}
        extra["secondary"]["what"] = ("bounce rays regenerated on the host from the primary camera hits (2^23 camera rays at full size) of this scene: origin on the surface, "
                                      "cosine-weighted direction (the Lambert lobe of the render at bounce 1)")
        tr.rays_upload(rays)
    if world > 1:
        barrier()

    if rank != 0:
        barrier()
        if dist:
            dist.destroy_process_group()
        return

    # ---- parity gate + CPU baseline + roofline inputs (rank 0, host cores) ----
    mismatches, max_ulp, brute_mismatches = None, None, None
    cpu = None
    V = T = None
    if not a.no_cpu_baseline:
        from oracle.pt_oracle import OracleScene, hw_threads
        orc = OracleScene(sc)
        ncpu = min(nrays, 1 << a.cpu_log2_rays)
        gh = gh_resident[:ncpu]
        t0 = time.perf_counter()
        oh, oi, cnt = orc.trace_closest(rays[:ncpu], counters=True)
        cpu_s = time.perf_counter() - t0
        mismatches = int((gh["prim"] != oh["prim"]).sum())                           # the device-resident (timed) path
        max_ulp = 0
        for c0 in range(0, ncpu, 1 << 22):                                          # chunked: no 64-bit temporaries of the whole set
            sl = slice(c0, min(ncpu, c0 + (1 << 22)))
            max_ulp = max(max_ulp, int(np.abs(gh["t"][sl].view(np.uint32).astype(np.int64) - oh["t"][sl].view(np.uint32).astype(np.int64)).max()))
        e2e_mismatches = int((e2e_hits["prim"][:ncpu] != oh["prim"]).sum()) + int((e2e_hits["t"][:ncpu].view(np.uint32) != oh["t"].view(np.uint32)).sum())
        mismatches += e2e_mismatches                                                 # the e2e path must agree too
        nb = 128
        bh, _ = orc.trace_closest(rays[:nb], brute=True)
        brute_mismatches = int((gh["prim"][:nb] != bh["prim"]).sum())
        V, T = float(cnt[0]) / ncpu, float(cnt[1]) / ncpu
        cores = hw_threads()
        cpu = {"value": ncpu / cpu_s / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first 2^{int(np.log2(ncpu))} rays of the same ray set, CPU oracle BVH8 traversal on {cores} host threads ({cpu_s:.1f} s); "
                         "substitutes for the lavapipe arm (no Vulkan loader/ICD in the image; the reference has no PT shader)"}
        # V / T split of the headline set, and the algorithmic bytes of the two extra sets (bounded oracle samples)
        if extra:
            hm = oh["prim"] != 0xFFFFFFFF
            nsub = min(ncpu, 1 << 22)
            for label, sel in (("hits", hm[:nsub]), ("misses", ~hm[:nsub])):
                sub = rays[:nsub][sel]
                if len(sub):
                    _, _, c2 = orc.trace_closest(sub, counters=True)
                    extra[f"V_T_{label}"] = [float(c2[0]) / len(sub), float(c2[1]) / len(sub)]
            for name in ("hits_only", "secondary"):
                rs = extra[name].pop("_rays")
                sub = rs[: 1 << 21]
                soh, _, c2 = orc.trace_closest(sub, counters=True)
                v2, t2 = float(c2[0]) / len(sub), float(c2[1]) / len(sub)
                bpr = 32.0 + 16.0 + 80.0 * v2 + 48.0 * t2
                ach = extra[name]["mrays_per_s"] * 1e6 * bpr / 1e9
                extra[name].update({"nodes_per_ray": v2, "tris_per_ray": t2, "bytes_per_ray": bpr, "hit_fraction": float((soh["prim"] != 0xFFFFFFFF).mean()),
                                    "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak}})
    for name in ("hits_only", "secondary"):
        if name in extra:
            extra[name].pop("_rays", None)
    roof = None
    if V is not None:
        bytes_ray = 32.0 + 16.0 + 80.0 * V + 48.0 * T
        kernel_ms = dev_ms / a.steps
        achieved = nrays * bytes_ray / (kernel_ms * 1e-3) / 1e9
        traffic = ncu_traffic()
        tb = traffic.get("dram_bytes_per_launch") if traffic else None
        roof = {"bound": "hbm", "kernel": "k_trace_rays<closest, flat>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "bytes_per_ray": bytes_ray, "nodes_per_ray": V, "tris_per_ray": T, "kernel_ms": kernel_ms,
                "traffic": tb, "traffic_source": traffic.get("source") if traffic else None,
                "dram_frac": (tb * (nrays / float(traffic.get("rays_per_launch", 1 << 26))) / (kernel_ms * 1e-3) / 1e9 / peak) if tb else None,
                "limiter": "latency of dependent random L2/HBM fetches at 32 warps/SM, co-limited by issue slots (ncu: profiles/); frac is on ALGORITHMIC bytes, "
                           "dram_frac on the DRAM bytes ncu measured (upper BVH levels hit L2)"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dev_ms_max / a.steps,
            "wall_ms_per_step": wall_ms_max / a.steps, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, world), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nrays * 32, "d2h_bytes_per_step": nrays * 16, "steps": e2e_steps,
                    "path": "foundation_pt_trace_closest(pinned host rays -> pinned host hits), chunked H2D / kernel / D2H on three streams"},
            "gpu_launches": launches_total, "roofline": roof, "cpu_baseline": cpu,
            "hit_id_mismatches": mismatches, "hit_t_max_ulp": max_ulp, "brute_force_mismatches": brute_mismatches,
            "spp_per_s": spp_per_s, "render_ms": render_ms, "render_mrays_per_s": (render_rays / (render_ms * 1e-3) / 1e6) if render_ms else None,
            "rays_per_path": rays_per_path, "render_modes": render, "n_gpu_vs_1_gpu_max_abs_diff": ngpu_diff,
            "frame_mean_radiance": frame_mean, **extra,
            "build": {"build_ms_first": build_first_ms, "build_ms": build_steady_ms, "sort_ms": sort_steady_ms, "nodes8": int(bs.num_nodes8),
                      "device_bytes": int(bs.device_bytes), "mtris_per_s": bs.num_triangles / (build_steady_ms * 1e-3) / 1e6,
                      "hbm_frac_at_450B_per_tri": bs.num_triangles * 450.0 / (build_steady_ms * 1e-3) / 1e9 / peak}}
    barrier()
    if dist:
        dist.destroy_process_group()
    print(json.dumps(line), flush=True)          # the ONE JSON line, last thing on stdout


if __name__ == "__main__":
    main()
