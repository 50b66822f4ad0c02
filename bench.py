#!/usr/bin/env python
"""bench.py — headline benchmark of the wavefront path tracer (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[2] — the synthetic 1 387 526-triangle instanced scene, Disney BSDF + NEE,
1920x1080, Independent sampler seed 19980810, path depth 10 — rendered in *steps* of SPP_PER_STEP samples per pixel (the
full config is 1024 spp = 4 such steps; step s renders sample indices [s*SPP, (s+1)*SPP)).  One step = one pass of the hot
path over one batch of 1920*1080*SPP camera samples.  For N > 1 the same frame is sharded by interleaved 32x32 pixel tiles
(strong scaling) and the raw film is sum-reduced to rank 0 once, after the K-th step, inside the timed region (configs[4]).

value  : Msamples/s, scene and path state resident in HBM (device timing bracketed by barrier + synchronize).
e2e    : Msamples/s through the C-ABI with HOST buffers — every step uploads the flattened scene from host memory
         (lrk_upload_scene), renders, and downloads the normalised film (lrk_download_film).
roofline: the closest-hit traversal kernel: algorithmic bytes (48 B per ray + 64 B per BVH node visited
         + 48 B per triangle tested + 64 B per instance entered, counted by the kernel's counting variant on the same
         deterministic workload) / CUDA-event time of its launches inside the timed region, against the measured HBM peak.
cpu_baseline: the CPU oracle (a port of the reference estimator; the reference's own `cpu` backend cannot be built here)
         on all host cores, on a bounded tile sample of the same frame.
--impl reference: times that CPU implementation as its own arm (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

METRIC = "Msamples/s (wavefront path tracing, 1.39M-triangle instanced scene, Disney + NEE, 1920x1080)"
UNIT = "Msamples/s"
WIDTH, HEIGHT = 1920, 1080
SPP_PER_STEP = 256
FULL_SPP = 1024


def build_scene():
    from luisarender_b200 import scenes
    from luisarender_b200.api import Scene

    return Scene.from_source(scenes.instanced_spheres(resolution=(WIDTH, HEIGHT), spp=FULL_SPP, seed=1), REPO)


def workload_config(n_gpus: int) -> dict:
    return {
        "workload": "BASELINE.json configs[2]: synthetic 1M-triangle instanced scene (1,387,526 tris), Disney BSDF + NEE, "
                    f"{WIDTH}x{HEIGHT}, depth 10, rr_depth 0; step = {SPP_PER_STEP} spp of the 1024-spp render",
        "samples_per_step": WIDTH * HEIGHT * SPP_PER_STEP,
        "spp_per_step": SPP_PER_STEP,
        "sharding": "single GPU" if n_gpus == 1 else f"interleaved 32x32 pixel tiles over {n_gpus} GPUs + one NCCL film reduce",
        "l2_policy": "per-pass path state (~25 GB for the 132.7 M paths of a 64-spp pass, four passes per step on one GPU) is far larger "
                     "than the 126 MB L2; no explicit flush",
        "host_buffers": "pageable (std::vector) for the e2e upload",
    }


class ClockSampler:
    """nvidia-smi clocks/throttle sampler running during the timed region."""

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.rows: list[list[str]] = []
        self.proc = None
        self.thread = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [r for r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
                "samples": len(rows), "reasons": reasons}


def measured_hbm_peak() -> tuple[float, str]:
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_oracle_rate(desc, target_seconds: float, spp: int, threads: int = 0):
    """Time the CPU oracle on a bounded tile sample of the frame. Returns (Msamples/s, samples, seconds, description)."""
    from oracle import binding as O

    # calibrate on a run long enough to amortise thread start-up (128 host threads on the GPU box), then size
    # the tile fraction so that the measured run lasts about `target_seconds`
    world, rate = 64, 0.0
    for _ in range(4):
        t0 = time.perf_counter()
        _, cnt = O.render(desc, 0, spp, threads=threads, rank=0, world=world, tile_size=32)
        dt = max(time.perf_counter() - t0, 1e-3)
        rate = cnt["samples"] / dt
        if dt >= 0.5 or world == 1:
            break
        world = max(1, world // 4)
    full = WIDTH * HEIGHT * spp
    world = int(min(512, max(1, round(full / max(rate * target_seconds, 1.0)))))
    t0 = time.perf_counter()
    _, cnt = O.render(desc, 0, spp, threads=threads, rank=0, world=world, tile_size=32)
    dt = time.perf_counter() - t0
    desc_s = f"1/{world} of the 32x32 tiles of the {WIDTH}x{HEIGHT} frame at {spp} spp ({cnt['samples']} samples)"
    return cnt["samples"] / dt * 1e-6, cnt["samples"], dt, desc_s, world


def reference_on_interpreter():
    """The UNMODIFIED reference renderer (oracle/_ref/bin/luisa-render-cli, built from /root/reference by oracle/ref) on its
    `interp` backend - a host AST interpreter - for a 96x54 @2spp view of the same scene, timed by the reference's own
    'Rendering finished in ... ms' line (src/base/integrator.cpp:111-112).  Reported for completeness only: an interpreter
    says nothing about the reference's LLVM `cpu` backend.  Returns None when oracle/_ref is absent."""
    import re
    import subprocess
    import tempfile

    cli = REPO / "oracle" / "_ref" / "bin" / "luisa-render-cli"
    if not cli.exists():
        return None
    try:
        from luisarender_b200 import scenes

        w, h, spp = 96, 54, 2
        with tempfile.TemporaryDirectory() as tmp:
            (Path(tmp) / "scene.luisa").write_text(scenes.instanced_spheres(resolution=(w, h), spp=spp, output="interp.exr"))
            log = subprocess.run([str(cli), "-b", "interp", "scene.luisa"], cwd=tmp, capture_output=True, text=True, timeout=600)
        ms = float(re.search(r"Rendering finished in ([0-9.eE+-]+) ms", log.stdout + log.stderr).group(1))
        return {"value": round(w * h * spp / ms * 1e-3, 6), "unit": UNIT, "kind": "reference",
                "sample": f"{w}x{h} @{spp} spp of the same scene, the reference's own render timer",
                "note": "unmodified reference renderer on oracle/ref's AST-interpreter backend (one host thread per small "
                        "dispatch); bit-identical output to the port, not representative of the reference's LLVM cpu backend"}
    except Exception as e:  # noqa: BLE001 - a reported extra, never fatal
        return {"unavailable": str(e)[:200]}


def run_reference(args, rank: int):
    """--impl reference: the CPU implementation of the path (oracle port) on the host cores, rank 0 only."""
    if rank != 0:
        return
    from oracle import binding as O

    cores = os.cpu_count() or 1
    scene = build_scene()
    desc = scene.desc()
    _, _, _, _, world = cpu_oracle_rate(desc, 1.5, SPP_PER_STEP)
    for w in range(args.warmup):
        O.render(desc, w * SPP_PER_STEP, (w + 1) * SPP_PER_STEP, rank=0, world=world, tile_size=32)
    samples = 0
    t0 = time.perf_counter()
    for s in range(args.steps):
        _, cnt = O.render(desc, s * SPP_PER_STEP, (s + 1) * SPP_PER_STEP, rank=0, world=world, tile_size=32)
        samples += cnt["samples"]
    dt = time.perf_counter() - t0
    value = samples / dt * 1e-6
    sample_desc = f"each step = 1/{world} of the 32x32 tiles of the frame at {SPP_PER_STEP} spp ({samples // args.steps} samples/step)"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_desc,
                         "note": "the reference's Rust/LLVM `cpu` backend + Embree cannot be built in this environment (SURVEY.md §8c); "
                                 "this is the oracle port of the same estimator on all host cores; the port's films are bit-identical to the unmodified "
                                 "reference renderer run through oracle/ref's interpreter backend (tests/test_ref_render.py)"},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    interp = reference_on_interpreter()
    if interp is not None:
        line["reference_on_interpreter"] = interp
    print(json.dumps(line), flush=True)


def scene_upload_bytes(desc) -> int:
    import ctypes as C
    from luisarender_b200 import _ffi as F

    return int(desc.vertex_count * 32 + desc.triangle_count * (12 + 8 + 4) + desc.mesh_count * C.sizeof(F.Mesh) +
               desc.instance_count * (16 + 48 + 64) + desc.bvh_node_count * 64 + desc.tri_slot_count * 48 +
               desc.surface_count * C.sizeof(F.Surface) + desc.light_count * (C.sizeof(F.Light) + 8) + C.sizeof(F.Camera))


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch

    from luisarender_b200 import distributed as D
    from luisarender_b200.api import Renderer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the radiance path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = D.init_process_group("nccl") if world > 1 else None

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    scene = build_scene()
    desc = scene.desc()
    r = Renderer(device_index=local_rank)
    r.upload(desc)
    r.set_shard(rank, world, D.TILE_SIZE)
    film_t = D.device_film_tensor(r, HEIGHT, WIDTH) if world > 1 else None
    K, W, S = args.steps, args.warmup, SPP_PER_STEP

    def step(s):
        r.render(s * S, (s + 1) * S)

    # ---- device-resident timing -------------------------------------------------------------------------
    for w in range(W):
        step(w)
    if film_t is not None:
        D.reduce_film(film_t)  # warm the NCCL communicator
    r.clear()
    r.set_option("time_kernels", 1)
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    t0 = time.perf_counter()
    for s in range(K):
        step(s)
    if film_t is not None:
        D.reduce_film(film_t)
    barrier()
    dt = time.perf_counter() - t0
    clock_info = clocks.stop() if rank == 0 else None
    st = r.stats()
    r.set_option("time_kernels", 0)
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        c = torch.tensor([st["closest_rays"], st["shadow_rays"], st["kernel_launches"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        total_closest, total_shadow, total_launches = (int(x) for x in c.tolist())
    else:
        total_closest, total_shadow, total_launches = st["closest_rays"], st["shadow_rays"], st["kernel_launches"]
    samples = WIDTH * HEIGHT * S * K
    value = samples / dt * 1e-6

    # ---- roofline of the dominant kernel (closest-hit traversal), rank 0's launches ------------------------
    r.clear()
    r.set_option("count_traversal", 1)
    for s in range(K):
        step(s)
    r.set_option("count_traversal", 0)
    cst = r.stats()
    alg_bytes = 48 * cst["closest_rays"] + 64 * cst["closest_nodes"] + 48 * cst["closest_tris"] + 64 * cst["closest_xforms"]
    trace_launches = st["passes"] * desc.integrator.max_depth  # one closest-hit launch per bounce per pass
    peak, peak_src = measured_hbm_peak()
    achieved = alg_bytes / max(st["trace_closest_ms"] * 1e-3, 1e-9) * 1e-9
    traffic = None
    tp = REPO / "profiles" / "traversal_traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "kernel": "trace_closest_kernel<false> (BVH2 closest-hit traversal)", "bound": "hbm", "achieved": round(achieved, 1),
        "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
        "algorithmic_bytes": int(alg_bytes), "launches": int(trace_launches), "kernel_ms_total": round(st["trace_closest_ms"], 3),
        "per_ray": {"nodes": round(cst["closest_nodes"] / max(cst["closest_rays"], 1), 2),
                    "tris": round(cst["closest_tris"] / max(cst["closest_rays"], 1), 2),
                    "xforms": round(cst["closest_xforms"] / max(cst["closest_rays"], 1), 2)},
        "share_of_step": round(st["trace_closest_ms"] / max(st["render_ms"], 1e-9), 4),
        "other_kernels_ms": {"trace_shadow": round(st["trace_shadow_ms"], 3), "shade": round(st["shade_ms"], 3), "other": round(st["other_ms"], 3)},
    }

    # ---- end to end through the C-ABI with host buffers ------------------------------------------------------
    e2e_steps = max(3, min(K, 8))
    h2d = scene_upload_bytes(desc)
    d2h = WIDTH * HEIGHT * 16
    r.upload(desc)
    r.set_shard(rank, world, D.TILE_SIZE)
    barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        r.upload(desc)  # host -> device copy of the step's inputs (flattened scene, camera, integrator)
        r.render(s * S, (s + 1) * S)
        if world > 1:
            D.reduce_film(D.device_film_tensor(r, HEIGHT, WIDTH))
            torch.cuda.synchronize()
        if rank == 0:
            img = r.film()  # device -> host read of the step's result (normalised film)
    barrier()
    e2e_dt = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e_value = WIDTH * HEIGHT * S * e2e_steps / e2e_dt * 1e-6
    if rank == 0:
        assert np.isfinite(img).all()

    # ---- CPU baseline (rank 0, single GPU run only) ------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, n, secs, sample_desc, _ = cpu_oracle_rate(desc, 15.0, SPP_PER_STEP)
        cpu = {"value": round(rate, 4), "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": sample_desc,
               "seconds": round(secs, 2),
               "pinned": "films bit-identical to the unmodified reference renderer on 21 scenes incl. this one at 96x54 (tests/test_ref_render.py)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(dt / K * 1e3, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "mrays_per_s": round((total_closest + total_shadow) / dt * 1e-6, 1),
            "rays": {"closest": total_closest, "shadow": total_shadow},
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": int(total_launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clock_info,
            "device_ms_per_step": round(st["render_ms"] / K, 3),
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    r.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    from luisarender_b200 import distributed as D

    rank, world, local_rank = D.env_world()
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py: --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
