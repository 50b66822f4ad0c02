"""Ad-hoc GPU sanity / timing script (run under gpurun): parity vs the oracle on small renders, then timing
of the two headline scenes with per-kernel-category CUDA-event times."""
from __future__ import annotations

import json
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))

from luisarender_b200 import scenes  # noqa: E402
from luisarender_b200.api import Renderer, Scene  # noqa: E402
from oracle import binding as O  # noqa: E402


def rel_l2(a, b):
    return float(np.linalg.norm(a[..., :3] - b[..., :3]) / max(np.linalg.norm(b[..., :3]), 1e-30))


def parity(name, src, spp, r):
    sc = Scene.from_source(src, REPO)
    d = sc.desc()
    r.upload(d)
    r.render(0, spp)
    gpu_raw = r.film(raw=True)
    t = time.time()
    cpu_raw, cnt = O.render(d, 0, spp)
    dt = time.time() - t
    diff = np.abs(gpu_raw - cpu_raw)
    nbad = int((diff[..., :3].max(axis=-1) > 1e-4 * np.maximum(1.0, np.abs(cpu_raw[..., :3]).max(axis=-1))).sum())
    st = r.stats()
    print(json.dumps({"parity": name, "rel_l2": rel_l2(gpu_raw, cpu_raw), "max_abs": float(diff.max()), "pixels_off": nbad,
                      "pixels": int(diff.shape[0] * diff.shape[1]), "gpu_rays": [st["closest_rays"], st["shadow_rays"]],
                      "cpu_rays": [cnt["closest_rays"], cnt["shadow_rays"]], "cpu_s": round(dt, 2),
                      "gpu_ms": round(st["render_ms"], 3)}), flush=True)


def timing(name, src, spp, r, count=False):
    sc = Scene.from_source(src, REPO)
    print(json.dumps({"scene": name, **{k: v for k, v in sc.info().items() if not k.startswith("world")}}), flush=True)
    d = sc.desc()
    r.upload(d)
    r.set_option("time_kernels", 0)
    r.render(0, min(spp, 4))  # warm-up
    r.clear()
    r.set_option("time_kernels", 1)
    r.render(0, spp)
    st = r.stats()
    w, h = d.camera.resolution[0], d.camera.resolution[1]
    out = {"scene": name, "spp": spp, "ms": round(st["render_ms"], 2),
           "Msamples_s": round(w * h * spp / st["render_ms"] * 1e-3, 2),
           "Mrays_s": round((st["closest_rays"] + st["shadow_rays"]) / st["render_ms"] * 1e-3, 2),
           "closest_rays": st["closest_rays"], "shadow_rays": st["shadow_rays"],
           "trace_closest_ms": round(st["trace_closest_ms"], 2), "trace_shadow_ms": round(st["trace_shadow_ms"], 2),
           "shade_ms": round(st["shade_ms"], 2), "other_ms": round(st["other_ms"], 2), "launches": st["kernel_launches"],
           "passes": st["passes"]}
    if count:
        r.clear()
        r.set_option("time_kernels", 0)
        r.set_option("count_traversal", 1)
        r.render(0, spp)
        st = r.stats()
        rays = st["closest_rays"] + st["shadow_rays"]
        for kind in ("closest", "shadow"):
            nr = max(st[kind + "_rays"], 1)
            out[kind + "_per_ray"] = [round(st[kind + "_nodes"] / nr, 2), round(st[kind + "_tris"] / nr, 2), round(st[kind + "_xforms"] / nr, 2)]
            alg = 48 * nr + 64 * st[kind + "_nodes"] + 48 * st[kind + "_tris"] + 64 * st[kind + "_xforms"]
            out[kind + "_alg_GBs"] = round(alg / (out["trace_" + kind + "_ms"] * 1e-3) * 1e-9, 1)
        r.set_option("count_traversal", 0)
    print(json.dumps(out), flush=True)
    return r.film()


def main():
    r = Renderer(device_index=0)
    parity("cornell_matte_64x64x16", scenes.cornell_box(resolution=(64, 64), spp=16), 16, r)
    parity("cornell_disney_64x64x16", scenes.cornell_box(resolution=(64, 64), spp=16, surface="Disney"), 16, r)
    parity("spheres_small_96x54x8", scenes.instanced_spheres(resolution=(96, 54), spp=8, big_subdivision=4, small_count=12), 8, r)
    img = timing("cornell_1024", scenes.cornell_box(resolution=(1024, 1024), spp=64), 64, r, count=True)
    out_dir = REPO / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    from luisarender_b200.api import save_image
    save_image(out_dir / "cornell_gpu.pfm", img)
    img = timing("spheres_1080p", scenes.instanced_spheres(resolution=(1920, 1080), spp=16), 16, r, count=True)
    save_image(out_dir / "spheres_gpu.pfm", img)


if __name__ == "__main__":
    main()
