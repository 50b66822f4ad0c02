"""Renders BASELINE.json's single-GPU configurations at FULL size on cuda:0 and prints one JSON line per config.

    python tools/run_configs.py [--configs C1,C2,C3,C4] [--spp-fraction 1.0]

C1 Cornell 512x512 @16 · C2 Cornell 1024x1024 @4096 (matte) · C3 1.39M-tri instanced, Disney + NEE, 1920x1080 @1024
· C4 = C3 scene + homogeneous medium, depth 8, 3840x2160 @4096.  Timing = device time of lrk_render (CUDA events inside
the library, excluding parse / BVH build / upload, like the reference's own "Rendering finished in" clock,
/root/reference/src/integrators/wave_path.cpp:503,565).  Size-independent checks: every pixel's weight equals spp, the
film is finite and non-negative, sample/ray counters are consistent.
"""
import argparse, json, sys, time
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from luisarender_b200 import scenes  # noqa: E402
from luisarender_b200.api import Renderer, Scene  # noqa: E402

CONFIGS = {
    "C1": lambda: (scenes.cornell_box(resolution=(512, 512), spp=16), 16),
    "C2": lambda: (scenes.cornell_box(resolution=(1024, 1024), spp=4096), 4096),
    "C3": lambda: (scenes.instanced_spheres(resolution=(1920, 1080), spp=1024), 1024),
    "C4": lambda: (scenes.instanced_spheres(resolution=(3840, 2160), spp=4096, medium=True, depth=8), 4096),
    # row f3 (not a BASELINE config): Mirror / Glass / rough Glass / Plastic / Metal spheres, mirror wall - the hit bucket of
    # the MicrofacetFamilyClosure next to the Matte bucket
    "F3": lambda: (scenes.materials_box(resolution=(1920, 1080), spp=256, depth=10, subdivision=5), 256),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C1,C2,C3,C4")
    ap.add_argument("--spp-fraction", type=float, default=1.0)
    args = ap.parse_args()
    r = Renderer(0)
    for name in args.configs.split(","):
        src, spp_full = CONFIGS[name]()
        t0 = time.time()
        sc = Scene.from_source(src, REPO)
        d = sc.desc()
        t_host = time.time() - t0
        r.upload(d)
        w, h = d.camera.resolution[0], d.camera.resolution[1]
        spp = max(1, int(round(spp_full * args.spp_fraction)))
        r.render(0, min(spp, 4))  # warm-up (allocations, clocks)
        r.clear()
        t0 = time.time()
        r.render(0, spp)
        wall = time.time() - t0
        st = r.stats()
        raw = r.film(raw=True)
        # surface integrators keep every sample; the volume estimator yields a few non-finite samples per thousand, which
        # the film drops together with their weight (color.cpp:107-130)
        full = float((raw[..., 3] == spp).mean())
        dropped = float(spp * raw.shape[0] * raw.shape[1] - raw[..., 3].astype(np.float64).sum())
        ok = bool((raw[..., 3] <= spp).all() and (full == 1.0 or name == "C4") and dropped <= 2e-3 * st["samples"]
                  and np.isfinite(raw).all() and raw[..., :3].min() >= 0)
        ms = st["render_ms"]
        print(json.dumps({"config": name, "resolution": [w, h], "spp": spp, "spp_full": spp_full, "samples": st["samples"],
                          "render_ms": round(ms, 2), "wall_ms": round(wall * 1e3, 1), "msamples_per_s": round(st["samples"] / ms / 1e3, 1),
                          "mrays_per_s": round((st["closest_rays"] + st["shadow_rays"]) / ms / 1e3, 1),
                          "closest_rays": st["closest_rays"], "shadow_rays": st["shadow_rays"], "passes": st.get("passes"),
                          "host_parse_bvh_s": round(t_host, 2), "mean_rgb": [float(x) for x in (raw[..., :3].mean(axis=(0, 1)) / spp)],
                          "dropped_samples": dropped, "checks_ok": ok}), flush=True)


if __name__ == "__main__":
    main()
