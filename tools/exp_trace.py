"""Kernel experiment driver (run under gpurun): one 64-spp pass of BASELINE config C3 (1920x1080, 1.39 M triangles) per setting,
per-kernel-category CUDA-event times.  Settings: run-time options given as name=v1,v2,... on the command line; the library build
is selected with LRK_DEVICE_LIB (tools/build_variants.sh).   python tools/exp_trace.py [--scene C3|C2|C4] [--spp 64] [opt=a,b,c ...]"""
import argparse, itertools, json, os, sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from luisarender_b200 import scenes  # noqa: E402
from luisarender_b200.api import Renderer, Scene  # noqa: E402

SCENES = {
    "C2": lambda: scenes.cornell_box(resolution=(1024, 1024), spp=4096),
    "C3": lambda: scenes.instanced_spheres(resolution=(1920, 1080), spp=1024),
    "C4": lambda: scenes.instanced_spheres(resolution=(3840, 2160), spp=4096, medium=True, depth=8),
    "F3": lambda: scenes.materials_box(resolution=(1920, 1080), spp=256, depth=10, subdivision=5),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="C3")
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--count", action="store_true")
    ap.add_argument("--device-bvh", action="store_true", help="build the hierarchy on the GPU (option device_bvh) instead of uploading the host's")
    ap.add_argument("opts", nargs="*")
    a = ap.parse_args()
    sweep = [(o.split("=")[0], [int(v) for v in o.split("=")[1].split(",")]) for o in a.opts]
    sc = Scene.from_source(SCENES[a.scene](), REPO)
    d = sc.desc()
    r = Renderer(0)
    if a.device_bvh:
        r.set_option("device_bvh", 1)
    import time
    t0 = time.perf_counter()
    r.upload(d)
    upload_ms = (time.perf_counter() - t0) * 1e3
    w, h = d.camera.resolution[0], d.camera.resolution[1]
    r.render(0, min(a.spp, 8))
    for combo in itertools.product(*[vals for _, vals in sweep]) if sweep else [()]:
        for (name, _), v in zip(sweep, combo):
            r.set_option(name, v)
        best = None
        for _ in range(a.repeat):
            r.clear()
            r.set_option("time_kernels", 1)
            r.render(0, a.spp)
            st = r.stats()
            if best is None or st["render_ms"] < best["render_ms"]:
                best = st
        out = {"lib": os.environ.get("LRK_DEVICE_LIB", "libb200pt.so"), "scene": a.scene, "spp": a.spp, "device_bvh": bool(a.device_bvh),
               "upload_ms": round(upload_ms, 1), "host_bvh_build_ms": round(sc.info()["bvh_build_ms"], 1),
               **{name: v for (name, _), v in zip(sweep, combo)},
               "ms": round(best["render_ms"], 2), "Msamples_s": round(w * h * a.spp / best["render_ms"] * 1e-3, 1),
               "closest_ms": round(best["trace_closest_ms"], 2), "shadow_ms": round(best["trace_shadow_ms"], 2),
               "shade_ms": round(best["shade_ms"], 2), "other_ms": round(best["other_ms"], 2), "passes": best["passes"],
               "rays": [best["closest_rays"], best["shadow_rays"]]}
        if a.count:
            r.clear()
            r.set_option("time_kernels", 0)
            r.set_option("count_traversal", 1)
            r.render(0, a.spp)
            st = r.stats()
            r.set_option("count_traversal", 0)
            for kind in ("closest", "shadow"):
                nr = max(st[kind + "_rays"], 1)
                out[kind + "_per_ray"] = [round(st[kind + "_nodes"] / nr, 2), round(st[kind + "_tris"] / nr, 2), round(st[kind + "_xforms"] / nr, 2)]
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
