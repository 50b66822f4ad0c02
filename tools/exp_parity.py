"""Parity metrics of the device library selected with LRK_DEVICE_LIB against the oracle (run under gpurun): three scenes, the share of
pixels that differ by more than 1e-4 relative, rel-L2 of all pixels and of the 99 % closest, and the image means."""
import json, os, sys
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from luisarender_b200 import scenes
from luisarender_b200.api import Renderer, Scene
from oracle import binding as O

SCENES = {
    "cornell_128_16spp": lambda: scenes.cornell_box(resolution=(128, 128), spp=16),
    "cornell_disney_96_8spp": lambda: scenes.cornell_box(resolution=(96, 96), spp=8, surface="Disney"),
    "spheres_disney_128x72_8spp": lambda: scenes.instanced_spheres(resolution=(128, 72), spp=8, depth=8, big_subdivision=3, small_subdivision=2, small_count=20),
}
r = Renderer(0)
for name, make in SCENES.items():
    sc = Scene.from_source(make(), REPO)
    d = sc.desc()
    spp = d.camera.spp
    r.upload(d); r.clear(); r.render(0, spp)
    g = r.film(raw=True); st = r.stats()
    c, cnt = O.render(d, 0, spp)
    diff = np.abs(g[..., :3] - c[..., :3]).max(axis=-1)
    scale = np.maximum(np.abs(c[..., :3]).max(axis=-1), 1.0)
    off = float((diff > 1e-4 * scale).mean())
    rel = float(np.linalg.norm(g[..., :3] - c[..., :3]) / np.linalg.norm(c[..., :3]))
    keep = diff <= np.quantile(diff, 0.99)
    rel99 = float(np.linalg.norm((g[..., :3] - c[..., :3])[keep]) / np.linalg.norm(c[..., :3][keep]))
    med = float(np.median(diff / scale))
    print(json.dumps({"lib": os.environ.get("LRK_DEVICE_LIB", "libb200pt.so"), "scene": name, "off_1e-4": round(off, 5), "rel_l2": round(rel, 6),
                      "rel_l2_best99": round(rel99, 7), "median_rel_diff": med, "mean_gpu": float(g[..., :3].mean()), "mean_cpu": float(c[..., :3].mean()),
                      "weights_equal": bool(np.array_equal(g[..., 3], c[..., 3])), "closest_rays": [st["closest_rays"], cnt["closest_rays"]]}), flush=True)
