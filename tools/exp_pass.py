"""Experiment: throughput vs pass size / spp per call on the C3 scene (and Cornell C2)."""
import json, sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from luisarender_b200 import scenes
from luisarender_b200.api import Renderer, Scene

def run(r, d, spp, max_paths, reps=3):
    r.set_option("max_paths_per_pass", max_paths)
    r.clear(); r.render(0, spp)
    best = 1e30
    for i in range(reps):
        r.clear(); r.set_option("time_kernels", 1); r.render(0, spp); st = r.stats(); best = min(best, st["render_ms"])
    w, h = d.camera.resolution[0], d.camera.resolution[1]
    print(json.dumps({"spp": spp, "max_paths_Mi": max_paths >> 20, "ms": round(best, 2), "Msamples_s": round(w*h*spp/best*1e-3, 1),
                      "passes": st["passes"], "closest_ms": round(st["trace_closest_ms"],2), "shadow_ms": round(st["trace_shadow_ms"],2),
                      "shade_ms": round(st["shade_ms"],2)}), flush=True)

r = Renderer(0)
for name, src in (("spheres", scenes.instanced_spheres(resolution=(1920,1080), spp=1024)), ("cornell", scenes.cornell_box(resolution=(1024,1024), spp=4096))):
    sc = Scene.from_source(src, REPO); d = sc.desc(); r.upload(d); print(name, flush=True)
    for spp, mp in ((16, 8<<20), (16, 16<<20), (16, 34<<20), (32, 34<<20), (64, 34<<20), (64, 68<<20), (64, 136<<20), (128, 136<<20)):
        run(r, d, spp, mp)
