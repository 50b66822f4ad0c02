import json, sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(REPO))
from luisarender_b200 import scenes
from luisarender_b200.api import Renderer, Scene
r = Renderer(0)
sc = Scene.from_source(scenes.instanced_spheres(resolution=(1920,1080), spp=1024), REPO); d = sc.desc(); r.upload(d)
spp = 32
for tri in (1, 6, 10, 14, 20, 33):
    for inner in (4, 8, 12):
        r.set_option("tri_min", tri); r.set_option("inner_min", inner)
        r.clear(); r.render(0, spp)
        best = None
        for i in range(2):
            r.clear(); r.set_option("time_kernels", 1); r.render(0, spp); st = r.stats()
            if best is None or st["render_ms"] < best["render_ms"]: best = st
        print(json.dumps({"tri_min": tri, "inner_min": inner, "ms": round(best["render_ms"], 2), "closest_ms": round(best["trace_closest_ms"], 2),
                          "shadow_ms": round(best["trace_shadow_ms"], 2)}), flush=True)
