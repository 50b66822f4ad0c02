import json, sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(REPO))
from luisarender_b200 import scenes
from luisarender_b200.api import Renderer, Scene
import numpy as np
r = Renderer(0)
for name, src, spp in (("spheres", scenes.instanced_spheres(resolution=(1920,1080), spp=1024), 64), ("cornell", scenes.cornell_box(resolution=(1024,1024), spp=4096), 64)):
    sc = Scene.from_source(src, REPO); d = sc.desc(); r.upload(d)
    films = {}
    for bins in (0, 1):
        r.set_option("bin_rays", bins)
        r.clear(); r.render(0, spp)
        best = None
        for i in range(3):
            r.clear(); r.set_option("time_kernels", 1); r.render(0, spp); st = r.stats()
            if best is None or st["render_ms"] < best["render_ms"]: best = st
        films[bins] = r.film(raw=True)
        w, h = d.camera.resolution[0], d.camera.resolution[1]
        print(json.dumps({"scene": name, "bin_rays": bins, "ms": round(best["render_ms"], 2), "Msamples_s": round(w*h*spp/best["render_ms"]*1e-3, 1),
                          "closest_ms": round(best["trace_closest_ms"], 2), "shadow_ms": round(best["trace_shadow_ms"], 2), "shade_ms": round(best["shade_ms"], 2),
                          "other_ms": round(best["other_ms"], 2), "launches": best["kernel_launches"]}), flush=True)
    print("films identical:", bool(np.array_equal(films[0], films[1])), flush=True)
