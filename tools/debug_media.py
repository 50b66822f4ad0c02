"""Debug aid (run under gpurun): GPU vs oracle on the shape-media scene, depth by depth."""
import sys
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from luisarender_b200 import scenes
from luisarender_b200.api import Renderer, Scene
from oracle import binding as O

r = Renderer(0)
for kw in ({}, {"environment_medium": True}):
    for depth in (1, 2, 3, 5):
        sc = Scene.from_source(scenes.media_box(resolution=(48, 48), spp=4, depth=depth, **kw), REPO)
        d = sc.desc()
        r.upload(d)
        r.clear()
        r.render(0, 4)
        g = r.film(raw=True)
        c, cnt = O.render(d, 0, 4)
        st = r.stats()
        err = np.abs(g[..., :3] - c[..., :3]).max(axis=-1)
        bad = err > 1e-4 * np.maximum(np.abs(c[..., :3]).max(axis=-1), 1.0)
        ys, xs = np.nonzero(bad)
        print(kw, "depth", depth, "off", round(float(bad.mean()), 4), "w_equal", bool(np.array_equal(g[..., 3], c[..., 3])),
              "rays gpu", st["closest_rays"], st["shadow_rays"], "cpu", cnt["closest_rays"], cnt["shadow_rays"],
              "bbox", (int(xs.min()), int(xs.max()), int(ys.min()), int(ys.max())) if bad.any() else None,
              "mean", float(g[..., :3].mean()), float(c[..., :3].mean()))
        if bad.any() and depth <= 2:
            k = 0
            for y, x in zip(ys[:6], xs[:6]):
                print("   px", x, y, "gpu", g[y, x], "cpu", c[y, x])
