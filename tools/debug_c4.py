import sys
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(REPO))
from luisarender_b200 import scenes
from luisarender_b200.api import Renderer, Scene
from oracle import binding as O
r = Renderer(0)
for res in ((3840, 2160), (1920, 1080)):
    sc = Scene.from_source(scenes.instanced_spheres(resolution=res, spp=4096, medium=True, depth=8), REPO); d = sc.desc(); r.upload(d)
    spp = 2
    r.render(0, spp); raw = r.film(raw=True); st = r.stats()
    print(res, "weights ok", (raw[..., 3] == spp).all(), "finite", np.isfinite(raw).all(), "min", raw[..., :3].min(), "nan count", np.isnan(raw).sum(), "inf", np.isinf(raw).sum())
    print("weight hist", np.unique(raw[..., 3], return_counts=True))
    print(st["samples"], st["closest_rays"], st["passes"])
    cpu_part, cnt = O.render(d, 0, spp, rank=11, world=256, tile_size=32)
    mask = cpu_part[..., 3] > 0
    g, c = raw[mask][:, :3], cpu_part[mask][:, :3]
    err = np.abs(g - c).max(axis=-1)
    print("mask", mask.sum(), "off frac", (err > 1e-4 * np.maximum(np.abs(c).max(axis=-1), 1.0)).mean())
    keep = err <= np.quantile(err, 0.99)
    print("rel kept", np.linalg.norm((g - c)[keep]) / np.linalg.norm(c[keep]), "cpu min", c.min(), "gpu min in mask", g.min())
    bad = np.argwhere(raw[..., :3].min(axis=-1) < 0)[:5]; print("neg px", bad, [raw[y, x] for y, x in bad])
