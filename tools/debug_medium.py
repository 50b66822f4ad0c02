import sys, numpy as np
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(REPO))
from luisarender_b200 import scenes
from luisarender_b200.api import Renderer, Scene
from oracle import binding as O
r = Renderer(0)
def cmp(tag, g, c):
    diff = np.abs(g[..., :3] - c[..., :3]).max(axis=-1)
    bad = np.argwhere(diff > 1e-4 * np.maximum(1, np.abs(c[..., :3]).max(axis=-1)))
    print(tag, "bad pixels", len(bad), "of", diff.size, "sum gpu", g[..., :3].sum(), "cpu", c[..., :3].sum(), flush=True)
sc = Scene.from_source(scenes.instanced_spheres(resolution=(32, 18), spp=4, big_subdivision=3, small_subdivision=2, small_count=10, medium=True, depth=8), REPO)
d = sc.desc(); r.upload(d)
for s in range(4):
    r.clear(); r.render(s, s + 1); g = r.film(raw=True); c, _ = O.render(d, s, s + 1); cmp(f"sample {s}", g, c)
r.clear(); r.render(0, 4); g = r.film(raw=True); c, _ = O.render(d, 0, 4); cmp("samples 0..4 one pass", g, c)
r.clear(); r.set_option("max_paths_per_pass", 576); r.render(0, 4); g = r.film(raw=True); cmp("samples 0..4 four passes", g, c)
r.set_option("max_paths_per_pass", 136 << 20)
r.clear(); r.render(0, 2); g = r.film(raw=True); c2, _ = O.render(d, 0, 2); cmp("samples 0..2", g, c2)
for s in (2, 3):
    r.clear(); r.render(s, s + 1); g = r.film(raw=True); c, _ = O.render(d, s, s + 1)
    diff = np.abs(g[..., :3] - c[..., :3]).max(axis=-1)
    for y, x in np.argwhere(diff > 1e-3):
        print("sample", s, "pixel", x, y, "gpu", g[y, x], "cpu", c[y, x], "oracle_li", O.li(d, int(x), int(y), s))
