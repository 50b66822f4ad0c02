import sys
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(REPO))
from luisarender_b200 import scenes
from luisarender_b200.api import Renderer, Scene
from oracle import binding as O
r = Renderer(0)
full = scenes.textured_room(resolution=(48, 32), spp=4, wrappers=True, depth=2)
i = full.index('normal_map : Image'); j = full.index('normal_map_strength { 0.8 }') + len('normal_map_strength { 0.8 }')
src = (full[:i] + full[j:]).replace('alpha : Image', 'alpha_unused : Image')   # cube opacity only
d = Scene.from_source(src, REPO).desc()
r.upload(d)
H = float.fromhex
rays = np.array([
    [H("0x0p+0"), H("0x1.666666p+0"), H("0x1.0cccccp+2"), 0, H("-0x1.7b193p-3"), H("-0x1.d6807p-6"), H("-0x1.f6efdep-1"), H("0x1.fffffep+127")],
    [H("-0x1.2b1878p+0"), H("0x1.37ffbp+0"), H("-0x1.fffep+0"), 0, H("0x1.5858f2p-2"), H("0x1.df1768p-2"), H("0x1.a2789cp-1"), H("0x1.432408p+1")],
    [H("-0x1.2b1878p+0"), H("0x1.37ffbp+0"), H("-0x1.fffep+0"), 0, H("0x1.8f743cp-2"), H("-0x1.1836eap-1"), H("0x1.7b1f0cp-1"), H("0x1.fffffep+127")],
    [H("-0x1.32dafep-2"), H("0x1p-16"), H("-0x1.6778a6p-2"), 0, H("0x1.160a8cp-5"), H("0x1.f4ccd4p-1"), H("0x1.a44f68p-3"), H("0x1.3a096ep+1")],
], np.float32)
for any_hit in (False, True):
    a, _ = O.trace(d, rays, any_hit=any_hit); b = r.trace(rays, any_hit=any_hit)
    print("any_hit", any_hit, "oracle", a["inst"].tolist(), a["prim"].tolist(), "gpu", b["inst"].tolist(), b["prim"].tolist())
# the same rays repeated 4096 times (full warps, refill active)
big = np.tile(rays, (4096, 1))
b = r.trace(big, any_hit=True)
print("tiled any-hit gpu unique per ray:", [np.unique(b["inst"][k::4]).tolist() for k in range(4)])
# per-sample render of that pixel at depth 2 with different pass sizes
for mp in (0, 1024, 64):
    if mp: r.set_option("max_paths_per_pass", mp)
    r.clear(); r.render(0, 1); g = r.film(raw=True)
    print("max_paths", mp, "pixel (15,8):", g[8, 15, :3].tolist())
print("oracle:", O.li(d, 15, 8, 0).tolist())
