import json, sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(REPO))
from luisarender_b200 import scenes
from luisarender_b200.api import Renderer, Scene
r = Renderer(0)
sc = Scene.from_source(scenes.instanced_spheres(resolution=(3840, 2160), spp=4096, medium=True, depth=8), REPO); d = sc.desc(); r.upload(d)
spp = 16
r.render(0, spp)
for opt in ({}, {"count_traversal": 1}):
    r.clear(); r.set_option("time_kernels", 1)
    for k, v in opt.items(): r.set_option(k, v)
    r.render(0, spp); st = r.stats()
    print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()}), flush=True)
