"""Per-category device times of the F3 materials scene (row f3) next to Cornell: python tools/exp_f3.py"""
import json, sys
from pathlib import Path
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from luisarender_b200 import scenes  # noqa: E402
from luisarender_b200.api import Renderer, Scene  # noqa: E402

r = Renderer(0)
r.set_option("time_kernels", 1)
for name, src, spp in (("F3", scenes.materials_box(resolution=(1920, 1080), spp=64, depth=10, subdivision=5), 64),
                       ("F3_no_rr_depth4", scenes.materials_box(resolution=(1920, 1080), spp=64, depth=4, rr_depth=99, subdivision=5), 64),
                       ("cornell", scenes.cornell_box(resolution=(1920, 1080), spp=64), 64)):
    d = Scene.from_source(src, REPO).desc()
    r.upload(d)
    r.render(0, 4)
    r.clear()
    r.render(0, spp)
    st = r.stats()
    print(json.dumps({"scene": name, **{k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()}}))
