#!/usr/bin/env python3
"""Generates tests/golden/ref_renders.npz: images rendered by the UNMODIFIED reference renderer.

oracle/ref builds LuisaRender's own scene parser, node plugins, integrators and `luisa-render-cli` from /root/reference and
a LuisaCompute backend (`-b interp`) that executes the recorded kernels with a host AST interpreter (oracle/ref/README.md).
This script writes small scene files with this repository's generators (luisarender_b200/scenes.py - the same text the
product parses), runs `luisa-render-cli -b interp <scene>` on each and stores scene text + the film the reference hands to
save_image.  Needs /root/reference (this container); the fixture is committed.

    make -C oracle/ref && python tools/gen_ref_renders.py
"""
from __future__ import annotations

import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
CLI = REPO / "oracle" / "_ref" / "bin" / "luisa-render-cli"
OUT = REPO / "tests" / "golden" / "ref_renders.npz"


def cases() -> dict[str, str]:
    from luisarender_b200 import scenes

    spheres = dict(big_subdivision=2, small_subdivision=1, small_count=12)
    c = {}
    # config C1 / C2: Cornell box, matte + area light; the wavefront integrator and the megakernel one
    c["cornell_wavepath"] = scenes.cornell_box(resolution=(24, 24), spp=4)
    c["cornell_megapath"] = scenes.cornell_box(resolution=(24, 24), spp=4).replace("integrator : WavePath", "integrator : MegaPath")
    c["cornell_russian_roulette"] = scenes.cornell_box(resolution=(24, 24), spp=4, depth=12, rr_depth=2, rr_threshold=0.95)
    # config C3: instanced Loop-subdivision spheres, Disney closures, two area lights (reduced triangle count)
    c["spheres_disney"] = scenes.instanced_spheres(resolution=(32, 18), spp=2, depth=6, **spheres)
    # config C4: + homogeneous medium, MegaVPTNaive
    c["spheres_medium"] = scenes.instanced_spheres(resolution=(32, 18), spp=2, depth=6, medium=True, **spheres)
    # row f3: Mirror, Glass (smooth / rough), Plastic and Metal closures; Russian roulette with the refraction eta scale
    c["materials_wavepath"] = scenes.materials_box(resolution=(32, 24), spp=4, depth=8)
    c["materials_megapath_rr"] = scenes.materials_box(resolution=(32, 24), spp=4, depth=10, rr_depth=2, rr_threshold=0.95,
                                                      integrator="MegaPath")
    return c


def read_f32(path: Path) -> np.ndarray:
    with open(path, "rb") as f:
        w, h, ch = map(int, f.readline().split())
        return np.frombuffer(f.read(), dtype=np.float32).reshape(h, w, ch).copy()


def render_with_reference(source: str, workdir: Path, name: str = "scene") -> np.ndarray:
    """Runs the reference CLI on `source`; returns the RGBA film (the camera's `file` + '.f32', oracle/ref/shim.cpp)."""
    import re

    scene_path = workdir / f"{name}.luisa"
    scene_path.write_text(source)
    out_name = re.search(r'file\s*\{\s*"([^"]+)"\s*\}', source).group(1)
    log = subprocess.run([str(CLI), "-b", "interp", scene_path.name], cwd=workdir, capture_output=True, text=True, timeout=3600)
    out = workdir / (out_name + ".f32")
    if not out.exists():
        raise RuntimeError(f"reference render of '{name}' failed:\n{log.stdout[-2000:]}\n{log.stderr[-2000:]}")
    return read_f32(out)


def main() -> int:
    if not CLI.exists():
        print(f"{CLI} is missing: run `make -C oracle/ref` (needs /root/reference)", file=sys.stderr)
        return 1
    data = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, source in cases().items():
            image = render_with_reference(source, Path(tmp), name)
            data[f"{name}/scene"] = np.frombuffer(source.encode(), dtype=np.uint8)
            data[f"{name}/image"] = image
            print(f"{name:28s} {image.shape[1]}x{image.shape[0]}  mean rgb = {image[..., :3].mean():.6f}")
    np.savez_compressed(OUT, **data)
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
