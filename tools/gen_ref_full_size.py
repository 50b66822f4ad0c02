#!/usr/bin/env python3
"""Large renders by the unmodified reference renderer, kept as digests (SHA-256 of the film + 32x32-pixel block means):

    c1          BASELINE.json configs[0] at FULL size: Cornell box, 512x512 @16 spp (4.19 M samples; ~7 min)
    c3_full_resolution  configs[2] at its full 1920x1080, 2 spp (4.15 M samples)
    c4_quarter  configs[3]'s scene (homogeneous medium, MegaVPTNaive, depth 8) at 480x270 @4 spp
    c3_quarter  the headline scene of configs[2] (1 387 526 instanced triangles, Disney + NEE, depth 10) at a quarter of its
                resolution: 480x270 @4 spp (0.52 M samples)

Runs `oracle/_ref/bin/luisa-render-cli -b interp` (oracle/ref/README.md) on the scene text of scenes.cornell_box(512x512, 16 spp)
with the MegaPath integrator (the estimator of WavePath, one sample per pixel per dispatch: the film's float atomics then
have a fixed order even on several interpreter threads) - about 10 minutes on 8 cores - and stores the SHA-256 of the
film plus 32x32 block means in tests/golden/ref_full_size.json.  tests/test_ref_render.py compares the oracle's film, bit
for bit, through the hash.

    make -C oracle/ref && python tools/gen_ref_full_size.py [case [film.exr]]
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tools"))
OUT = REPO / "tests" / "golden" / "ref_full_size.json"


def c1_scene() -> str:
    from luisarender_b200 import scenes

    return scenes.cornell_box(resolution=(512, 512), spp=16).replace("integrator : WavePath", "integrator : MegaPath")


def c3_quarter_scene() -> str:
    from luisarender_b200 import scenes

    return scenes.instanced_spheres(resolution=(480, 270), spp=4, output="c3q.exr").replace("integrator : WavePath", "integrator : MegaPath")


def c3_full_resolution_scene() -> str:
    from luisarender_b200 import scenes

    return scenes.instanced_spheres(resolution=(1920, 1080), spp=2, output="c3f.exr").replace("integrator : WavePath", "integrator : MegaPath")


def c4_quarter_scene() -> str:
    from luisarender_b200 import scenes

    return scenes.instanced_spheres(resolution=(480, 270), spp=4, medium=True, depth=8, output="c4q.exr")


def _megapath(source: str) -> str:
    assert "integrator : WavePath" in source
    return source.replace("integrator : WavePath", "integrator : MegaPath")


def c3_full_resolution_wavepath_scene() -> str:
    # WavePath itself at full resolution: with 2 spp the film's two float atomic adds per pixel commute, so the film does
    # not depend on the interpreter's thread schedule
    from luisarender_b200 import scenes

    return scenes.instanced_spheres(resolution=(1920, 1080), spp=2, output="c3w.exr")


def c2_full_resolution_scene() -> str:
    from luisarender_b200 import scenes

    return scenes.cornell_box(resolution=(1024, 1024), spp=2, output="c2.exr")


def materials_large_scene() -> str:
    from luisarender_b200 import scenes

    return _megapath(scenes.materials_box(resolution=(320, 240), spp=4, depth=10, rr_depth=2, subdivision=4, mix=True, output="materials_large.exr"))


def textured_large_scene() -> str:
    from luisarender_b200 import scenes

    return _megapath(scenes.textured_room(resolution=(320, 240), spp=4, mesh_files=False, wrappers=True, output="textured_large.exr"))


def flatten_large_scene() -> str:
    from luisarender_b200 import scenes

    return _megapath(scenes.flatten_stress(resolution=(320, 240), spp=4, output="flatten_large.exr"))


def _small_case(name: str, resolution: tuple, new_resolution: tuple, spp: int, new_spp: int, output: str) -> str:
    """One of tools/gen_ref_renders.py's scenes at a larger film."""
    import re

    import gen_ref_renders as G

    source = G.cases()[name]
    old = f"resolution {{ {resolution[0]}, {resolution[1]} }}"
    assert source.count(old) == 1 and source.count(f"spp {{ {spp} }}") == 1
    source = source.replace(old, f"resolution {{ {new_resolution[0]}, {new_resolution[1]} }}").replace(f"spp {{ {spp} }}", f"spp {{ {new_spp} }}")
    source, n = re.subn(r'(Camera\b.*?\bfile\s*\{\s*")[^"]+("\s*\})', rf"\g<1>{output}\g<2>", source, count=1, flags=re.S)
    assert n == 1
    return source.replace("integrator : WavePath", "integrator : MegaPath")


def materials_textured_large_scene() -> str:
    return _small_case("materials_textured", (48, 30), (320, 200), 4, 4, "materials_textured_large.exr")


def named_metals_large_scene() -> str:
    return _small_case("materials_named_metals", (32, 24), (320, 240), 4, 4, "named_metals_large.exr")


def layered_large_scene() -> str:
    return _small_case("materials_layered", (32, 24), (192, 144), 4, 4, "layered_large.exr")


def sampler_pmj02bn_large_scene() -> str:
    return _small_case("spheres_sampler_pmj02bn", (32, 18), (333, 187), 8, 8, "sampler_pmj02bn_large.exr")


def sampler_sobol_large_scene() -> str:
    return _small_case("spheres_sampler_sobol", (32, 18), (333, 187), 3, 3, "sampler_sobol_large.exr")


def sampler_paddedsobol_large_scene() -> str:
    return _small_case("spheres_sampler_paddedsobol", (32, 18), (333, 187), 3, 3, "sampler_paddedsobol_large.exr")


def sampler_zsobol_large_scene() -> str:
    return _small_case("spheres_sampler_zsobol", (32, 18), (333, 187), 2, 2, "sampler_zsobol_large.exr")


def disney_transmissive_large_scene() -> str:
    return _small_case("spheres_disney_transmissive", (32, 18), (480, 270), 4, 4, "disney_transmissive_large.exr")


def media_shapes_large_scene() -> str:
    return _small_case("media_shapes", (32, 32), (160, 160), 4, 4, "media_shapes_large.exr")


def media_nested_large_scene() -> str:
    return _small_case("media_nested_in_environment_medium", (32, 32), (160, 160), 4, 4, "media_nested_large.exr")


def media_quirk_large_scene() -> str:
    return _small_case("media_true_hit_quirk", (32, 32), (160, 160), 4, 4, "media_quirk_large.exr")


def disney_thin_large_scene() -> str:
    return _small_case("spheres_disney_thin", (32, 18), (480, 270), 4, 4, "thin_large.exr")


def disney_thin_medium_large_scene() -> str:
    return _small_case("spheres_medium_disney_thin", (32, 18), (320, 180), 4, 4, "thin_medium_large.exr")


def media_disney_shells_large_scene() -> str:
    return _small_case("media_disney_shells", (32, 32), (192, 192), 4, 4, "shells_large.exr")


def textured_light_large_scene() -> str:
    return _small_case("textured_light", (32, 24), (320, 240), 4, 4, "texlight_large.exr")


def image_formats_large_scene() -> str:
    return _small_case("image_formats", (80, 48), (400, 240), 2, 2, "formats_large.exr")


def disney_lobes_large_scene() -> str:
    return _small_case("spheres_disney_all_lobes", (32, 18), (480, 270), 2, 4, "lobes_large.exr")


def cornell_rr_gaussian_large_scene() -> str:
    return _small_case("cornell_russian_roulette", (24, 24), (512, 512), 4, 4, "rr_large.exr").replace(
        "filter : Box { radius { 0.5 } }", "filter : Gaussian { radius { 1.5 } }")


def cornell_mitchell_large_scene() -> str:
    return _small_case("cornell_filter_mitchell", (24, 24), (256, 256), 4, 4, "mitchell_large.exr")


def cornell_options_large_scene() -> str:
    return _small_case("cornell_film_and_light_options", (24, 24), (256, 256), 4, 4, "options_large.exr")


def medium_channels_large_scene() -> str:
    return _small_case("spheres_medium_isotropic", (32, 18), (320, 180), 2, 4, "iso_large.exr")


def medium_hg_large_scene() -> str:
    return _small_case("spheres_medium", (32, 18), (320, 180), 2, 4, "hg_large.exr")


def materials_wavepath_large_scene() -> str:
    from luisarender_b200 import scenes

    return scenes.materials_box(resolution=(320, 240), spp=2, depth=10, rr_depth=2, subdivision=4, mix=True, output="materials_wavepath_large.exr")


def textured_wavepath_large_scene() -> str:
    from luisarender_b200 import scenes

    return scenes.textured_room(resolution=(320, 240), spp=2, mesh_files=False, wrappers=True, output="textured_wavepath_large.exr")


def cornell_disney_odd_scene() -> str:
    from luisarender_b200 import scenes

    return _megapath(scenes.cornell_box(resolution=(333, 187), spp=3, surface="Disney", output="cornell_disney_odd.exr"))


def subdivision_large_scene() -> str:
    from luisarender_b200 import scenes

    return _megapath(scenes.subdivision_scene(resolution=(320, 240), spp=4, output="sd.exr"))


def swizzle_large_scene() -> str:
    from luisarender_b200 import scenes

    return _megapath(scenes.swizzle_scene(resolution=(320, 240), spp=4, output="sw.exr"))


def checkerboard_large_scene() -> str:
    from luisarender_b200 import scenes

    return _megapath(scenes.checkerboard_scene(resolution=(320, 240), spp=4, output="ck.exr"))


def environment_large_scene() -> str:
    return _small_case("environment_image", (32, 20), (320, 200), 2, 4, "env_large.exr")


CASES = {
    "c1": (c1_scene, 16, "BASELINE.json configs[0]: Cornell box 512x512 @16 spp (MegaPath), rendered by luisa-render-cli -b interp"),
    "c3_quarter": (c3_quarter_scene, 4, "config C3's scene (BASELINE.json configs[2]: instanced Disney spheres) at 480x270 @4 spp (MegaPath), rendered by luisa-render-cli -b interp"),
    "c3_full_resolution": (c3_full_resolution_scene, 2, "config C3 (BASELINE.json configs[2]) at its full 1920x1080 resolution, 2 of its spp (4.15 M samples; MegaPath), rendered by luisa-render-cli -b interp"),
    "c4_quarter": (c4_quarter_scene, 4, "config C4's scene (BASELINE.json configs[3]: the same spheres in a homogeneous medium, MegaVPTNaive, depth 8) at 480x270 @4 spp, rendered by luisa-render-cli -b interp (GCC build: see DESIGN.md section 4 on homogeneous.cpp:91)"),
    "c3_full_resolution_wavepath": (c3_full_resolution_wavepath_scene, 2, "as c3_full_resolution, but the WavePath integrator itself"),
    "c2_full_resolution": (c2_full_resolution_scene, 2, "config C2 (BASELINE.json configs[1]): Cornell box at its full 1024x1024, 2 of its spp, WavePath"),
    "materials_large": (materials_large_scene, 4, "row f3: the materials box (Mirror, Glass, rough Glass, Plastic, Metal, Mix; level-4 spheres) 320x240 @4 spp, depth 10, Russian roulette from depth 2, MegaPath"),
    "textured_large": (textured_large_scene, 4, "row f1: the image-textured room with the surface wrappers (normal map, alpha cut-out, opacity) 320x240 @4 spp, MegaPath"),
    "flatten_large": (flatten_large_scene, 4, "row a23: the flattening stress scene 320x240 @4 spp, MegaPath"),
    "materials_textured_large": (materials_textured_large_scene, 4, "row f3: image-textured Mirror / Glass / Plastic / Metal parameters, 320x200 @4 spp, MegaPath"),
    "named_metals_large": (named_metals_large_scene, 4, "row f3: Metal with the reference's measured spectra (Cu, Gold, aluminium fallback), 320x240 @4 spp, MegaPath"),
    "layered_large": (layered_large_scene, 4, "row f3: the Layered surface (Glass over Matte / Mirror, Plastic over Matte), 192x144 @4 spp, MegaPath (GCC build)"),
    "sampler_pmj02bn_large": (sampler_pmj02bn_large_scene, 8, "row f2: PMJ02BN at an odd film size, 333x187 @8 spp, MegaPath"),
    "sampler_sobol_large": (sampler_sobol_large_scene, 3, "row f2: Sobol at an odd film size, 333x187 @3 spp, MegaPath"),
    "sampler_paddedsobol_large": (sampler_paddedsobol_large_scene, 3, "row f2: PaddedSobol at an odd film size, 333x187 @3 spp, MegaPath"),
    "sampler_zsobol_large": (sampler_zsobol_large_scene, 2, "row f2: ZSobol at an odd film size (more base-4 digits), 333x187 @2 spp, MegaPath"),
    "disney_transmissive_large": (disney_transmissive_large_scene, 4, "row a15: transmissive next to opaque Disney nodes, 480x270 @4 spp, MegaPath"),
    "media_shapes_large": (media_shapes_large_scene, 4, "row a22: media bound to Glass shells, 160x160 @4 spp, MegaVPTNaive (GCC build)"),
    "media_nested_large": (media_nested_large_scene, 4, "row a22: the same inside an environment medium, Russian roulette, 160x160 @4 spp, MegaVPTNaive (GCC build)"),
    "media_quirk_large": (media_quirk_large_scene, 4, "row a22: the true_hit(tag) quirk, 160x160 @4 spp, MegaVPTNaive (GCC build)"),
    "disney_thin_large": (disney_thin_large_scene, 4, "row a15: the sphere scene with thin, transmissive and opaque Disney nodes, 480x270 @4 spp, depth 8, Russian roulette from depth 2, MegaPath"),
    "disney_thin_medium_large": (disney_thin_medium_large_scene, 4, "rows a15 / a22: the same three Disney classes inside a homogeneous medium, 320x180 @4 spp, MegaVPTNaive (GCC build)"),
    "media_disney_shells_large": (media_disney_shells_large_scene, 4, "row a22: media bound to shapes behind thin / transmissive Disney shells in an environment medium, 192x192 @4 spp, MegaVPTNaive (GCC build)"),
    "textured_light_large": (textured_light_large_scene, 4, "row f1: an area light with an image emission in the textured room, 320x240 @4 spp, MegaPath"),
    "image_formats_large": (image_formats_large_scene, 2, "row f1: one panel per BMP / TGA storage variant, 400x240 @2 spp, MegaPath"),
    "disney_lobes_large": (disney_lobes_large_scene, 4, "row a15: the Disney sphere scene with EVERY Disney parameter set, 480x270 @4 spp, MegaPath"),
    "cornell_rr_gaussian_large": (cornell_rr_gaussian_large_scene, 4, "Cornell 512x512 @4 spp, depth 12, Russian roulette from depth 2, Gaussian filter (row a2), MegaPath"),
    "cornell_mitchell_large": (cornell_mitchell_large_scene, 4, "Cornell 256x256 @4 spp, Mitchell filter (negative lobes), MegaPath"),
    "cornell_options_large": (cornell_options_large_scene, 4, "Cornell 256x256 @4 spp, film exposure + clamp, two-sided scaled light, MegaPath"),
    "medium_channels_large": (medium_channels_large_scene, 4, "row a22: isotropic medium with per-channel coefficients, 320x180 @4 spp, MegaVPTNaive (GCC build)"),
    "medium_hg_large": (medium_hg_large_scene, 4, "row a22: Henyey-Greenstein medium g = 0.3, 320x180 @4 spp, MegaVPTNaive (GCC build)"),
    "materials_wavepath_large": (materials_wavepath_large_scene, 2, "row f3 through WavePath: the materials box + Mix, 320x240 @2 spp, depth 10, Russian roulette from depth 2"),
    "textured_wavepath_large": (textured_wavepath_large_scene, 2, "row f1 through WavePath: textures + wrappers, 320x240 @2 spp"),
    "cornell_disney_odd": (cornell_disney_odd_scene, 3, "Cornell box with Disney surfaces at an odd film size, 333x187 @3 spp, MegaPath"),
    "subdivision_large": (subdivision_large_scene, 4, "the LoopSubdiv shape (closed / open / valence-3 base meshes, limit normals) 320x240 @4 spp, MegaPath"),
    "swizzle_large": (swizzle_large_scene, 4, "the Swizzle texture (image channels reordered / picked, constants, nesting) on the textured room, 320x240 @4 spp, MegaPath"),
    "checkerboard_large": (checkerboard_large_scene, 4, "the Checkerboard texture with constant squares on the textured room, 320x240 @4 spp, MegaPath"),
    "environment_large": (environment_large_scene, 4, "row a12: image-lit Spherical environment (importance map, MIS compensation) + area light, 320x200 @4 spp, MegaPath (~20 min: the 2048x1024 importance-map kernels run on the interpreter)"),
}


def film_digest(film: np.ndarray) -> dict:
    film = np.ascontiguousarray(film, dtype=np.float32)
    h, w = film.shape[:2]
    bh, bw = h // 32, w // 32  # whole blocks only (270 rows: 8 blocks, the last 14 rows are covered by the hash alone)
    blocks = film[: bh * 32, : bw * 32, :3].reshape(bh, 32, bw, 32, 3).mean(axis=(1, 3))
    return {"sha256": hashlib.sha256(film.tobytes()).hexdigest(), "resolution": [w, h], "mean_rgb": [float(x) for x in film[..., :3].mean(axis=(0, 1))],
            "block_means_32x32": [[[round(float(c), 6) for c in px] for px in row] for row in blocks]}


def main() -> int:
    import gen_ref_renders as G

    if not G.CLI.exists():
        print(f"{G.CLI} is missing: run `make -C oracle/ref` (needs /root/reference)", file=sys.stderr)
        return 1
    names = [sys.argv[1]] if len(sys.argv) > 1 else list(CASES)
    data = json.loads(OUT.read_text()) if OUT.exists() else {}
    for name in names:
        scene, spp, what = CASES[name]
        source = scene()
        t0 = time.time()
        if len(sys.argv) > 2:  # an already rendered film (EXR written by the reference CLI)
            film = G.read_image(Path(sys.argv[2]))
        else:
            keep = os.environ.get("LRK_REF_KEEP")  # a directory to render into and keep the EXR files in (for diffing)
            with tempfile.TemporaryDirectory() as tmp:
                film = G.render_with_reference(source, Path(keep or tmp), name, timeout=3600, quiet=True)
        digest = film_digest(film)
        digest["config"] = what
        digest["spp"] = spp
        digest["scene_sha256"] = hashlib.sha256(source.encode()).hexdigest()
        digest["render_seconds"] = round(time.time() - t0, 1)
        data[name] = digest
        print(f"{name}: sha256 {digest['sha256']}, mean rgb {digest['mean_rgb']}")
    merged = json.loads(OUT.read_text()) if OUT.exists() else {}  # (another case may have been written meanwhile)
    merged.update({name: data[name] for name in names})
    OUT.write_text(json.dumps(merged, indent=1) + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
