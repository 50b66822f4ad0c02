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
    assets_early = "tests/golden/assets"
    c = {}
    # config C1 / C2: Cornell box, matte + area light; the wavefront integrator and the megakernel one
    c["cornell_wavepath"] = scenes.cornell_box(resolution=(24, 24), spp=4)
    c["cornell_megapath"] = scenes.cornell_box(resolution=(24, 24), spp=4).replace("integrator : WavePath", "integrator : MegaPath")
    c["cornell_russian_roulette"] = scenes.cornell_box(resolution=(24, 24), spp=4, depth=12, rr_depth=2, rr_threshold=0.95)
    # row a2 (filter importance sampling through the 64-tap LUT + alias table) with the four non-trivial reconstruction
    # filters; rows a19 / a20: film clamp and exposure; a9: a two-sided, scaled area light
    base = scenes.cornell_box(resolution=(24, 24), spp=4)
    for filt, radius in (("Gaussian", 1.5), ("Triangle", 1.0), ("Mitchell", 2.0), ("LanczosSinc", 1.5)):
        c[f"cornell_filter_{filt.lower()}"] = base.replace("filter : Box { radius { 0.5 } }", f"filter : {filt} {{ radius {{ {radius} }} }}")
    c["cornell_film_and_light_options"] = (
        base.replace("film : Color { resolution { 24, 24 } }", "film : Color { resolution { 24, 24 } exposure { 1.5 } clamp { 1.25 } }")
            .replace("emission : Constant { v { 17.0, 12.0, 4.0 } } }", "emission : Constant { v { 17.0, 12.0, 4.0 } } two_sided { true } scale { 0.75 } }"))
    # config C3: instanced Loop-subdivision spheres, Disney closures, two area lights (reduced triangle count)
    c["spheres_disney"] = scenes.instanced_spheres(resolution=(32, 18), spp=2, depth=6, **spheres)
    # config C4: + homogeneous medium, MegaVPTNaive
    c["spheres_medium"] = scenes.instanced_spheres(resolution=(32, 18), spp=2, depth=6, medium=True, **spheres)
    # row f3: Mirror, Glass (smooth / rough), Plastic and Metal closures; Russian roulette with the refraction eta scale
    c["materials_wavepath"] = scenes.materials_box(resolution=(32, 24), spp=4, depth=8)
    c["materials_megapath_rr"] = scenes.materials_box(resolution=(32, 24), spp=4, depth=10, rr_depth=2, rr_threshold=0.95,
                                                      integrator="MegaPath")
    c["materials_mix"] = scenes.materials_box(resolution=(32, 24), spp=4, depth=10, rr_depth=2, mix=True, output="mix.exr")
    # Metal with the reference's built-in measured spectra: gold, copper (on the mirror ball's place) and an unknown name (-> aluminium)
    named = (scenes.materials_box(resolution=(32, 24), spp=4, depth=6, output="metals.exr")
             .replace('Surface m_mirror : Mirror { Kd : Constant { v { 0.95, 0.8, 0.6 } } }',
                      'Surface m_mirror : Metal { eta { "Cu" } roughness : Constant { v { 0.15 } } }')
             .replace('Surface m_glass : Glass { eta { "BK7" } }', 'Surface m_glass : Metal { eta { "unobtainium" } roughness : Constant { v { 0.3 } } }'))
    import re as _re
    named = _re.sub(r'Surface m_metal : Metal \{ eta \{ [0-9., ]+\} ', 'Surface m_metal : Metal { eta { "Gold" } ', named)
    assert named.count('eta { "Gold" }') == 1 and named.count('eta { "Cu" }') == 1
    c["materials_named_metals"] = named
    # row f3 with image-textured parameters of Mirror / Glass / Plastic / Metal: the closure contexts are derived per hit
    c["materials_textured"] = scenes.textured_materials(resolution=(48, 30), spp=4, depth=6, assets=assets_early)
    # row f3's Layered surface: Glass over Matte with a scattering slab, Glass over Mirror with a black (attenuating) slab, Plastic over
    # Matte with the node's defaults; with Russian roulette in the second case
    c["materials_layered"] = scenes.layered_box(resolution=(32, 24), spp=4, depth=6, output="layered.exr")
    c["materials_layered_rr"] = scenes.layered_box(resolution=(32, 24), spp=4, depth=10, rr_depth=2, rr_threshold=0.95, integrator="MegaPath", output="layered_rr.exr")
    c["flatten_stress"] = scenes.flatten_stress()
    # the LoopSubdiv shape: closed, open (boundary / corner rules) and valence-3 base meshes, limit normals, level 0 pass-through
    c["subdivision"] = scenes.subdivision_scene(resolution=(64, 48), spp=4)
    # every Disney parameter (fake subsurface via flatness, anisotropy, sheen tint, clearcoat gloss, eta) on the sphere scene
    extra = ("  anisotropic : Constant { v { 0.6 } }\n  sheen_tint : Constant { v { 0.7 } }\n  clearcoat_gloss : Constant { v { 0.3 } }\n"
             "  flatness : Constant { v { 0.4 } }\n  eta : Constant { v { 1.33 } }\n}")
    c["spheres_disney_all_lobes"] = "\n".join(
        (line if not (line == "}" and prev.strip().startswith("sheen : Constant")) else extra)
        for prev, line in zip([""] + c["spheres_disney"].split("\n"), c["spheres_disney"].split("\n"))).replace('"spheres.exr"', '"lobes.exr"')
    # transmissive Disney surfaces (closure class "disney_trans": specular-transmission lobe, enter / exit events and the
    # Russian-roulette eta scale) next to opaque ones: every other Disney node of the sphere scene gets specular_trans
    def _transmissive(src):
        out, k = [], 0
        for line in src.split("\n"):
            out.append(line)
            if line.strip().startswith("sheen : Constant"):
                if k % 2 == 0:
                    out.append("  specular_trans : Constant { v { %s } }" % ("0.85" if k % 4 == 0 else "1.0"))
                    out.append("  eta : Constant { v { 1.45 } }")
                k += 1
        return "\n".join(out)
    c["spheres_disney_transmissive"] = _transmissive(
        scenes.instanced_spheres(resolution=(32, 18), spp=4, depth=8, rr_depth=2, **spheres)).replace('"spheres.exr"', '"trans.exr"')
    # thin Disney surfaces (closure class "disney_thin": five techniques, rescaled transmission distribution, Lambertian diffuse
    # transmission, "through" events) next to transmissive and opaque ones - all three closure classes, each with its own lobe union;
    # under the wavefront integrator with Russian roulette, and under the volume integrator (a through event keeps the medium)
    def _thin(src):
        out, k = [], 0
        for line in src.split("\n"):
            out.append(line)
            if line.strip().startswith("sheen : Constant"):
                if k % 4 == 0:
                    out += ["  thin { true }", "  specular_trans : Constant { v { 0.8 } }", "  diffuse_trans : Constant { v { 0.5 } }",
                            "  flatness : Constant { v { 0.3 } }", "  eta : Constant { v { 1.45 } }"]
                elif k % 4 == 1:
                    out += ["  thin { true }", "  diffuse_trans : Constant { v { 0.7 } }"]
                elif k % 4 == 2:
                    out += ["  specular_trans : Constant { v { 0.9 } }", "  eta : Constant { v { 1.33 } }"]
                k += 1
        assert k >= 4
        return "\n".join(out)
    c["spheres_disney_thin"] = _thin(
        scenes.instanced_spheres(resolution=(32, 18), spp=4, depth=8, rr_depth=2, **spheres)).replace('"spheres.exr"', '"thin.exr"')
    c["spheres_medium_disney_thin"] = _thin(
        scenes.instanced_spheres(resolution=(32, 18), spp=4, depth=6, medium=True, **spheres)).replace('"spheres.exr"', '"thinvpt.exr"')
    c["spheres_disney_thin_paddedsobol"] = (c["spheres_disney_thin"].replace("sampler : Independent", "sampler : PaddedSobol")
                                            .replace("spp { 4 }", "spp { 3 }").replace('"thin.exr"', '"thinps.exr"'))
    # row f2: the table-driven samplers.  Cornell @4 spp with each; the sphere scene with sample counts that are NOT the samplers'
    # favourite powers (pmj02bn: 8 is no power of 4 -> its pixel-sample sorting skips entries; Sobol' / PaddedSobol: 3 is no power of
    # 2; ZSobol: log2(2) is odd -> the half-digit branch), depth 6 = 30 dimensions per path
    for smp in ("PMJ02BN", "Sobol", "PaddedSobol", "ZSobol"):
        c[f"cornell_sampler_{smp.lower()}"] = base.replace("sampler : Independent", f"sampler : {smp}").replace('"cornell.exr"', f'"{smp.lower()}.exr"')
    for smp, n in (("PMJ02BN", 8), ("Sobol", 3), ("PaddedSobol", 3), ("ZSobol", 2)):
        c[f"spheres_sampler_{smp.lower()}"] = (scenes.instanced_spheres(resolution=(32, 18), spp=n, depth=6, **spheres)
                                               .replace("sampler : Independent", f"sampler : {smp}").replace('"spheres.exr"', f'"s{smp.lower()}.exr"'))
    # the medium path with an isotropic phase function (|g| < 1e-3 branch) and per-channel coefficients
    c["spheres_medium_isotropic"] = (c["spheres_medium"].replace("g { 0.3 }", "g { 0.0 }")
                                     .replace("sigma_a : Constant { v { 0.01, 0.01, 0.01 } }", "sigma_a : Constant { v { 0.02, 0.01, 0.005 } }")
                                     .replace("sigma_s : Constant { v { 0.05, 0.05, 0.05 } }", "sigma_s : Constant { v { 0.03, 0.06, 0.12 } }")
                                     .replace('"spheres.exr"', '"iso.exr"'))
    # row a22 beyond config C4: media bound to shapes - the medium tracker, enter / exit events at Glass shells, the transmittance
    # walk through them; with an environment medium around (three media, nested); the true_hit(tag) quirk of the reference
    c["media_shapes"] = scenes.media_box(resolution=(32, 32), spp=4)
    c["media_nested_in_environment_medium"] = scenes.media_box(resolution=(32, 32), spp=4, environment_medium=True, rr_depth=2, output="nested.exr")
    c["media_true_hit_quirk"] = scenes.media_box(resolution=(32, 32), spp=4, skip_quirk=True, output="quirk.exr")
    # row f1: image textures (8 / 16-bit PNG, grey, palette; all address modes, point + bilinear, sRGB / linear / gamma) on
    # Matte and Disney parameters; with wrappers: normal map, alpha-tested cut-out (ray queries), constant opacity.
    # mesh_files=False: the `Mesh` plugin of the reference needs assimp, which is not built
    assets = "tests/golden/assets"
    c["textured"] = scenes.textured_room(resolution=(32, 24), spp=2, mesh_files=False, assets=assets)
    c["textured_wrappers"] = scenes.textured_room(resolution=(32, 24), spp=2, mesh_files=False, assets=assets, wrappers=True)
    # an area light with an image emission, read at the uv of emitter hits and of sampled light points; the camera looks up at the
    # lamp so that primary rays hit it too
    c["textured_light"] = (scenes.textured_room(resolution=(32, 24), spp=4, mesh_files=False, assets=assets, textured_light=True, output="texlight.exr")
                           .replace("position { 0.0, 1.4, 4.2 }", "position { 0.0, 0.6, 4.2 }").replace("front { 0.0, -0.2, -1.0 }", "front { 0.0, 0.25, -1.0 }"))
    # BMP and TGA textures in every storage variant the host readers accept, against what stb_image hands the reference for them
    c["image_formats"] = scenes.image_formats_scene(resolution=(80, 48), spp=2, assets=assets)
    # JPEG textures (lossy: the texels are the reference's only if the decoder's arithmetic is stb_image's - csrc/host/jpegload.cpp)
    c["jpeg_formats"] = scenes.image_formats_scene(resolution=(120, 72), spp=2, assets=assets, files=scenes.JPEG_FORMAT_FILES, output="jpegs.exr")
    # a thin Disney surface whose diffuse_trans (slot 15) and colour are image textures, next to a constant specular_trans
    def _textured_thin(src):
        old = "  metallic : Constant { v { 0.2 } }"
        assert src.count(old) == 1
        return src.replace(old, old + f'''
  thin {{ true }}
  diffuse_trans : Image {{ file {{ "{assets}/rough_gray8.png" }} encoding {{ "linear" }} address {{ "mirror" }} uv_scale {{ 1.5 }} }}
  specular_trans : Constant {{ v {{ 0.4 }} }}''')
    c["textured_disney_thin"] = _textured_thin(scenes.textured_room(resolution=(32, 24), spp=4, mesh_files=False, assets=assets, output="texthin.exr"))
    # feature crossings: the table-driven samplers under closures that draw a different number of dimensions per bounce
    c["materials_mix_sobol"] = (scenes.materials_box(resolution=(32, 24), spp=3, depth=6, rr_depth=2, mix=True, output="mixsobol.exr")
                                .replace("sampler : Independent", "sampler : Sobol"))
    c["materials_layered_pmj02bn"] = (scenes.layered_box(resolution=(32, 24), spp=4, depth=6, output="layeredpmj.exr")
                                      .replace("sampler : Independent", "sampler : PMJ02BN"))
    c["textured_materials_zsobol"] = (scenes.textured_materials(resolution=(48, 30), spp=2, depth=6, assets=assets_early, output="texzsobol.exr")
                                      .replace("sampler : Independent", "sampler : ZSobol"))
    # media bound to shapes whose shells are Disney surfaces: a thin one (its "through" events never enter the medium behind it) and
    # a transmissive one (enter / exit with the eta scale), inside an environment medium
    c["media_disney_shells"] = (scenes.media_box(resolution=(32, 32), spp=4, environment_medium=True, rr_depth=2, output="shells.exr")
                                .replace('Surface shell_smooth : Glass { eta { "BK7" } }',
                                         'Surface shell_smooth : Disney { color : Constant { v { 0.9, 0.8, 0.6 } } thin { true } specular_trans : Constant { v { 0.7 } } '
                                         'diffuse_trans : Constant { v { 0.6 } } roughness : Constant { v { 0.3 } } }')
                                .replace('Surface shell_rough : Glass { Kr : Constant { v { 1.0, 0.95, 0.9 } } Kt : Constant { v { 0.9, 0.95, 1.0 } } '
                                         'roughness : Constant { v { 0.25 } } eta { 1.33 } }',
                                         'Surface shell_rough : Disney { color : Constant { v { 0.6, 0.8, 0.9 } } specular_trans : Constant { v { 0.9 } } '
                                         'roughness : Constant { v { 0.25 } } eta : Constant { v { 1.33 } } }'))
    assert "thin { true }" in c["media_disney_shells"] and c["media_disney_shells"].count("Glass") == 0
    # the Swizzle texture: reordered image channels, one channel as a scalar parameter, swizzled constants, nesting
    c["swizzle"] = scenes.swizzle_scene(resolution=(64, 48), spp=4, assets=assets)
    # the Checkerboard texture with constant squares (baked into a point-sampled, repeating 2x2 image by the host)
    c["checkerboard"] = scenes.checkerboard_scene(resolution=(64, 48), spp=4, assets=assets)
    # row a12: Spherical environment with an image emission (importance map, MIS compensation) next to an area light
    c["environment_image"] = scenes.environment_scene(resolution=(32, 20), spp=2, emission="image", assets=assets, sky_file="sky.exr")
    # the volume integrator under an image-lit environment: environment misses / environment NEE from inside a homogeneous
    # environment medium, next to an area light, with a thin Disney ball (through events)
    c["environment_medium_thin"] = (
        scenes.environment_scene(resolution=(32, 20), spp=2, emission="image", assets=assets, sky_file="sky.exr", output="envvpt.exr")
        .replace("integrator : WavePath {", "integrator : MegaVPTNaive {")
        .replace("  cameras { @camera }", "  environment_medium : Homogeneous {\n    sigma_a : Constant { v { 0.02, 0.03, 0.05 } }\n    sigma_s : Constant { v { 0.15, 0.12, 0.1 } }\n"
                 "    phasefunction : HenyeyGreenstein { g { 0.4 } }\n  }\n  cameras { @camera }")
        .replace("Surface ball_s : Matte { Kd : Constant { v { 0.8, 0.8, 0.8 } } }",
                 "Surface ball_s : Disney { color : Constant { v { 0.8, 0.7, 0.5 } } thin { true } diffuse_trans : Constant { v { 0.8 } } specular_trans : Constant { v { 0.3 } } }"))
    assert "MegaVPTNaive" in c["environment_medium_thin"] and "thin { true }" in c["environment_medium_thin"] and "environment_medium" in c["environment_medium_thin"]
    # the headline scenes at full geometric size (1 387 526 instanced triangles: Loop-subdivision spheres at level 7 and 3
    # built by the reference's own Sphere plugin), low resolution: config C3 (Disney + NEE) and config C4 (+ medium, depth 8)
    c["config_c3_full_scene"] = scenes.instanced_spheres(resolution=(96, 54), spp=2, output="c3.exr")
    c["config_c4_full_scene"] = scenes.instanced_spheres(resolution=(96, 54), spp=2, medium=True, depth=8, output="c4.exr")
    return c


def read_image(path: Path) -> np.ndarray:
    """The film the reference wrote (EXR, fp32 ZIP, through its tinyexr) decoded by this repository's own reader."""
    import ctypes as C

    from luisarender_b200 import _ffi as F

    host = F.host_lib()
    host.lrh_load_image.argtypes = [C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_void_p, C.c_uint64]
    w, h, ch = C.c_uint32(), C.c_uint32(), C.c_uint32()
    if host.lrh_load_image(str(path).encode(), C.byref(w), C.byref(h), C.byref(ch), None, 0) != 0:
        raise RuntimeError(host.lrh_last_error().decode())
    out = np.zeros((h.value, w.value, 4), dtype=np.float32)
    if host.lrh_load_image(str(path).encode(), C.byref(w), C.byref(h), C.byref(ch), out.ctypes.data, out.size) != 0:
        raise RuntimeError(host.lrh_last_error().decode())
    return out


def render_with_reference(source: str, workdir: Path, name: str = "scene", timeout: int = 1200, quiet: bool = False) -> np.ndarray:
    """Runs the reference CLI on `source`; returns the RGBA film it saved (the camera's `file`, an fp32 EXR)."""
    import re

    scene_path = workdir / f"{name}.luisa"
    # asset paths in the scene text are relative to the repository root (the product resolves them the same way)
    scene_path.write_text(source.replace('"tests/golden/assets/', f'"{REPO}/tests/golden/assets/'))
    out_name = re.search(r'Camera\b.*?\bfile\s*\{\s*"([^"]+)"\s*\}', source, re.S).group(1)  # the camera's output file
    # (the interp backend runs small dispatches - everything that touches the film - on one thread: reproducible atomics)
    # (large medium renders: MegaVPTNaive logs every path vertex through device_log - do not keep gigabytes of it)
    log = subprocess.run([str(CLI), "-b", "interp", scene_path.name], cwd=workdir, text=True, timeout=timeout,
                         **({"stdout": subprocess.DEVNULL, "stderr": subprocess.DEVNULL} if quiet else {"capture_output": True}))
    out = workdir / out_name
    if not out.exists():
        raise RuntimeError(f"reference render of '{name}' failed:\n{(log.stdout or '')[-2000:]}\n{(log.stderr or '')[-2000:]}")
    return read_image(out)


def main() -> int:
    """python tools/gen_ref_renders.py [name ...]: without names every case is rendered (the environment case alone takes
    ~17 min: its 2048x1024 importance-map kernels run on the interpreter); with names only those, merged into the fixture."""
    if not CLI.exists():
        print(f"{CLI} is missing: run `make -C oracle/ref` (needs /root/reference)", file=sys.stderr)
        return 1
    only = set(sys.argv[1:])
    data = dict(np.load(OUT)) if only and OUT.exists() else {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, source in cases().items():
            if only and name not in only:
                continue
            image = render_with_reference(source, Path(tmp), name)
            data[f"{name}/scene"] = np.frombuffer(source.encode(), dtype=np.uint8)
            data[f"{name}/image"] = image
            print(f"{name:28s} {image.shape[1]}x{image.shape[0]}  mean rgb = {image[..., :3].mean():.6f}")
    np.savez_compressed(OUT, **data)
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
