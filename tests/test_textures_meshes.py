"""mesh files and image textures — host loaders, the software texture sampler and the per-hit
parameter evaluation, checked on the CPU (oracle + host library).  The CUDA side is compared with the oracle in
tests/test_gpu_parity.py (scene "textured")."""
from __future__ import annotations

import ctypes as C
import struct
import sys
import zlib
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests" / "golden"))

import make_assets  # noqa: E402
from luisarender_b200 import _ffi as F  # noqa: E402
from luisarender_b200 import scenes  # noqa: E402
from luisarender_b200.api import Scene  # noqa: E402
from oracle import binding as O  # noqa: E402

ASSETS = REPO / "tests" / "golden" / "assets"
f32 = np.float32


def _matte_scene(texture_props: str, mesh: str = "") -> str:
    """One textured quad (or mesh file) + light + camera: the smallest scene that carries a texture / mesh through the loader."""
    shape = mesh or """Shape quad : InlineMesh {
  positions { -1.0, 0.0, 1.0,  1.0, 0.0, 1.0,  1.0, 0.0, -1.0,  -1.0, 0.0, -1.0 }
  uvs { 0.0, 0.0,  1.0, 0.0,  1.0, 1.0,  0.0, 1.0 }
  indices { 0, 1, 2, 0, 2, 3 }
  surface { @s }
}"""
    return f"""
Surface s : Matte {{ Kd : Image {{ {texture_props} }} }}
Light l : Diffuse {{ emission : Constant {{ v {{ 5.0 }} }} }}
{shape}
Shape lamp : InlineMesh {{
  positions {{ -0.5, 2.0, 0.5,  -0.5, 2.0, -0.5,  0.5, 2.0, -0.5,  0.5, 2.0, 0.5 }}
  indices {{ 0, 1, 2, 0, 2, 3 }}
  light {{ @l }}
}}
Camera c : Pinhole {{ position {{ 0.0, 3.0, 0.1 }} front {{ 0.0, -1.0, -0.02 }} up {{ 0.0, 0.0, -1.0 }} fov {{ 40.0 }} spp {{ 1 }}
  film : Color {{ resolution {{ 8, 8 }} }} }}
render {{ integrator : WavePath {{ depth {{ 2 }} }} cameras {{ @c }} shapes {{ @quad, @lamp }} }}
"""


def _texels(desc, i=0) -> np.ndarray:
    t = desc.textures[i]
    a = np.ctypeslib.as_array(desc.texels, shape=(desc.texel_count * 4,))
    return a[t.texel_offset * 4:(t.texel_offset + t.width * t.height) * 4].reshape(t.height, t.width, 4).copy()


# ---- image loaders ------------------------------------------------------------------------------------------------------
def test_png_loader_matches_source_arrays():
    for name, src, norm in [("checker_rgb8.png", make_assets.checker_rgb8(), 255.0), ("ramp_rgba16.png", make_assets.ramp_rgba16(), 65535.0)]:
        sc = Scene.from_source(_matte_scene(f'file {{ "{ASSETS / name}" }}'), REPO)
        d = sc.desc()
        tex = _texels(d)
        assert d.textures[0].channels == 4 and tex.shape[:2] == src.shape[:2]
        c = src.shape[2]
        assert np.array_equal(tex[..., :c], src.astype(f32) / f32(norm))  # x / 255, x / 65535 (cpu_texture.h:63)
        if c == 3:
            assert (tex[..., 3] == 1).all()
    grey = Scene.from_source(_matte_scene(f'file {{ "{ASSETS / "rough_gray8.png"}" }}'), REPO).desc()
    assert grey.textures[0].channels == 1
    assert np.array_equal(_texels(grey)[..., 0], make_assets.rough_gray8().astype(f32) / f32(255))
    pal = _texels(Scene.from_source(_matte_scene(f'file {{ "{ASSETS / "palette4.png"}" }}'), REPO).desc())
    colours = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0]], f32)
    idx = np.add.outer(np.arange(8), np.arange(8)) % 4
    assert np.array_equal(pal[..., :3], colours[idx])


def test_bmp_and_tga_loaders_match_source_arrays():
    """Every BMP / TGA storage variant the readers accept (tests/golden/make_assets.py: palettes of 1 / 4 / 8 bits, core header, 16-bit
    5-5-5 and 5-6-5 bit fields, 24 / 32 bits, top-down, an all-zero alpha byte; TGA raw / run-length with runs across rows, grey,
    grey + alpha, colour-mapped, 15-bit) against the picture it was written from; channel counts as the reference's storage policy
    has them (stb_image's component count: 1, 2, or RGBA).  The renders of tests/golden/ref_renders.npz: image_formats pin the
    same files against the reference's own loader."""
    pictures = make_assets.bmp_tga_pictures()
    for name, want in pictures.items():
        if name.startswith("_"):
            continue
        ext = name.split("_")[0]
        d = Scene.from_source(_matte_scene(f'file {{ "{ASSETS / (name + "." + ext)}" }} encoding {{ "linear" }}'), REPO).desc()
        tex = _texels(d)
        c = want.shape[2]
        assert d.textures[0].channels == c and tex.shape[:2] == want.shape[:2], name
        assert np.array_equal(tex[..., :c] if c != 2 else tex[..., :2], want.astype(f32) / f32(255)), name
        if c < 4:
            assert (tex[..., 3] == 1).all(), name


def test_jpeg_reader_matches_the_references_decoder_byte_for_byte():
    """A JPEG decode is lossy: the texels only equal the reference's if the arithmetic after the entropy decoder is stb_image's (csrc/host/
    jpegload.cpp).  tests/golden/jpeg_texels.npz is what the reference's own stb_image returns (tools/gen_jpeg_pins.py) for the fixtures of
    tests/golden/make_assets.py: write_jpegs - grey, 4:4:4, 4:2:2, 4:2:0, 4:4:0, 4:1:1, one-pixel-wide pictures, restart intervals, optimised
    tables, quality 12 and 100, progressive (interleaved DC, refinement scans, restarts), RGB stored as such (by ids / by Adobe marker), CMYK and YCCK."""
    want_all = np.load(REPO / "tests" / "golden" / "jpeg_texels.npz")
    assert len(want_all.files) >= 20
    for name in want_all.files:
        want = want_all[name]
        d = Scene.from_source(_matte_scene(f'file {{ "{ASSETS / (name + ".jpg")}" }} encoding {{ "linear" }}'), REPO).desc()
        tex = _texels(d)
        c = want.shape[2]
        assert c in (1, 4) and d.textures[0].channels == c and tex.shape[:2] == want.shape[:2], name
        assert np.array_equal(tex[..., :c], want.astype(f32) / f32(255)), (name, int((tex[..., :c] != want.astype(f32) / f32(255)).sum()))
        if c == 1:
            assert (tex[..., 3] == 1).all(), name


def test_jpeg_pins_are_what_the_reference_decodes_now():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_jpeg_pins", REPO / "tools" / "gen_jpeg_pins.py")
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    if not gen.LIB.exists():
        pytest.skip("oracle/_ref is not built here (the reference's sources are not on this machine)")
    now, kept = gen.decode_all(), np.load(REPO / "tests" / "golden" / "jpeg_texels.npz")
    assert sorted(now) == sorted(kept.files)
    for name, a in now.items():
        assert np.array_equal(a, kept[name]), name


def test_jpeg_reader_refuses_what_it_does_not_read(tmp_path):
    src = bytearray((ASSETS / "jpg_444.jpg").read_bytes())
    sof = src.index(b"\xff\xc0")
    arithmetic = bytearray(src)
    arithmetic[sof + 1] = 0xC9
    (tmp_path / "arithmetic.jpg").write_bytes(arithmetic)
    with pytest.raises(Exception, match="unsupported JPEG process"):
        Scene.from_source(_matte_scene(f'file {{ "{tmp_path / "arithmetic.jpg"}" }}'), REPO)
    (tmp_path / "cut.jpg").write_bytes(src[:sof + 6])
    with pytest.raises(Exception, match="truncated JPEG"):
        Scene.from_source(_matte_scene(f'file {{ "{tmp_path / "cut.jpg"}" }}'), REPO)
    huge = bytearray(src)
    huge[sof + 5:sof + 9] = b"\xff\xff\xff\xff"  # a 65535 x 65535 frame header must not allocate its 25 GB of coefficients
    (tmp_path / "huge.jpg").write_bytes(huge)
    with pytest.raises(Exception, match="larger than"):
        Scene.from_source(_matte_scene(f'file {{ "{tmp_path / "huge.jpg"}" }}'), REPO)
    (tmp_path / "not.jpg").write_bytes(b"\x89PNG\r\n\x1a\n")
    with pytest.raises(Exception, match="not a JPEG"):
        Scene.from_source(_matte_scene(f'file {{ "{tmp_path / "not.jpg"}" }}'), REPO)


def test_bmp_and_tga_readers_refuse_what_the_reference_refuses(tmp_path):
    rle = bytearray((ASSETS / "bmp_pal8.bmp").read_bytes())
    rle[30] = 1  # BI_RLE8
    (tmp_path / "rle.bmp").write_bytes(rle)
    with pytest.raises(RuntimeError, match="RLE"):
        Scene.from_source(_matte_scene(f'file {{ "{tmp_path / "rle.bmp"}" }}'), REPO)
    cut = (ASSETS / "tga_rgb24.tga").read_bytes()[:200]
    (tmp_path / "cut.tga").write_bytes(cut)
    with pytest.raises(RuntimeError, match="truncated"):
        Scene.from_source(_matte_scene(f'file {{ "{tmp_path / "cut.tga"}" }}'), REPO)
    (tmp_path / "x.gif").write_bytes(b"GIF89a")
    with pytest.raises(RuntimeError, match="unsupported image format"):
        Scene.from_source(_matte_scene(f'file {{ "{tmp_path / "x.gif"}" }}'), REPO)


def test_png_filters_and_low_bit_depths(tmp_path):
    """Sub / Up / Average / Paeth scanline filters and a 1-bit greyscale file, written by hand."""
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (6, 5, 3), dtype=np.uint8)
    h, w, c = img.shape
    rows = b""
    prev = np.zeros(w * c, np.int32)
    for y in range(h):
        cur = img[y].reshape(-1).astype(np.int32)
        ft = 1 + (y % 4)
        left = np.concatenate([np.zeros(c, np.int32), cur[:-c]])
        upleft = np.concatenate([np.zeros(c, np.int32), prev[:-c]])
        if ft == 1:
            pred = left
        elif ft == 2:
            pred = prev
        elif ft == 3:
            pred = (left + prev) // 2
        else:
            p = left + prev - upleft
            pa, pb, pc = abs(p - left), abs(p - prev), abs(p - upleft)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, upleft))
        rows += bytes([ft]) + ((cur - pred) % 256).astype(np.uint8).tobytes()
        prev = cur

    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xFFFFFFFF)

    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(rows)) + chunk(b"IEND", b"")
    (tmp_path / "filtered.png").write_bytes(data)
    tex = _texels(Scene.from_source(_matte_scene(f'file {{ "{tmp_path / "filtered.png"}" }}'), REPO).desc())
    assert np.array_equal(tex[..., :3], img.astype(f32) / f32(255))
    bits = rng.integers(0, 2, (3, 10), dtype=np.uint8)
    packed = b"".join(b"\x00" + np.packbits(bits[y]).tobytes() for y in range(3))
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 10, 3, 1, 0, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(packed)) + chunk(b"IEND", b"")
    (tmp_path / "bits.png").write_bytes(data)
    d = Scene.from_source(_matte_scene(f'file {{ "{tmp_path / "bits.png"}" }}'), REPO).desc()
    assert d.textures[0].channels == 1 and np.array_equal(_texels(d)[..., 0], bits.astype(f32))


def test_float_image_formats(tmp_path):
    pic = (np.arange(4 * 2 * 3, dtype=f32).reshape(2, 4, 3) / f32(23.0))
    pfm = _texels(Scene.from_source(_matte_scene(f'file {{ "{ASSETS / "tiny.pfm"}" }}'), REPO).desc())
    assert np.array_equal(pfm[..., :3], pic)  # bottom-up storage flipped to top-first
    ppm = _texels(Scene.from_source(_matte_scene(f'file {{ "{ASSETS / "tiny.ppm"}" }}'), REPO).desc())
    assert np.array_equal(ppm[..., :3], np.round(pic * 255).astype(np.uint8).astype(f32) / f32(255))
    # EXR / HDR written by this repo's own film writer, read back by the texture loader
    from luisarender_b200.api import save_image
    rgba = np.concatenate([np.random.default_rng(1).uniform(0, 4, (5, 7, 3)).astype(f32), np.ones((5, 7, 1), f32)], axis=-1)
    save_image(tmp_path / "t.exr", rgba)
    exr = _texels(Scene.from_source(_matte_scene(f'file {{ "{tmp_path / "t.exr"}" }}'), REPO).desc())
    assert np.array_equal(exr, rgba)
    save_image(tmp_path / "t.hdr", rgba)
    d = Scene.from_source(_matte_scene(f'file {{ "{tmp_path / "t.hdr"}" }}'), REPO).desc()
    hdr = _texels(d)
    assert d.textures[0].encoding == F.TEX_ENCODING_LINEAR  # .hdr / .exr default to linear (image.cpp:93-98)
    # RGBE: 8-bit mantissas under the pixel's shared exponent
    assert (np.abs(hdr[..., :3] - rgba[..., :3]) <= rgba[..., :3].max(axis=-1, keepdims=True) / 100).all()
    assert np.array_equal(hdr[..., :3], hdr[..., :3].astype(np.float16).astype(f32))  # stored as HALF4 (imageio.cpp:383)
    (tmp_path / "nothing.gif").write_bytes(b"GIF89a")
    with pytest.raises(RuntimeError, match="unsupported image format"):
        Scene.from_source(_matte_scene('file { "nothing.gif" }'), tmp_path)
    with pytest.raises(RuntimeError, match="cannot open"):
        Scene.from_source(_matte_scene('file { "missing.png" }'), tmp_path)


# ---- sampler: numpy re-derivation of cpu_texture.h:418-464,489-493 in float32 ------------------------------------------------
def _coord_point(address, uv, s):
    one_minus_eps = f32(np.nextafter(f32(1), f32(0)))
    if address == F.TEX_ADDRESS_EDGE:
        return f32(min(max(uv, f32(0)), one_minus_eps) * s)
    if address == F.TEX_ADDRESS_REPEAT:
        return f32(f32(uv - np.floor(uv)) * s)
    if address == F.TEX_ADDRESS_MIRROR:
        uv = f32(np.fmod(abs(uv), f32(2)))
        uv = uv if uv < 1 else f32(f32(2) - uv)
        return f32(min(uv, one_minus_eps) * s)
    return f32(65536.0) if (uv < 0 or uv >= 1) else f32(uv * s)


def _read(tex, x, y):
    h, w = tex.shape[:2]
    return tex[y, x] if (x < w and y < h) else np.zeros(4, f32)


def _sample(tex, address, linear, u, v):
    h, w = tex.shape[:2]
    sx, sy = f32(w), f32(h)
    if not linear:
        return _read(tex, int(_coord_point(address, u, sx)), int(_coord_point(address, v, sy)))
    ix, iy = f32(1) / sx, f32(1) / sy
    ax, bx = _coord_point(address, f32(u - f32(.5) * ix), sx), _coord_point(address, f32(u + f32(.5) * ix), sx)
    ay, by = _coord_point(address, f32(v - f32(.5) * iy), sy), _coord_point(address, f32(v + f32(.5) * iy), sy)
    x0, x1, y0, y1 = min(ax, bx), max(ax, bx), min(ay, by), max(ay, by)
    tx, ty = f32(x1 - np.floor(x1)), f32(y1 - np.floor(y1))
    lerp = lambda a, b, t: (t * (b - a) + a).astype(f32)
    v00, v01, v10, v11 = _read(tex, int(x0), int(y0)), _read(tex, int(x1), int(y0)), _read(tex, int(x0), int(y1)), _read(tex, int(x1), int(y1))
    return lerp(lerp(v00, v01, tx), lerp(v10, v11, tx), ty)


@pytest.mark.parametrize("address", ["edge", "repeat", "mirror", "zero"])
@pytest.mark.parametrize("filt", ["point", "bilinear"])
def test_texture_sampling_matches_the_reference_sampler_exactly(address, filt):
    props = f'file {{ "{ASSETS / "checker_rgb8.png"}" }} address {{ "{address}" }} filter {{ "{filt}" }} encoding {{ "linear" }}'
    d = Scene.from_source(_matte_scene(props), REPO).desc()
    tex = _texels(d)
    t = d.textures[0]
    rng = np.random.default_rng(11)
    uvs = np.concatenate([rng.uniform(-2.5, 3.5, (300, 2)), [[0, 0], [1, 1], [0.5, 0.5], [1 - 1e-7, 0.25], [-1e-8, 0.999999], [1.0, 0.0]]]).astype(f32)
    lib = O.lib()
    for u, v in uvs:
        out = np.zeros(4, f32)
        lib.oracle_texture_evaluate(C.byref(d), 0, np.array([u, v], f32).ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_float)))
        ref = _sample(tex, t.address, filt == "bilinear", u, v)
        assert np.array_equal(out, ref), (address, filt, u, v, out, ref)


def test_texture_decode_scale_and_uv_transform():
    lib = O.lib()

    def evaluate(props, u, v):
        d = Scene.from_source(_matte_scene(props), REPO).desc()
        out = np.zeros(4, f32)
        lib.oracle_texture_evaluate(C.byref(d), 0, np.array([u, v], f32).ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_float)))
        return out, d

    file = f'file {{ "{ASSETS / "checker_rgb8.png"}" }} filter {{ "point" }}'
    lin, d = evaluate(file + ' encoding { "linear" }', 0.3, 0.7)
    srgb, _ = evaluate(file, 0.3, 0.7)  # .png defaults to sRGB (image.cpp:93-98)
    expect = np.where(lin <= 0.04045, lin / 12.92, ((lin + 0.055) / 1.055) ** 2.4)
    assert np.allclose(srgb, expect, rtol=2e-6)
    gam, _ = evaluate(file + ' encoding { "gamma" } gamma { 2.0 } scale { 0.5 }', 0.3, 0.7)
    assert np.allclose(gam, 0.5 * lin.astype(np.float64) ** 2, rtol=2e-6)
    # uv' = uv * uv_scale + uv_offset (image.cpp:136-141)
    moved, _ = evaluate(file + ' encoding { "linear" } uv_scale { 0.5, 2.0 } uv_offset { 0.05, 0.1 }', 0.5, 0.3)
    direct, _ = evaluate(file + ' encoding { "linear" }', f32(0.5) * f32(0.5) + f32(0.05), f32(0.3) * f32(2.0) + f32(0.1))
    assert np.array_equal(moved, direct)


def test_resolved_surface_parameters():
    """Colour slots: saturate + luminance; Matte sigma: saturate(x) * 90; Disney roughness remap (disney.cpp:932-956)."""
    d = Scene.from_source(scenes.textured_room(), REPO).desc()
    lib = O.lib()
    uv = np.array([0.37, 0.61], f32)
    uvp = uv.ctypes.data_as(C.POINTER(C.c_float))

    def tex(i):
        out = np.zeros(4, f32)
        lib.oracle_texture_evaluate(C.byref(d), i, uvp, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    s = F.Surface()
    lib.oracle_resolve_surface(C.byref(d), 1, uvp, C.byref(s))  # wall: Kd = 16-bit RGBA ramp * 0.9, sigma = grey texture
    assert np.allclose(list(s.p)[:3], np.clip(tex(1)[:3], 0, 1)) and s.p[3] == pytest.approx(float(np.clip(tex(2)[0], 0, 1)) * 90.0)
    lib.oracle_resolve_surface(C.byref(d), 2, uvp, C.byref(s))  # cube: Disney colour + roughness textures
    rgb = np.clip(tex(3)[:3], 0, 1)
    assert np.allclose(list(s.p)[:3], rgb)
    assert s.p[3] == pytest.approx(0.212671 * rgb[0] + 0.715160 * rgb[1] + 0.072169 * rgb[2], rel=1e-6)
    assert s.p[6] == pytest.approx(max(float(tex(4)[0]) ** 2, 1e-4), rel=1e-6) and s.p[4] == pytest.approx(0.2) and s.p[11] == pytest.approx(0.5)
    assert d.surfaces[2].flags & F.SURFACE_REMAP_ROUGHNESS and d.surfaces[2].flags & F.SURFACE_HAS_TEXTURES
    assert d.surfaces[2].lobes & 16  # clearcoat lobe enabled by the constant


# ---- mesh files ---------------------------------------------------------------------------------------------------------------
def _mesh_scene(file: str, extra: str = "") -> str:
    return _matte_scene(f'file {{ "{ASSETS / "checker_rgb8.png"}" }}', mesh=f'Shape quad : Mesh {{ file {{ "{file}" }} {extra} surface {{ @s }} }}')


def _mesh_arrays(desc, mesh_index=0):
    m = desc.meshes[mesh_index]
    verts = np.array([[*desc.vertices[m.vertex_offset + i].p, *desc.vertices[m.vertex_offset + i].n, *desc.vertices[m.vertex_offset + i].uv]
                      for i in range(m.vertex_count)], f32)
    tris = np.array([[desc.triangles[m.triangle_offset + i].i0, desc.triangles[m.triangle_offset + i].i1, desc.triangles[m.triangle_offset + i].i2]
                     for i in range(m.triangle_count)])
    return verts, tris


def test_obj_loader_triangulates_joins_and_generates_creased_normals():
    d = Scene.from_source(_mesh_scene(ASSETS / "cube.obj"), REPO).desc()
    verts, tris = _mesh_arrays(d)
    assert len(tris) == 12 and len(verts) == 24  # 6 quads -> 12 triangles; 90-degree creases keep 4 vertices per face
    for t in tris:
        p = verts[t, :3]
        face_n = np.cross(p[1] - p[0], p[2] - p[0])
        face_n /= np.linalg.norm(face_n)
        assert np.allclose(verts[t, 3:6], face_n, atol=1e-6)  # 45-degree smoothing limit: faceted cube
        assert np.allclose(np.abs(face_n).max(), 1.0) and (p @ face_n > 0).all()  # outward winding preserved
    assert set(map(tuple, verts[:, 6:8])) == {(0.0, 1.0), (1.0, 1.0), (1.0, 0.0), (0.0, 0.0)}
    inst0 = d.instances[0]
    assert inst0.handle[0] & 1023 & F.SHAPE_HAS_VERTEX_NORMAL and inst0.handle[0] & 1023 & F.SHAPE_HAS_VERTEX_UV
    # v is flipped unless flip_uv is set (mesh.cpp:62): corner 1 of the first face has vt (0,0) -> (0,1)
    raw = _mesh_arrays(Scene.from_source(_mesh_scene(ASSETS / "cube.obj", "flip_uv { true }"), REPO).desc())[0]
    assert np.array_equal(raw[:, 6], verts[:, 6]) and np.allclose(raw[:, 7], 1 - verts[:, 7])
    dropped = Scene.from_source(_mesh_scene(ASSETS / "cube.obj", "drop_normal { true } drop_uv { true }"), REPO).desc()
    v2, t2 = _mesh_arrays(dropped)
    assert len(v2) == 8 and len(t2) == 12 and (v2[:, 3:6] == [0, 0, 1]).all() and (v2[:, 6:] == 0).all()
    assert dropped.instances[0].handle[0] & 1023 & (F.SHAPE_HAS_VERTEX_NORMAL | F.SHAPE_HAS_VERTEX_UV) == 0


def test_obj_smooth_normals_negative_indices_and_errors(tmp_path):
    # a shallow pyramid (faces 20 degrees apart): normals are averaged across the apex; indices given relative to the end
    (tmp_path / "pyr.obj").write_text("v 0 0.1 0\nv 1 0 1\nv 1 0 -1\nv -1 0 -1\nv -1 0 1\n"
                                      "f -5 -4 -3\nf -5 -3 -2\nf -5 -2 -1\nf -5 -1 -4\n")
    verts, tris = _mesh_arrays(Scene.from_source(_mesh_scene(tmp_path / "pyr.obj"), REPO).desc())
    assert len(tris) == 4
    apex = verts[np.all(np.isclose(verts[:, :3], [0, 0.1, 0]), axis=1)]
    assert len(apex) == 1 and np.allclose(apex[0, 3:6], [0, 1, 0], atol=1e-6)
    with pytest.raises(RuntimeError, match="index out of range"):
        (tmp_path / "bad.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 4\n")
        Scene.from_source(_mesh_scene(tmp_path / "bad.obj"), REPO)
    with pytest.raises(RuntimeError, match="unsupported mesh format"):
        Scene.from_source(_mesh_scene(tmp_path / "mesh.fbx"), REPO)
    with pytest.raises(RuntimeError, match="subdivision"):
        Scene.from_source(_mesh_scene(ASSETS / "cube.obj", "subdivision { 1 }"), REPO)


def test_ply_ascii_and_binary_agree():
    a = _mesh_arrays(Scene.from_source(_mesh_scene(ASSETS / "tetra_ascii.ply"), REPO).desc())
    b = _mesh_arrays(Scene.from_source(_mesh_scene(ASSETS / "tetra_binary.ply"), REPO).desc())
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    verts, tris = a
    assert len(tris) == 4 and len(verts) == 12  # all dihedral angles of a tetrahedron exceed 45 degrees
    assert np.allclose(np.linalg.norm(verts[:, 3:6], axis=1), 1.0, atol=1e-6)


def test_identical_mesh_files_are_deduplicated_and_instanced():
    src = _matte_scene(f'file {{ "{ASSETS / "checker_rgb8.png"}" }}', mesh=f"""
Shape a : Mesh {{ file {{ "{ASSETS / "cube.obj"}" }} surface {{ @s }} }}
Shape b : Mesh {{ file {{ "{ASSETS / "cube.obj"}" }} surface {{ @s }} transform : SRT {{ translate {{ 2.0, 0.0, 0.0 }} }} }}
Shape quad : Group {{ shapes {{ @a, @b }} }}""")
    info = Scene.from_source(src, REPO).info()
    assert info["instances"] == 3 and info["meshes"] == 2  # two cube instances share one mesh record (geometry.cpp:29-60)


# ---- the whole path on the oracle ------------------------------------------------------------------------------------------------
def test_textured_scene_renders_what_the_textures_say():
    d = Scene.from_source(scenes.textured_room(resolution=(48, 32), spp=16), REPO).desc()
    raw, cnt = O.render(d, 0, 16)
    img = O.convert_film(d, raw)[..., :3]
    assert np.isfinite(img).all() and cnt["samples"] == 48 * 32 * 16
    floor = img[26:, :, :]
    # the floor alternates between a red-ish and a blue-ish checker cell
    redish, blueish = (floor[..., 0] > 1.3 * floor[..., 2]).mean(), (floor[..., 2] > 1.3 * floor[..., 0]).mean()
    assert redish > 0.2 and blueish > 0.2
    # same scene with every texture scaled to zero is black except for the light itself
    dark = scenes.textured_room(resolution=(48, 32), spp=4).replace('address { "repeat" } filter { "bilinear" }', 'address { "repeat" } filter { "bilinear" } scale { 0.0 }')
    d2 = Scene.from_source(dark, REPO).desc()
    img2 = O.convert_film(d2, O.render(d2, 0, 4)[0])[..., :3]
    assert img2[28:, :, :].max() < img[28:, :, :].mean()


# ---- surface wrappers: opacity (stochastic alpha test in traversal) and normal maps (src/base/surface.h:160-275) -------------
def test_alpha_test_is_a_hash_of_the_hit_and_matches_the_mask_statistics():
    d = Scene.from_source(scenes.textured_room(wrappers=True), REPO).desc()
    screen = next(i for i in range(d.instance_count) if d.instances[i].handle[0] & 1023 & F.SHAPE_MAYBE_NON_OPAQUE
                  and d.surfaces[(d.instances[i].handle[1] >> 12) & 4095].opacity_tex)
    # rays from the camera side straight at random points of the screen quad
    rng = np.random.default_rng(5)
    n = 4000
    st = rng.uniform(0.02, 0.98, (n, 2)).astype(f32)
    p = np.array([0.2, 0.0, 1.0], f32) + st[:, :1] * np.array([1.4, 0.0, -0.4], f32) + st[:, 1:] * np.array([0.0, 1.3, 0.0], f32)
    o = p + np.array([0.3, 0.1, 1.0], f32)
    dirs = (p - o) / np.linalg.norm(p - o, axis=1, keepdims=True)
    rays = np.zeros((n, 8), f32)
    rays[:, :3], rays[:, 4:7], rays[:, 7] = o, dirs, 1e30
    hits, _ = O.trace(d, rays)
    on_screen = hits["inst"] == screen
    # expected hit fraction = mean of the (linear-encoded, bilinear, uv_scale 2) alpha texture over the quad
    lib = O.lib()
    tex_id = d.surfaces[(d.instances[screen].handle[1] >> 12) & 4095].opacity_tex - 1
    alpha = np.zeros(n, f32)
    out = np.zeros(4, f32)
    for k in range(n):
        lib.oracle_texture_evaluate(C.byref(d), tex_id, st[k].ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_float)))
        alpha[k] = min(max(out[0], 0.0), 1.0)
    assert 0.2 < alpha.mean() < 0.8
    assert abs(on_screen.mean() - alpha.mean()) < 0.03
    assert (alpha[on_screen] > 0).all() and not on_screen[alpha == 0].any()  # fully transparent texels never stop a ray
    # deterministic (hash of inst, prim, barycentrics), identical in the BVH and the brute-force traversal, and in any-hit mode
    again, _ = O.trace(d, rays)
    brute, _ = O.trace(d, rays, brute=True)
    assert np.array_equal(hits, again) and np.array_equal(hits["inst"], brute["inst"]) and np.array_equal(hits["prim"], brute["prim"])
    shadow = rays.copy()
    shadow[:, 7] = np.linalg.norm(p - o, axis=1) * 1.001  # just past the screen: occluded <=> the screen (or the cube) stopped the ray
    occluded, _ = O.trace(d, shadow, any_hit=True)
    closest_before = hits["inst"] != 0xFFFFFFFF
    t_hit_is_screen = on_screen
    assert (occluded["inst"][t_hit_is_screen] != 0xFFFFFFFF).all()
    assert closest_before.any()


def test_constant_opacity_and_neutral_normal_map():
    base = scenes.textured_room(resolution=(40, 28), spp=32)
    d0 = Scene.from_source(base, REPO).desc()
    img0 = O.convert_film(d0, O.render(d0, 0, 32)[0])[..., :3]
    # a constant normal map of (0.5, 0.5, 1) decodes to the unperturbed normal: same image up to rounding in the frame rebuild
    flat = base.replace("Surface floor_s : Matte {", "Surface floor_s : Matte {\n  normal_map : Constant { v { 0.5, 0.5, 1.0 } }")
    d1 = Scene.from_source(flat, REPO).desc()
    assert d1.surfaces[0].flags & F.SURFACE_HAS_NORMAL_MAP and d1.surfaces[0].normal_tex == 0
    img1 = O.convert_film(d1, O.render(d1, 0, 32)[0])[..., :3]
    assert img1.mean() == pytest.approx(img0.mean(), rel=2e-3)
    assert np.median(np.abs(img1 - img0)) < 1e-5
    # a tilted constant normal changes the floor's shading
    tilted = base.replace("Surface floor_s : Matte {", "Surface floor_s : Matte {\n  normal_map : Constant { v { 0.9, 0.5, 0.8 } }")
    d2 = Scene.from_source(tilted, REPO).desc()
    img2 = O.convert_film(d2, O.render(d2, 0, 32)[0])[..., :3]
    assert np.abs(img2[22:] - img0[22:]).mean() > 0.05 * img0[22:].mean()
    # opacity 1 is opaque (no flag); opacity 0 makes the cube vanish: the pixels behind it show the wall / floor
    opaque = Scene.from_source(base.replace("clearcoat : Constant { v { 0.5 } }", "clearcoat : Constant { v { 0.5 } } opacity : Constant { v { 1.0 } }"), REPO).desc()
    assert not any(opaque.instances[i].handle[0] & 1023 & F.SHAPE_MAYBE_NON_OPAQUE for i in range(opaque.instance_count))
    gone = Scene.from_source(base.replace("clearcoat : Constant { v { 0.5 } }", "clearcoat : Constant { v { 0.5 } } alpha : Constant { v { 0.0 } }"), REPO).desc()
    assert gone.surfaces[2].flags & F.SURFACE_MAYBE_NON_OPAQUE and gone.surfaces[2].opacity == 0.0
    rays = np.array([[0.0, 1.4, 4.2, 0.0, -0.16, -0.25, -0.95, 1e30]], f32)
    rays[0, 4:7] /= np.linalg.norm(rays[0, 4:7])
    cube_inst = 2
    assert O.trace(d0, rays)[0]["inst"][0] == cube_inst and O.trace(gone, rays)[0]["inst"][0] != cube_inst
