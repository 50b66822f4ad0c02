"""Host front-end: scene-description grammar (text + JSON), node/property semantics, flattening rules and
BVH validity.  Expected behaviours cite the reference file:line they mirror."""
from __future__ import annotations

import ctypes as C
import json

import numpy as np
import pytest

from luisarender_b200 import _ffi as F
from luisarender_b200 import scenes
from luisarender_b200.api import Scene

MINIMAL = """
Surface white : Matte { Kd : Constant { v { 0.5, 0.5, 0.5 } } }
Light lamp : Diffuse { emission : Constant { v { 1, 2, 3 } } scale { 2 } }
Shape quad : InlineMesh {
  positions { -1, 0, -1,  1, 0, -1,  1, 0, 1,  -1, 0, 1 }
  indices { 0, 1, 2, 0, 2, 3 }
  surface { @white }
}
Shape lamp_quad : InlineMesh {
  positions { -1, 2, -1,  -1, 2, 1,  1, 2, 1,  1, 2, -1 }
  indices { 0, 1, 2, 0, 2, 3 }
  light { @lamp }
}
Camera cam : Pinhole {
  position { 0, 1, 4 }  fov { 40 }  spp { #SPP }
  film : Color { resolution { 32, 16 } exposure { 1 } }
}
render {
  integrator : WavePath { depth { 5 } rr_depth { 2 } }
  cameras { @cam }
  shapes { @quad, @lamp_quad }
}
"""


def _arr(ptr, n, dtype, shape=None):
    a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dtype).itemsize,)).view(dtype)
    return a.reshape(shape) if shape else a


def test_text_scene_with_macros_and_defaults():
    sc = Scene.from_source("define SPP 7 // local macro\n" + MINIMAL)
    d = sc.desc()
    assert d.camera.spp == 7
    # CLI macros win over local defines (src/sdl/scene_parser.cpp:420-451)
    d2 = Scene.from_source("define SPP 7\n" + MINIMAL, macros={"SPP": "11"}).desc()
    assert d2.camera.spp == 11
    assert list(d.camera.resolution) == [32, 16]
    assert d.integrator.max_depth == 5 and d.integrator.rr_depth == 2
    assert d.integrator.rr_threshold == pytest.approx(0.95)  # wave_path.cpp:45
    assert d.integrator.sampler_seed == 19980810  # sampler.cpp:11
    assert list(d.film.scale) == [2.0, 2.0, 2.0] and d.film.clamp == 256.0  # color.cpp:38-41
    assert d.camera.filter_radius == 0.5  # Box default, camera.cpp:19-20 / filter.cpp:13
    assert d.camera.tan_half_fov == pytest.approx(np.tan(np.radians(40.0) / 2), rel=1e-6)
    assert d.light_count == 1 and d.surface_count == 1 and d.instance_count == 2
    light = d.lights[0]
    assert list(light.emission) == [1.0, 2.0, 3.0] and light.scale == 2.0 and light.two_sided == 0
    assert list(d.surfaces[0].p[:4]) == [0.5, 0.5, 0.5, 0.0]


def test_undefined_macro_and_unknown_plugin_are_hard_errors():
    with pytest.raises(RuntimeError, match="Undefined macro"):
        Scene.from_source(MINIMAL)
    bad = MINIMAL.replace("#SPP", "1").replace("Matte {", "Hair {")  # a surface implementation this library does not have
    with pytest.raises(RuntimeError, match="luisa-render-surface-hair"):
        Scene.from_source(bad)
    with pytest.raises(RuntimeError, match="Redefinition of property"):
        Scene.from_source(MINIMAL.replace("#SPP", "1").replace("fov { 40 }", "fov { 40 } fov { 41 }"))


def test_import_pulls_in_text_and_json_files_relative_to_the_importing_file(tmp_path):
    """scene_parser.cpp:76-82: `import "file"` parses another scene file (by its extension: text or JSON) into the same description,
    paths relative to the directory of the file that imports; the result equals the scene written in one piece."""
    whole = MINIMAL.replace("#SPP", "2")
    head, rest = whole.split("Shape lamp_quad", 1)
    assert "Surface white" in head and "Light lamp" in head
    (tmp_path / "parts").mkdir()
    (tmp_path / "parts" / "materials.luisa").write_text(head + 'import "lamp.json"\n')
    lamp_body, tail = rest.split("Camera cam", 1)
    (tmp_path / "parts" / "lamp.json").write_text(json.dumps({"lamp_quad": {"type": "Shape", "impl": "InlineMesh", "prop": {
        "positions": [-1, 2, -1, -1, 2, 1, 1, 2, 1, 1, 2, -1], "indices": [0, 1, 2, 0, 2, 3], "light": "@lamp"}}}))
    (tmp_path / "main.luisa").write_text('import "parts/materials.luisa"\nCamera cam' + tail)
    a, b = Scene.from_file(tmp_path / "main.luisa").desc(), Scene.from_source(whole).desc()
    assert a.triangle_count == b.triangle_count and a.instance_count == b.instance_count and a.light_count == b.light_count == 1
    n = a.triangle_count * 3
    assert np.array_equal(_arr(a.triangles, n, np.uint32), _arr(b.triangles, n, np.uint32))
    with pytest.raises(RuntimeError):
        (tmp_path / "broken.luisa").write_text('import "parts/missing.luisa"\nCamera cam' + tail)
        Scene.from_file(tmp_path / "broken.luisa")


def test_film_option_that_would_change_the_film_is_refused_not_ignored():
    """color.cpp:124-129: warn_nan { true } makes the reference overwrite a pixel with (inf, 0, 0, 1) on a NaN / infinite sample; the
    accumulate kernel has no such pass, so the option is a load error (its default, false, loads)."""
    from luisarender_b200 import scenes as S
    src = S.cornell_box(resolution=(16, 16), spp=1)
    assert "film : Color {" in src
    Scene.from_source(src.replace("film : Color {", "film : Color { warn_nan { false }"))
    with pytest.raises(RuntimeError, match="warn_nan"):
        Scene.from_source(src.replace("film : Color {", "film : Color { warn_nan { true }"))


def test_base_node_inheritance_and_tag_aliases():
    src = MINIMAL.replace("#SPP", "1") + """
    """
    src = src.replace("Surface white : Matte { Kd : Constant { v { 0.5, 0.5, 0.5 } } }",
                      "Surf base_mat : Matte { Kd : Constant { v { 0.25 } } sigma : Constant { v { 0.5 } } }\n"
                      "surface white : Matte (@base_mat) { Kd : Constant { v { 0.5, 0.5, 0.5 } } }")
    d = Scene.from_source(src).desc()
    # Kd overridden, sigma inherited from the base node (scene_node_desc.h:280-293); sigma = saturate(v)*90 (matte.cpp:126)
    assert list(d.surfaces[0].p[:4]) == [0.5, 0.5, 0.5, 45.0]


def test_json_scene_matches_text_scene():
    text = MINIMAL.replace("#SPP", "3")
    js = {
        "white": {"type": "Surface", "impl": "Matte", "prop": {"Kd": {"impl": "Constant", "prop": {"v": [0.5, 0.5, 0.5]}}}},
        "lamp": {"type": "Light", "impl": "Diffuse", "prop": {"emission": {"impl": "Constant", "prop": {"v": [1, 2, 3]}}, "scale": 2}},
        "quad": {"type": "Shape", "impl": "InlineMesh", "prop": {
            "positions": [-1, 0, -1, 1, 0, -1, 1, 0, 1, -1, 0, 1], "indices": [0, 1, 2, 0, 2, 3], "surface": "@white"}},
        "lamp_quad": {"type": "Shape", "impl": "InlineMesh", "prop": {
            "positions": [-1, 2, -1, -1, 2, 1, 1, 2, 1, 1, 2, -1], "indices": [0, 1, 2, 0, 2, 3], "light": "@lamp"}},
        "cam": {"type": "Camera", "impl": "Pinhole", "prop": {
            "position": [0, 1, 4], "fov": 40, "spp": 3,
            "film": {"impl": "Color", "prop": {"resolution": [32, 16], "exposure": 1}}}},
        "render": {"integrator": {"impl": "WavePath", "prop": {"depth": 5, "rr_depth": 2}},
                   "cameras": ["@cam"], "shapes": ["@quad", "@lamp_quad"]},
    }
    a = Scene.from_source(text).desc()
    b = Scene.from_source("// comments are allowed (scene_parser_json.cpp:28)\n" + json.dumps(js), json=True).desc()
    for name in ("camera", "film", "integrator"):
        assert bytes(getattr(a, name)) == bytes(getattr(b, name)), name
    assert a.instance_count == b.instance_count and a.triangle_count == b.triangle_count
    assert bytes(_arr(a.vertices, a.vertex_count, np.uint8, None)[: a.vertex_count * 32]) == \
        bytes(_arr(b.vertices, b.vertex_count, np.uint8, None)[: b.vertex_count * 32])


def test_cornell_flattening(cornell_small):
    d = cornell_small.desc()
    info = cornell_small.info()
    assert info["unique_triangles"] == 32 and info["instances"] == 8 and info["lights"] == 1 and info["surfaces"] == 3
    from oracle import binding as O
    lib = O.lib()
    # instance order = order of render.shapes (geometry.cpp:106); the light is the last shape, has no surface
    u = (C.c_uint32 * 6)()
    f = (C.c_float * 2)()
    lib.oracle_decode_handle(d.instances[7].handle, u, f)
    buffer_base, flags, surface_tag, light_tag, medium_tag, tri_count = list(u)
    assert buffer_base == 7 * 4 and flags == 8 and tri_count == 2 and light_tag == 0
    assert f[1] == 1.0  # default intersection offset factor (shape.cpp:90-92)
    lib.oracle_decode_handle(d.instances[3].handle, u, f)  # right wall: second registered surface (green)
    assert u[1] == 4 and u[2] == 1
    assert [d.light_handles[0].instance_id, d.light_handles[0].light_tag] == [7, 0]
    # camera "position/front/up" compatibility path builds a View transform (camera.cpp:30-50)
    c2w = np.array(list(d.camera.camera_to_world)).reshape(3, 4)
    assert np.allclose(c2w[:, 3], [-0.01, 0.995, 5.0]) and np.allclose(c2w[:, :3], np.eye(3))


def test_uniform_light_sampler_counts_light_nodes_not_instances():
    """two shapes sharing ONE light node => n = 1 and only the first instance is sampled."""
    src = MINIMAL.replace("#SPP", "1").replace("shapes { @quad, @lamp_quad }", "shapes { @quad, @lamp_quad, @lamp_quad2 }") + """
Shape lamp_quad2 : InlineMesh {
  positions { -1, 3, -1,  -1, 3, 1,  1, 3, 1,  1, 3, -1 }
  indices { 0, 1, 2, 0, 2, 3 }
  light { @lamp }
}
"""
    d = Scene.from_source(src).desc()
    assert d.instance_count == 3 and d.light_count == 1
    assert d.light_handles[0].instance_id == 1


def test_instancing_overrides_and_transform_chain():
    src = """
Surface a : Matte { Kd : Constant { v { 0.1 } } }
Surface b : Matte { Kd : Constant { v { 0.9 } } }
Light lamp : Diffuse { emission : Constant { v { 1 } } }
Shape tri : InlineMesh { positions { 0,0,0, 1,0,0, 0,1,0 } indices { 0,1,2 } surface { @a } }
Shape inner : Instance { shape { @tri } transform : SRT { translate { 1, 0, 0 } } }
Shape outer : Group { shapes { @inner, @tri } surface { @b } transform : SRT { scale { 2 } } }
Shape lamp_tri : InlineMesh { positions { 0,5,0, 0,5,1, 1,5,0 } indices { 0,1,2 } light { @lamp } visible { false } }
Camera cam : Pinhole { film : Color { resolution { 8 } } }
render { integrator : MegaPath {} cameras { @cam } shapes { @outer, @tri, @lamp_tri } }
"""
    sc = Scene.from_source(src)
    d = sc.desc()
    assert d.instance_count == 4 and d.mesh_count == 2  # identical mesh bytes share one BLAS (geometry.cpp:53-57)
    from oracle import binding as O
    u = (C.c_uint32 * 6)()
    f = (C.c_float * 2)()
    tags = []
    for i in range(3):
        O.lib().oracle_decode_handle(d.instances[i].handle, u, f)
        tags.append(u[2])
    # a group's non-null surface overrides its children's (geometry.cpp:36-38): b is registered first -> tag 0
    assert tags == [0, 0, 1]
    m0 = np.array(list(d.instances[0].object_to_world)).reshape(3, 4)
    assert np.allclose(m0, [[2, 0, 0, 2], [0, 2, 0, 0], [0, 0, 2, 0]])  # scale(2) * translate(1,0,0)
    m1 = np.array(list(d.instances[1].object_to_world)).reshape(3, 4)
    assert np.allclose(m1, [[2, 0, 0, 0], [0, 2, 0, 0], [0, 0, 2, 0]])
    w2o = np.array(list(d.instances[0].world_to_object)).reshape(3, 4)
    assert np.allclose(w2o @ np.vstack([m0, [0, 0, 0, 1]]), np.hstack([np.eye(3), np.zeros((3, 1))]), atol=1e-6)
    assert d.instances[3].visible == 0 and d.resolution if False else True
    assert list(d.camera.resolution) == [8, 8]  # single value -> square (color.cpp:29-32)


def test_sphere_geometry_levels():
    for level, tris in ((0, 20), (1, 80), (3, 1280)):
        src = MINIMAL.replace("#SPP", "1").replace("shapes { @quad, @lamp_quad }", "shapes { @quad, @lamp_quad, @ball }") + \
            f"Shape ball : Sphere {{ subdivision {{ {level} }} surface {{ @white }} }}\n"
        d = Scene.from_source(src).desc()
        mesh = d.meshes[2]
        assert mesh.triangle_count == tris  # 20 * 4^level (sphere.cpp:13,88-101)
        v = _arr(C.cast(d.vertices, C.c_void_p).value + 32 * mesh.vertex_offset, mesh.vertex_count * 8, np.float32, (-1, 8))
        assert mesh.vertex_count == 10 * 4 ** level + 2
        assert np.allclose(np.linalg.norm(v[:, :3], axis=1), 1.0, atol=1e-6)
        assert np.array_equal(v[:, :3], v[:, 3:6])  # n = p
        assert (v[:, 6:] >= 0).all() and (v[:, 6:] <= 1).all()  # fract(x) = x - floor(x) may round to 1.0
        t = _arr(C.cast(d.triangles, C.c_void_p).value + 12 * mesh.triangle_offset, tris * 3, np.uint32, (-1, 3))
        p = v[:, :3]
        n = np.cross(p[t[:, 1]] - p[t[:, 0]], p[t[:, 2]] - p[t[:, 0]])
        assert (np.einsum("ij,ij->i", n, p[t[:, 0]]) > 0).all()  # outward-facing winding
        # closed manifold: every edge shared by exactly two triangles
        e = np.sort(np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]]), axis=1)
        _, counts = np.unique(e, axis=0, return_counts=True)
        assert (counts == 2).all()


def test_alias_tables_and_pdf(spheres_small):
    d = spheres_small.desc()
    for mi in range(d.mesh_count):
        m = d.meshes[mi]
        pdf = _arr(C.cast(d.pdf, C.c_void_p).value + 4 * m.triangle_offset, m.triangle_count, np.float32)
        al = _arr(C.cast(d.alias, C.c_void_p).value + 8 * m.triangle_offset, m.triangle_count * 2, np.uint32, (-1, 2))
        prob = al[:, 0].view(np.float32)
        assert pdf.sum() == pytest.approx(1.0, abs=1e-4)
        assert (prob >= 0).all() and (prob <= 1.0 + 1e-6).all() and (al[:, 1] < m.triangle_count).all()
        # the alias table reproduces the pdf: P(i) = (prob_i + sum_{j: alias_j = i} (1 - prob_j)) / n
        recon = prob.astype(np.float64).copy()
        np.add.at(recon, al[:, 1], 1.0 - prob.astype(np.float64))
        assert np.allclose(recon / m.triangle_count, pdf, atol=2e-6)


def test_bvh_is_valid(spheres_small):
    d = spheres_small.desc()
    nodes = _arr(d.bvh_nodes, d.bvh_node_count * 16, np.uint32, (-1, 16))
    boxes = nodes[:, :12].view(np.float32)
    tv = _arr(d.tri_verts, d.tri_slot_count * 12, np.float32, (-1, 12))
    LEAF = 0x80000000
    covered = np.zeros(d.tri_slot_count, dtype=np.int32)

    def child_bounds(ref, in_blas, xform=None):
        if ref == 0xFFFFFFFF:
            return None
        if ref & LEAF:
            if in_blas:
                first, count = ref & 0x0FFFFFFF, ((ref >> 28) & 7) + 1
                assert count <= 4
                covered[first:first + count] += 1
                pts = tv[first:first + count][:, [0, 1, 2, 4, 5, 6, 8, 9, 10]].reshape(-1, 3)
                return pts.min(0), pts.max(0)
            return None
        lo = np.minimum(boxes[ref, 0:3], boxes[ref, 6:9])
        hi = np.maximum(boxes[ref, 3:6], boxes[ref, 9:12])
        return lo, hi

    # every BLAS: child boxes contain their subtree; every triangle slot referenced exactly once
    for mi in range(d.mesh_count):
        m = d.meshes[mi]
        stack = [m.bvh_root]
        while stack:
            n = stack.pop()
            for c, (lo_s, hi_s) in enumerate(((slice(0, 3), slice(3, 6)), (slice(6, 9), slice(9, 12)))):
                ref = int(nodes[n, 12 + c])
                b = child_bounds(ref, True)
                if b is not None:
                    assert (boxes[n, lo_s] <= b[0] + 1e-6).all() and (boxes[n, hi_s] >= b[1] - 1e-6).all()
                    if ref & LEAF:
                        # stored boxes are PADDED (bvh.cpp: the rounded slab test must not cull what the triangle test accepts),
                        # by less than the spawned-ray offset: 5e-7 x the mesh's extent
                        extent = np.abs(tv[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]]).max() * 2.0
                        assert (boxes[n, lo_s] < b[0]).all() and (boxes[n, hi_s] > b[1]).all()
                        assert (b[0] - boxes[n, lo_s] <= 1e-6 * extent).all() and (boxes[n, hi_s] - b[1] <= 1e-6 * extent).all()
                if ref != 0xFFFFFFFF and not (ref & LEAF):
                    assert int(nodes[ref, 14]) == n  # parent link
                    stack.append(ref)
    assert (covered == 1).all()
    # TLAS leaves reference every visible instance exactly once
    seen = []
    stack = [d.tlas_root]
    while stack:
        n = stack.pop()
        for c in range(2):
            ref = int(nodes[n, 12 + c])
            if ref == 0xFFFFFFFF:
                continue
            if ref & LEAF:
                seen.append(ref & 0x7FFFFFFF)
            else:
                stack.append(ref)
    assert sorted(seen) == list(range(d.instance_count))


def test_filter_tables_box_and_gaussian():
    d = Scene.from_source(MINIMAL.replace("#SPP", "1")).desc()
    lut = np.array(list(d.camera.filter_lut))
    assert np.allclose(lut, 1.0 / 63.0)  # Box: lut normalised by the 63 mid-point sums (filter.cpp:30-41)
    assert np.allclose(list(d.camera.filter_pdf)[:63], 1.0 / 63.0) and np.allclose(list(d.camera.filter_alias_probs)[:63], 1.0)
    g = Scene.from_source(MINIMAL.replace("#SPP", "1").replace("fov { 40 }", "fov { 40 } filter : Gaussian { radius { 1.5 } }")).desc()
    lut = np.array(list(g.camera.filter_lut))
    assert lut[0] == pytest.approx(0.0, abs=1e-7) and lut[31] == pytest.approx(lut[32], rel=1e-3) and lut.argmax() in (31, 32)
    assert sum(list(g.camera.filter_pdf)[:63]) == pytest.approx(1.0, abs=1e-5)


def test_scene_generators_match_baseline_configs():
    c1 = Scene.from_source(scenes.cornell_box()).desc()
    assert list(c1.camera.resolution) == [512, 512] and c1.camera.spp == 16 and c1.triangle_count == 32
    info = Scene.from_source(scenes.instanced_spheres(big_subdivision=5)).info()  # full size is built by the GPU tests / bench
    assert info["instances"] == 67 and info["surfaces"] == 9 and info["lights"] == 2
    assert info["instanced_triangles"] == 4 * 20 * 4 ** 5 + 60 * 1280 + 2 + 4


def test_loop_subdiv_shape():
    """src/shapes/loop_subdiv.cpp: `mesh` / `shape` / `base` name the base shape, which must be a mesh; the level is clamped
    to 10, level 0 passes the base mesh (and its vertex properties) through; a subdivided mesh has normals and no uvs."""
    from luisarender_b200 import scenes

    src = scenes.subdivision_scene(resolution=(8, 6), spp=1)
    d = Scene.from_source(src).desc()
    counts = sorted(int(d.meshes[i].triangle_count) for i in range(d.mesh_count))
    assert counts == sorted([2, 2, 12 * 4 ** 3, 4 * 4 ** 2, 2 * 4 ** 2, 5 * 4, 1])
    with pytest.raises(RuntimeError, match="LoopSubdiv only supports mesh shapes"):
        Scene.from_source(src.replace("shape : InlineMesh {\n    positions { 0.0, 1.0, 0.0,  -0.9, -0.4, 0.55,  0.9, -0.4, 0.55,  0.0, -0.4, -1.0 }\n"
                                      "    indices { 0, 1, 2,  0, 2, 3,  0, 3, 1,  1, 3, 2 }\n  }",
                                      "shape : Group { shapes { @floor } }"))
    with pytest.raises(RuntimeError, match="base"):
        Scene.from_source(src.replace("  level { 0 }", "  level { 0 }").replace("Shape passthrough : LoopSubdiv {\n  mesh : InlineMesh", "Shape passthrough : LoopSubdiv {\n  cage : InlineMesh"))


def test_swizzle_texture():
    """src/textures/swizzle.cpp: channel letters or an index list, at most four, each < 4; constants fold (nesting included);
    an image base stays an image texture with permuted texels and the swizzle's channel count."""
    from pathlib import Path

    REPO = Path(__file__).resolve().parent.parent
    src = scenes.swizzle_scene(resolution=(8, 6), spp=1)
    d = Scene.from_source(src, REPO).desc()
    channels = sorted(int(d.textures[i].channels) for i in range(d.texture_count))
    assert 1 in channels and 3 in channels  # sigma <- one channel of the RGBA16 ramp; "bgr" / "gbr" stay three-channel
    with pytest.raises(RuntimeError, match="Invalid swizzle channel 'q'"):
        Scene.from_source(src.replace('swizzle { "bgr" }', 'swizzle { "bqr" }'), REPO)
    with pytest.raises(RuntimeError, match="out of range"):
        Scene.from_source(src.replace("swizzle { 2, 0, 1 }", "swizzle { 2, 0, 7 }"), REPO)


def test_checkerboard_texture():
    """src/textures/checkerboard.cpp with constant squares: baked into a 2x2 point-sampled repeating image at uv * scale / 2;
    image squares and the ambiguous absent 'on' (1 as a scalar, black as a colour in the reference) are refused."""
    from pathlib import Path

    REPO = Path(__file__).resolve().parent.parent
    src = scenes.checkerboard_scene(resolution=(8, 6), spp=1)
    d = Scene.from_source(src, REPO).desc()
    baked = [d.textures[i] for i in range(d.texture_count) if d.textures[i].width == 2 and d.textures[i].height == 2]
    assert len(baked) == 3
    assert sorted((round(t.uv_scale[0], 3), round(t.uv_scale[1], 3), int(t.channels)) for t in baked) == [(1.25, 1.25, 3), (1.5, 1.5, 3), (2.5, 3.5, 3)]
    with pytest.raises(RuntimeError, match="'on' must be given"):
        Scene.from_source(src.replace("on : Constant { v { 0.9, 0.85, 0.3 } } scale { 2.5 }", "scale { 2.5 }"), REPO)
    with pytest.raises(RuntimeError, match="only constant"):
        Scene.from_source(src.replace("on : Constant { v { 35.0 } }", 'on : Image { file { "tests/golden/assets/rough_gray8.png" } }'), REPO)


def test_checkerboard_extends_each_square_by_its_own_channel_count():
    """checkerboard.cpp:76-105 decodes `on` and `off` separately (extend_color_to_rgb per child): a grey square next to an RGB
    square stays grey, and a one-channel square used as a colour is (a, a, a) - not (a, 0, 0)."""
    from pathlib import Path

    REPO = Path(__file__).resolve().parent.parent
    src = scenes.checkerboard_scene(resolution=(8, 6), spp=1)
    old = "Kd : Checkerboard { on : Constant { v { 0.8, 0.2, 0.2 } } off : Constant { v { 0.1, 0.1, 0.6 } } scale { 5.0, 7.0 } }"
    assert src.count(old) == 1
    d = Scene.from_source(src.replace(old, "Kd : Checkerboard { on : Constant { v { 0.8, 0.2, 0.2 } } off : Constant { v { 0.3 } } scale { 5.0, 7.0 } }"),
                          REPO).desc()
    t = next(d.textures[i] for i in range(d.texture_count) if d.textures[i].width == 2 and round(d.textures[i].uv_scale[1], 3) == 3.5)
    texels = np.array([d.texels[4 * t.texel_offset + k] for k in range(16)], dtype=np.float32).reshape(4, 4)
    assert int(t.channels) == 3
    np.testing.assert_array_equal(texels[0, :3], np.float32([0.8, 0.2, 0.2]))
    np.testing.assert_array_equal(texels[1, :3], np.float32([0.3, 0.3, 0.3]))
    with pytest.raises(RuntimeError, match="Cannot convert property 'swizzle'"):
        Scene.from_source(src.replace(old, "Kd : Swizzle { base : Constant { v { 0.8, 0.2, 0.2 } } swizzle { 0, 1.7, 2 } }"), REPO)


def test_swizzle_of_checkerboard_composes():
    """Swizzle { Checkerboard { on : Swizzle { Constant } } }: the baked 2x2 texels carry the composed permutation (this
    composition was also rendered with the reference once: bit-identical to the oracle)."""
    from pathlib import Path

    REPO = Path(__file__).resolve().parent.parent
    src = scenes.checkerboard_scene(resolution=(8, 6), spp=1)
    old = "Kd : Checkerboard { on : Constant { v { 0.8, 0.2, 0.2 } } off : Constant { v { 0.1, 0.1, 0.6 } } scale { 5.0, 7.0 } }"
    assert src.count(old) == 1
    src = src.replace(old, 'Kd : Swizzle { base : Checkerboard { on : Swizzle { base : Constant { v { 0.8, 0.2, 0.2, 0.5 } } swizzle { "xyz" } } '
                           'off : Constant { v { 0.1, 0.1, 0.6 } } scale { 5.0, 7.0 } } swizzle { "brg" } }')
    d = Scene.from_source(src, REPO).desc()
    t = next(d.textures[i] for i in range(d.texture_count) if d.textures[i].width == 2 and round(d.textures[i].uv_scale[1], 3) == 3.5)
    texels = np.array([d.texels[4 * t.texel_offset + k] for k in range(16)], dtype=np.float32).reshape(4, 4)
    assert int(t.channels) == 3
    np.testing.assert_array_equal(texels[0, :3], np.float32([0.2, 0.8, 0.2]))  # on, "brg"
    np.testing.assert_array_equal(texels[1, :3], np.float32([0.6, 0.1, 0.1]))  # off, "brg"
    np.testing.assert_array_equal(texels[3, :3], texels[0, :3])


def test_disney_closure_classes_and_their_lobe_unions():
    """src/surfaces/disney.cpp:61-75,925-930,966-995: a node is "disney_thin" when `thin` is set AND one of the two transmissions is
    not black, "disney_trans" when not thin with a non-black specular_trans, else "disney"; each class ORs the lobes of its own
    nodes only.  The scene is the fixture tests/golden/ref_renders.npz: spheres_disney_thin, which the reference itself rendered."""
    import sys
    from pathlib import Path

    REPO = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(REPO / "tools"))
    import gen_ref_renders as G

    source = G.cases()["spheres_disney_thin"]
    d = Scene.from_source(source, REPO).desc()
    recs = [d.surfaces[i] for i in range(d.surface_count) if d.surfaces[i].type == 1]  # LRK_SURFACE_DISNEY
    thin = [s for s in recs if s.flags & 64]
    trans = [s for s in recs if s.flags & 16]
    opaque = [s for s in recs if not s.flags & (16 | 64)]
    assert thin and trans and opaque and not any(s.flags & 16 for s in thin)
    assert {s.lobes for s in thin} == {1 | 2 | 4 | 8 | 16 | 32 | 64 | 128}  # flatness (4), spec_trans and diff_trans come from thin nodes only
    assert {s.lobes for s in trans} == {1 | 2 | 8 | 16 | 32 | 128}
    assert {s.lobes for s in opaque} == {1 | 2 | 8 | 16 | 32}
    assert sorted({round(s.p[15], 3) for s in thin}) == [0.5, 0.7] and all(s.p[15] == 0.0 for s in trans + opaque)
    # `thin` without any transmission: an ordinary opaque record (diffuse_trans is built but black)
    plain = source.replace("diffuse_trans : Constant { v { 0.7 } }", "diffuse_trans : Constant { v { 0.0 } }")
    d2 = Scene.from_source(plain, REPO).desc()
    assert sum(1 for i in range(d2.surface_count) if d2.surfaces[i].flags & 64) == len(thin) // 2
    assert {d2.surfaces[i].lobes for i in range(d2.surface_count) if d2.surfaces[i].flags & 64} == {255}  # the remaining thin nodes still carry both transmissions


def test_light_with_image_emission_flattens_to_a_texture_slot():
    """src/lights/diffuse.cpp:23-26,74: `emission` is any texture; an image is evaluated per point (lrk_light.emission_tex), a
    constant is folded into lrk_light.emission."""
    from pathlib import Path

    repo = Path(__file__).resolve().parent.parent
    d = Scene.from_source(scenes.textured_room(resolution=(16, 12), spp=1, mesh_files=False, textured_light=True), repo).desc()
    assert d.light_count == 1 and 1 <= d.lights[0].emission_tex <= d.texture_count and list(d.lights[0].emission) == [0.0, 0.0, 0.0]
    tex = d.textures[d.lights[0].emission_tex - 1]
    assert tex.scale == 20.0 and list(tex.uv_scale) == [2.0, 2.0] and tex.address == 1  # repeat
    d = Scene.from_source(scenes.textured_room(resolution=(16, 12), spp=1, mesh_files=False), repo).desc()
    assert d.lights[0].emission_tex == 0 and list(d.lights[0].emission) == [14.0, 13.0, 11.0]
