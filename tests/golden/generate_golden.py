"""Generates tests/golden/oracle_kat.json — known-answer vectors for every function on the hot path.

The reference ships no golden vectors for this path and cannot be run here, so these
vectors are produced by the CPU oracle itself (oracle/oracle.cpp) and committed: they pin the oracle against
regressions and give the CUDA kernels fixed targets.  Independent pins (Python re-derivations of the RNG,
alias tables, warps; brute-force intersection; closure integrals) live in tests/test_oracle.py.

    python tests/golden/generate_golden.py            # rewrites oracle_kat.json
"""
from __future__ import annotations

import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))

from luisarender_b200 import _ffi as F  # noqa: E402
from luisarender_b200 import scenes  # noqa: E402
from luisarender_b200.api import Scene  # noqa: E402
from oracle import binding as O  # noqa: E402

OUT = Path(__file__).resolve().parent / "oracle_kat.json"


def f32(a):
    return [float(np.float32(x)) for x in np.asarray(a, dtype=np.float32).ravel()]


def fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def golden_scenes():
    return {
        "cornell": scenes.cornell_box(resolution=(32, 32), spp=4),
        "cornell_disney": scenes.cornell_box(resolution=(32, 32), spp=4, surface="Disney"),
        "spheres": scenes.instanced_spheres(resolution=(32, 18), spp=2, big_subdivision=3, small_subdivision=2, small_count=10),
        # config C4 in miniature: homogeneous environment medium + MegaVPTNaive semantics, depth 8
        "spheres_medium": scenes.instanced_spheres(resolution=(32, 18), spp=4, big_subdivision=3, small_subdivision=2, small_count=10,
                                                   medium=True, depth=8),
        # image-textured Matte / Disney parameters on an InlineMesh with uvs and on OBJ / PLY mesh files
        "textured": scenes.textured_room(resolution=(48, 32), spp=4),
        # + the surface wrappers: normal map, alpha-tested cut-out (stochastic alpha inside traversal), constant opacity
        "textured_wrappers": scenes.textured_room(resolution=(48, 32), spp=4, wrappers=True),
        # rows a12 / f3: image-textured Spherical environment (importance map, MIS in the miss stage) next to an area light
        "environment": scenes.environment_scene(resolution=(48, 30), spp=4),
    }


def disney_surface(seed: int) -> F.Surface:
    rng = np.random.default_rng(seed)
    s = F.Surface()
    s.type = 1
    s.lobes = 1 | 2 | 8 | 16 | 32
    color = rng.uniform(0.05, 1.0, 3)
    lum = 0.212671 * color[0] + 0.715160 * color[1] + 0.072169 * color[2]
    vals = [*color, lum, rng.uniform(), 1.5, max(rng.uniform() ** 2, 1e-4), rng.uniform(), rng.uniform() * 0.5, rng.uniform(),
            rng.uniform(), rng.uniform(), rng.uniform(), 0.0, 0.0, 0.0]
    for i, v in enumerate(vals):
        s.p[i] = v
    return s


def matte_surface(sigma: float) -> F.Surface:
    s = F.Surface()
    s.type = 0
    for i, v in enumerate([0.7, 0.5, 0.3, sigma]):
        s.p[i] = v
    return s


def unit(v):
    v = np.asarray(v, dtype=np.float64)
    return (v / np.linalg.norm(v)).astype(np.float32)


def surface_cases():
    rng = np.random.default_rng(7)
    cases = []
    for k in range(12):
        surf = matte_surface(0.0 if k % 3 == 0 else 30.0) if k < 4 else disney_surface(k)
        ng = unit(rng.normal(size=3))
        ns = unit(ng + 0.2 * rng.normal(size=3))
        dpdu = unit(np.cross(ng, rng.normal(size=3)))
        wo = unit(ng * rng.uniform(0.1, 1.0) + 0.8 * rng.normal(size=3))
        if np.dot(wo, ng) < 0:
            wo = -wo
        wi = unit(ng * rng.uniform(0.1, 1.0) + 0.8 * rng.normal(size=3))
        u = rng.uniform(size=3).astype(np.float32)
        cases.append((surf, ng, ns, dpdu, wo, wi, u))
    return cases


def generate() -> dict:
    lib = O.lib()
    out: dict = {}
    out["xxhash32_uint4"] = [[a, b, c, d, int(lib.oracle_xxhash32_uint4(a, b, c, d))]
                             for a, b, c, d in [(0, 0, 0, 0), (1, 2, 3, 4), (511, 257, 19980810, 0), (1023, 1023, 19980810, 4095),
                                                (0xFFFFFFFF, 0x12345678, 0xDEADBEEF, 0x9E3779B9)]]
    seq = []
    st = C.c_uint32(12345)
    for _ in range(8):
        u = lib.oracle_lcg(C.byref(st))
        seq.append([int(st.value), float(u)])
    out["lcg_from_12345"] = seq
    cases = []
    for p, n in [((0.5, -2.0, 3.0), (0.0, 1.0, 0.0)), ((0.01, 0.02, -0.03), (0.6, 0.0, 0.8)), ((-7.25, 100.0, 1e-3), (-0.57735, 0.57735, 0.57735)),
                 ((1.0, 1.0, 1.0), (0.0, 0.0, -1.0))]:
        pa, na, oa = np.array(p, np.float32), np.array(n, np.float32), np.zeros(3, np.float32)
        lib.oracle_offset_ray_origin(fp(pa), fp(na), fp(oa))
        cases.append({"p": f32(pa), "n": f32(na), "out": f32(oa)})
    out["offset_ray_origin"] = cases
    warps = []
    for u in [(0.1, 0.7), (0.5, 0.5000001), (0.9, 0.2), (0.25, 0.25), (0.999, 0.001)]:
        ua, w, t = np.array(u, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32)
        lib.oracle_sample_cosine_hemisphere(fp(ua), fp(w))
        lib.oracle_sample_uniform_triangle(fp(ua), fp(t))
        warps.append({"u": f32(ua), "cosine_hemisphere": f32(w), "uniform_triangle": f32(t)})
    out["warps"] = warps
    surf_out = []
    for surf, ng, ns, dpdu, wo, wi, u in surface_cases():
        f, pdf = np.zeros(3, np.float32), C.c_float()
        lib.oracle_surface_evaluate(C.byref(surf), fp(ng), fp(ns), fp(dpdu), fp(wo), fp(wi), fp(f), C.byref(pdf))
        swi, sf, spdf = np.zeros(3, np.float32), np.zeros(3, np.float32), C.c_float()
        lib.oracle_surface_sample(C.byref(surf), fp(ng), fp(ns), fp(dpdu), fp(wo), float(u[0]), fp(u[1:].copy()), fp(swi), fp(sf), C.byref(spdf))
        surf_out.append({"type": int(surf.type), "lobes": int(surf.lobes), "p": f32(list(surf.p)), "ng": f32(ng), "ns": f32(ns),
                         "dpdu": f32(dpdu), "wo": f32(wo), "wi": f32(wi), "u": f32(u), "eval_f": f32(f), "eval_pdf": float(pdf.value),
                         "sample_wi": f32(swi), "sample_f": f32(sf), "sample_pdf": float(spdf.value)})
    out["surfaces"] = surf_out
    per_scene = {}
    for name, src in golden_scenes().items():
        sc = Scene.from_source(src, REPO)
        d = sc.desc()
        w, h = d.camera.resolution[0], d.camera.resolution[1]
        entry: dict = {}
        rays = []
        for px, py, s in [(0, 0, 0), (w // 2, h // 2, 1), (w - 1, h - 1, 3), (5, 11 % h, 2)]:
            ray, weight, state = O.generate_ray(d, px, py, s)
            rays.append({"pixel": [px, py], "sample": s, "ray": f32(ray), "weight": f32(weight), "state": int(state)})
        entry["camera_rays"] = rays
        flt = []
        for u in [(0.3, 0.8), (0.0, 0.999), (0.5, 0.5)]:
            ua, off, wgt = np.array(u, np.float32), np.zeros(2, np.float32), C.c_float()
            lib.oracle_sample_filter(C.byref(d), fp(ua), fp(off), C.byref(wgt))
            flt.append({"u": f32(ua), "offset": f32(off), "weight": float(wgt.value)})
        entry["filter"] = flt
        # hit reconstruction + light sampling at the primary hits of a few pixels
        its = []
        for px, py in [(w // 2, h // 2), (w // 4, (3 * h) // 4), (w - 3, 2)]:
            ray, _, _ = O.generate_ray(d, px, py, 0)
            hits, _ = O.trace(d, ray[None, :])
            rec = {"pixel": [px, py], "hit": [int(hits["inst"][0]), int(hits["prim"][0])], "bary": f32(hits["bary"][0])}
            if hits["inst"][0] != 0xFFFFFFFF:
                it = np.zeros(19, np.float32)
                lib.oracle_interaction(C.byref(d), ray.ctypes.data, hits.ctypes.data, fp(it))
                ls = np.zeros(12, np.float32)
                ul = np.array([0.35, 0.65], np.float32)
                lib.oracle_sample_light(C.byref(d), ray.ctypes.data, hits.ctypes.data, C.c_float(0.6), fp(ul), fp(ls))
                rec["interaction"] = f32(it)
                rec["light_sample"] = f32(ls)
            its.append(rec)
        entry["interactions"] = its
        entry["li"] = [{"pixel": [px, py], "sample": s, "rgb": f32(O.li(d, px, py, s))}
                       for px, py, s in [(w // 2, h // 2, 0), (w // 2, h // 2, 1), (3, h - 2, 0), (w - 2, 1, 2), (w // 3, h // 3, 3)]]
        raw, cnt = O.render(d, 0, d.camera.spp, threads=2)
        entry["film_sum_rgb"] = f32(raw[..., :3].reshape(-1, 3).astype(np.float64).sum(axis=0))
        entry["film_row0"] = f32(raw[0, :, :3])
        entry["counters"] = {k: cnt[k] for k in ("samples", "closest_rays", "shadow_rays", "path_vertices")}
        per_scene[name] = entry
    out["scenes"] = per_scene
    return out


if __name__ == "__main__":
    data = generate()
    OUT.write_text(json.dumps(data, indent=1))
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes)")
