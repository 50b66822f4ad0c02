#!/usr/bin/env python3
"""Generates tests/golden/ref_pins.npz: seeded inputs and the outputs of the REFERENCE's own functions for them.

The reference functions (LuisaRender src/util/{rng,sampling,scattering,frame}.cpp, written in the LuisaCompute DSL) are
compiled from /root/reference by oracle/ref/Makefile and executed by the AST interpreter oracle/ref/interp.cpp; this script
drives oracle/_ref/librefpins.so through ctypes.  It only runs where /root/reference exists (this container); the resulting
fixture is committed so that the pins travel to machines without the reference (tests/test_ref_pins.py).

    make -C oracle/ref && python tools/gen_ref_pins.py
"""
from __future__ import annotations

import ctypes
import sys
import zlib
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
LIB = REPO / "oracle" / "_ref" / "librefpins.so"
OUT = REPO / "tests" / "golden" / "ref_pins.npz"
N = 192  # cases per pin

# argument generators -----------------------------------------------------------------------------------------------


def _u(r, n):  # uniform [0, 1)
    return r.random((n, 1), dtype=np.float32)


def _u2(r, n):
    return r.random((n, 2), dtype=np.float32)


def _bits(r, n, k=1):
    return r.integers(0, 2**32, size=(n, k), dtype=np.uint32).view(np.float32)


def _dir(r, n):  # unit vectors, both hemispheres, a few grazing / axis-aligned ones
    v = r.normal(size=(n, 3)).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True).astype(np.float32)
    v[0] = (0, 0, 1)
    v[1] = (0, 0, -1)
    v[2] = (1, 0, 0)
    v[3] = np.array([0.6, 0.0, 0.8], dtype=np.float32)
    return v.astype(np.float32)


def _vec(r, n):  # arbitrary (non-unit) vectors
    return (r.normal(size=(n, 3)) * 2.0).astype(np.float32)


def _color(r, n):
    return r.random((n, 3), dtype=np.float32)


def _alpha(r, n):
    a = (1e-3 + r.random((n, 2), dtype=np.float32) ** 2).astype(np.float32)
    a[: n // 2, 1] = a[: n // 2, 0]  # isotropic half
    return a


def _eta(r, n):
    return (1.05 + 1.5 * r.random((n, 1), dtype=np.float32)).astype(np.float32)


def _one(r, n):
    return np.ones((n, 1), dtype=np.float32)


def _eta3(r, n):
    return (0.1 + 3.0 * r.random((n, 3), dtype=np.float32)).astype(np.float32)


def _k3(r, n):
    return (0.5 + 5.0 * r.random((n, 3), dtype=np.float32)).astype(np.float32)


def _cos(r, n):
    c = (2.0 * r.random((n, 1), dtype=np.float32) - 1.0).astype(np.float32)
    c[0], c[1], c[2] = 1.0, -1.0, 0.0
    return c


def _pdf(r, n):
    p = (10.0 * r.random((n, 1), dtype=np.float32) ** 3).astype(np.float32)
    p[0] = 0.0
    return p


def _sigma(r, n):
    return (90.0 * r.random((n, 1), dtype=np.float32)).astype(np.float32)


def _ratio(r, n):
    return r.random((n, 1), dtype=np.float32)


def _eta_any(r, n):  # relative index for refract(): both < 1 and > 1
    return (0.4 + 1.6 * r.random((n, 1), dtype=np.float32)).astype(np.float32)


def _count(r, n):
    return np.full((n, 1), 37, dtype=np.uint32).view(np.float32)


def _frame(r, n):  # geometric normal, a shading normal on its side, a tangent that is not parallel to it
    ng = _dir(r, n)
    ns = ng + 0.4 * r.normal(size=(n, 3)).astype(np.float32)
    ns /= np.linalg.norm(ns, axis=1, keepdims=True)
    ns[: n // 4] = ng[: n // 4]  # flat-shaded quarter
    tg = _vec(r, n)
    return np.concatenate([ng, ns.astype(np.float32), tg], axis=1).astype(np.float32)


def _disney_ctx(r, n):  # lrk_surface.p[0..14] of a Disney node (include/lrk.h)
    c = r.random((n, 15), dtype=np.float32)
    c[:, 3] = 0.212671 * c[:, 0] + 0.715160 * c[:, 1] + 0.072169 * c[:, 2]  # color_lum
    c[:, 5] = 1.1 + 0.9 * c[:, 5]  # eta
    c[:, 6] = np.maximum(c[:, 6] ** 2, 1e-4)  # roughness remapped to alpha
    c[:, 13] = 0.0  # specular_trans (opaque closure)
    c[: n // 8, 0:4] = 0.0  # black base colour: tint_weight = 1 branch
    c[n // 8: n // 4, 4] = 1.0  # fully metallic
    return c.astype(np.float32)


def _disney_trans_ctx(r, n):  # a transmissive Disney node: specular_trans in [0, 1], with pure-transmission and opaque-like rows
    c = _disney_ctx(r, n)
    c[:, 13] = r.random(n, dtype=np.float32)
    c[: n // 6, 13] = 1.0
    c[n // 6: n // 4, 4] = 0.0  # non-metallic rows keep the transmission lobe alive
    return c.astype(np.float32)


def _disney_thin_ctx(r, n):  # a thin Disney node: 16 context values, the last one diffuse_trans (lrk_surface.p[15])
    c = np.concatenate([_disney_trans_ctx(r, n), r.random((n, 1), dtype=np.float32)], axis=1)
    c[n // 3: n // 2, 15] = 0.0  # no diffuse transmission
    c[n // 2: 2 * n // 3, 13] = 0.0  # no specular transmission
    c[-n // 8:, 15] = 1.0  # everything diffuse goes through
    return c.astype(np.float32)


def _sigma_a(r, n):
    a = (2.0 * r.random((n, 3), dtype=np.float32)).astype(np.float32)
    a[: n // 3] = 0.0
    return a


def _kd_scaled(r, n):
    return (0.05 + 1.2 * r.random((n, 3), dtype=np.float32)).astype(np.float32)


_EVAL_OUT = "ffff"
_SAMPLE_OUT = "fffffffu"

# pin name -> (argument generators in the callable's argument order, output kinds: 'f' float words / 'u' integer words)
PINS = {
    "xxhash32_4": ([lambda r, n: _bits(r, n, 4)], "u"),
    "uniform_uint_to_float": ([_bits], "f"),
    "lcg": ([_bits], "fu"),
    "pcg32_seq": ([lambda r, n: _bits(r, n, 2)], "uuuuffuuuu"),
    "sample_uniform_triangle": ([_u2], "fff"),
    "sample_uniform_disk_concentric": ([_u2], "ff"),
    "sample_cosine_hemisphere": ([_u2], "fff"),
    "cosine_hemisphere_pdf": ([_cos], "f"),
    "sample_uniform_sphere": ([_u2], "fff"),
    "balance_heuristic": ([_pdf, _pdf], "f"),
    "sample_alias_table": ([_count, _u], "uf"),
    "frame_make_n": ([_dir], "f" * 9),
    "frame_make_ns": ([_dir, _vec], "f" * 9),
    "frame_local_to_world": ([_dir, _dir, _dir, _vec], "fff"),
    "frame_world_to_local": ([_dir, _dir, _dir, _vec], "fff"),
    "clamp_shading_normal": ([_dir, _dir, _dir], "fff"),
    "refract": ([_dir, _dir, _eta_any], "ffff"),
    "face_forward": ([_vec, _vec], "fff"),
    "spherical_direction": ([_u, _u, lambda r, n: (6.28 * _u(r, n)).astype(np.float32)], "fff"),
    "spherical_theta": ([_dir], "f"),
    "spherical_phi": ([_dir], "f"),
    "tr_roughness_to_alpha": ([_u], "f"),
    "tr_D": ([_alpha, _dir], "f"),
    "tr_Lambda": ([_alpha, _dir], "f"),
    "tr_G1": ([_alpha, _dir], "f"),
    "tr_G": ([_alpha, _dir, _dir], "f"),
    "tr_sample_wh": ([_alpha, _dir, _u2], "fff"),
    "tr_pdf": ([_alpha, _dir, _dir], "f"),
    "fresnel_dielectric": ([_cos, _one, _eta], "f"),
    "fresnel_conductor": ([_cos, _one, _eta3, _k3], "fff"),
    "fresnel_dielectric_integral": ([lambda r, n: (0.5 + 2.0 * _u(r, n)).astype(np.float32)], "f"),
    "lambert_reflection_evaluate": ([_color, _dir, _dir], "fff"),
    "lambert_reflection_sample": ([_color, _dir, _u2], "f" * 7),
    "lambert_reflection_pdf": ([_color, _dir, _dir], "f"),
    "oren_nayar_evaluate": ([_color, _sigma, _dir, _dir], "fff"),
    "microfacet_reflection_dielectric_evaluate": ([_color, _alpha, _one, _eta, _dir, _dir], "fff"),
    "microfacet_reflection_dielectric_sample": ([_color, _alpha, _one, _eta, _dir, _u2], "f" * 7),
    "microfacet_reflection_dielectric_pdf": ([_color, _alpha, _one, _eta, _dir, _dir], "f"),
    "microfacet_reflection_conductor_evaluate": ([_color, _alpha, _eta3, _k3, _dir, _dir], "fff"),
    "microfacet_reflection_conductor_sample": ([_color, _alpha, _eta3, _k3, _dir, _u2], "f" * 7),
    "microfacet_transmission_evaluate": ([_color, _alpha, _one, _eta, _dir, _dir], "fff"),
    "microfacet_transmission_sample": ([_color, _alpha, _one, _eta, _dir, _u2], "f" * 7),
    "microfacet_transmission_pdf": ([_color, _alpha, _one, _eta, _dir, _dir], "f"),
    "fresnel_blend_evaluate": ([_color, _color, _alpha, _ratio, _dir, _dir], "fff"),
    "fresnel_blend_sample": ([_color, _color, _alpha, _ratio, _dir, _u2], "f" * 7),
    "fresnel_blend_pdf": ([_color, _color, _alpha, _ratio, _dir, _dir], "f"),
    # the reference's surface closures through Surface::Closure::{evaluate,sample} (oracle/ref/pin_<surface>.cpp)
    "matte_evaluate": ([_color, _sigma, _frame, _dir, _dir], _EVAL_OUT),
    "matte_sample": ([_color, _sigma, _frame, _dir, _u, _u2], _SAMPLE_OUT),
    "mirror_evaluate": ([_color, _alpha, _frame, _dir, _dir], _EVAL_OUT),
    "mirror_sample": ([_color, _alpha, _frame, _dir, _u, _u2], _SAMPLE_OUT),
    "glass_evaluate": ([_color, _color, _eta, _alpha, _ratio, _frame, _dir, _dir], _EVAL_OUT),
    "glass_sample": ([_color, _color, _eta, _alpha, _ratio, _frame, _dir, _u, _u2], _SAMPLE_OUT),
    "plastic_evaluate": ([_kd_scaled, _ratio, _sigma_a, _eta, _alpha, _frame, _dir, _dir], _EVAL_OUT),
    "plastic_sample": ([_kd_scaled, _ratio, _sigma_a, _eta, _alpha, _frame, _dir, _u, _u2], _SAMPLE_OUT),
    "metal_evaluate": ([_eta3, _k3, _color, _alpha, _frame, _dir, _dir], _EVAL_OUT),
    "metal_sample": ([_eta3, _k3, _color, _alpha, _frame, _dir, _u, _u2], _SAMPLE_OUT),
}
for _mask in (35, 59, 63, 32):  # lobe masks: diffuse+retro+specular, + sheen + clearcoat, + fake subsurface, specular only
    PINS[f"disney_evaluate_{_mask}"] = ([_disney_ctx, _frame, _dir, _dir], _EVAL_OUT)
    PINS[f"disney_sample_{_mask}"] = ([_disney_ctx, _frame, _dir, _u, _u2], _SAMPLE_OUT)


for _mask in (163, 191):  # transmissive closure: diffuse+retro+specular+spec_trans, and every lobe + spec_trans
    PINS[f"disneytrans_evaluate_{_mask}"] = ([_disney_trans_ctx, _frame, _dir, _dir], _EVAL_OUT)
    PINS[f"disneytrans_sample_{_mask}"] = ([_disney_trans_ctx, _frame, _dir, _u, _u2], _SAMPLE_OUT)


for _mask in (227, 255, 108):  # thin closure: base + both transmissions, every lobe, and a mask without the diffuse bits
    PINS[f"disneythin_evaluate_{_mask}"] = ([_disney_thin_ctx, _frame, _dir, _dir], _EVAL_OUT)
    PINS[f"disneythin_sample_{_mask}"] = ([_disney_thin_ctx, _frame, _dir, _u, _u2], _SAMPLE_OUT)


def alias_table_values(seed: int = 7, n: int = 37) -> np.ndarray:
    r = np.random.default_rng(seed)
    v = r.random(n, dtype=np.float32) ** 3
    v[5] = 0.0
    v[11] = 4.0
    return v.astype(np.float32)


def make_inputs(name: str, n: int = N) -> np.ndarray:
    gens, _ = PINS[name]
    r = np.random.default_rng(zlib.crc32(name.encode()))
    cols = [g(r, n) for g in gens]
    return np.ascontiguousarray(np.concatenate([c.view(np.uint32) for c in cols], axis=1))


class RefPins:
    def __init__(self, path: Path = LIB):
        self.lib = ctypes.CDLL(str(path))
        self.lib.refpin_last_error.restype = ctypes.c_char_p
        self.lib.refpin_name.restype = ctypes.c_char_p
        self.lib.refpin_eval.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_uint64]

    def names(self):
        return [self.lib.refpin_name(i).decode() for i in range(self.lib.refpin_count())]

    def signature(self, name: str):
        ni, no = ctypes.c_int(), ctypes.c_int()
        if self.lib.refpin_signature(name.encode(), ctypes.byref(ni), ctypes.byref(no)) != 0:
            raise RuntimeError(self.lib.refpin_last_error().decode())
        return ni.value, no.value

    def eval(self, name: str, inputs: np.ndarray, buffer: np.ndarray | None = None, buffer_count: int = 0) -> np.ndarray:
        ni, no = self.signature(name)
        assert inputs.dtype == np.uint32 and inputs.shape[1] == ni, (name, inputs.shape, ni)
        out = np.zeros((inputs.shape[0], no), dtype=np.uint32)
        bptr = buffer.ctypes.data_as(ctypes.c_void_p) if buffer is not None else None
        rc = self.lib.refpin_eval(name.encode(), inputs.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p),
                                  inputs.shape[0], bptr, buffer_count)
        if rc != 0:
            raise RuntimeError(f"{name}: {self.lib.refpin_last_error().decode()}")
        return out

    def create_alias_table(self, values: np.ndarray):
        n = len(values)
        prob = np.zeros(n, dtype=np.float32)
        alias = np.zeros(n, dtype=np.uint32)
        pdf = np.zeros(n, dtype=np.float32)
        self.lib.refpin_create_alias_table(values.ctypes.data_as(ctypes.c_void_p), n, prob.ctypes.data_as(ctypes.c_void_p),
                                           alias.ctypes.data_as(ctypes.c_void_p), pdf.ctypes.data_as(ctypes.c_void_p))
        return prob, alias, pdf


def alias_buffer(prob: np.ndarray, alias: np.ndarray) -> np.ndarray:
    """AliasEntry{float prob; uint alias} records (src/util/sampling.h:29-32)."""
    buf = np.zeros((len(prob), 2), dtype=np.uint32)
    buf[:, 0] = prob.view(np.uint32)
    buf[:, 1] = alias
    return np.ascontiguousarray(buf)


def main() -> int:
    if not LIB.exists():
        print(f"{LIB} is missing: run `make -C oracle/ref` (needs /root/reference)", file=sys.stderr)
        return 1
    ref = RefPins()
    known = set(ref.names())
    data = {}
    values = alias_table_values()
    prob, alias, pdf = ref.create_alias_table(values)
    data["create_alias_table/values"] = values
    data["create_alias_table/prob"] = prob
    data["create_alias_table/alias"] = alias
    data["create_alias_table/pdf"] = pdf
    table = alias_buffer(prob, alias)
    for name, (_, kinds) in PINS.items():
        assert name in known, name
        inp = make_inputs(name)
        buf = table if name == "sample_alias_table" else None
        out = ref.eval(name, inp, buf, len(values) if buf is not None else 0)
        assert out.shape[1] == len(kinds), (name, out.shape, kinds)
        data[f"{name}/in"] = inp
        data[f"{name}/out"] = out
        print(f"{name:48s} {inp.shape[1]:3d} -> {out.shape[1]:2d} words x {inp.shape[0]}")
    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT, **data)
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
