"""the Spherical environment light (src/environments/spherical.cpp) with the uniform light
sampler's environment handling (src/lightsamplers/uniform.cpp:40-101) — host importance map, oracle sampling / evaluation
and the estimator's miss + NEE branches, on the CPU.  GPU parity: tests/test_gpu_parity.py (scenes "environment*")."""
from __future__ import annotations

import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))

from luisarender_b200 import scenes  # noqa: E402
from luisarender_b200.api import Scene  # noqa: E402
from oracle import binding as O  # noqa: E402

f32 = np.float32


@pytest.fixture(scope="module")
def env_image_scene():
    return Scene.from_source(scenes.environment_scene(resolution=(48, 30), spp=16), REPO)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _sample(d, u):
    out = np.zeros(7, f32)
    O.lib().oracle_environment_sample(C.byref(d), _fp(np.asarray(u, f32)), _fp(out))
    return out


def _evaluate(d, w):
    out = np.zeros(4, f32)
    O.lib().oracle_environment_evaluate(C.byref(d), _fp(np.asarray(w, f32)), _fp(out))
    return out


def test_importance_map_is_a_normalised_density_with_mis_compensation(env_image_scene):
    d = env_image_scene.desc()
    e = d.environment
    assert e.present == 1 and e.emission_tex != 0 and (e.map_width, e.map_height) == (2048, 1024)
    assert e.env_prob == pytest.approx(0.5)  # area lights present: clamp(environment_weight, 0.01, 0.99)
    pdf = np.ctypeslib.as_array(e.pdf, shape=(1024, 2048))
    assert pdf.mean() == pytest.approx(1.0, rel=1e-4)  # pdf = p(x, y) * pixel_count (spherical.cpp:207-215)
    assert (pdf >= 0).all() and 0.3 < (pdf == 0).mean() < 0.95  # MIS compensation zeroes everything below the average
    # the brightest cells sit where the sun texel is (texture row 4 of 16, column 22-23 of 32)
    y, x = np.unravel_index(np.argmax(pdf), pdf.shape)
    assert 4 / 16 <= (y + 0.5) / 1024 <= 5 / 16 and 22 / 32 <= (x + 0.5) / 2048 <= 24 / 32
    # alias tables: marginal over rows, then one conditional table per row; sampling reproduces the density
    alias = np.ctypeslib.as_array(C.cast(e.alias, C.POINTER(C.c_uint32)), shape=(1024 + 1024 * 2048, 2))
    prob = alias[:, 0].view(f32)
    assert ((prob >= 0) & (prob <= 1.0 + 1e-5)).all() and (alias[:1024, 1] < 1024).all() and (alias[1024:, 1] < 2048).all()
    rng = np.random.default_rng(3)
    n = 200_000
    u = rng.uniform(size=(n, 2)).astype(f32)

    def draw(table_prob, table_alias, count, uu):
        s = uu * f32(count)
        i = np.minimum(s.astype(np.uint32), count - 1)
        r = s - np.floor(s)
        return np.where(r < table_prob[i], i, table_alias[i])

    iy = draw(prob[:1024], alias[:1024, 1], 1024, u[:, 1])
    base = 1024 + iy.astype(np.int64) * 2048
    s = u[:, 0] * f32(2048)
    ix0 = np.minimum(s.astype(np.uint32), 2047)
    r = s - np.floor(s)
    ix = np.where(r < prob[base + ix0], ix0, alias[base + ix0, 1])
    hist = np.zeros((16, 32))
    np.add.at(hist, (iy // 64, ix // 64), 1)
    expect = pdf.reshape(16, 64, 32, 64).mean(axis=(1, 3)) / (16 * 32) * n
    big = expect > 200
    assert big.sum() >= 3 and (np.abs(hist[big] - expect[big]) < 5 * np.sqrt(expect[big])).all()  # within 5 sigma per bin


def test_sample_and_evaluate_agree(env_image_scene):
    d = env_image_scene.desc()
    rng = np.random.default_rng(9)
    ratios, ok = [], 0
    for u in rng.uniform(size=(400, 2)):
        s = _sample(d, u)
        assert np.isfinite(s).all() and abs(np.linalg.norm(s[4:7]) - 1) < 1e-5 and s[3] > 0
        e = _evaluate(d, s[4:7])
        ratios.append(e[3] / s[3])
        ok += np.allclose(e[:3], s[:3], rtol=0.05, atol=1e-3)
    # the direction round-trips through acos / atan2, so a sample on a cell border may land in the neighbouring cell
    assert np.median(ratios) == pytest.approx(1.0, rel=1e-3) and np.mean(np.abs(np.array(ratios) - 1) < 1e-2) > 0.9
    assert ok > 0.9 * 400
    # the sampling density integrates to one over the sphere (uniform directions, importance weight 4 pi)
    w = rng.normal(size=(60000, 3))
    w /= np.linalg.norm(w, axis=1, keepdims=True)
    integral = np.mean([_evaluate(d, x)[3] for x in w]) * 4 * np.pi
    assert integral == pytest.approx(1.0, rel=0.08)
    # world <-> environment rotation: the sun sits at uv ~ (22.7/32, 4.5/16) of the texture, rotated by 40 degrees about +y
    sun = np.array([_sample(d, u)[4:7] for u in rng.uniform(size=(300, 2))])
    phi, theta = 2 * np.pi * (1 - 22.7 / 32), np.pi * 4.5 / 16
    local = np.array([np.sin(phi) * np.sin(theta), np.cos(theta), np.cos(phi) * np.sin(theta)])
    a = np.radians(40.0)
    world = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]) @ local
    assert (sun @ world > 0.99).mean() > 0.3  # a large share of the samples goes to the sun


def test_constant_environment_is_a_furnace_for_a_convex_diffuse_object():
    """White constant environment, a lone Matte sphere (Kd = 0.8, convex: every bounce escapes): radiance = 0.8 on the sphere,
    1 on the background — checks the miss term, uniform-sphere NEE and their MIS weights together."""
    src = scenes.environment_scene(resolution=(40, 40), spp=64, emission=(1.0, 1.0, 1.0), area_light=False, depth=12)
    src = src.replace("shapes { @ball, @floor }", "shapes { @ball }").replace("position { 0.0, 1.2, 4.0 }", "position { 0.0, 0.7, 4.0 }").replace(
        "front { 0.0, -0.12, -1.0 }", "front { 0.0, 0.0, -1.0 }")
    d = Scene.from_source(src, REPO).desc()
    assert d.environment.env_prob == 1.0 and d.environment.emission_tex == 0 and d.light_count == 0
    img = O.convert_film(d, O.render(d, 0, 64)[0])[..., :3]
    assert np.allclose(img[:3, :, :], 1.0, atol=1e-5)  # background rows: pure miss radiance
    centre = img[17:23, 17:23, :]
    assert centre.mean() == pytest.approx(0.8, rel=0.02)


def test_environment_weight_and_lighting_checks():
    none = scenes.environment_scene(area_light=True).replace("environment : Spherical", "environment_unused : Spherical")
    assert Scene.from_source(none, REPO).desc().environment.present == 0
    only_env = Scene.from_source(scenes.environment_scene(emission=(0.5, 0.6, 0.7), area_light=False), REPO).desc()
    assert only_env.environment.env_prob == 1.0 and only_env.light_count == 0
    w = Scene.from_source(scenes.environment_scene(emission=(0.5, 0.6, 0.7), environment_weight=0.999), REPO).desc()
    assert w.environment.env_prob == pytest.approx(0.99)
    black = Scene.from_source(scenes.environment_scene(emission=(0.0, 0.0, 0.0)), REPO).desc()
    assert black.environment.present == 0  # is_black environments are not built (pipeline.cpp)
    # the estimator with and without the environment: the area-light-only image is darker everywhere the sky is visible
    a = Scene.from_source(scenes.environment_scene(resolution=(32, 20), spp=8, emission=(0.3, 0.3, 0.3)), REPO).desc()
    b = Scene.from_source(none.replace("resolution { 64, 40 }", "resolution { 32, 20 }"), REPO).desc()
    ia, ib = O.render(a, 0, 8)[0][..., :3], O.render(b, 0, 8)[0][..., :3]
    assert ia[:4].mean() > ib[:4].mean() + 0.2 * 8
