"""pytest configuration: the `gpu` marker and shared scene fixtures.

CPU tests (-m "not gpu") cover the oracle against its golden vectors, the host front-end and the C-ABI
surface; GPU tests (-m gpu) are the parity tests proper and call through the C-ABI of libb200pt.so.
"""
from __future__ import annotations

import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _native_build():
    """Build the in-tree libraries once (no-op when up to date)."""
    from luisarender_b200 import build as B
    from oracle import binding as O

    B.build_all()
    O.build()


@pytest.fixture(scope="session")
def cornell_small():
    from luisarender_b200 import scenes
    from luisarender_b200.api import Scene

    return Scene.from_source(scenes.cornell_box(resolution=(48, 48), spp=8), REPO)


@pytest.fixture(scope="session")
def cornell_disney_small():
    from luisarender_b200 import scenes
    from luisarender_b200.api import Scene

    return Scene.from_source(scenes.cornell_box(resolution=(48, 48), spp=8, surface="Disney"), REPO)


@pytest.fixture(scope="session")
def spheres_small():
    from luisarender_b200 import scenes
    from luisarender_b200.api import Scene

    return Scene.from_source(
        scenes.instanced_spheres(resolution=(64, 36), spp=4, big_subdivision=4, small_subdivision=2, small_count=12), REPO)


@pytest.fixture(scope="session")
def textured_wrappers_small():
    """Row f1 scene: mesh files, image textures, normal map, alpha-tested and half-transparent surfaces."""
    from luisarender_b200 import scenes
    from luisarender_b200.api import Scene

    return Scene.from_source(scenes.textured_room(resolution=(48, 32), spp=4, wrappers=True), REPO)


@pytest.fixture(scope="session", params=["strict_math", "fast_math"])
def gpu_renderer(request):
    """The device library in its two arithmetic configurations (csrc/device/shade.cu): `fast_math` is the product default - the
    Matte / Disney / volume closure kernels compiled with the fast-math arithmetic the reference's own CUDA backend uses - and
    `strict_math` (lrk_set_option("strict_math", 1)) runs every closure kernel in IEEE arithmetic like the oracle.  Every GPU test
    that takes this fixture runs in both; `gpu_renderer.fast` tells a test which tolerance to state.  Traversal, ray generation
    and the film are the same IEEE code in both."""
    from luisarender_b200.api import Renderer

    r = Renderer(device_index=0)  # raises when there is no CUDA device: GPU tests must not silently fall back
    r.fast = request.param == "fast_math"
    r.set_option("strict_math", 0 if r.fast else 1)
    yield r
    r.close()
