"""The C-ABI surface: both libraries load without a GPU, export every symbol the headers declare, and the
ctypes mirrors have the same layout as the C structs (checked against gcc's sizeof/offsetof)."""
from __future__ import annotations

import ctypes as C
import re
import subprocess
from pathlib import Path

from luisarender_b200 import _ffi as F

REPO = Path(__file__).resolve().parents[1]


def _declared(header: str, prefix: str) -> set[str]:
    text = (REPO / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"static inline[^{]*\{.*?\n\}", "", text, flags=re.S)  # header-only helpers are not exports
    return set(re.findall(rf"\b({prefix}_[a-z_0-9]+)\s*\(", text))


def test_device_library_exports_every_declared_symbol():
    declared = _declared("lrk.h", "lrk")
    assert declared == set(F.LRK_SYMBOLS)
    lib = F.device_lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.lrk_abi_version() == F.LRK_ABI_VERSION


def test_host_library_exports_every_declared_symbol():
    declared = _declared("lrh.h", "lrh")
    assert declared == set(F.LRH_SYMBOLS)
    lib = F.host_lib()
    for name in declared:
        assert getattr(lib, name) is not None


def test_struct_layouts_match_c(tmp_path):
    structs = {
        "lrk_vertex": F.Vertex, "lrk_triangle": F.Triangle, "lrk_alias_entry": F.AliasEntry, "lrk_ray": F.Ray,
        "lrk_hit": F.Hit, "lrk_mesh": F.Mesh, "lrk_bvh_node": F.BvhNode, "lrk_instance": F.Instance,
        "lrk_surface": F.Surface, "lrk_texture": F.Texture, "lrk_light": F.Light, "lrk_light_handle": F.LightHandle, "lrk_camera": F.Camera,
        "lrk_film": F.Film, "lrk_integrator": F.Integrator, "lrk_medium": F.Medium, "lrk_environment": F.Environment, "lrk_sampler": F.Sampler, "lrk_scene_desc": F.SceneDesc,
        "lrk_device_cfg": F.DeviceCfg, "lrk_stats": F.Stats, "lrh_scene_info": F.SceneInfo,
    }
    offsets = [("lrk_scene_desc", "camera"), ("lrk_scene_desc", "integrator"), ("lrk_scene_desc", "environment_medium"),
               ("lrk_scene_desc", "light_handles"), ("lrk_scene_desc", "texels"), ("lrk_scene_desc", "environment"), ("lrk_scene_desc", "sampler"), ("lrk_sampler", "zsobol_hash"), ("lrk_environment", "alias"), ("lrk_surface", "tex"), ("lrk_stats", "trace_closest_ms"), ("lrk_camera", "filter_alias_indices")]
    src = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{REPO / "include" / "lrh.h"}"', "int main(void){"]
    for name in structs:
        src.append(f'printf("{name} %zu\\n", sizeof({name}));')
    for s, f in offsets:
        src.append(f'printf("{s}.{f} %zu\\n", offsetof({s}, {f}));')
    src.append("return 0;}")
    c_file = tmp_path / "sizes.c"
    c_file.write_text("\n".join(src))
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", str(c_file), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for name, cls in structs.items():
        assert int(out[name]) == C.sizeof(cls), name
    for s, f in offsets:
        assert int(out[f"{s}.{f}"]) == getattr(structs[s], f).offset, (s, f)
    # layouts the kernels rely on
    assert C.sizeof(F.Vertex) == 32 and C.sizeof(F.Triangle) == 12 and C.sizeof(F.Ray) == 32 and C.sizeof(F.Hit) == 16
    assert C.sizeof(F.BvhNode) == 64 and C.sizeof(F.AliasEntry) == 8 and C.sizeof(F.LightHandle) == 8


def test_device_library_fails_loudly_without_a_gpu():
    """No CPU fallback: without a CUDA device lrk_create must return an error, not a working context."""
    import torch

    if torch.cuda.is_available():
        return
    lib = F.device_lib()
    ctx = C.c_void_p()
    cfg = F.DeviceCfg(-1, 0, 0)
    rc = lib.lrk_create(C.byref(cfg), C.byref(ctx))
    assert rc != 0 and not ctx.value


def test_cli_help_and_backend_policy():
    cli = F.LIB_DIR / "luisa-render-cli"
    assert cli.exists()
    r = subprocess.run([str(cli), "-h"], capture_output=True, text=True)
    assert r.returncode == 0 and "--backend" in r.stdout and "--define" in r.stdout
    r = subprocess.run([str(cli)], capture_output=True, text=True)
    assert r.returncode != 0  # scene file not specified


def test_assign_tiles_is_longest_processing_time_first():
    """lrk_assign_tiles (host only): every tile gets exactly one owner, loads are balanced to within one tile's cost, the result is
    deterministic, and a uniform cost degenerates to equal counts."""
    import ctypes as C

    import numpy as np

    lib = F.device_lib()
    rng = np.random.default_rng(5)
    cost = (rng.pareto(1.5, 2040) * 1000 + 500).astype(np.uint32)  # heavy-tailed like a frame with sky and geometry
    for world in (1, 2, 3, 8):
        owner = np.full(2040, 99, np.uint32)
        assert lib.lrk_assign_tiles(cost.ctypes.data, 2040, world, owner.ctypes.data) == 0
        assert owner.max() < world
        load = np.bincount(owner, weights=cost.astype(np.float64), minlength=world)
        assert load.max() - load.min() <= cost.max()
        assert load.max() / load.mean() <= 1.002 or world == 1 or cost.max() > 0.002 * load.mean()
        again = np.zeros(2040, np.uint32)
        lib.lrk_assign_tiles(cost.ctypes.data, 2040, world, again.ctypes.data)
        assert np.array_equal(owner, again)
    flat = np.full(64, 7, np.uint32)
    owner = np.zeros(64, np.uint32)
    lib.lrk_assign_tiles(flat.ctypes.data, 64, 8, owner.ctypes.data)
    assert (np.bincount(owner, minlength=8) == 8).all()
    assert lib.lrk_assign_tiles(None, 4, 2, owner.ctypes.data) != 0
