"""Two GPUs, one process each: tile sharding + lrk_reduce_film on real hardware.

Skipped on a single-GPU box (`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu` runs it).  The reduced film must
be bit-identical to the single-GPU film: every pixel has one owner (lrk_tile_owner) and the other ranks add +0.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO / "tools"))
CLI = REPO / "luisarender_b200" / "lib" / "luisa-render-cli"

pytestmark = pytest.mark.gpu


def _gpus() -> int:
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, os.environ["LRB_REPO"])
import torch
from luisarender_b200 import scenes, distributed as D
from luisarender_b200.api import Renderer, Scene

rank, world, local = D.env_world()
torch.cuda.set_device(local)
dist = D.init_process_group("nccl")
scene = Scene.from_source(scenes.instanced_spheres(resolution=(640, 360), spp=8), os.environ["LRB_REPO"])
r = Renderer(device_index=local)
r.upload(scene.desc())
D.init_film_comm(r, rank, world)
r.set_shard(rank, world, D.TILE_SIZE)
r.render(0, 8)
mine = r.film(raw=True).copy()
assert ((mine[..., 3] > 0) == D.owned_pixel_mask(640, 360, rank, world)).all()
r.reduce_film(0)
if rank == 0:
    reduced = r.film(raw=True).copy()
    r.set_shard(0, 1, D.TILE_SIZE)
    r.clear()
    r.render(0, 8)
    single = r.film(raw=True)
    assert np.array_equal(reduced, single), float(np.abs(reduced - single).max())
    assert r.stats()["reduce_ms"] == 0.0  # clear() reset the stats; the reduce above was timed before it
# cost-balanced shards (lrk_balance_shards): every rank probes the frame on its own and must arrive at the same table
r.balance_shards(rank, world, D.TILE_SIZE, 1)
r.render(0, 8)
mine = r.film(raw=True).copy()
owned = torch.from_numpy((mine[..., 3] > 0).astype(np.int32)).cuda()
dist.all_reduce(owned)
assert (owned.cpu().numpy() == 1).all()           # a partition of the film: no tile rendered twice, none missing
r.reduce_film(0)
if rank == 0:
    assert np.array_equal(r.film(raw=True), single)  # and the same film, bit for bit
dist.barrier()
dist.destroy_process_group()
"""


def test_two_rank_reduce_through_the_c_abi_is_bit_identical(tmp_path):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, LRB_REPO=str(REPO))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_cli_on_two_gpus_writes_the_single_gpu_image(tmp_path):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import gen_ref_renders as G
    from luisarender_b200 import scenes

    src = scenes.instanced_spheres(resolution=(640, 360), spp=8)
    images = {}
    for gpus in (1, 2):
        d = tmp_path / f"g{gpus}"
        d.mkdir()
        (d / "scene.luisa").write_text(src)
        cmd = [str(CLI), "-b", "cuda", str(d / "scene.luisa")] + (["--gpus", "2"] if gpus == 2 else [])
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        if gpus == 2:
            assert "Rank 1 of 2" in r.stdout and "film reduce" in r.stdout
        out = [p for p in d.iterdir() if p.suffix in (".exr", ".png", ".hdr")]
        assert len(out) == 1, list(d.iterdir())
        images[gpus] = G.read_image(out[0])
    assert images[1].shape == images[2].shape
    assert np.array_equal(images[1], images[2])
