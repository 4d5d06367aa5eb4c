"""CPU tests (world_size 2, gloo) of the multi-GPU host logic in pathfinder-cpp_b200/sharding.py.

The renderer on each rank is the ORACLE here (no GPU in this container): what is under test is the partitioning
claim of SURVEY.md section 8e -- strip k rendered alone, from the same scene translated by -y0, equals rows
[y0, y1) of the full frame -- and the gather plumbing that bench.py uses with NCCL on the GPU box.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "oracle", "pathfinder-cpp_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))

import scenes  # noqa: E402
import sharding  # noqa: E402

SIZE, N_PATHS = 192, 120


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _render(scene, lut):
    import pforacle

    fr = pforacle.Frame(scene, lut)
    px = fr.render()
    fr.close()
    return px


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lut = np.load(os.path.join(ROOT, "tests", "golden", "area_lut.npz"))["lut"]
    paths, colors = scenes.synthetic_paths(N_PATHS, SIZE)
    # translucent paints so that stacking order and partial coverage both matter
    colors = colors.copy()
    colors[:, 3] = 160
    for p in paths:
        p["opaque"] = False
    y0, y1 = sharding.strip_bounds(SIZE, world, rank)
    rows = sharding.strip_rows(SIZE, world)
    local = torch.zeros((rows, SIZE, 4), dtype=torch.uint8)
    if y1 > y0:
        strip = scenes.build_scene_from_outlines(SIZE, SIZE, paths, colors, strip=(y0, y1))
        local[: y1 - y0] = torch.from_numpy(_render(strip, lut))
    full = sharding.gather_strips(local)[:SIZE]
    # scene sharding: every scene of a batch lands on exactly one rank
    share = list(sharding.scene_share(7, world, rank))
    shares = [None] * world
    dist.all_gather_object(shares, share)
    if rank == 0:
        want = _render(scenes.build_scene_from_outlines(SIZE, SIZE, paths, colors), lut)
        np.save(out, np.stack([full.numpy(), want]))
        assert sorted(sum(shares, [])) == list(range(7))
    dist.destroy_process_group()


def test_strip_bounds_cover_the_canvas():
    for height in (16, 100, 512, 8192, 8200):
        for world in (1, 2, 3, 4, 8):
            edges = [sharding.strip_bounds(height, world, r) for r in range(world)]
            assert edges[0][0] == 0 and max(e[1] for e in edges) == height
            for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
                assert a1 == b0 or (b0 == b1 == height)
            assert all(y0 % 16 == 0 or y0 == height for y0, _ in edges)
            assert all(y1 - y0 <= sharding.strip_rows(height, world) for y0, y1 in edges)


def test_scene_share_partitions():
    for n in (0, 1, 7, 4096):
        for world in (1, 2, 8):
            got = sum((list(sharding.scene_share(n, world, r)) for r in range(world)), [])
            assert got == list(range(n))


@pytest.mark.timeout(300)
def test_strips_gathered_over_gloo_equal_the_full_frame(tmp_path):
    out = str(tmp_path / "frames.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    full, want = np.load(out)
    assert np.array_equal(full, want), "strip-sharded frame differs from the single-view frame"
