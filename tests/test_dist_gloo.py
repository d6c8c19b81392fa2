"""world_size-2 gloo test of the N>1 path's host logic: rig-blob broadcast and frame sharding."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from surround360_b200 import dist as sdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    payload = np.arange(1000, dtype=np.uint8) if rank == 0 else np.zeros(0, np.uint8)
    got = sdist.broadcast_bytes(payload, src=0)
    mine = sdist.frames_for_rank(64, rank, world)
    counts = sdist.gather_frame_counts(len(mine))
    np.save(os.path.join(out_dir, f"r{rank}.npy"), got)
    np.save(os.path.join(out_dir, f"f{rank}.npy"), np.array(mine))
    np.save(os.path.join(out_dir, f"c{rank}.npy"), np.array(counts))
    dist.destroy_process_group()


def test_broadcast_and_sharding_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ref = np.arange(1000, dtype=np.uint8)
    frames = []
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"r{r}.npy"), ref)
        assert np.load(tmp_path / f"c{r}.npy").tolist() == [32, 32]
        frames += np.load(tmp_path / f"f{r}.npy").tolist()
    assert sorted(frames) == list(range(64))     # every frame rendered exactly once


def test_sharding_is_a_partition():
    from surround360_b200 import dist as sdist
    for world in (1, 2, 3, 4, 8):
        allf = sum((sdist.frames_for_rank(64, r, world) for r in range(world)), [])
        assert sorted(allf) == list(range(64))
        assert max(len(sdist.frames_for_rank(64, r, world)) for r in range(world)) - \
            min(len(sdist.frames_for_rank(64, r, world)) for r in range(world)) <= 1


def test_video_chain_ranges_are_contiguous_and_cover():
    from surround360_b200 import dist as sdist
    for n, world, chains in ((64, 8, 1), (1000, 8, 12), (17, 2, 4), (5, 2, 4)):
        r = sdist.chain_ranges(n, world, chains)
        flat = [rng for row in r for rng in row]
        assert len(r) == world and all(len(row) == chains for row in r)
        assert flat[0][0] == 0 and flat[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(flat, flat[1:]))          # contiguous, in order, no gaps
        sizes = [b - a for a, b in flat]
        assert max(sizes) - min(sizes) <= 1
