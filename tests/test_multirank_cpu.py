"""CPU suite: the N>1 host logic (sample split + film reduce) with world_size 2 on
the gloo backend."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from raytracingproject_b200 import multigpu


def test_ranges_partition_the_sample_set():
    for world in (1, 2, 3, 4, 8):
        for total in (1, 7, 64, 1024, 1025):
            got = []
            for r in range(world):
                b, n = multigpu.strong_range(r, world, total, start_sample=5)
                got += list(range(b, b + n))
            assert got == list(range(5, 5 + total))
        ranges = [multigpu.weak_range(r, 16) for r in range(world)]
        assert [b for b, _ in ranges] == [16 * r for r in range(world)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sample_film(s, shape):
    rng = np.random.default_rng(1000 + s)
    return rng.random(shape, dtype=np.float32)


def _worker(rank, world, port, total, use_oracle, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    begin, n = multigpu.strong_range(rank, world, total)
    if use_oracle:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from oracle import cycles_ref
        from raytracingproject_b200 import scenes
        rs = cycles_ref.build_scene(scenes.cornell(32, 18, materials="diffuse"), threads=1)
        film_np, _ = rs.render(begin, n, tile_size=16)
        rs.close()
    else:
        film_np = np.zeros((9, 16, 4), np.float32)
        for s in range(begin, begin + n):
            film_np += _sample_film(s, film_np.shape)
    film = torch.from_numpy(film_np.copy())
    multigpu.reduce_film(film)
    if rank == 0:
        np.save(out, film.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("use_oracle", [False, True])
def test_two_rank_film_reduce(tmp_path, use_oracle):
    total = 6
    if use_oracle:
        from oracle import cycles_ref
        if not cycles_ref.available():
            pytest.skip("oracle/_ref not built")
    out = str(tmp_path / "film.npy")
    mp.spawn(_worker, args=(2, _free_port(), total, use_oracle, out), nprocs=2, join=True)
    got = np.load(out)
    if use_oracle:
        from oracle import cycles_ref
        from raytracingproject_b200 import scenes
        rs = cycles_ref.build_scene(scenes.cornell(32, 18, materials="diffuse"), threads=1)
        want, _ = rs.render(0, total, tile_size=16)
        rs.close()
    else:
        want = np.zeros((9, 16, 4), np.float32)
        for s in range(total):
            want += _sample_film(s, want.shape)
    # same samples, only the fp32 addition order differs ((a+b+c)+(d+e+f) vs a+b+..+f)
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-6)
    assert abs(multigpu.display_scale(total) * total - 1.0) < 1e-7
