"""N>1 host logic on CPU: world_size-2 gloo process group, sharded sim ranges, one integer all-reduce.
The per-rank counts come from the scalar native mirror (the GPU kernel cannot run here); what is under test is
the product's sharding + collective code in monte-carlo-gp_b200/distributed.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_cases as gc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_sims, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests")]
    import mcgp_b200
    from oracle import pyoracle as po
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, mc, seed, _ = gc.get_case("sprint19")
    params = po.make_params(cfg, mc)

    def producer(begin, count, hist):
        hist[0] += torch.from_numpy(po.run_native(params, 2024, count, sim_begin=begin)["hist"])

    D = importlib_distributed()
    hist = D.run_sharded(n_sims, 1, params.n_drivers, producer)
    np.save(os.path.join(out_dir, f"hist{rank}.npy"), hist.numpy())
    dist.destroy_process_group()


def importlib_distributed():
    import importlib
    return importlib.import_module("monte-carlo-gp_b200.distributed")


def test_shard_ranges_tile_exactly():
    D = importlib_distributed()
    for n in (0, 1, 7, 10 ** 9 + 7):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (b0, c0), (b1, _) in zip(spans, spans[1:]):
                assert b0 + c0 == b1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        D.shard_range(10, 2, 2)


def test_world2_allreduce_equals_single_process(tmp_path, oracle):
    n_sims, world = 6001, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_sims, str(tmp_path)), nprocs=world, join=True)
    cfg, mc, seed, _ = gc.get_case("sprint19")
    single = oracle.run_native(oracle.make_params(cfg, mc), 2024, n_sims)["hist"]
    for r in range(world):
        got = np.load(tmp_path / f"hist{r}.npy")
        assert got.shape == (1, 20, 20) and np.array_equal(got[0], single)


def test_tallies():
    D = importlib_distributed()
    h = np.arange(16).reshape(4, 4)
    t = D.tallies(h)
    assert t["win"].tolist() == [0, 4, 8, 12] and t["podium"].tolist() == [3, 15, 27, 39] and t["points"].tolist() == h.sum(1).tolist()
