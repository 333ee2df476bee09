"""N > 1 host logic on CPU: world-size-2 gloo processes shard a configuration batch and an edge batch,
answer their blocks (oracle-backed stand-in device) and gather; the result must equal the single-process
answer.  Also covers ragged shard sizes."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multirobot_pathplanning_benchmark_b200 import scene as S
from multirobot_pathplanning_benchmark_b200.dist import gather_csr, gather_rows, shard_range, shard_sizes, sharded_map
from multirobot_pathplanning_benchmark_b200.scenes import SCENES


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 64, 1000003):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert sum(shard_sizes(n, w)) == n and max(shard_sizes(n, w)) - min(shard_sizes(n, w)) <= 1


def _worker(rank, world, port, B, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.fakes import OracleSceneDevice
    mk, kw = SCENES["2d_handover"]
    sc = mk()
    cs = S.compile_blob(sc, kw["tol"])
    dev = OracleSceneDevice()
    dev.set_mode(0, cs)
    lim = sc.limits()
    q = np.random.RandomState(0).uniform(lim[0], lim[1], (B, sc.dof)).astype(np.float32)
    flags = sharded_map(lambda s, e: torch.from_numpy(dev.check_configs(0, q[s:e])), B)
    q2 = np.random.RandomState(1).uniform(lim[0], lim[1], (B // 8, sc.dof)).astype(np.float32)
    first = sharded_map(lambda s, e: torch.from_numpy(dev.check_edges(0, q[s:e], q2[s:e], 0.01)[1]), B // 8)
    if rank == 0:
        np.savez(out_path, flags=flags.numpy(), first=first.numpy())
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tmp_path):
    from tests.fakes import OracleSceneDevice
    B = 4001  # ragged on purpose
    out = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(2, 29531 + os.getpid() % 200, B, out), nprocs=2, join=True)
    got = np.load(out)
    mk, kw = SCENES["2d_handover"]
    sc = mk()
    cs = S.compile_blob(sc, kw["tol"])
    dev = OracleSceneDevice()
    dev.set_mode(0, cs)
    lim = sc.limits()
    q = np.random.RandomState(0).uniform(lim[0], lim[1], (B, sc.dof)).astype(np.float32)
    q2 = np.random.RandomState(1).uniform(lim[0], lim[1], (B // 8, sc.dof)).astype(np.float32)
    assert np.array_equal(got["flags"], dev.check_configs(0, q))
    assert np.array_equal(got["first"], dev.check_edges(0, q[:B // 8], q2, 0.01)[1])


def test_gather_rows_single_process_is_identity():
    x = torch.arange(10)
    assert torch.equal(gather_rows(x, 10), x)


def _radius_csr(queries, corpus, r):
    """reference-style r-disc rows (np.where order, prm_graph.py:479-500) as CSR"""
    from oracle import oracle_abstract as OA
    sl = np.array([[0, 2], [2, 4]])
    rows = [np.nonzero(OA.batch_config_dist(q, corpus, sl, "max_euclidean") < r)[0] for q in queries]
    off = np.zeros(len(rows) + 1, np.int64)
    off[1:] = np.cumsum([len(x) for x in rows])
    return off, (np.concatenate(rows) if rows else np.zeros(0, np.int64)).astype(np.int32)


def _csr_worker(rank, world, port, Q, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    corpus = np.random.RandomState(2).uniform(-2, 2, (300, 4))
    queries = corpus[:Q]
    s, e = shard_range(Q, rank, world)
    off, idx = _radius_csr(queries[s:e], corpus, 0.9)
    goff, gidx = gather_csr(torch.from_numpy(off), torch.from_numpy(idx), Q)
    if rank == 1:
        np.savez(out_path, off=goff.numpy(), idx=gidx.numpy())
    dist.destroy_process_group()


def test_ragged_radius_rows_gather_across_two_ranks(tmp_path):
    Q = 101
    out = str(tmp_path / "csr.npz")
    mp.spawn(_csr_worker, args=(2, 29731 + os.getpid() % 200, Q, out), nprocs=2, join=True)
    got = np.load(out)
    corpus = np.random.RandomState(2).uniform(-2, 2, (300, 4))
    off, idx = _radius_csr(corpus[:Q], corpus, 0.9)
    assert np.array_equal(got["off"], off) and np.array_equal(got["idx"], idx)
    assert idx.size > Q  # every row holds at least itself, most hold more
