"""world_size-2 gloo test of the multi-GPU host logic: contiguous sharding of the batch and the
single statistics reduction.  The per-rank 'solve' here is the CPU oracle (the test is about the
plumbing, the engine itself needs a GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nonlin_b200 import distributed as D
from nonlin_b200 import workloads as W
from nonlin_b200._lib import NLB_STAT_COUNT, NLB_STAT_NAMES


def test_shard_range_covers_batch_without_overlap():
    for B in (0, 1, 7, 8, 1000, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            ranges = [D.shard_range(B, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == B
            for (a, b), (c, d) in zip(ranges, ranges[1:]):
                assert b == c and a <= b
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_range(10, 2, 2)


def test_combine_stats():
    a = np.arange(NLB_STAT_COUNT, dtype=np.int64); b = 2 * a
    out = D.combine_stats([a, b])
    assert out[0] == 0 and out[6] == 18 and out[D.STAT_MAX_INDEX] == 18 and out[8] == 24


def _stats_from(ib, status):
    v = np.zeros(NLB_STAT_COUNT, dtype=np.int64)
    v[0] = status.size; v[1] = (status == 0).sum(); v[5] = (status != 0).sum()
    v[2] = ib["converge_on_fcn"].sum(); v[3] = ib["converge_on_chng"].sum(); v[4] = ib["converge_on_zero_diff"].sum()
    v[6] = ib["iter_count"].sum(); v[7] = ib["fcn_count"].sum(); v[8] = ib["jacobian_count"].sum()
    v[9] = ib["iter_count"].max() if status.size else 0
    return v


def _worker(rank, world, port, B, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.nl_oracle import Oracle

        o = Oracle()
        w = W.c2_broyden_2x2(B)
        lo, hi = D.shard_range(B, rank, world)
        x0 = D.shard_soa(w["x0"], lo, hi)
        x, f, ib, st = o.solve_batch(w["solver"], w["fcn"], x0, nthreads=1)
        stats = torch.from_numpy(_stats_from(ib, st))
        D.allreduce_stats(stats)
        np.save(os.path.join(out_dir, "stats_%d.npy" % rank), stats.numpy())
        np.save(os.path.join(out_dir, "x_%d.npy" % rank), x)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_solve_and_stats_reduction(tmp_path, oracle):
    B, world = 1001, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    w = W.c2_broyden_2x2(B)
    x, f, ib, st = oracle.solve_batch(w["solver"], w["fcn"], w["x0"])
    expect = _stats_from(ib, st)
    s0 = np.load(tmp_path / "stats_0.npy"); s1 = np.load(tmp_path / "stats_1.npy")
    assert np.array_equal(s0, s1) and np.array_equal(s0, expect)
    assert D.stats_dict(s0)["systems"] == B and NLB_STAT_NAMES[0] == "systems"
    # shards concatenate to the unsharded result bit for bit
    xs = np.concatenate([np.load(tmp_path / "x_0.npy"), np.load(tmp_path / "x_1.npy")], axis=1)
    assert np.array_equal(xs, x)
