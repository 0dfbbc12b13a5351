"""CPU tier: the N>1 host path with world_size 2 over gloo — utterance sharding (afe_shard_utterances), the corpus
statistics record and its merge rule (SUM | MIN | MAX, exactly what afe_normalizer_allreduce issues over NCCL), and
the finalize formulas (afe_cmvn_finalize_host). The per-rank statistics are produced with NumPy here (no GPU); the
GPU tier checks that the fused kernel produces the same record (test_corpus_cmvn_two_shards_equal_one)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _features(n_utts, seed=5):
    """Ragged synthetic 'feature' utterances (rows x 39), the same on every rank."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(50, 400, size=n_utts)
    return [(rng.standard_normal((int(n), 39)) * rng.uniform(0.5, 4) + rng.uniform(-5, 5)).astype(np.float32) for n in lens]


def _record(x):
    """stats record of afe_cuda.h: sum[w], sumsq[w], count, min[w], max[w] (double sums of float values / float products)"""
    x64 = x.astype(np.float64)
    sq = (x * x).astype(np.float64)  # float product widened, normalizercpu.cpp:45
    return np.concatenate([x64.sum(0), sq.sum(0), [float(len(x))], x64.min(0), x64.max(0)])


def _worker(rank, world, port, norm, out_dir):
    sys.path.insert(0, ROOT)
    import afe_loader
    afe = afe_loader.load()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    feats = _features(37)
    lens = np.array([len(f) * 160 + 240 for f in feats], np.int64)      # sample lengths the sharder balances on
    starts = afe.shard_utterances(lens, world)
    mine = feats[starts[rank]:starts[rank + 1]]
    w = 39
    rec = _record(np.concatenate(mine)) if mine else np.concatenate([np.zeros(2 * w + 1), np.full(w, np.finfo(np.float32).max), np.full(w, -np.finfo(np.float32).max)])
    t_sum = torch.from_numpy(rec[:2 * w + 1].copy())
    t_min = torch.from_numpy(rec[2 * w + 1:3 * w + 1].copy())
    t_max = torch.from_numpy(rec[3 * w + 1:].copy())
    dist.all_reduce(t_sum, op=dist.ReduceOp.SUM)
    dist.all_reduce(t_min, op=dist.ReduceOp.MIN)
    dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    merged = np.concatenate([t_sum.numpy(), t_min.numpy(), t_max.numpy()])
    mean, scale = afe.cmvn_finalize_host(norm, w, merged)
    local = np.concatenate(mine) if mine else np.zeros((0, w), np.float32)
    y = (local - mean) if norm == 1 else (local - mean) * scale
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), y)
    np.save(os.path.join(out_dir, f"starts{rank}.npy"), starts)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("norm", [1, 2, 3])
def test_two_rank_corpus_cmvn_equals_single_process(tmp_path, norm):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, norm, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    import afe_loader
    afe = afe_loader.load()
    feats = _features(37)
    allx = np.concatenate(feats)
    mean, scale = afe.cmvn_finalize_host(norm, 39, _record(allx))
    want = (allx - mean) if norm == 1 else (allx - mean) * scale
    s0, s1 = np.load(tmp_path / "starts0.npy"), np.load(tmp_path / "starts1.npy")
    np.testing.assert_array_equal(s0, s1)                       # every rank derives the same shard plan
    assert s0[0] == 0 and s0[-1] == 37 and 0 < s0[1] < 37
    got = np.concatenate([np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")])
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)    # sums differ only by double rounding of the merge
    # and the finalize formulas are the reference's (normalizercpu.cpp:31-66)
    x = allx.astype(np.float64)
    np.testing.assert_allclose(mean, x.mean(0), atol=1e-6)
    if norm == 2:
        np.testing.assert_allclose(scale, 1 / x.std(0, ddof=1), rtol=1e-5)
