"""N>1 host logic on CPU with the gloo backend (world_size 2): tensor-sharded merge layout +
all-gather, and the Gram buffer reduction.  The CUDA kernels are replaced by numpy here; what is
under test is the partition / exchange / reassembly code the GPU path uses verbatim."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vl_merging_b200.gram import reduce_gram_buffers
from vl_merging_b200.merge import ShardLayout


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return dict(ret)


SIZES = [2304 * 8, 64, 8, 8, 256, 8, 1024, 32, 1024, 8, 8, 8, 8, 13, 1, 777]


def _merge_job(rank, world):
    layout = ShardLayout(SIZES, world)
    rng = np.random.default_rng(7)
    srcs = [(rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)) for n in SIZES]
    local = torch.zeros(max(layout.shard_size[rank], 1))
    for i in layout.mine(rank):  # "the kernel": 0.5*a + 0.5*b on this rank's targets only
        a, b = srcs[i]
        local[layout.offset[i]: layout.offset[i] + SIZES[i]] = torch.from_numpy(np.float32(0.5) * a + np.float32(0.5) * b)
    full = layout.gather(local, rank, None)
    ok = all(np.array_equal(layout.view(full, i, (SIZES[i],)).numpy(), np.float32(0.5) * a + np.float32(0.5) * b)
             for i, (a, b) in enumerate(srcs))
    return ok, sorted(layout.mine(rank)), layout.shard_size


def test_sharded_merge_all_gather_world2():
    res = _spawn(_merge_job)
    assert res[0][0] and res[1][0]
    assert sorted(res[0][1] + res[1][1]) == list(range(len(SIZES)))  # every target has exactly one owner
    assert not set(res[0][1]) & set(res[1][1])
    sizes = res[0][2]
    assert res[1][2] == sizes and abs(sizes[0] - sizes[1]) <= max(SIZES)  # balanced, same view on both ranks


def _gram_job(rank, world):
    d1, d2 = 8, 16
    arena = torch.zeros(d1 * d1 + d2 * d2)
    buffers = {"a": arena[: d1 * d1].view(d1, d1), "b": arena[d1 * d1:].view(d2, d2), "loose": torch.zeros(4, 4)}
    rng = np.random.default_rng(100 + rank)
    xs = {"a": rng.standard_normal((5 + rank, d1)), "b": rng.standard_normal((3, d2))}
    calls, rows = {"a": 0, "b": 0, "loose": 0}, {"a": 0, "b": 0, "loose": 0}
    for k, x in xs.items():
        buffers[k] += torch.from_numpy(x.T @ x).float()
        calls[k] += 1
        rows[k] += x.shape[0]
    if rank == 1:  # a module only one rank saw
        buffers["loose"] += 1.0
        calls["loose"], rows["loose"] = 1, 2
    reduce_gram_buffers(buffers, [arena], calls, rows, None)
    return {k: v.clone().numpy() for k, v in buffers.items()}, calls, rows


def test_gram_all_reduce_world2():
    res = _spawn(_gram_job)
    want = {}
    for rank in range(2):
        rng = np.random.default_rng(100 + rank)
        for k, shape in (("a", (5 + rank, 8)), ("b", (3, 16))):
            x = rng.standard_normal(shape)
            want[k] = want.get(k, 0) + x.T @ x
    for rank in range(2):
        bufs, calls, rows = res[rank]
        for k in ("a", "b"):
            assert np.allclose(bufs[k], want[k], rtol=1e-5, atol=1e-5)  # = sum of the per-shard reference Grams
        assert np.array_equal(bufs["loose"], np.ones((4, 4), np.float32))
        assert calls == {"a": 2, "b": 2, "loose": 1} and rows == {"a": 11, "b": 6, "loose": 2}


def test_layout_single_rank_is_identity():
    layout = ShardLayout(SIZES, 1)
    assert layout.mine(0) == list(range(len(SIZES)))
    flat = torch.arange(layout.shard_size[0], dtype=torch.float32)
    assert layout.gather(flat, 0, None) is flat
    assert all(layout.offset[i] % 4 == 0 for i in range(len(SIZES)))  # 16-byte aligned segments


def _ragged_job(rank, world):
    """Buffers created lazily exist only where the module fired: rank 0 holds {a, only0}, rank 1 holds {a, only1}.
    The reduction first agrees on the union, so both ranks issue the same collectives and end up with the same keys."""
    buffers = {"a": torch.full((4, 4), float(rank + 1))}
    calls, rows = {"a": 1}, {"a": 3}
    buffers[f"only{rank}"] = torch.full((2 + rank, 2 + rank), 10.0 * (rank + 1))
    calls[f"only{rank}"], rows[f"only{rank}"] = 1, 5
    from collections import defaultdict
    calls, rows = defaultdict(int, calls), defaultdict(int, rows)
    reduce_gram_buffers(buffers, [], calls, rows, None)
    return {k: v.numpy().copy() for k, v in buffers.items()}, dict(calls), dict(rows)


def test_ranks_with_different_buffer_sets_agree_first():
    res = _spawn(_ragged_job)
    for rank in range(2):
        bufs, calls, rows = res[rank]
        assert sorted(bufs) == ["a", "only0", "only1"]
        assert np.array_equal(bufs["a"], np.full((4, 4), 3.0, np.float32))
        assert np.array_equal(bufs["only0"], np.full((2, 2), 10.0, np.float32))
        assert np.array_equal(bufs["only1"], np.full((3, 3), 20.0, np.float32))
        assert calls == {"a": 2, "only0": 1, "only1": 1} and rows == {"a": 6, "only0": 5, "only1": 5}


def _mismatch_job(rank, world):
    buffers = {"a": torch.zeros(4 + rank, 4 + rank)}
    try:
        reduce_gram_buffers(buffers, [], {"a": 0}, {"a": 0}, None)
    except RuntimeError as e:
        return str(e)
    return None


def test_width_mismatch_between_ranks_raises_instead_of_hanging():
    res = _spawn(_mismatch_job)
    assert all("width" in (res[r] or "") for r in range(2)), res


def _empty_rank_job(rank, world):
    """One rank saw no hooked module at all: it still takes part in the digest exchange and ends up with zero buffers
    for the other rank's names (no rank skips a collective)."""
    from collections import defaultdict

    buffers = {} if rank == 0 else {"a": torch.full((3, 3), 2.0), "b": torch.full((2, 2), 5.0)}
    calls = defaultdict(int, {} if rank == 0 else {"a": 1, "b": 2})
    rows = defaultdict(int, {} if rank == 0 else {"a": 4, "b": 6})
    reduce_gram_buffers(buffers, [], calls, rows, None)
    return {k: v.numpy().copy() for k, v in buffers.items()}, dict(calls), dict(rows)


def test_a_rank_without_buffers_takes_part_in_the_exchange():
    res = _spawn(_empty_rank_job)
    for rank in range(2):
        bufs, calls, rows = res[rank]
        assert sorted(bufs) == ["a", "b"]
        assert np.array_equal(bufs["a"], np.full((3, 3), 2.0, np.float32)) and np.array_equal(bufs["b"], np.full((2, 2), 5.0, np.float32))
        assert calls == {"a": 1, "b": 2} and rows == {"a": 4, "b": 6}
