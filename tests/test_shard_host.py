"""Host logic of the sharded multi-GPU apply, on CPU: the cyclic distribution of a work vector (mrx_shard_cyclic, the layout
mrx_apply_sharded uses), the contiguous weighted partition (mrx_shard_partition) and the exchange protocol built on the
cyclic layout (rank-major padded segments, one all-gather, identical split decisions), run with world_size 2 over gloo.
The device part of the same path is covered by tests/test_gpu_sharded.py."""
import os
import socket

import numpy as np
import pytest


def test_partition_properties(libs):
    mw, _ = libs
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            cost = rng.integers(1, 2000, size=n)
            b = mw.shard_partition(cost, world)
            assert len(b) == world + 1 and b[0] == 0 and b[-1] == n
            assert all(b[r] <= b[r + 1] for r in range(world))  # contiguous, order preserving, covers everything
            if n >= 8 * world:
                loads = [int(cost[b[r]:b[r + 1]].sum()) for r in range(world)]
                assert max(loads) <= cost.sum() / world + cost.max()  # balanced to within one item
    # degenerate costs (nodes without any band) still spread by count
    b = mw.shard_partition(np.ones(10, dtype=np.int64), 2)
    assert b == [0, 5, 10]


def test_cyclic_distribution_properties(libs):
    mw, _ = libs
    B = mw.shard_block()  # items dealt out together (MRX_SHARD_BLOCK)
    assert B >= 1
    for n in (0, 1, 2, 7, 8, 9, 64, 1001):
        for world in (1, 2, 3, 4, 8):
            counts = [mw.shard_cyclic(n, world, r) for r in range(world)]
            rows = counts[0][1]
            assert all(c[1] == rows for c in counts) and rows == -(-(-(-n // B)) // world) * B
            assert sum(c[0] for c in counts) == n                       # every item has exactly one owner
            assert max(c[0] for c in counts) - min(c[0] for c in counts) <= B  # balanced to within one block
            seen = set()
            for i in range(n):
                row = mw.shard_cyclic_row(i, n, world)
                r, j = divmod(row, rows)
                assert r == (i // B) % world and j == (i // (B * world)) * B + i % B and j < counts[r][0]
                seen.add(row)
            assert len(seen) == n                                       # rows are distinct: the unpack is a permutation


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, seed, q):
    import torch
    import torch.distributed as dist
    import mrcpp_b200 as mw
    from mrcpp_b200 import _lib
    _lib.init(-1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    cost = rng.integers(1, 500, size=n)            # identical on every rank (replicated topology)
    truth = np.sqrt(np.arange(n * 8, dtype=np.float64) + 1.0).reshape(n, 8)  # stands for the component norms
    begin = mw.shard_partition(cost, world)
    # every rank "computes" only its items (i % world == rank) into its segment of a rank-major, padded buffer; ONE
    # all-gather of equal segments completes it; the host reads item i at row (i % world) * rows + i / world
    cnt, rows = mw.shard_cyclic(n, world, rank)
    mine = torch.zeros(rows, 8, dtype=torch.float64)
    B = mw.shard_block()
    owned = [i for i in range(n) if (i // B) % world == rank]  # in work-vector order = local order
    assert len(owned) == cnt
    mine[:cnt] = torch.from_numpy(truth[owned])
    segs = [torch.zeros(rows, 8, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(segs, mine)
    full = torch.cat(segs)
    normsW = torch.stack([full[mw.shard_cyclic_row(i, n, world)] for i in range(n)])
    ok = bool(np.array_equal(normsW.numpy(), truth))
    # the split decision is a pure function of the exchanged norms -> identical on all ranks
    split = (normsW[:, 1:].pow(2).sum(1).sqrt() > 20.0).numpy()
    gathered = [None] * world
    dist.all_gather_object(gathered, split.tobytes())
    same = all(g == gathered[0] for g in gathered)
    q.put((rank, ok, same, begin))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_protocol_gloo_world2(libs):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, 257, 3, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok and same for _, ok, same, _ in res)
    assert res[0][3] == res[1][3]


def test_shared_host_mirror_without_an_arena_falls_back_to_the_rank_mirror(libs):
    """mrx_tree_set_shared_host_mirror with no communicator / no arena: returns 1 and behaves like mrx_tree_set_host_mirror(tree, 1)
    (the result of a sharded apply is then downloaded over the calling rank's link alone); host logic, no device needed"""
    mw, _ = libs
    from mrcpp_b200 import _lib
    L = _lib.load()
    mra = mw.MultiResolutionAnalysis(5, -2, (-1, -1, -1), (2, 2, 2), 20)
    t = mw.FunctionTree(mra)
    assert L.mrx_tree_set_shared_host_mirror(t._h, None) == 1
    assert L.mrx_tree_set_host_mirror(t._h, 0) == 0
