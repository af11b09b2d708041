"""CPU: the multi-GPU host logic (contiguous work-balanced sharding + final gather of result arrays) on a
world_size-2 gloo group.  The data path has no collective; only the fixed-pitch result arrays are gathered."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_ranges_cover_and_balance():
    from bfa_b200.sharding import shard_ranges, work_estimate
    rng = np.random.default_rng(0)
    T = rng.integers(60, 1800, 1000); N = np.maximum(4, T // rng.integers(5, 20, 1000))
    for world in (1, 2, 3, 8):
        r = shard_ranges(T, N, world)
        assert r[0][0] == 0 and r[-1][1] == 1000 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        w = work_estimate(T, N)
        loads = [w[s:e].sum() for s, e in r]
        assert max(loads) <= 1.1 * (sum(loads) / world) + w.max()
    assert shard_ranges([], [], 4) == [(0, 0)] * 4
    assert shard_ranges([100], [10], 2)[0][1] + 0 >= 0


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bfa_b200.sharding import gather_results, shard_ranges
    T = [100 + 7 * i for i in range(11)]; N = [5 + i for i in range(11)]
    ranges = shard_ranges(T, N, world)
    counts = [e - s for s, e in ranges]
    s0, e0 = ranges[rank]
    P = 12
    # fake per-rank results that encode the global utterance index
    stamps = torch.zeros((counts[rank], P, 4), dtype=torch.int32)
    conf = torch.zeros((counts[rank], P), dtype=torch.float32)
    for i, u in enumerate(range(s0, e0)):
        stamps[i, :, 0] = u; conf[i] = u + 0.5
    n_st = torch.arange(s0, e0, dtype=torch.int32); st = torch.full((counts[rank],), rank, dtype=torch.int32)
    g = gather_results(stamps, conf, n_st, st, counts)
    ok = (g[0].shape == (11, P, 4) and bool((g[0][:, 0, 0] == torch.arange(11)).all()) and bool((g[1][:, 3] == torch.arange(11) + 0.5).all())
          and bool((g[2] == torch.arange(11)).all()) and g[3].tolist() == sum([[r] * counts[r] for r in range(world)], []))
    # the same through the single-collective path: the result arrays are views of one allocation
    from bfa_b200.sharding import gather_packed
    from bfa_b200.aligner import BatchResult, result_arena_words
    bp = max(counts[rank], 1)
    w = result_arena_words(bp, P, True, True)
    arena = torch.zeros(w["total"], dtype=torch.int32)
    a_st = arena[w["stamps"]:w["stamps"] + bp * P * 4].view(bp, P, 4); a_st[:counts[rank]] = stamps
    a_cf = arena[w["conf"]:w["conf"] + bp * P].view(torch.float32).view(bp, P); a_cf[:counts[rank]] = conf
    a_ns = arena[w["n_stamps"]:w["n_stamps"] + bp]; a_ns[:counts[rank]] = n_st
    a_ss = arena[w["status"]:w["status"] + bp]; a_ss[:counts[rank]] = st
    a_dp = arena[w["dp_final"]:w["dp_final"] + bp].view(torch.float32); a_dp[:counts[rank]] = -n_st.float()
    res = BatchResult(None, None, None, a_dp, a_ss, a_st, a_cf, a_ns, None, P)
    res.arena = arena
    h = gather_packed(res, counts)
    ok = ok and all(bool((x == y).all()) for x, y in zip(h[:4], g)) and bool((h[4] == -torch.arange(11).float()).all())
    # gather to ONE rank (north_star: "NCCL used only for the final gather of timestamp arrays"): rank 0 gets everything
    h0 = gather_packed(res, counts, dst=0)
    if rank == 0:
        ok = ok and all(bool((x == y).all()) for x, y in zip(h0[:4], g))
    else:
        ok = ok and h0 is None
    # ragged corpora: ranks whose longest targets differ would pick different default stamp pitches; the pitch is agreed on
    # with one all-reduce, and a gather over mismatched layouts raises instead of mis-slicing
    from bfa_b200.sharding import global_stamp_pitch
    ok = ok and global_stamp_pitch(10 + 20 * rank, 500, True) == 38 and global_stamp_pitch(10, 300 + rank, False) == 301
    P2 = P + 3 * rank
    w2 = result_arena_words(bp, P2, True, True)
    arena2 = torch.zeros(w2["total"], dtype=torch.int32)
    res2 = BatchResult(None, None, None, arena2[w2["dp_final"]:w2["dp_final"] + bp].view(torch.float32), arena2[w2["status"]:w2["status"] + bp],
                       arena2[w2["stamps"]:w2["stamps"] + bp * P2 * 4].view(bp, P2, 4), arena2[w2["conf"]:w2["conf"] + bp * P2].view(torch.float32).view(bp, P2),
                       arena2[w2["n_stamps"]:w2["n_stamps"] + bp], None, P2)
    res2.arena = arena2
    try:
        gather_packed(res2, counts)
        ok = False
    except ValueError as e:
        ok = ok and "different result layouts" in str(e)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gather_results_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs: p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
