"""GPU: sharding.PeerArena -- the alignment kernels of every rank write their packed result arrays straight into the gathering
rank's symmetric-memory buffer (align_batch(arena=...)).  One process per visible GPU (at most 2; with one GPU the group has a
single rank and the "peer" slot is local memory: the same code path end to end)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    try:
        _work(rank, world, port, q)
    except Exception as e:   # noqa: BLE001  (the parent must not wait for a result that never comes)
        import traceback
        q.put((rank, "error", traceback.format_exc()[-1500:]))


def _work(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import bfa_b200
        from bfa_b200 import synth
        from bfa_b200.aligner import result_arena_words
        from bfa_b200.sharding import PeerArena
        Cc, B, T, N = 66, 256, 240, 20
        lp, tgt, _ = synth.planted_batch(B, T, N, Cc, seed=500 + rank, device=dev)
        tg = tgt.to(torch.int32).reshape(-1).contiguous()
        dec = bfa_b200.AlignmentUtils(Cc - 1, 0).viterbi_decoder
        row_off = torch.arange(B, dtype=torch.int64, device=dev) * T * Cc
        p = dec._params(True, True, True)
        plan = dec.plan_batch([T] * B, [N] * B, Cc, params=p, device=dev)
        words = result_arena_words(B, plan.max_stamps, True, True)["total"]
        try:
            pa = PeerArena(words, dev, buffers=2, dst=0)
        except Exception as e:   # noqa: BLE001  (no symmetric memory on this box / in this torch build)
            q.put((rank, "unavailable", str(e)[:200]))
            return
        # reference: the same batch into a zero-initialised local arena
        ref = dec.align_batch(lp, row_off, [T] * B, Cc, tg, [N] * B, params=p, plan=plan, arena=torch.zeros(words, dtype=torch.int32, device=dev))
        outs = [dec.align_batch(lp, row_off, [T] * B, Cc, tg, [N] * B, params=p, plan=plan, arena=pa.arena(i)) for i in range(2)]
        outs[1] = dec.align_batch(lp, row_off, [T] * B, Cc, tg, [N] * B, params=p, plan=plan, out=outs[1])     # reuse keeps the peer slot
        torch.cuda.synchronize()
        dist.barrier()
        mine = ref.arena.clone()
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        ok = True
        if rank == 0:
            for i in range(2):
                got = pa.recv[i].view(world, -1)
                for r in range(world):
                    ok = ok and torch.equal(got[r, :words], allr[r])
        ok = ok and int((ref.status[:B] != 0).sum()) == 0 and int((ref.n_stamps[:B] != N).sum()) == 0
        dist.barrier()
        q.put((rank, "ok" if ok else "mismatch", ""))
    finally:
        dist.destroy_process_group()


def test_peer_arena_results_land_on_rank0():
    import torch.multiprocessing as mp
    world = min(2, torch.cuda.device_count())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    if any(s == "unavailable" for _, s, _ in res):
        pytest.skip("torch symmetric memory unavailable: " + "; ".join(m for _, s, m in res if s == "unavailable"))
    assert all(s == "ok" for _, s, _ in res), res
