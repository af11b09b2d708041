"""Multi-GPU sharding of a corpus of utterances.

Every utterance is an independent DP problem (the reference's batch loop, forced_alignment.py:885-905,
carries nothing between iterations), so ranks take contiguous utterance ranges balanced by estimated work
and never exchange posteriors.  The only collective is the final gather of the fixed-pitch result arrays
(stamps / confidences / counts / status).  Works with any torch.distributed backend: NCCL on GPUs, gloo in
the CPU tests."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def work_estimate(T: Sequence[int], N: Sequence[int]) -> np.ndarray:
    """Relative cost of aligning each utterance: frames x live states (band-limited path width)."""
    T = np.asarray(T, np.int64); N = np.asarray(N, np.int64)
    L = 4 * N + 1
    band = np.where(L > 60, np.maximum(L // 4, 20), L)
    return T * np.minimum(L, 2 * band + 1) + 64


def shard_ranges(T: Sequence[int], N: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous [start, end) utterance ranges per rank, balanced by cumulative work."""
    w = work_estimate(T, N)
    B = len(w)
    if world <= 1 or B == 0:
        return [(0, B)] + [(B, B)] * (max(world, 1) - 1)
    cum = np.concatenate([[0], np.cumsum(w)])
    bounds = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        bounds.append(int(np.searchsorted(cum, target, side="left")))
    bounds.append(B)
    bounds = np.maximum.accumulate(np.minimum(bounds, B))
    return [(int(bounds[r]), int(bounds[r + 1])) for r in range(world)]


def gather_results(stamps: torch.Tensor, conf: torch.Tensor, n_stamps: torch.Tensor, status: torch.Tensor, counts: Sequence[int],
                   group=None):
    """All-gather the per-rank result arrays.  Ranks may own different numbers of utterances (`counts[r]`);
    rows are padded to max(counts) for the collective and trimmed afterwards.
    stamps [B_r, P, 4] i32, conf [B_r, P] f32, n_stamps/status [B_r] i32 -> tensors over all utterances."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bmax = int(max(counts)) if len(counts) else 0
    P = stamps.shape[1]

    def pad(x, shape_tail):
        out = x.new_zeros((bmax,) + tuple(shape_tail))
        out[: counts[rank]] = x[: counts[rank]]
        return out

    packs = [pad(stamps, (P, 4)), pad(conf, (P,)), pad(n_stamps, ()), pad(status, ())]
    outs = []
    for x in packs:
        buf = x.new_empty((world * bmax,) + tuple(x.shape[1:]))   # concatenated layout: accepted by NCCL and gloo
        dist.all_gather_into_tensor(buf, x.contiguous(), group=group)
        buf = buf.view((world, bmax) + tuple(x.shape[1:]))
        outs.append(torch.cat([buf[r, : counts[r]] for r in range(world)], dim=0))
    return tuple(outs)


def global_stamp_pitch(max_N: int, max_T: int, ignore_noise: bool, group=None, device=None) -> int:
    """The stamp pitch (BfaShape.max_stamps) every rank must use so that the packed result arenas of all ranks have one
    layout: the default pitch of align_batch depends on the rank's own longest target, and ranks own different utterances.
    One small all-reduce (MAX) before the first batch; pass the result as `max_stamps` to plan_batch / align_batch."""
    mine = (max_N + 8) if ignore_noise else max(max_T, 1)
    t = torch.tensor([mine], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def gather_packed(result, counts: Sequence[int], group=None, dst=None):
    """The same gather as gather_results in ONE collective: `result` is a BatchResult whose per-utterance arrays are
    views of a single allocation (`result.arena`, see aligner.result_arena_words).  Returns
    (stamps [sum B_r, P, 4] i32, conf [sum B_r, P] f32 or None, n_stamps, status, dp_final) over all utterances.
    Every rank must have used the same stamp pitch (global_stamp_pitch); this is checked, not assumed.
    dst = None: all-gather (every rank gets everything).  dst = r: gather to rank r only (north_star's "final gather of the
    timestamp arrays"): the other ranks return None."""
    from .aligner import result_arena_words
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    ms = result.max_stamps
    want_conf = result.conf is not None
    # the layouts must agree: pitch, confidence section, and this rank's count as the caller sees it
    meta = torch.tensor([ms, int(want_conf), int(counts[rank])], dtype=torch.int64, device=result.arena.device)
    metas = meta.new_empty(world * 3)
    dist.all_gather_into_tensor(metas, meta, group=group)
    metas = metas.view(world, 3).cpu()
    if not (bool((metas[:, 0] == ms).all()) and bool((metas[:, 1] == int(want_conf)).all())):
        raise ValueError(f"gather_packed: ranks used different result layouts (max_stamps per rank {metas[:, 0].tolist()}); "
                         "agree on one pitch first (sharding.global_stamp_pitch)")
    if metas[:, 2].tolist() != [int(c) for c in counts]:
        raise ValueError(f"gather_packed: counts {list(counts)} do not match what the ranks hold {metas[:, 2].tolist()}")
    sizes = [result_arena_words(max(int(c), 1), ms, True, want_conf) for c in counts]
    wmax = max(sz["total"] for sz in sizes)
    arena = result.arena
    if arena.numel() < wmax:
        arena = torch.cat([arena, arena.new_zeros(wmax - arena.numel())])
    if dst is None:
        buf = arena.new_empty(world * wmax)
        dist.all_gather_into_tensor(buf, arena[:wmax].contiguous(), group=group)
    else:
        parts = [arena.new_empty(wmax) for _ in range(world)] if rank == dst else None
        dist.gather(arena[:wmax].contiguous(), parts, dst=dst, group=group)
        if rank != dst:
            return None
        buf = torch.cat(parts)
    buf = buf.view(world, wmax)
    st, cf, ns, ss, dp = [], [], [], [], []
    for r in range(world):
        b, sz = int(counts[r]), sizes[r]
        bp = max(b, 1)
        st.append(buf[r, sz["stamps"]:sz["stamps"] + bp * ms * 4].view(bp, ms, 4)[:b])
        if want_conf:
            cf.append(buf[r, sz["conf"]:sz["conf"] + bp * ms].view(torch.float32).view(bp, ms)[:b])
        ns.append(buf[r, sz["n_stamps"]:sz["n_stamps"] + bp][:b])
        ss.append(buf[r, sz["status"]:sz["status"] + bp][:b])
        dp.append(buf[r, sz["dp_final"]:sz["dp_final"] + bp].view(torch.float32)[:b])
    return (torch.cat(st), torch.cat(cf) if want_conf else None, torch.cat(ns), torch.cat(ss), torch.cat(dp))


class PushGather:
    """The same gather without taking SMs from the aligner: every rank PUSHES its packed result arena into all peers'
    receive buffers with copy-engine peer-to-peer writes over NVLink (torch symmetric memory), on a side stream.
    The banded kernel needs every SM of the GPU for one task per warp pair; a collective kernel that grabs a few SMs
    delays it by the collective's whole duration, the copy engines do not.

        pg = PushGather(result.arena.numel(), device)          # collective: every rank of the group calls it
        pg.wait(i); r = align_batch(..., out=results[i]); pg.push(i, r.arena)      # i = batch & 1 (two receive buffers)
        pg.wait(i) ... then a barrier of the group: pg.recv[i].view(world, -1)[q] is rank q's arena of that batch

    `wait(i)` makes the CURRENT stream wait until this rank's pushes of buffer i have left (the arena may be overwritten);
    that the PEERS' pushes have landed is known after a barrier of the group (or any later collective).
    Raises if symmetric memory is not available (callers fall back to gather_packed / all_gather_into_tensor)."""

    def __init__(self, n_words: int, device, group=None, buffers: int = 2, dst=None):
        """dst = None: push to every peer (an all-gather).  dst = r: push to rank r only -- the final gather to one place,
        streamed: with `buffers` = the number of batches, rank r ends up holding recv[i].view(world, -1)[q] = rank q's arena of
        batch i for every batch, and each rank has sent 1 / (world - 1) of what the all-gather sends."""
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank, self.n = dist.get_world_size(group), dist.get_rank(group), int(n_words)
        self.dst = dst
        # every rank must push arenas of one size: agree on it (ranks own different utterances)
        sz = torch.tensor([self.n], dtype=torch.int64, device=device)
        szs = sz.new_empty(self.world)
        dist.all_gather_into_tensor(szs, sz, group=group)
        if not bool((szs == self.n).all()):
            raise ValueError(f"PushGather: ranks have different arena sizes {szs.tolist()}; use one stamp pitch and one batch size "
                             "(sharding.global_stamp_pitch) or pad the arenas")
        self.stream = torch.cuda.Stream(device=device)
        self.recv, self._peers, self._done = [], [], []
        per = self.world * self.n
        big = symm.empty(buffers * per, dtype=torch.int32, device=device)        # one symmetric allocation, one rendezvous
        hdl = symm.rendezvous(big, group)
        peers = [hdl.get_buffer(q, (buffers * per,), torch.int32) for q in range(self.world)]
        for b in range(buffers):
            self.recv.append(big[b * per:(b + 1) * per])
            self._peers.append([pq[b * per:(b + 1) * per] for pq in peers])
            self._done.append(None)

    def push(self, i: int, arena: torch.Tensor) -> None:
        ready = torch.cuda.Event()
        ready.record()                                        # the batch's kernels on the current stream
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            if arena.numel() != self.n:
                raise ValueError(f"PushGather.push: arena of {arena.numel()} words, expected {self.n}")
            targets = [(self.rank + k) % self.world for k in range(self.world)] if self.dst is None else [self.dst]
            for q in targets:                                 # all peers: start with the next rank, spreads the load over the links
                self._peers[i][q][self.rank * self.n:(self.rank + 1) * self.n].copy_(arena, non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        self._done[i] = done

    def wait(self, i: int) -> None:
        if self._done[i] is not None:
            torch.cuda.current_stream().wait_event(self._done[i])
            self._done[i] = None


class PeerArena:
    """The gather without a gather step: every rank's alignment kernels write their packed result arrays (stamps | conf |
    n_stamps | status | dp_final, `align_batch(..., arena=...)`) STRAIGHT into the gathering rank's memory -- ordinary stores to a
    peer-mapped address, carried by NVLink while the kernel computes.  No copy, no side stream, no event between two launches (so
    back-to-back launches keep overlapping their ramp-up and tail), no SM taken from the aligner.

        pa = PeerArena(n_words, device, buffers=K, dst=0)        # collective: one symmetric allocation, one rendezvous
        r = align_batch(..., arena=pa.arena(i))                  # first call for buffer i; later: out=r
        ... torch.cuda.synchronize(); barrier of the group ...   # every rank's kernels are done: their stores have landed
        pa.recv[i].view(world, -1)[q]                            # on rank dst: rank q's arena of buffer i

    The traffic is what the result arrays weigh (4 MB per 4096 utterances), written once.  Raises if symmetric memory is not
    available (callers fall back to PushGather / gather_packed)."""

    def __init__(self, n_words: int, device, group=None, buffers: int = 2, dst: int = 0):
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank, self.dst = dist.get_world_size(group), dist.get_rank(group), int(dst)
        self.n = (int(n_words) + 3) // 4 * 4                      # 16-byte pitch: the kernels store timestamps 16 bytes at a time
        sz = torch.tensor([self.n], dtype=torch.int64, device=device)
        szs = sz.new_empty(self.world)
        dist.all_gather_into_tensor(szs, sz, group=group)
        if not bool((szs == self.n).all()):
            raise ValueError(f"PeerArena: ranks have different arena sizes {szs.tolist()}; use one stamp pitch and one batch size "
                             "(sharding.global_stamp_pitch)")
        per = self.world * self.n
        big = symm.empty(buffers * per, dtype=torch.int32, device=device)
        hdl = symm.rendezvous(big, group)
        big.zero_()                                               # unused stamp slots stay zero (checksums over whole arenas)
        torch.cuda.synchronize(device)
        dist.barrier(group)                                       # nobody stores into rank dst's buffer before it is zeroed
        self._big = big
        self._at_dst = big if self.rank == self.dst else hdl.get_buffer(self.dst, (buffers * per,), torch.int32)
        self.recv = [big[b * per:(b + 1) * per] for b in range(buffers)]     # meaningful on rank dst
        self._per = per

    def arena(self, i: int) -> torch.Tensor:
        """This rank's slot of buffer i in rank dst's memory (a peer-mapped tensor on every rank but dst)."""
        o = i * self._per + self.rank * self.n
        return self._at_dst[o:o + self.n]
