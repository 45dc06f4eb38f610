"""2-ply lookahead and 1-ply greedy on the device (K5), built from K1 (legal plays), K4 (tensor-core MLP
leaf evaluation fused with the encoder) and two segmented reductions.

The reference's own 2-ply (src/moves/expect_minmax.py:1-206) is commented-out code; this implements the
definition of SURVEY.md 8(c) on the reference's live primitives:

    A      = get_all_possible_moves(me, board, roll)                       (afterstates, reference order)
    score_i = +win_reward(A_i)                      if me has borne off 15 in A_i
            = - sum_{r in 21 rolls} p_r * v_r       otherwise, with
    v_r    = max_j leaf(B_ij),  B_i. = get_all_possible_moves(opp, A_i, r);  V(encode(A_i, opp)) if no reply
    leaf(B) = win_reward(B) if opp has borne off 15 else V(encode(B, flag=opp))
    choice = argmax_i score_i (lowest index on ties)
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import BgError, check, lib
from .engine import MovegenWorkspace, _stream, legal_moves
from .value_net import ValueNet


def segment_argmax(scores: torch.Tensor, starts: torch.Tensor, counts: torch.Tensor):
    B = counts.shape[0]
    best = torch.empty(B, dtype=torch.int32, device=scores.device)
    best_score = torch.empty(B, dtype=torch.float32, device=scores.device)
    with torch.cuda.device(scores.device):
        check(lib().bg_segment_argmax(scores.data_ptr(), starts.data_ptr(), counts.data_ptr(), B, best.data_ptr(),
                                      best_score.data_ptr(), _stream()), "bg_segment_argmax")
    return best, best_score


def greedy_actions(env, net: ValueNet):
    """1-ply greedy (BASELINE config 3): argmax_i V(encode(A_i, flag=mover)) over each game's legal plays,
    straight from the env's ragged afterstate buffer.  -> (actions (N,) i32 (-1 where no legal play), values)"""
    vals = net.values(env.after52, env.row_players, n_rows_dev=env.alloc_rows)
    return segment_argmax(vals, env.legal_starts, env.legal_counts)


class TwoPlySearch:
    """Persistent workspaces (no allocation and no host synchronisation inside the chunk loop): the only host
    syncs of search() are the row count after the root move generation and one status read at the end.
    max_afterstates_per_chunk bounds the workspace: 21 x 40 reply rows of 57 B per afterstate = 4.7 GB at the default
    98,304 (one B200 has 180 GB; larger chunks amortise the overflow tiers' latency: 32,768 -> 81,920 was +3.5 %)."""

    def __init__(self, net: ValueNet, max_afterstates_per_chunk: int = 98304, replies_per_position: int = 40,
                 overlap: bool = True):
        self.net = net
        self.chunk = int(max_afterstates_per_chunk)
        self.rpp = int(replies_per_position)
        self.leaves_evaluated = 0
        self._bufs = None
        self._side = None
        self.overlap = bool(overlap)

    def _workspace(self, dev):
        if self._bufs is None or self._bufs["dev"] != dev or self._bufs["rpp"] != self.rpp:
            W = self.chunk * 21
            cap = max(W * self.rpp, 4096)
            z = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)
            self._bufs = dict(dev=dev, rpp=self.rpp, cap=cap, ws=MovegenWorkspace(W, dev), counts=z(W, torch.int32),
                              starts=z(W, torch.int64), alloc=torch.zeros(1, dtype=torch.int64, device=dev),
                              replies=z((cap, 52), torch.int8), rowp=z(cap, torch.int8), leaf_v=z(cap, torch.float32),
                              pass_v=z(self.chunk, torch.float32), leaves=torch.zeros(1, dtype=torch.int64, device=dev))
        return self._bufs

    def _score_chunk(self, A: torch.Tensor, movers: torch.Tensor, out: torch.Tensor):
        dev = A.device
        M = A.shape[0]
        b = self._workspace(dev)
        L = lib()
        b["alloc"].zero_()
        net = self.net
        if self._side is None and self.overlap:
            self._side = torch.cuda.Stream(device=dev)
        with torch.cuda.device(dev):
            # K1 (replies to the 21 rolls) + K4 (leaf values, pass values) in one call; K4 overlaps K1's overflow tiers
            check(L.bg_twoply_replies_values(A.data_ptr(), movers.data_ptr(), M, b["replies"].data_ptr(), b["cap"],
                                             b["rowp"].data_ptr(), b["counts"].data_ptr(), b["starts"].data_ptr(),
                                             b["alloc"].data_ptr(), b["ws"].status.data_ptr(), b["ws"].buf.data_ptr(),
                                             b["ws"].nbytes, net.w1_bf16.data_ptr(), None, net.wv.data_ptr(),
                                             net.bv, b["leaf_v"].data_ptr(), b["pass_v"].data_ptr(),
                                             self._side.cuda_stream if self.overlap else None, _stream()),
                  "bg_twoply_replies_values")
        b["leaves"].add_(b["alloc"])
        with torch.cuda.device(dev):
            check(L.bg_twoply_scores(b["leaf_v"].data_ptr(), b["starts"].data_ptr(), b["counts"].data_ptr(),
                                     b["pass_v"].data_ptr(), A.data_ptr(), movers.data_ptr(), M, out.data_ptr(), _stream()),
                  "bg_twoply_scores")

    def score_afterstates(self, after52: torch.Tensor, movers: torch.Tensor) -> torch.Tensor:
        """2-ply score of every root afterstate (movers[i] = the player who made play i)."""
        M = after52.shape[0]
        dev = after52.device
        out = torch.empty(M, dtype=torch.float32, device=dev)
        if M == 0:
            return out
        after52, movers = after52.contiguous(), movers.contiguous()
        while True:
            b = self._workspace(dev)
            b["ws"].status.zero_()
            b["leaves"].zero_()
            n_chunks = -(-M // self.chunk)
            per = -(-M // n_chunks)                               # equal chunks: the fixed latencies of K1's overflow tiers amortise best
            for m0 in range(0, M, per):
                m1 = min(M, m0 + per)
                self._score_chunk(after52[m0:m1], movers[m0:m1], out[m0:m1])
            st = int(b["ws"].status.item())                       # the one synchronisation
            if st & 4:                                            # reply buffer too small: grow and redo (never dropped silently)
                self.rpp *= 2
                continue
            if st:
                raise BgError(f"2-ply reply generation: device status {st}: {_lib.status_message(st)}")
            self.leaves_evaluated += int(b["leaves"].item())
            return out

    def search_unfused(self, boards52: torch.Tensor, players: torch.Tensor, dice: torch.Tensor):
        """The search through the unfused pipeline (replies materialised in HBM, chunked): what bg_twoply is tested
        against.  Same return value as search()."""
        counts, offsets, A, rowp = legal_moves(boards52, players, dice, with_row_players=True)
        scores = self.score_afterstates(A, rowp)
        best, _ = segment_argmax(scores, offsets[:-1].contiguous(), counts.clamp(min=0).contiguous())
        return best, scores, offsets, A

    # ------------------------------------------------------------------ the fused search: ONE C call, no host sync
    def _fused_buffers(self, B, dev, rows_per_root):
        cap = int(rows_per_root) * B + 4096
        f = getattr(self, "_fb", None)
        if f is None or f["dev"] != dev or f["B"] < B or f["cap"] < cap:
            z = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)
            nbytes = int(lib().bg_twoply_workspace_bytes(B, cap))
            f = self._fb = dict(dev=dev, B=B, cap=cap, after=z((cap, 52), torch.int8), scores=z(cap, torch.float32),
                                counts=z(B, torch.int32), starts=z(B, torch.int64), best=z(B, torch.int32),
                                best_score=z(B, torch.float32), stats=torch.zeros(4, dtype=torch.int64, device=dev),
                                status=torch.zeros(1, dtype=torch.int32, device=dev), ws=z(nbytes, torch.uint8), nbytes=nbytes)
        return f

    def search_device(self, boards52: torch.Tensor, players: torch.Tensor, dice: torch.Tensor, rows_per_root: int = 64):
        """bg_twoply: K1 on the roots -> fused on-chip expansion / leaf evaluation / max -> scores -> argmax, one C call,
        nothing synchronised.  -> dict of persistent device buffers: best (B,), best_score (B,), counts (B,), starts (B,)
        (root b's plays are rows starts[b] .. +counts[b] of `after` / `scores`, reference legal_moves order), stats
        (4,) i64 = (root afterstates, leaves evaluated, overflow items, overflow leaves), status (1,) i32 (check with check_status)."""
        boards52 = boards52.reshape(-1, 52).to(torch.int8).contiguous()
        B, dev = boards52.shape[0], boards52.device
        players = players.to(torch.int8).contiguous()
        dice = dice.to(torch.int8).reshape(B, 2).contiguous()
        f = self._fused_buffers(B, dev, rows_per_root)
        f["status"].zero_()
        net = self.net
        with torch.cuda.device(dev):
            check(lib().bg_twoply(boards52.data_ptr(), players.data_ptr(), dice.data_ptr(), B, net.w1_bf16.data_ptr(),
                                  net.wv.data_ptr(), net.bv, f["cap"], f["after"].data_ptr(), f["scores"].data_ptr(),
                                  f["counts"].data_ptr(), f["starts"].data_ptr(), f["best"].data_ptr(),
                                  f["best_score"].data_ptr(), f["stats"].data_ptr(), f["status"].data_ptr(),
                                  f["ws"].data_ptr(), f["nbytes"], _stream()), "bg_twoply")
        f["n_roots"] = B
        return f

    def search(self, boards52: torch.Tensor, players: torch.Tensor, dice: torch.Tensor, rows_per_root: int = 64):
        """-> best (B,) i32 index into each root's legal plays (-1 if none), scores (total,) f32,
        offsets (B+1,) i64, afterstates (total,52) i8 (reference legal_moves order, roots in order).
        The search itself is search_device (one call, no sync); this wrapper reads the status, grows the buffers if the
        roots had more plays than expected, and gathers the slab-ordered rows into CSR order."""
        B = boards52.reshape(-1, 52).shape[0]
        while True:
            f = self.search_device(boards52, players, dice, rows_per_root)
            st = int(f["status"].item())
            if st & 4:                                            # more plays than room: grow and redo (never dropped silently)
                rows_per_root *= 2
                continue
            if st:
                raise BgError(f"2-ply search: device status {st}: {_lib.status_message(st)}")
            break
        n_after, n_leaves, self.overflow_items, self.overflow_leaves = (int(x) for x in f["stats"].tolist())
        self.leaves_evaluated += n_leaves
        counts = f["counts"][:B].to(torch.int64)
        offsets = torch.zeros(B + 1, dtype=torch.int64, device=counts.device)
        torch.cumsum(counts, 0, out=offsets[1:])
        roots = torch.repeat_interleave(torch.arange(B, device=counts.device), counts, output_size=n_after)
        rows = f["starts"][:B][roots] + (torch.arange(n_after, device=counts.device) - offsets[roots])
        return f["best"][:B].clone(), f["scores"][rows], offsets, f["after"][rows]
