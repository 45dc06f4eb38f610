"""Function-level drop-ins over the C ABI: batched legal-move generation and feature encoding on
torch CUDA tensors.  Mirrors the reference's free functions

    get_all_possible_moves(player, board, roll)              src/moves/get_all_moves.py:9-70
    get_board_features_batch_from_tensors(boards, player)    src/ai/batching.py:78-147

for a whole batch at once.  Positions use the packed `board52` layout (see include/bg_b200.h);
`to_board52` / `from_board52` convert from / to the reference's (4, 24) int8 tensors.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import BgError, check, lib

FEATURES = 198
LD_BF16 = 208


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise BgError("bg_b200 kernels need CUDA tensors (there is no CPU fallback)")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def to_board52(boards_4x24: torch.Tensor) -> torch.Tensor:
    """(B,4,24) int8 reference layout (board/immutable_board.py:20-27) -> (B,52) int8."""
    b = boards_4x24.reshape(-1, 4, 24).to(torch.int8)
    return torch.cat([b[:, 0], b[:, 1], b[:, 2, 0:2], b[:, 3, 0:2]], dim=1).contiguous()


def as_board52(boards: torch.Tensor) -> torch.Tensor:
    """Packed (B,52) rows from either layout: the reference's (B,4,24) / (4,24) tensors are converted, (B,52) / (52,)
    pass through."""
    if boards.dim() >= 2 and tuple(boards.shape[-2:]) == (4, 24):
        return to_board52(boards)
    return boards.reshape(-1, 52).to(torch.int8).contiguous()


def from_board52(b52: torch.Tensor) -> torch.Tensor:
    b52 = b52.reshape(-1, 52)
    out = torch.zeros((b52.shape[0], 4, 24), dtype=torch.int8, device=b52.device)
    out[:, 0] = b52[:, 0:24]
    out[:, 1] = b52[:, 24:48]
    out[:, 2, 0:2] = b52[:, 48:50]
    out[:, 3, 0:2] = b52[:, 50:52]
    return out


def initial_board52(n: int, device) -> torch.Tensor:
    b = torch.zeros((n, 52), dtype=torch.int8, device=device)
    for p, c in ((0, 2), (11, 5), (16, 3), (18, 5), (24 + 23, 2), (24 + 12, 5), (24 + 7, 3), (24 + 5, 5)):
        b[:, p] = c
    return b


def render_board52(b52, tokens=("X", "O")) -> str:
    """ASCII dump of one position in the layout of BackgammonEnv.render (environment/backgammon_env.py:253-355): points
    12..23 on top (PLAYER2's home board 18..23 on the right), 11..0 below.  The reference indexes the bar / borne-off
    men as `tensor[player, 24]` / `[player, 25]`, columns that do not exist in its (4,24) tensor (so its render
    raises); here they are read from rows 2 and 3 where the board keeps them (board/immutable_board.py:20-27)."""
    b = [int(v) for v in (b52.tolist() if hasattr(b52, "tolist") else b52)]
    pts, col = [], []
    for i in range(24):
        c1, c2 = b[i], b[24 + i]
        if c1 > 0 and c2 > 0:
            pts.append(0); col.append("?")
        elif c1 > 0:
            pts.append(c1); col.append(tokens[0])
        elif c2 > 0:
            pts.append(c2); col.append(tokens[1])
        else:
            pts.append(0); col.append(" ")

    def half(points, colors, player):
        bar, off = b[48 + player], b[50 + player]
        lines = []
        for i in range(max([0] + points + [bar, off])):
            row = [c if n > i else " " for n, c in zip(points, colors)]
            lines.append("|  " + " | ".join(f"{r:^3}" for r in row[:6]) + f" | {(tokens[player] if bar > i else ' '):^3} | "
                         + " | ".join(f"{r:^3}" for r in row[6:]) + f" | {(tokens[player] if off > i else ' '):^3} |")
        return lines
    out = ["| 12 | 13 | 14 | 15 | 16 | 17 | BAR | 18 | 19 | 20 | 21 | 22 | 23 | OFF |",
           f"|------------Outer Board-------------|     |-----------P={tokens[1]} Home Board----------|     |"]
    out += half(pts[12:], col[12:], 1)
    out.append("|------------------------------------|     |-----------------------------------|     |")
    out += half(pts[:12][::-1], col[:12][::-1], 0)
    out += [f"|------------Outer Board-------------|     |-----------P={tokens[0]} Home Board----------|     |",
            "| 11 | 10 | 9  | 8  | 7  | 6  | BAR | 5  | 4  | 3  | 2  | 1  | 0  | OFF |", ""]
    return "\n".join(out)


def read_status(status: torch.Tensor, what: str = "kernel"):
    s = int(status.item())
    if s:
        raise BgError(f"{what}: device status {s}: {_lib.status_message(s)}")


class MovegenWorkspace:
    """Reusable device scratch for K1 (no allocation inside the C ABI)."""

    def __init__(self, max_batch: int, device):
        self.max_batch = int(max_batch)
        self.nbytes = int(lib().bg_movegen_workspace_bytes(self.max_batch))
        self.buf = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)


def legal_moves(boards52: torch.Tensor, players: torch.Tensor, dice: torch.Tensor, max_rows_per_board: int = 0,
                workspace: MovegenWorkspace | None = None, with_row_players: bool = False, check_status: bool = True):
    """All legal plays of B positions -> (counts_true (B,) i32, offsets (B+1,) i64, afterstates (total,52) i8).

    Deterministic two-pass CSR form (bg_movegen_count -> exclusive scan -> bg_movegen_write); rows of a
    position are in the reference's legal_moves order.  max_rows_per_board > 0 keeps only the first
    that many rows per position (the env's max_legal_moves truncation, backgammon_env.py:218-223);
    counts_true is never truncated.
    """
    _require_cuda(boards52, players, dice)
    boards52 = as_board52(boards52)                       # the reference's (B,4,24) layout is accepted too
    B = boards52.shape[0]
    players = players.to(torch.int8).contiguous()
    dice = dice.to(torch.int8).reshape(B, 2).contiguous()
    dev = boards52.device
    ws = workspace or MovegenWorkspace(max(B, 1), dev)
    if ws.max_batch < B:
        raise BgError("workspace too small for this batch")
    counts_true = torch.empty(B, dtype=torch.int32, device=dev)
    L = lib()
    with torch.cuda.device(dev):
        check(L.bg_movegen_count(boards52.data_ptr(), players.data_ptr(), dice.data_ptr(), B, counts_true.data_ptr(),
                                 ws.status.data_ptr(), ws.buf.data_ptr(), ws.nbytes, _stream()), "bg_movegen_count")
        kept = counts_true.clamp(min=0)
        if max_rows_per_board > 0:
            kept = kept.clamp(max=max_rows_per_board)
        offsets = torch.zeros(B + 1, dtype=torch.int64, device=dev)
        torch.cumsum(kept, 0, out=offsets[1:])
        total = int(offsets[-1].item())
        after = torch.empty((max(total, 1), 52), dtype=torch.int8, device=dev)
        rowp = torch.empty(max(total, 1), dtype=torch.int8, device=dev) if with_row_players else None
        check(L.bg_movegen_write(boards52.data_ptr(), players.data_ptr(), dice.data_ptr(), B, offsets.data_ptr(),
                                 int(max_rows_per_board), after.data_ptr(), total,
                                 rowp.data_ptr() if rowp is not None else None, None, None, None,
                                 ws.status.data_ptr(), ws.buf.data_ptr(), ws.nbytes, _stream()), "bg_movegen_write")
    if check_status:
        read_status(ws.status, "legal_moves")
    after = after[:total]
    if with_row_players:
        return counts_true, offsets, after, rowp[:total]
    return counts_true, offsets, after


def encode(boards52: torch.Tensor, flags, dtype=torch.float32, ld: int | None = None, out: torch.Tensor | None = None,
           n_rows_dev: torch.Tensor | None = None) -> torch.Tensor:
    """198-feature encoding of B positions.  flags: int (0/1 for all rows) or (B,) int8 tensor = the player
    whose turn flag is set.  f32 -> (B,198) exactly the reference's tensor; bf16 -> (B, ld) with ld=208 by
    default, columns 198.. zero (the MLP's K padding).  `out` reuses a buffer; `n_rows_dev` (1-element int64
    CUDA tensor) bounds the rows on the device, e.g. the env's alloc_rows counter."""
    _require_cuda(boards52)
    boards52 = as_board52(boards52)                       # the reference's (B,4,24) layout is accepted too
    B = boards52.shape[0]
    dev = boards52.device
    if isinstance(flags, torch.Tensor):
        fl = flags.to(torch.int8).contiguous()
        fptr, fall = fl.data_ptr(), 0
    else:
        fl, fptr, fall = None, None, int(flags)
    L = lib()
    nptr = n_rows_dev.data_ptr() if n_rows_dev is not None else None
    if out is not None and (not out.is_contiguous() or out.shape[0] < B):
        raise BgError("encode: `out` must be contiguous with at least B rows")
    if out is not None:
        ld = out.shape[1]
    with torch.cuda.device(dev):
        if dtype == torch.float32:
            ld = ld or FEATURES
            if out is None:
                out = torch.empty((B, ld), dtype=torch.float32, device=dev)
            check(L.bg_encode_f32(boards52.data_ptr(), fptr, fall, B, nptr, out.data_ptr(), ld, _stream()), "bg_encode_f32")
        elif dtype == torch.bfloat16:
            ld = ld or LD_BF16
            if out is None:
                out = torch.empty((B, ld), dtype=torch.bfloat16, device=dev)
            check(L.bg_encode_bf16(boards52.data_ptr(), fptr, fall, B, nptr, out.data_ptr(), ld, _stream()), "bg_encode_bf16")
        else:
            raise BgError("encode: dtype must be float32 or bfloat16")
    return out
