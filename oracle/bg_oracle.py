"""ctypes binding of the CPU oracle (oracle/bg_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of bg_oracle.c.  Importable from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never from the product package.

All boards here use the reference layout: int8 (4, 24)
(/root/reference/src/board/immutable_board.py:20-27).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbg_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "bg_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-std=c11", "-shared", "-o", _SO, src, "-lm"], cwd=_HERE
        )
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        i8p, i32p, i64p, f32p = (C.POINTER(C.c_int8), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_float))
        L.bg_legal_moves.restype = C.c_int
        L.bg_legal_moves.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.bg_legal_moves_batch.restype = None
        L.bg_legal_moves_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bg_encode_batch.restype = None
        L.bg_encode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.bg_win_score.restype = C.c_int
        L.bg_win_score.argtypes = [C.c_void_p, C.c_int, f32p]
        L.bg_philox_dice.restype = None
        L.bg_philox_dice.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
        L.bg_philox_action.restype = C.c_uint32
        L.bg_philox_action.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        L.bg_env_new.restype = C.c_void_p
        L.bg_env_new.argtypes = [C.c_int, C.c_int]
        L.bg_env_free.restype = None
        L.bg_env_free.argtypes = [C.c_void_p]
        L.bg_env_set_external_dice.restype = None
        L.bg_env_set_external_dice.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.bg_env_set_philox.restype = None
        L.bg_env_set_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.bg_env_dice_consumed.restype = C.c_int64
        L.bg_env_dice_consumed.argtypes = [C.c_void_p]
        L.bg_env_reset.restype = None
        L.bg_env_reset.argtypes = [C.c_void_p]
        L.bg_env_step.restype = C.c_int
        L.bg_env_step.argtypes = [C.c_void_p, C.c_int, f32p, i32p, i32p, i32p, i32p]
        L.bg_env_get.restype = None
        L.bg_env_get.argtypes = [C.c_void_p, C.c_void_p, i32p, C.c_void_p, i32p, i32p, C.c_void_p, i32p, i32p]
        L.bg_env_set_position.restype = None
        L.bg_env_set_position.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.bg_env_get_afterstates.restype = None
        L.bg_env_get_afterstates.argtypes = [C.c_void_p, C.c_void_p]
        L.bg_env_observation.restype = None
        L.bg_env_observation.argtypes = [C.c_void_p, C.c_void_p]
        L.bg_env_random_rollout.restype = C.c_int64
        L.bg_env_random_rollout.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_int, C.POINTER(C.c_double), i64p, i64p]
        L.bg_mlp_value_batch.restype = None
        L.bg_mlp_value_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p]
        L.bg_mlp_value_bf16.restype = C.c_float
        L.bg_mlp_value_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int]
        L.bg_twoply.restype = C.c_int
        L.bg_twoply.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                C.c_int, C.c_void_p, C.c_int, i32p, i64p, C.c_int]
        L.bg_pack52.restype = None
        L.bg_pack52.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.bg_unpack52.restype = None
        L.bg_unpack52.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def initial_board() -> np.ndarray:
    """immutable_board.py:25-40"""
    b = np.zeros((4, 24), np.int8)
    b[0, 0], b[0, 11], b[0, 16], b[0, 18] = 2, 5, 3, 5
    b[1, 23], b[1, 12], b[1, 7], b[1, 5] = 2, 5, 3, 5
    return b


def legal_moves(board: np.ndarray, player: int, d0: int, d1: int, with_moves: bool = False):
    """get_all_possible_moves (moves/get_all_moves.py:9-70): afterstates (n,4,24) int8 in reference list order."""
    board = np.ascontiguousarray(board, np.int8).reshape(4, 24)
    n = lib().bg_legal_moves(_p(board), int(player), int(d0), int(d1), None, None, None, 0)
    after = np.zeros((n, 4, 24), np.int8)
    nsub = np.zeros((n,), np.int32)
    subs = np.zeros((n, 4, 3), np.int8)
    if n:
        lib().bg_legal_moves(_p(board), int(player), int(d0), int(d1), _p(after), _p(nsub), _p(subs), n)
    if with_moves:
        return after, nsub, subs
    return after


def legal_moves_batch(boards: np.ndarray, players: np.ndarray, dice: np.ndarray):
    """-> counts (B,) i32, offsets (B+1,) i64, afterstates (total,4,24) i8 (reference order per board)."""
    boards = np.ascontiguousarray(boards, np.int8).reshape(-1, 4, 24)
    B = boards.shape[0]
    players = np.ascontiguousarray(players, np.int8)
    dice = np.ascontiguousarray(dice, np.int8).reshape(B, 2)
    counts = np.zeros((B,), np.int32)
    lib().bg_legal_moves_batch(_p(boards), _p(players), _p(dice), B, _p(counts), None, None)
    offsets = np.zeros((B + 1,), np.int64)
    np.cumsum(counts, out=offsets[1:])
    after = np.zeros((int(offsets[-1]), 4, 24), np.int8)
    if offsets[-1]:
        lib().bg_legal_moves_batch(_p(boards), _p(players), _p(dice), B, _p(counts), _p(offsets), _p(after))
    return counts, offsets, after


def encode(boards: np.ndarray, flags) -> np.ndarray:
    """get_board_features_batch_from_tensors (ai/batching.py:78-147): (B,198) f32."""
    boards = np.ascontiguousarray(boards, np.int8).reshape(-1, 4, 24)
    B = boards.shape[0]
    flags = np.ascontiguousarray(np.broadcast_to(np.asarray(flags, np.int8), (B,)))
    out = np.zeros((B, 198), np.float32)
    lib().bg_encode_batch(_p(boards), _p(flags), B, _p(out))
    return out


def win_score(board: np.ndarray, player: int):
    board = np.ascontiguousarray(board, np.int8)
    r = C.c_float(0)
    s = lib().bg_win_score(_p(board), int(player), C.byref(r))
    return s, float(r.value)


def philox_dice(seed: int, stream: int, draw: int):
    d = np.zeros(2, np.int8)
    lib().bg_philox_dice(seed, stream, draw, _p(d))
    return int(d[0]), int(d[1])


def philox_action(seed: int, stream: int, draw: int, n: int) -> int:
    return int(lib().bg_philox_action(seed, stream, draw, n))


def pack52(boards: np.ndarray) -> np.ndarray:
    boards = np.ascontiguousarray(boards, np.int8).reshape(-1, 4, 24)
    out = np.zeros((boards.shape[0], 52), np.int8)
    lib().bg_pack52(_p(boards), boards.shape[0], _p(out))
    return out


def unpack52(b52: np.ndarray) -> np.ndarray:
    b52 = np.ascontiguousarray(b52, np.int8).reshape(-1, 52)
    out = np.zeros((b52.shape[0], 4, 24), np.int8)
    lib().bg_unpack52(_p(b52), b52.shape[0], _p(out))
    return out


class Env:
    """BackgammonEnv (environment/backgammon_env.py) with injectable dice."""

    def __init__(self, match_length=15, max_legal_moves=500):
        self._e = lib().bg_env_new(match_length, max_legal_moves)
        self._dice = None

    def __del__(self):
        try:
            lib().bg_env_free(self._e)
        except Exception:
            pass

    def set_external_dice(self, dice: np.ndarray):
        self._dice = np.ascontiguousarray(dice, np.int8).reshape(-1, 2)
        lib().bg_env_set_external_dice(self._e, _p(self._dice), self._dice.shape[0])

    def set_philox(self, seed: int, stream: int):
        lib().bg_env_set_philox(self._e, seed, stream)

    def dice_consumed(self) -> int:
        return int(lib().bg_env_dice_consumed(self._e))

    def reset(self):
        lib().bg_env_reset(self._e)
        return self.observation()

    def step(self, action: int):
        r, ip, fl, w, gs = C.c_float(0), C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int32(0)
        done = lib().bg_env_step(self._e, int(action), C.byref(r), C.byref(ip), C.byref(fl), C.byref(w), C.byref(gs))
        info = {"current_player": ip.value, "passed": bool(fl.value & 1), "invalid": bool(fl.value & 2),
                "winner": w.value, "game_score": gs.value}
        return float(r.value), bool(done), info

    def state(self):
        b = np.zeros((4, 24), np.int8)
        roll = np.zeros(2, np.int32)
        sc = np.zeros(2, np.int32)
        pl, n, nt, go, mo = (C.c_int32(0) for _ in range(5))
        lib().bg_env_get(self._e, _p(b), C.byref(pl), _p(roll), C.byref(n), C.byref(nt), _p(sc), C.byref(go), C.byref(mo))
        return {"board": b, "player": pl.value, "roll": roll, "n_legal": n.value, "n_legal_true": nt.value,
                "scores": sc, "game_over": bool(go.value), "match_over": bool(mo.value)}

    def set_position(self, board, player, d0, d1):
        board = np.ascontiguousarray(board, np.int8)
        lib().bg_env_set_position(self._e, _p(board), int(player), int(d0), int(d1))

    def afterstates(self) -> np.ndarray:
        n = self.state()["n_legal"]
        a = np.zeros((n, 4, 24), np.int8)
        if n:
            lib().bg_env_get_afterstates(self._e, _p(a))
        return a

    def observation(self) -> np.ndarray:
        f = np.zeros(198, np.float32)
        lib().bg_env_observation(self._e, _p(f))
        return f

    def random_rollout(self, steps: int, act_seed: int, encode: bool = True):
        cs, gd, ps = C.c_double(0), C.c_int64(0), C.c_int64(0)
        n = lib().bg_env_random_rollout(self._e, steps, act_seed, int(encode), C.byref(cs), C.byref(gd), C.byref(ps))
        return int(n), float(cs.value), int(gd.value), int(ps.value)


def mlp_value(x: np.ndarray, W1, b1, wv, bv: float) -> np.ndarray:
    """BackgammonPolicyNetwork value head (agent/policy_network.py:58-75), f32."""
    x = np.ascontiguousarray(x, np.float32).reshape(-1, 198)
    W1 = np.ascontiguousarray(W1, np.float32)
    b1 = np.ascontiguousarray(b1, np.float32)
    wv = np.ascontiguousarray(wv, np.float32).reshape(-1)
    out = np.zeros((x.shape[0],), np.float32)
    lib().bg_mlp_value_batch(_p(x), x.shape[0], _p(W1), _p(b1), _p(wv), float(bv), W1.shape[0], _p(out))
    return out


def mlp_value_bf16(x: np.ndarray, W1, b1, wv, bv: float) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32).reshape(-1, 198)
    W1 = np.ascontiguousarray(W1, np.float32)
    b1 = np.ascontiguousarray(b1, np.float32)
    wv = np.ascontiguousarray(wv, np.float32).reshape(-1)
    return np.array([lib().bg_mlp_value_bf16(_p(x[i]), _p(W1), _p(b1), _p(wv), float(bv), W1.shape[0])
                     for i in range(x.shape[0])], np.float32)


def twoply(board, me, d0, d1, W1, b1, wv, bv, use_bf16=False):
    """2-ply per SURVEY.md 8(c): -> scores (n,) f32 in reference play order, best index, leaf count."""
    board = np.ascontiguousarray(board, np.int8)
    W1 = np.ascontiguousarray(W1, np.float32)
    b1 = np.ascontiguousarray(b1, np.float32)
    wv = np.ascontiguousarray(wv, np.float32).reshape(-1)
    n = lib().bg_legal_moves(_p(board), int(me), int(d0), int(d1), None, None, None, 0)
    scores = np.zeros((max(n, 1),), np.float32)
    best, leaves = C.c_int32(-1), C.c_int64(0)
    lib().bg_twoply(_p(board), int(me), int(d0), int(d1), _p(W1), _p(b1), _p(wv), float(bv), W1.shape[0],
                    _p(scores), n, C.byref(best), C.byref(leaves), int(use_bf16))
    return scores[:n], int(best.value), int(leaves.value)
