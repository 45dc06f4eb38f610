"""ctypes binding of libbg_b200.so (include/bg_b200.h).  No CPU fallback: if the CUDA library is
missing or fails to load, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbg_b200.so")

BG_STATUS = {1: "BAD_INPUT (count outside 0..15 or die outside 1..6)",
             2: "SCRATCH_OVERFLOW (more boards per level than the large scratch holds)",
             4: "OUTPUT_OVERFLOW (afterstate buffer too small)",
             8: "DICE_EXHAUSTED (external dice stream ran out)"}


class BgError(RuntimeError):
    pass


class EnvState(C.Structure):
    _fields_ = [("n_games", C.c_longlong), ("boards52", C.c_void_p), ("players", C.c_void_p), ("dice", C.c_void_p),
                ("scores", C.c_void_p), ("draws", C.c_void_p), ("match_over", C.c_void_p),
                ("afterstates52", C.c_void_p), ("starts", C.c_void_p), ("counts", C.c_void_p),
                ("seed", C.c_ulonglong), ("stream_base", C.c_ulonglong), ("ext_dice", C.c_void_p),
                ("ext_len", C.c_longlong), ("match_length", C.c_int32), ("no_auto_reset", C.c_int32),
                ("game_over", C.c_void_p)]


class StepOut(C.Structure):
    _fields_ = [("rewards", C.c_void_p), ("dones", C.c_void_p), ("info_player", C.c_void_p),
                ("winner", C.c_void_p), ("game_score", C.c_void_p), ("flags", C.c_void_p)]


_V, _LL, _I, _SZ, _U64, _U32, _F = C.c_void_p, C.c_longlong, C.c_int, C.c_size_t, C.c_ulonglong, C.c_uint32, C.c_float

# name -> (restype, argtypes); must list every function include/bg_b200.h declares (tests check this)
SIGNATURES = {
    "bg_last_error": (C.c_char_p, []),
    "bg_version": (_I, []),
    "bg_movegen_workspace_bytes": (_SZ, [_LL]),
    "bg_movegen_count": (_I, [_V, _V, _V, _LL, _V, _V, _V, _SZ, _V]),
    "bg_movegen_write": (_I, [_V, _V, _V, _LL, _V, _I, _V, _LL, _V, _V, _V, _V, _V, _V, _SZ, _V]),
    "bg_movegen_slab": (_I, [_V, _V, _V, _LL, _I, _V, _LL, _V, _V, _V, _V, _V, _V, _V, _V, _SZ, _V]),
    "bg_ppo_encode_block": (_I, [_V, _V, _V, _LL, _I, _V, _V]),
    "bg_record_state": (_I, [_V, _V, _V, _V, _V]),
    "bg_set_team_threads": (_I, [_I, _I]),
    "bg_encode_f32": (_I, [_V, _V, _I, _LL, _V, _V, _LL, _V]),
    "bg_encode_bf16": (_I, [_V, _V, _I, _LL, _V, _V, _LL, _V]),
    "bg_update_legal_plays": (_I, [_V, _V, _V, _LL, _I, _V, _LL, _V, _V, _V, _V, _V, _V, _V, _SZ, _V, _LL, _V, _LL, _V, _V, _V, _V]),
    "bg_env_reset": (_I, [C.POINTER(EnvState), _V, _V, _V]),
    "bg_env_step": (_I, [C.POINTER(EnvState), _V, C.POINTER(StepOut), _V, _V]),
    "bg_random_actions": (_I, [_V, _LL, _U64, _U64, _U32, _V, _V]),
    "bg_env_step_random": (_I, [C.POINTER(EnvState), _U64, _U32, _V, C.POINTER(StepOut), _V, _V]),
    "bg_copy_actions_async": (_I, [_V, _V, _LL, _V]),
    "bg_movegen_replies_slab": (_I, [_V, _V, _LL, _I, _V, _LL, _V, _V, _V, _V, _V, _V, _V, _V, _SZ, _V]),
    "bg_twoply_replies_values": (_I, [_V, _V, _LL, _V, _LL, _V, _V, _V, _V, _V, _V, _SZ, _V, _V, _V, _F, _V, _V, _V, _V]),
    "bg_twoply_scores": (_I, [_V, _V, _V, _V, _V, _V, _LL, _V, _V]),
    "bg_workspace_bytes": (_SZ, [_I, _LL]),
    "bg_twoply_workspace_bytes": (_SZ, [_LL, _LL]),
    "bg_twoply": (_I, [_V, _V, _V, _LL, _V, _V, _F, _LL, _V, _V, _V, _V, _V, _V, _V, _V, _V, _SZ, _V]),
    "bg_segment_argmax": (_I, [_V, _V, _V, _LL, _V, _V, _V]),
    "bg_pack_wa": (_I, [_V, _V, _V]),
    "bg_policy_workspace_bytes": (_SZ, [_LL]),
    "bg_policy_sample": (_I, [_V, _V, _I, _LL, _V, _V, _V, _V, _V, _V, _F, _U64, _U64, _U32, _I, _V, _V, _V, _V, _V, _SZ, _V]),
    "bg_gae": (_I, [_V, _V, _V, _V, _I, _LL, _F, _F, _V, _V, _V]),
    "bg_ppo_loss_grad": (_I, [_V, _I, _LL, _V, _V, _V, _V, _V, _V, _LL, _F, _F, _F, _V, _V, _V, _V, _V]),
    "bg_ppo_pack_weights": (_I, [_V, _V, _V, _V, _V, _V, _V]),
    "bg_ppo_gemm_nt": (_I, [_I, _V, _LL, _LL, _V, _V, _V, _V, _V]),
    "bg_ppo_gemm_tn": (_I, [_I, _V, _V, _LL, _LL, _V, _V, _V]),
    "bg_ppo_loss_grad_classes": (_I, [_V, _V, _V, _V, _LL, _LL, _LL, _V, _V, _V, _V, _V, _F, _F, _F, _V, _V, _I, _V]),
    "bg_ppo_logits_loss_a": (_I, [_V, _LL, _LL, _V, _V, _V, _V, _V, _V, _V, _F, _F, _F, _V, _V, _V, _V]),
    "bg_ppo_gather_block": (_I, [_V, _LL, _V, _LL, _I, _I, _V, _V]),
    "bg_ppo_gemm_debug": (_I, [_I]),
    "bg_adam_step": (_I, [_V, _V, _V, _V, _LL, _F, _F, _F, _F, _I, _F, _V]),
    "bg_pack_w1": (_I, [_V, _V, _V, _V]),
    "bg_mlp_value": (_I, [_V, _V, _I, _I, _LL, _V, _V, _V, _V, _F, _I, _V, _V]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BgError(f"{LIB_PATH} is missing: build it with `python mlp-ppo-2ply-p3_b200/build.py` "
                          "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        raise BgError(f"{what} failed ({rc}): {lib().bg_last_error().decode()}")


def status_message(status: int) -> str:
    return "; ".join(msg for bit, msg in BG_STATUS.items() if status & bit)
