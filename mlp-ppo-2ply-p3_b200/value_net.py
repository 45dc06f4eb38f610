"""ValueNet -- the leaf evaluator: value head of the reference's BackgammonPolicyNetwork
(src/agent/policy_network.py:44-75) running as K4 (tcgen05 bf16 GEMM fused with the encoder).

Weights load from the reference's own state_dict keys (fc1.weight (128,198), fc1.bias (128),
value_head.weight (1,128), value_head.bias (1)); action_head.* is ignored here (policy head = "next" row N1).
"""
from __future__ import annotations

import torch

from ._lib import BgError, check, lib
from .engine import _stream

HIDDEN = 128
FEATURES = 198
LD = 208


class ValueNet:
    def __init__(self, fc1_weight, fc1_bias, value_weight, value_bias, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise BgError("ValueNet needs a CUDA device (there is no CPU fallback)")
        w = torch.as_tensor(fc1_weight, dtype=torch.float32).to(device).contiguous()
        if tuple(w.shape) != (HIDDEN, FEATURES):
            raise BgError(f"fc1.weight must be ({HIDDEN},{FEATURES})")
        self.device = device
        self.fc1_weight_f32 = w
        self.b1 = torch.as_tensor(fc1_bias, dtype=torch.float32).to(device).reshape(HIDDEN).contiguous()
        self.wv = torch.as_tensor(value_weight, dtype=torch.float32).to(device).reshape(HIDDEN).contiguous()
        self.bv = float(torch.as_tensor(value_bias).reshape(-1)[0])
        self.w1_bf16 = torch.empty((HIDDEN, LD), dtype=torch.bfloat16, device=device)
        with torch.cuda.device(device):
            check(lib().bg_pack_w1(w.data_ptr(), self.b1.data_ptr(), self.w1_bf16.data_ptr(), _stream()), "bg_pack_w1")   # bias folded into columns 198/199

    @classmethod
    def from_state_dict(cls, sd, device):
        """sd: the reference checkpoint (ppo_agent.py:377-403 saves policy_network.state_dict())."""
        return cls(sd["fc1.weight"], sd["fc1.bias"], sd["value_head.weight"], sd["value_head.bias"], device)

    @classmethod
    def random_init(cls, device, seed=0):
        """torch.manual_seed(seed) + nn.Linear default init, as BackgammonPolicyNetwork(198,128,500) does."""
        g = torch.Generator().manual_seed(seed)

        def lin(out_f, in_f):
            k = 1.0 / in_f ** 0.5
            return (torch.rand((out_f, in_f), generator=g) * 2 - 1) * k, (torch.rand((out_f,), generator=g) * 2 - 1) * k
        w1, b1 = lin(HIDDEN, FEATURES)
        lin(500, HIDDEN)
        wv, bv = lin(1, HIDDEN)
        return cls(w1, b1, wv, bv, device)

    def values(self, boards52: torch.Tensor, flags, flip_flags: bool = False, terminal_aware: bool = False,
               n_rows_dev: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
        """V(encode(board, flag)) for B positions -> (B,) f32.  flags: int or (B,) int8 tensor; flip_flags
        evaluates with the other player's flag; terminal_aware returns the win reward for rows whose flag
        player has borne off all 15 men."""
        if not boards52.is_cuda:
            raise BgError("ValueNet.values needs CUDA tensors")
        boards52 = boards52.reshape(-1, 52)
        if not boards52.is_contiguous():
            boards52 = boards52.contiguous()
        B = boards52.shape[0]
        if isinstance(flags, torch.Tensor):
            fl = flags.to(torch.int8).contiguous()
            fptr, fall = fl.data_ptr(), 0
        else:
            fl, fptr, fall = None, None, int(flags)
        if out is None:
            out = torch.empty(B, dtype=torch.float32, device=boards52.device)
        with torch.cuda.device(boards52.device):
            check(lib().bg_mlp_value(boards52.data_ptr(), fptr, fall, int(flip_flags), B,
                                     n_rows_dev.data_ptr() if n_rows_dev is not None else None,
                                     self.w1_bf16.data_ptr(), None, self.wv.data_ptr(), self.bv,
                                     int(terminal_aware), out.data_ptr(), _stream()), "bg_mlp_value")
        return out
