"""PolicyValueNet -- the reference's BackgammonPolicyNetwork (src/agent/policy_network.py:44-75: fc1 198->128 + ReLU,
action_head 128->500, value_head 128->1) with the rollout-side forward of BackgammonPPOAgent.select_action
(src/agent/ppo_agent.py:138-191) running as ONE kernel (csrc/policy.cu): feature encoding, both GEMMs on the
tcgen05 tensor cores, prefix action mask, softmax, sampling, log-prob and value.

Master weights are f32 torch tensors under the reference's state_dict keys (fc1.*, action_head.*, value_head.*),
so reference checkpoints load and save unchanged (ppo_agent.py:377-403); `sync()` re-packs the bf16 operand tiles
after an optimiser step.
"""
from __future__ import annotations

import torch

from ._lib import BgError, check, lib
from .engine import _stream
from .value_net import FEATURES, HIDDEN, LD, ValueNet

ACTIONS = 500
ACT_PAD = 512
KEYS = ("fc1.weight", "fc1.bias", "action_head.weight", "action_head.bias", "value_head.weight", "value_head.bias")


class PolicyValueNet(ValueNet):
    def __init__(self, state_dict, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise BgError("PolicyValueNet needs a CUDA device (there is no CPU fallback)")
        sd = {k: torch.as_tensor(state_dict[k], dtype=torch.float32).detach().to(device).contiguous().clone() for k in KEYS}
        if tuple(sd["action_head.weight"].shape) != (ACTIONS, HIDDEN) or tuple(sd["fc1.weight"].shape) != (HIDDEN, FEATURES):
            raise BgError("expected the reference architecture 198-128-{500,1}")
        self.params = sd                                        # f32 master weights (the learner updates these in place)
        self.device = device
        self.w1_bf16 = torch.empty((HIDDEN, LD), dtype=torch.bfloat16, device=device)
        self.wa_bf16 = torch.empty((ACT_PAD, HIDDEN), dtype=torch.bfloat16, device=device)
        self.sync()

    # ------------------------------------------------------------------ weights
    def sync(self):
        """Re-pack the bf16 tensor-core operands from the f32 master weights (call after an optimiser step)."""
        p = self.params
        self.fc1_weight_f32 = p["fc1.weight"]
        self.b1 = p["fc1.bias"]
        self.ba = p["action_head.bias"]
        self.wv = p["value_head.weight"].reshape(HIDDEN)
        self.bv = float(p["value_head.bias"].reshape(-1)[0])    # (one small D2H read per sync)
        with torch.cuda.device(self.device):
            check(lib().bg_pack_w1(p["fc1.weight"].data_ptr(), p["fc1.bias"].data_ptr(), self.w1_bf16.data_ptr(), _stream()), "bg_pack_w1")
            check(lib().bg_pack_wa(p["action_head.weight"].data_ptr(), self.wa_bf16.data_ptr(), _stream()), "bg_pack_wa")

    @classmethod
    def from_state_dict(cls, sd, device):
        return cls(sd, device)

    def state_dict(self):
        """The reference's checkpoint format: torch.save(net.state_dict(), path) loads into BackgammonPolicyNetwork."""
        return {k: v.detach().clone().cpu() for k, v in self.params.items()}

    @classmethod
    def random_init(cls, device, seed=0):
        """nn.Linear default init in the reference's construction order (fc1, action_head, value_head)."""
        g = torch.Generator().manual_seed(seed)

        def lin(out_f, in_f):
            k = 1.0 / in_f ** 0.5
            return (torch.rand((out_f, in_f), generator=g) * 2 - 1) * k, (torch.rand((out_f,), generator=g) * 2 - 1) * k
        w1, b1 = lin(HIDDEN, FEATURES)
        wa, ba = lin(ACTIONS, HIDDEN)
        wv, bv = lin(1, HIDDEN)
        return cls({"fc1.weight": w1, "fc1.bias": b1, "action_head.weight": wa, "action_head.bias": ba,
                    "value_head.weight": wv, "value_head.bias": bv}, device)

    # ------------------------------------------------------------------ rollout forward
    def act(self, boards52: torch.Tensor, flags, legal_counts: torch.Tensor | None, seed: int = 0, stream_base: int = 0,
            step: int = 0, greedy: bool = False, want_logits: bool = False, out=None):
        """select_action for B positions -> (actions (B,) i32, log_probs (B,) f32, values (B,) f32[, logits (B,500)]).
        flags: int or (B,) int8 (player to move); legal_counts: (B,) i32 number of legal slots (prefix mask) or None."""
        if not boards52.is_cuda:
            raise BgError("PolicyValueNet.act needs CUDA tensors")
        boards52 = boards52.reshape(-1, 52)
        if not boards52.is_contiguous():
            boards52 = boards52.contiguous()
        B, dev = boards52.shape[0], boards52.device
        if isinstance(flags, torch.Tensor):
            fl = flags.to(torch.int8).contiguous()
            fptr, fall = fl.data_ptr(), 0
        else:
            fl, fptr, fall = None, None, int(flags)
        cptr = None
        if legal_counts is not None:
            legal_counts = legal_counts.to(torch.int32).contiguous()
            cptr = legal_counts.data_ptr()
        if out is None:
            out = (torch.empty(B, dtype=torch.int32, device=dev), torch.empty(B, dtype=torch.float32, device=dev),
                   torch.empty(B, dtype=torch.float32, device=dev))
        actions, logp, values = out
        logits = torch.empty((B, ACTIONS), dtype=torch.float32, device=dev) if want_logits else None
        ws = getattr(self, "_ws", None)                           # row-class lists (persistent scratch)
        need_of = self.__dict__.setdefault("_ws_need", {})        # the C ABI's own requirement (row lists + class B partials), per batch size
        need = need_of.get(B)
        if need is None:
            need = need_of[B] = int(lib().bg_policy_workspace_bytes(max(B, 1)))
        if ws is None or ws.numel() < need or ws.device != dev:
            ws = self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib().bg_policy_sample(boards52.data_ptr(), fptr, fall, B, cptr, self.w1_bf16.data_ptr(), None,
                                         self.wa_bf16.data_ptr(), self.ba.data_ptr(), self.wv.data_ptr(), self.bv,
                                         int(seed), int(stream_base), int(step), int(greedy), actions.data_ptr(),
                                         logp.data_ptr(), values.data_ptr(), logits.data_ptr() if want_logits else None,
                                         ws.data_ptr(), ws.numel(), _stream()), "bg_policy_sample")
        return (actions, logp, values, logits) if want_logits else (actions, logp, values)
