"""PPO self-play on the device: rollout buffer, returns/GAE, the clipped-surrogate update and the data-parallel
trainer (BASELINE config 5).

Mirrors the reference's learner:
    BackgammonPPOAgent.select_action   src/agent/ppo_agent.py:138-191   -> PolicyValueNet.act (one fused kernel)
    per-sample python `memory` dicts   src/agent/ppo_agent.py:175-186   -> RolloutBuffer, [T][N] device tensors
    compute_returns                    src/agent/ppo_agent.py:206-216   -> bg_gae (per game; lambda = 1 = the reference)
    update (4 full-batch epochs)       src/agent/ppo_agent.py:218-366   -> ppo_loss + PPOLearner.update
    train_agent                        src/agent/train.py:30-123        -> PPOTrainer.train
    hyper-parameters                   src/agent/config.py:4-22         -> PPOConfig defaults

Games shard across GPUs by game id (no data-path collective); the ONLY collective is the gradient all-reduce of
one flat 90,101-element f32 bucket per optimiser step (NCCL over NVLink), plus three scalars for the return
normalisation.  The rollout side (env step, legal moves, policy forward, sampling, returns) is hand-written CUDA;
the update is ManualUpdate: five library GEMMs on bf16 operands (what the reference's `torch.amp.autocast`,
ppo_agent.py:269, makes of its linear layers) around one kernel of ours (bg_ppo_loss_grad) for the loss and its
gradient; `manual_backward=False` keeps torch autograd over the same math.

Documented divergences from the reference: returns are computed per game, not over the interleaved
memory of all envs as one sequence (`returns_mode="interleaved"` reproduces that for parity tests); games in flight
are kept across updates unless `reset_each_update=True` (train.py:40 resets them); no GradScaler (bf16 needs none).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
import os

import torch
import torch.nn.functional as F

from ._lib import BgError, StepOut, check, lib
from .engine import _stream, encode
from .policy_net import ACTIONS, KEYS, PolicyValueNet

MASK_LOG = -103.27893          # log(1e-45) in f32: the reference's log(mask + 1e-45) for an illegal slot


@dataclasses.dataclass
class PPOConfig:                                   # src/agent/config.py:4-22
    t_horizon: int = 512
    num_epochs: int = 4
    learning_rate: float = 1e-3
    gamma: float = 0.99
    lam: float = 1.0                               # 1.0 = Monte-Carlo returns as the reference (no GAE there)
    bootstrap: bool = False                        # reference: R = 0 at the end of the memory
    eps_clip: float = 0.25
    value_loss_coef: float = 0.5
    entropy_coef_start: float = 0.15
    entropy_coef_end: float = 0.01
    entropy_anneal_episodes: int = 400_000
    num_minibatches: int = 1                       # reference: full batch
    returns_mode: str = "per_game"                 # or "interleaved" (the reference's walk, for parity tests)
    reset_each_update: bool = False                # reference: True (train.py:40)
    autocast: bool = True
    manual_backward: bool = True                   # CUDA + autocast: explicit GEMMs + bg_ppo_loss_grad instead of torch autograd
    update_impl: str = "tcgen05"                   # the explicit path's GEMMs: "tcgen05" = our kernels (TensorCoreUpdate, with our Adam
                                                   # kernel), "cublas" = library GEMMs (ManualUpdate, torch.optim.Adam): kept for comparison


# --------------------------------------------------------------------------------------------- pure torch math

def policy_value_forward(params, x):
    """BackgammonPolicyNetwork.forward (policy_network.py:58-75) on features x (B,198+) -> logits (B,500), values (B,)"""
    w1 = params["fc1.weight"]
    if x.shape[1] > w1.shape[1]:                   # K3's bf16 rows are padded to 208 columns (zeros): pad the weight, not a copy of x
        w1 = F.pad(w1, (0, x.shape[1] - w1.shape[1]))
    h = F.relu(F.linear(x, w1, params["fc1.bias"]))
    logits = F.linear(h, params["action_head.weight"], params["action_head.bias"])
    values = F.linear(h, params["value_head.weight"], params["value_head.bias"]).squeeze(-1)
    return logits, values


class _FusedPPOLoss(torch.autograd.Function):
    """bg_ppo_loss_grad: the masked-softmax / clipped-surrogate / MSE / entropy loss and its gradient w.r.t. the logits
    and values in one pass over the logits (csrc/ppo.cu)."""

    @staticmethod
    def forward(ctx, logits, values, counts, actions, old_logp, adv, returns, eps_clip, value_coef, entropy_coef):
        if logits.dtype not in (torch.bfloat16, torch.float32):
            logits = logits.float()
        logits = logits.contiguous()
        B = logits.shape[0]
        dev = logits.device
        v32 = values.float().contiguous()
        dlogits, dvalues = torch.empty_like(logits), torch.empty(B, dtype=torch.float32, device=dev)
        sums = torch.zeros(3, dtype=torch.float32, device=dev)
        f = lambda t, dt: t.to(dt).contiguous()
        with torch.cuda.device(dev):
            check(lib().bg_ppo_loss_grad(logits.data_ptr(), int(logits.dtype == torch.bfloat16), logits.shape[1], v32.data_ptr(),
                                         f(counts, torch.int32).data_ptr(), f(actions, torch.int32).data_ptr(),
                                         f(old_logp, torch.float32).data_ptr(), f(adv, torch.float32).data_ptr(),
                                         f(returns, torch.float32).data_ptr(), B, float(eps_clip), float(value_coef),
                                         float(entropy_coef), dlogits.data_ptr(), dvalues.data_ptr(), None, sums.data_ptr(), _stream()),
                  "bg_ppo_loss_grad")
        ctx.save_for_backward(dlogits, dvalues)
        ctx.values_dtype = values.dtype
        means = sums / B
        loss = means[0] + value_coef * means[1] - entropy_coef * means[2]
        return loss, means[0], means[1], means[2]

    @staticmethod
    def backward(ctx, g, *_unused):
        dlogits, dvalues = ctx.saved_tensors
        return (dlogits * g.to(dlogits.dtype), (dvalues * g).to(ctx.values_dtype), None, None, None, None, None, None, None, None)


def ppo_loss(params, x, counts, actions, old_logp, returns, advantages, eps_clip, value_coef, entropy_coef,
             autocast=True, fused=None):
    """The loss of one epoch of BackgammonPPOAgent.update (ppo_agent.py:268-299).  counts = legal slots per sample
    (prefix mask, backgammon_env.py:228-231).  On CUDA the part after the two linear layers is ONE kernel
    (bg_ppo_loss_grad; fused=False keeps the torch chain, which is also the CPU path of the host-logic tests)."""
    dev_type = x.device.type
    if fused is None:
        fused = x.is_cuda
    if fused:
        with torch.autocast(device_type=dev_type, dtype=torch.bfloat16, enabled=autocast):
            logits, values = policy_value_forward(params, x)
        loss, pl, vl, ent = _FusedPPOLoss.apply(logits, values, counts, actions, old_logp, advantages, returns, eps_clip,
                                                value_coef, entropy_coef)
        return loss, pl.detach(), vl.detach(), ent.detach()
    with torch.autocast(device_type=dev_type, dtype=torch.bfloat16, enabled=autocast):
        logits, values = policy_value_forward(params, x)
        slot = torch.arange(ACTIONS, device=x.device)[None, :]
        masked = logits.float() + torch.where(slot < counts[:, None], 0.0, MASK_LOG)       # logits + log(mask + 1e-45)
        logp_all = torch.log_softmax(masked, dim=-1)
        new_logp = logp_all.gather(1, actions.long()[:, None])[:, 0]
        ratios = torch.exp(new_logp - old_logp)
        surr1 = ratios * advantages
        surr2 = torch.clamp(ratios, 1 - eps_clip, 1 + eps_clip) * advantages
        policy_loss = -torch.min(surr1, surr2).mean()
        value_loss = F.mse_loss(values.float(), returns)
        p = logp_all.exp()
        entropy = -(p * logp_all).sum(-1).mean()
        loss = policy_loss + value_coef * value_loss - entropy_coef * entropy
    return loss, policy_loss.detach(), value_loss.detach(), entropy.detach()


class ManualUpdate:
    """One epoch of BackgammonPPOAgent.update (ppo_agent.py:268-305: forward, loss, backward) without autograd, for bf16
    feature rows x (B,208) from K3 whose spare column 198 has been set to 1.0:

        h      = relu(x @ W1p^T)                 W1p (128,208) bf16: fc1.weight | fc1.bias in column 198
        logits = h @ Wap^T + bap                 Wap (512,128) bf16: action_head.weight | value_head.weight as row 500
        bg_ppo_loss_grad(values = NULL)          value = column 500; writes d loss / d logits (bf16), its column sums (f32)
        dWap = dlogits^T @ h   dh = dlogits @ Wap   dpre = dh * (h > 0)   dW1p = dpre^T @ x

    five library GEMMs (bf16 operands as under the reference's autocast; the two weight gradients accumulate and are
    returned in f32), one elementwise pass and our loss kernel.  The value head and both biases ride inside the GEMMs /
    the loss kernel, so nothing else touches a (B, .) tensor: no dlogits * grad_output pass, no bias column sums over
    (B,500), no separate value-head matmuls."""

    LD = 512
    ONE_COL = 198

    def __init__(self, device):
        self.device = device
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=device)
        self.w1p, self.wap, self.bap = z((128, 208), torch.bfloat16), z((self.LD, 128), torch.bfloat16), z((self.LD,), torch.bfloat16)
        self.dbias, self.sums = z((self.LD,), torch.float32), z((3,), torch.float32)
        self.B = -1
        self.f32_out = None                      # torch.mm(..., out_dtype=f32) available?
        self._zero_key = None                    # rows the zero part of self.dlogits is valid for

    def _buffers(self, B):
        if B != self.B:
            e = lambda shape: torch.empty(shape, dtype=torch.bfloat16, device=self.device)
            self.h, self.logits, self.dlogits, self.dh = e((B, 128)), e((B, self.LD)), e((B, self.LD)), e((B, 128))
            self.B = B
            self._zero_key = None

    def invalidate(self):
        """New rollout data: the zeros kept in dlogits belong to the previous rows."""
        self._zero_key = None

    def _mm_f32(self, a, b):
        """a @ b for bf16 operands with the f32 accumulator returned as is (torch.mm(..., out_dtype=float32), torch >= 2.8);
        older torch: the bf16 result widened, as autograd under autocast would give."""
        if self.f32_out is None:
            try:
                r = torch.mm(a, b, out_dtype=torch.float32)
                self.f32_out = True
                return r
            except (TypeError, RuntimeError):
                self.f32_out = False
        return torch.mm(a, b, out_dtype=torch.float32) if self.f32_out else torch.mm(a, b).float()

    @torch.no_grad()
    def epoch(self, params, grads, x, counts, actions, old_logp, adv, returns, eps_clip, value_coef, entropy_coef):
        """params / grads: dicts of f32 tensors under the reference's keys; grads are overwritten.  counts / actions i32,
        the other per-sample vectors f32, all contiguous.  -> f32 tensor (policy_loss, value_loss, entropy, total_loss)."""
        B = x.shape[0]
        self._buffers(B)
        w1p, wap, bap = self.w1p, self.wap, self.bap
        w1p[:, :198].copy_(params["fc1.weight"]); w1p[:, self.ONE_COL].copy_(params["fc1.bias"])
        wap[:ACTIONS].copy_(params["action_head.weight"]); wap[ACTIONS].copy_(params["value_head.weight"][0])
        bap[:ACTIONS].copy_(params["action_head.bias"]); bap[ACTIONS:ACTIONS + 1].copy_(params["value_head.bias"])
        torch.mm(x, w1p.t(), out=self.h)
        self.h.relu_()
        torch.addmm(bap, self.h, wap.t(), out=self.logits)
        self.dbias.zero_(); self.sums.zero_()
        # dlogits is mostly zeros (a row with n <= 128 legal slots has 128 live columns + the value column): they are written
        # once per set of rows (a fill at 7 TB/s) and left alone in the following epochs (BG_LOSS_DLOGITS_PREZEROED)
        key = (counts.data_ptr(), B)
        if key != self._zero_key:
            self.dlogits.view(torch.int32).zero_()
            self._zero_key = key
        with torch.cuda.device(self.device):
            check(lib().bg_ppo_loss_grad(self.logits.data_ptr(), 1 | 2, self.LD, None, counts.data_ptr(), actions.data_ptr(),
                                         old_logp.data_ptr(), adv.data_ptr(), returns.data_ptr(), B, float(eps_clip),
                                         float(value_coef), float(entropy_coef), self.dlogits.data_ptr(), None,
                                         self.dbias.data_ptr(), self.sums.data_ptr(), _stream()), "bg_ppo_loss_grad")
        dwap = self._mm_f32(self.dlogits.t(), self.h)                       # (512,128)
        torch.mm(self.dlogits, wap, out=self.dh)                            # (B,128)
        dpre = torch.ops.aten.threshold_backward(self.dh, self.h, 0)        # relu backward
        dw1p = self._mm_f32(dpre.t(), x)                                    # (128,208)
        grads["fc1.weight"].copy_(dw1p[:, :198]); grads["fc1.bias"].copy_(dw1p[:, self.ONE_COL])
        grads["action_head.weight"].copy_(dwap[:ACTIONS]); grads["action_head.bias"].copy_(self.dbias[:ACTIONS])
        grads["value_head.weight"].copy_(dwap[ACTIONS:ACTIONS + 1]); grads["value_head.bias"].copy_(self.dbias[ACTIONS:ACTIONS + 1])
        m = self.sums / B
        return torch.stack([m[0], m[1], m[2], m[0] + value_coef * m[1] - entropy_coef * m[2]])


class TensorCoreUpdate:
    """One epoch of BackgammonPPOAgent.update (ppo_agent.py:268-305: forward, loss, backward) with NO library GEMM: every
    product runs on the tcgen05 tensor cores in kernels of ours (csrc/ppo_gemm.cu) around the loss kernels of csrc/ppo.cu.

    The work follows the action mask: the samples are sorted once per rollout into class A (1..128 legal slots and the
    stored action among them -- ~94 % of a self-play batch) and class B (passes, whose reference arithmetic is a softmax over
    all 500 slots, and rows with more slots), each class padded to whole 128-row tiles with zero rows.  Class A rows only
    touch slots 0..127, so their logits / dlogits are 144 columns wide (128 slots + the value head in column 128) instead
    of 512: a quarter of the head's FLOPs and traffic.  Every activation matrix is kept TILE-BLOCKED in HBM (the tensor
    cores' shared-memory operand image, see csrc/ppo_gemm.cu), so the kernels stream whole tiles with bulk copies.

        xb      = x[perm] blocked, bias column set              bg_ppo_gather_block        (once per rollout: prepare())
        h       = relu(xb W1p^T)                                   bg_ppo_gemm_nt HIDDEN
        logits  = h Wap^T + b          per class                   bg_ppo_gemm_nt LOGITS_A / _B (value head = one more row of Wap)
        dlogits = d loss / d logits    per class                   bg_ppo_loss_grad_classes
        dpre    = (dlogits Wap) [h>0]  per class                   bg_ppo_gemm_nt DPRE_A / _B
        dWap   += dlogits^T h          per class                   bg_ppo_gemm_tn GRAD_WA_A / _B   -> flat f32 gradient
        dW1p   += dpre^T xb                                        bg_ppo_gemm_tn GRAD_W1
    """

    ONE_COL = 198

    def __init__(self, device, fuse_loss: bool = True):
        self.device = device
        self.fuse_loss = bool(fuse_loss)          # class A: loss in the logits GEMM's epilogue (False: separate packed loss kernel)
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=device)
        self.w1p, self.wap_a, self.wap_b = z((26 * 128 * 8,), torch.bfloat16), z((16 * 144 * 8,), torch.bfloat16), z((16 * 512 * 8,), torch.bfloat16)
        self.bias_a, self.bias_b = z((144,), torch.float32), z((512,), torch.float32)
        self.gflat, self.pflat = z((90101,), torch.float32), z((90101,), torch.float32)
        self.dbias, self.sums = z((512,), torch.float32), z((3,), torch.float32)
        self.w1t = z((199 * 128,), torch.float32)              # scratch of GRAD_W1 (dW1p transposed)
        self._cap = (0, 0)
        self._p = None
        self._scratch = {}

    # the class of a sample: csrc/ppo.cu loss_row_is_packed
    @staticmethod
    def class_a(counts, actions):
        return (counts >= 1) & (counts <= 128) & (actions >= 0) & (actions < counts)

    @staticmethod
    def to_blocked(m):
        """(R, C) row-major, R % 128 == 0, C % 8 == 0 -> the tile-blocked layout (flat); for tests and debugging"""
        R, C = m.shape
        return m.view(R // 128, 128, C // 8, 8).permute(0, 2, 1, 3).contiguous().view(-1)

    @staticmethod
    def from_blocked(b, R, C):
        return b.view(R // 128, C // 8, 128, 8).permute(0, 2, 1, 3).reshape(R, C)

    def prepare(self, x, counts, actions, old_logp, adv, returns, boards=None):
        """Once per rollout: sort the samples by class (class A first, by legal count; original order inside class B), pad each class to whole
        tiles, gather the per-sample vectors and build the blocked x (with the bias column).  One host read (the class sizes)."""
        B = counts.shape[0]                              # (boards = (boards52 (B,52) int8, movers (B,) int8): x may be None, see below)
        dev = self.device
        self._p = None
        a = self.class_a(counts, actions)
        # class A first, and inside class A by the number of legal slots: most tiles then hold only rows with <= 32 slots, and the
        # fused logits / loss kernel skips the blocks of slots that are illegal for all 32 rows of a warp
        # (persistent scratch for everything that is B long: no allocation on this path after the first rollout of a size)
        sc = self._scratch
        if sc.get("B") != B:
            sc.clear()
            rows_cap = (-(-B // 128) + 1) * 128
            sc.update(B=B, key=torch.empty(B, dtype=torch.int32, device=dev), skey=torch.empty(B, dtype=torch.int32, device=dev),
                      order=torch.empty(B, dtype=torch.int64, device=dev), perm=torch.empty(rows_cap, dtype=torch.int32, device=dev),
                      idx=torch.empty(rows_cap, dtype=torch.int64, device=dev),
                      counts=torch.empty(rows_cap, dtype=torch.int32, device=dev), actions=torch.empty(rows_cap, dtype=torch.int32, device=dev),
                      old_logp=torch.empty(rows_cap, dtype=torch.float32, device=dev), adv=torch.empty(rows_cap, dtype=torch.float32, device=dev),
                      returns=torch.empty(rows_cap, dtype=torch.float32, device=dev))
        key = sc["key"]
        key.copy_(counts)
        key.masked_fill_(~a, 1 << 20)
        torch.sort(key, stable=True, out=(sc["skey"], sc["order"]))
        order = sc["order"]
        n_a = int(a.sum().item())
        n_b = B - n_a
        TA, TB = -(-n_a // 128), -(-n_b // 128)
        rows = (TA + TB) * 128
        perm = sc["perm"][:max(rows, 1)]
        perm.fill_(-1)
        perm[:n_a] = order[:n_a]
        perm[TA * 128:TA * 128 + n_b] = order[n_a:]
        idx = sc["idx"][:max(rows, 1)]
        idx.copy_(perm)
        idx.clamp_(min=0)
        def g(t, dt, name):
            src = t if t.dtype == dt else t.to(dt)
            return torch.index_select(src, 0, idx, out=sc[name][:idx.shape[0]])
        p = dict(B=B, n_a=n_a, n_b=n_b, TA=TA, TB=TB, rows=rows, perm=perm, counts=g(counts, torch.int32, "counts"),
                 actions=g(actions, torch.int32, "actions"), old_logp=g(old_logp, torch.float32, "old_logp"),
                 adv=g(adv, torch.float32, "adv"), returns=g(returns, torch.float32, "returns"))
        if self._cap[0] < TA + TB or self._cap[1] < TB:
            # sized for ANY split of B samples into the two classes (TA + TB <= ceil(B / 128) + 1), so that the next rollout of the
            # same size never reallocates (a reallocation of these buffers is a 20 ms hiccup at 4 M samples)
            ct, cb = max(-(-B // 128) + 1, self._cap[0]), max(TB + TB // 4 + 1, self._cap[1], (TA + TB) // 8 + 1)
            e = lambda n: torch.zeros(n, dtype=torch.bfloat16, device=dev)
            self.xb, self.h, self.dpre = e(ct * 128 * 208), e(ct * 128 * 128), e(ct * 128 * 128)
            self.la, self.dla = e(ct * 128 * 144), e(ct * 128 * 144)
            self.lb, self.dlb = e(cb * 128 * 512), e(cb * 128 * 512)
            self._cap = (ct, cb)
        # the padding rows of the last tile of each class must read as zeros in dlogits (the loss kernels never write them)
        if TA:
            self.dla[(TA - 1) * 128 * 144:TA * 128 * 144].zero_()
        if TB:
            self.dlb[(TB - 1) * 128 * 512:TB * 128 * 512].zero_()
        with torch.cuda.device(dev):
            if boards is not None:
                # K3 and the gather in one pass: features straight from the stored boards into the blocked layout (no row-major x)
                b52, movers = boards
                check(lib().bg_ppo_encode_block(b52.data_ptr(), movers.data_ptr(), perm.data_ptr(), rows, self.ONE_COL, self.xb.data_ptr(),
                                                _stream()), "bg_ppo_encode_block")
            else:
                x = x.contiguous()
                check(lib().bg_ppo_gather_block(x.data_ptr(), x.shape[1], perm.data_ptr(), rows, 208, self.ONE_COL, self.xb.data_ptr(), _stream()),
                      "bg_ppo_gather_block")
        self._p = p
        return p

    @torch.no_grad()
    def epoch(self, params, grads, x, counts, actions, old_logp, adv, returns, eps_clip, value_coef, entropy_coef, prepared=False,
              flat_params=None, flat_grads=None):
        """Same contract as ManualUpdate.epoch.  prepared: prepare() has been called for this rollout (the learner does it once
        for all epochs); else it is done here.  flat_params / flat_grads: the learner's flat f32 buffers (KEYS order), read /
        written in place instead of through the dicts."""
        p = self._p if prepared else self.prepare(x, counts, actions, old_logp, adv, returns)
        B, n_a, n_b, TA, TB = p["B"], p["n_a"], p["n_b"], p["TA"], p["TB"]
        TT = TA + TB
        L = lib()
        if flat_params is None:
            o = 0
            for k in KEYS:
                n = params[k].numel()
                self.pflat[o:o + n].copy_(params[k].reshape(-1))
                o += n
            flat_params = self.pflat
        gflat = flat_grads if flat_grads is not None else self.gflat
        gflat.zero_(); self.dbias.zero_(); self.sums.zero_()
        st = _stream()
        xp, hp, dp = self.xb.data_ptr(), self.h.data_ptr(), self.dpre.data_ptr()
        # class B buffers are indexed from their own tile 0: pass them offset by -TA tiles (only tiles in range are touched)
        lb_off, dlb_off = self.lb.data_ptr() - TA * 128 * 512 * 2, self.dlb.data_ptr() - TA * 128 * 512 * 2
        with torch.cuda.device(self.device):
            check(L.bg_ppo_pack_weights(flat_params.data_ptr(), self.w1p.data_ptr(), self.wap_a.data_ptr(), self.wap_b.data_ptr(),
                                        self.bias_a.data_ptr(), self.bias_b.data_ptr(), st), "bg_ppo_pack_weights")
            check(L.bg_ppo_gemm_nt(0, xp, 0, TT, self.w1p.data_ptr(), None, None, hp, st), "ppo gemm HIDDEN")
            fused = self.fuse_loss
            if fused:      # class A: the loss is the epilogue of the logits GEMM (the logits never reach HBM)
                check(L.bg_ppo_logits_loss_a(hp, n_a, B, self.wap_a.data_ptr(), self.bias_a.data_ptr(), p["counts"].data_ptr(),
                                             p["actions"].data_ptr(), p["old_logp"].data_ptr(), p["adv"].data_ptr(), p["returns"].data_ptr(),
                                             float(eps_clip), float(value_coef), float(entropy_coef), self.dla.data_ptr(),
                                             self.dbias.data_ptr(), self.sums.data_ptr(), st), "bg_ppo_logits_loss_a")
            else:
                check(L.bg_ppo_gemm_nt(1, hp, 0, TA, self.wap_a.data_ptr(), self.bias_a.data_ptr(), None, self.la.data_ptr(), st), "ppo gemm LOGITS_A")
            check(L.bg_ppo_gemm_nt(2, hp, TA, TT, self.wap_b.data_ptr(), self.bias_b.data_ptr(), None, lb_off, st), "ppo gemm LOGITS_B")
            check(L.bg_ppo_loss_grad_classes(self.la.data_ptr(), self.dla.data_ptr(), self.lb.data_ptr(), self.dlb.data_ptr(), n_a, n_b,
                                             TA * 128, p["counts"].data_ptr(), p["actions"].data_ptr(), p["old_logp"].data_ptr(),
                                             p["adv"].data_ptr(), p["returns"].data_ptr(), float(eps_clip), float(value_coef),
                                             float(entropy_coef), self.dbias.data_ptr(), self.sums.data_ptr(), int(fused), st),
                  "bg_ppo_loss_grad_classes")
            check(L.bg_ppo_gemm_nt(3, self.dla.data_ptr(), 0, TA, self.wap_a.data_ptr(), None, hp, dp, st), "ppo gemm DPRE_A")
            check(L.bg_ppo_gemm_nt(4, dlb_off, TA, TT, self.wap_b.data_ptr(), None, hp, dp, st), "ppo gemm DPRE_B")
            check(L.bg_ppo_gemm_tn(5, hp, self.dla.data_ptr(), 0, TA, gflat.data_ptr(), None, st), "ppo gemm GRAD_WA_A")
            check(L.bg_ppo_gemm_tn(6, hp, dlb_off, TA, TT, gflat.data_ptr(), None, st), "ppo gemm GRAD_WA_B")
            check(L.bg_ppo_gemm_tn(7, dp, xp, 0, TT, gflat.data_ptr(), self.w1t.data_ptr(), st), "ppo gemm GRAD_W1")
        # bias gradients = column sums of dlogits, accumulated by the loss kernels
        OFF_BA, OFF_BV = 128 * 198 + 128 + 500 * 128, 128 * 198 + 128 + 500 * 128 + 500 + 128
        gflat[OFF_BA:OFF_BA + ACTIONS].copy_(self.dbias[:ACTIONS])
        gflat[OFF_BV:OFF_BV + 1].copy_(self.dbias[ACTIONS:ACTIONS + 1])
        if flat_grads is None:
            o = 0
            for k in KEYS:
                n = grads[k].numel()
                grads[k].copy_(gflat[o:o + n].view(grads[k].shape))
                o += n
        m = self.sums / B
        return torch.stack([m[0], m[1], m[2], m[0] + value_coef * m[1] - entropy_coef * m[2]])


def discounted_returns(rewards, dones, values, last_values, gamma, lam):
    """bg_gae on CUDA tensors ([T][N], step-major) -> (returns, advantages)."""
    T, N = rewards.shape
    ret, adv = torch.empty_like(rewards), torch.empty_like(rewards)
    with torch.cuda.device(rewards.device):
        check(lib().bg_gae(rewards.data_ptr(), dones.data_ptr(), values.data_ptr() if values is not None else None,
                           last_values.data_ptr() if last_values is not None else None, T, N, float(gamma), float(lam),
                           ret.data_ptr(), adv.data_ptr(), _stream()), "bg_gae")
    return ret, adv


class FlatParams:
    """The network's f32 master weights as views into ONE flat buffer (and their grads into another), so that the
    data-parallel gradient exchange is a single all-reduce of 90,101 floats."""

    def __init__(self, state_dict, device):
        shapes = [(k, tuple(state_dict[k].shape)) for k in KEYS]
        n = sum(math.prod(s) for _, s in shapes)
        self.flat = torch.zeros(n, dtype=torch.float32, device=device)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=device)
        self.params, o = {}, 0
        for k, s in shapes:
            m = math.prod(s)
            self.flat[o:o + m].copy_(torch.as_tensor(state_dict[k], dtype=torch.float32).reshape(-1))
            p = self.flat[o:o + m].view(s).requires_grad_(True)
            p.grad = self.flat_grad[o:o + m].view(s)
            self.params[k] = p
            o += m
        self.numel = n

    def all_reduce_grads(self, dist=None):
        if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat_grad)                     # one flat bucket (sum) ...
            self.flat_grad.div_(dist.get_world_size())          # ... averaged: every rank holds the same shard size


class PPOLearner:
    """Clipped-surrogate PPO update over a finished rollout (ppo_agent.py:218-366), data parallel."""

    def __init__(self, state_dict, device, cfg: PPOConfig | None = None, dist=None, host_logic_test: bool = False):
        self.cfg = cfg or PPOConfig()
        self.device = torch.device(device)
        if self.device.type != "cuda" and not host_logic_test:
            # no CPU fallback: the learner runs on the GPU.  The CPU suite exercises the host logic (flat bucket, gloo
            # all-reduce, return normalisation, the reference's loss arithmetic in torch) by saying so explicitly.
            raise BgError("PPOLearner needs a CUDA device (host_logic_test=True is for the CPU test-suite only)")
        self.dist = dist
        self.fp = FlatParams(state_dict, self.device)
        self.optimizer = torch.optim.Adam(list(self.fp.params.values()), lr=self.cfg.learning_rate)   # ppo_agent.py:83
        self.total_episodes = 0
        self.total_steps = 0
        self.entropy_coef = self.cfg.entropy_coef_start
        self.last = {}
        self._manual = None

    def _adam_step(self):
        """torch.optim.Adam's step (ppo_agent.py:83,301-305) as ONE kernel over the flat bucket, after the single flat
        all-reduce (the division by the world size is folded into the kernel)."""
        world = 1
        if self.dist is not None and self.dist.is_initialized() and self.dist.get_world_size() > 1:
            self.dist.all_reduce(self.fp.flat_grad)
            world = self.dist.get_world_size()
        if not hasattr(self, "_adam_m"):
            self._adam_m, self._adam_v, self._adam_t = torch.zeros_like(self.fp.flat), torch.zeros_like(self.fp.flat), 0
        self._adam_t += 1
        with torch.cuda.device(self.device):
            check(lib().bg_adam_step(self.fp.flat.data_ptr(), self.fp.flat_grad.data_ptr(), self._adam_m.data_ptr(),
                                     self._adam_v.data_ptr(), self.fp.numel, float(self.cfg.learning_rate), 0.9, 0.999, 1e-8,
                                     self._adam_t, 1.0 / world, _stream()), "bg_adam_step")

    def update_entropy_coef(self):                                                                   # ppo_agent.py:193-204
        c = self.cfg
        progress = min(1.0, self.total_episodes / c.entropy_anneal_episodes)
        self.entropy_coef = c.entropy_coef_start - progress * (c.entropy_coef_start - c.entropy_coef_end)

    def _global_mean_std(self, x):
        """mean and unbiased std of x over ALL ranks (returns.mean() / returns.std(), ppo_agent.py:256)."""
        s = torch.stack([x.sum().double(), (x.double() ** 2).sum(), torch.tensor(float(x.numel()), device=x.device, dtype=torch.float64)])
        if self.dist is not None and self.dist.is_initialized() and self.dist.get_world_size() > 1:
            self.dist.all_reduce(s)
        n = s[2]
        mean = s[0] / n
        var = (s[1] - n * mean * mean) / torch.clamp(n - 1, min=1.0)
        return mean.float(), var.clamp(min=0).sqrt().float()

    def tensor_core_path(self) -> bool:
        """True if update() will run the hand-written tcgen05 update (then it can take the stored boards instead of features)"""
        c = self.cfg
        return bool(c.manual_backward and c.autocast and c.update_impl == "tcgen05" and max(1, int(c.num_minibatches)) == 1
                    and torch.device(self.device).type == "cuda")

    def update(self, x, counts, actions, old_logp, old_values, returns, boards=None):
        """x (B,198+) features, the rest (B,).  Normalises the returns, advantages = returns - V_old (ppo_agent.py:256-259),
        then num_epochs passes over the batch (full batch, or num_minibatches slices).
        boards = (boards52 (B,52) int8, movers (B,) int8) with x = None: only on the tensor-core path (tensor_core_path()), which
        then encodes the features straight into its blocked operand layout."""
        c = self.cfg
        mean, std = self._global_mean_std(returns)
        returns = (returns - mean) / (std + 1e-5)
        adv = returns - old_values
        B = counts.shape[0]
        mb = max(1, int(c.num_minibatches))
        stats = torch.zeros(4, device=counts.device)
        if x is None:
            if boards is None or not self.tensor_core_path():
                raise BgError("PPOLearner.update: x = None needs boards and the tensor-core update path")
            manual = True
        else:
            boards = None
            manual = (c.manual_backward and c.autocast and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 2
                      and x.shape[1] == 208 and x.is_contiguous())
        tc = manual and c.update_impl == "tcgen05" and mb == 1
        if manual:
            if self._manual is None or isinstance(self._manual, TensorCoreUpdate) != tc:
                self._manual = (TensorCoreUpdate(self.device, fuse_loss=os.environ.get("BG_PPO_FUSE_LOSS", "1") != "0") if tc
                                else ManualUpdate(self.device))
            if x is not None:
                x[:, ManualUpdate.ONE_COL] = 1.0              # spare (zero) column of K3's rows: carries fc1.bias through the GEMMs
            counts, actions = counts.to(torch.int32).contiguous(), actions.to(torch.int32).contiguous()
            old_logp, returns, adv = old_logp.float().contiguous(), returns.float().contiguous(), adv.float().contiguous()
            grads = {k: p.grad for k, p in self.fp.params.items()}
            if tc:
                self._manual.prepare(x, counts, actions, old_logp, adv, returns, boards=boards)   # class sort + blocked x, once for all epochs
            else:
                self._manual.invalidate()
        for _ in range(c.num_epochs):
            for k in range(mb):
                sl = slice(k * B // mb, (k + 1) * B // mb)
                if tc:
                    st = self._manual.epoch(self.fp.params, grads, x, counts, actions, old_logp, adv, returns, c.eps_clip,
                                            c.value_loss_coef, self.entropy_coef, prepared=True, flat_params=self.fp.flat,
                                            flat_grads=self.fp.flat_grad)
                elif manual:
                    st = self._manual.epoch(self.fp.params, grads, x[sl], counts[sl], actions[sl], old_logp[sl], adv[sl], returns[sl],
                                            c.eps_clip, c.value_loss_coef, self.entropy_coef)
                else:
                    loss, pl, vl, ent = ppo_loss(self.fp.params, x[sl], counts[sl], actions[sl], old_logp[sl], returns[sl], adv[sl],
                                                 c.eps_clip, c.value_loss_coef, self.entropy_coef, autocast=c.autocast)
                    self.fp.flat_grad.zero_()
                    loss.backward()
                    st = torch.stack([pl.float(), vl.float(), ent.float(), loss.detach().float()])
                if tc:
                    self._adam_step()                         # all-reduce (sum) + our Adam kernel (the 1 / world average folded in)
                else:
                    self.fp.all_reduce_grads(self.dist)
                    self.optimizer.step()
                stats += st
                self.total_steps += 1
        stats /= c.num_epochs * mb
        self.last = dict(zip(("policy_loss", "value_loss", "entropy", "total_loss"), stats.tolist()))
        self.update_entropy_coef()
        return self.last

    def state_dict(self):
        return {k: v.detach().clone() for k, v in self.fp.params.items()}


class RolloutBuffer:
    """[T][N] device tensors replacing the per-sample python dicts of ppo_agent.py:175-186.  Observations are kept as
    board52 + player (53 B instead of 792 B of f32 features); the learner re-encodes them with K3."""

    def __init__(self, T, N, device):
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=device)
        self.T, self.N = T, N
        self.boards, self.players = z((T, N, 52), torch.int8), z((T, N), torch.int8)
        self.counts, self.actions = z((T, N), torch.int32), z((T, N), torch.int32)
        self.logp, self.values = z((T, N), torch.float32), z((T, N), torch.float32)
        self.rewards, self.dones = z((T, N), torch.float32), z((T, N), torch.uint8)
        self.winner, self.info_player = z((T, N), torch.int8), z((T, N), torch.int8)
        self.game_score, self.flags = z((T, N), torch.int8), z((T, N), torch.uint8)


class PPOTrainer:
    """Self-play PPO (train.py:30-123) with every game of this rank resident on its GPU."""

    def __init__(self, env, net: PolicyValueNet, cfg: PPOConfig | None = None, dist=None, seed: int = 0):
        self.env, self.net, self.cfg, self.dist = env, net, cfg or PPOConfig(), dist
        self.device = env.device
        self.learner = PPOLearner(net.params, self.device, self.cfg, dist)
        # the net's master weights ARE the learner's parameters from here on
        net.params = {k: v.detach() for k, v in self.learner.fp.params.items()}
        net.sync()
        self.buf = RolloutBuffer(self.cfg.t_horizon, env.num_envs, self.device)
        self.seed = int(seed)
        self.global_step = 0
        self.episodes = 0
        self.history = []

    def collect(self):
        """T steps of self-play: policy kernel -> K2 -> K1, results written straight into the rollout buffer."""
        env, buf, net = self.env, self.buf, self.net
        L = lib()
        with torch.cuda.device(self.device):
            if self.cfg.reset_each_update:
                env.reset()
            for t in range(buf.T):
                st = env._state()
                check(L.bg_record_state(C.byref(st), buf.boards[t].data_ptr(), buf.players[t].data_ptr(), buf.counts[t].data_ptr(), _stream()),
                      "bg_record_state")                           # boards, movers, legal counts of this step in one launch
                net.act(env.boards52, env.players, env.legal_counts, seed=self.seed, stream_base=env.stream_base,
                        step=self.global_step, out=(buf.actions[t], buf.logp[t], buf.values[t]))
                out = StepOut(buf.rewards[t].data_ptr(), buf.dones[t].data_ptr(), buf.info_player[t].data_ptr(),
                              buf.winner[t].data_ptr(), buf.game_score[t].data_ptr(), buf.flags[t].data_ptr())
                check(L.bg_env_step(C.byref(st), buf.actions[t].data_ptr(), C.byref(out), env.status.data_ptr(), _stream()),
                      "bg_env_step")
                env._refresh_legal_moves()
                self.global_step += 1
            last_v = net.values(env.boards52, env.players) if self.cfg.bootstrap else None
            if self.cfg.returns_mode == "interleaved":
                T, N = buf.T, buf.N
                ret, _ = discounted_returns(buf.rewards.view(T * N, 1), buf.dones.view(T * N, 1), None, None, self.cfg.gamma, 1.0)
                ret = ret.view(T, N)
            else:
                ret, _ = discounted_returns(buf.rewards, buf.dones, buf.values if self.cfg.lam < 1.0 or self.cfg.bootstrap else None,
                                            last_v, self.cfg.gamma, self.cfg.lam)
        return ret

    def update(self, returns):
        buf = self.buf
        B = buf.T * buf.N
        if self.learner.tensor_core_path():               # the update encodes the features itself, straight into its operand layout
            stats = self.learner.update(None, buf.counts.view(B), buf.actions.view(B), buf.logp.view(B), buf.values.view(B), returns.reshape(B),
                                        boards=(buf.boards.view(B, 52), buf.players.view(B)))
            self.net.sync()
            return stats
        dt = torch.bfloat16 if self.cfg.autocast else torch.float32
        xb = getattr(self, "_x", None)                    # persistent: a fresh 1.7 GB tensor per update made the caching allocator split and
        if xb is None or xb.dtype != dt or xb.shape[0] != B:   # re-grow its largest block every few updates (20-100 ms cudaMalloc hiccups)
            self._x = None
            xb = self._x = torch.empty((B, 208 if self.cfg.autocast else 198), dtype=dt, device=self.device)
        x = encode(buf.boards.view(B, 52), buf.players.view(B), dtype=dt, out=xb)
        if not self.cfg.autocast:
            x = x[:, :198]
        stats = self.learner.update(x, buf.counts.view(B), buf.actions.view(B), buf.logp.view(B), buf.values.view(B), returns.reshape(B))
        self.net.sync()
        return stats

    def count_episodes(self):
        """Finished games and reward sum of the rollout just collected, summed over ALL ranks (the reference anneals the
        entropy coefficient on its global episode count, ppo_agent.py:193, train.py:73): every rank then holds the same
        total_episodes and uses the same entropy_coef for the same all-reduced step.  -> (n_done, reward_sum, world)"""
        t = torch.stack([self.buf.dones.sum(dtype=torch.float64), self.buf.rewards.sum(dtype=torch.float64)])
        world = 1
        if self.dist is not None and self.dist.is_initialized() and self.dist.get_world_size() > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            world = self.dist.get_world_size()
        n_done, reward_sum = int(t[0].item()), float(t[1].item())
        self.episodes += n_done
        self.learner.total_episodes = self.episodes
        return n_done, reward_sum, world

    def train(self, num_updates: int, log_every: int = 1, log=print):
        for u in range(num_updates):
            ret = self.collect()
            n_done, reward_sum, world = self.count_episodes()
            stats = self.update(ret)
            stats.update(update=u, episodes=self.episodes, steps=self.global_step * self.env.num_envs * world,
                         mean_reward_per_game=reward_sum / max(1, n_done), entropy_coef=self.learner.entropy_coef)
            self.env.check_status()
            self.history.append(stats)
            if log and u % log_every == 0:
                log(stats)
        return self.history


@torch.no_grad()
def evaluate_vs_random(net: PolicyValueNet, num_games: int = 4096, device=None, seed: int = 1234, max_steps: int = 2000,
                       greedy: bool = True):
    """Win rate (and mean points) of the policy as PLAYER1 against a uniform-random PLAYER2: each env plays exactly one
    game; a learning-curve metric for parity with the reference's `Win Rate` scalar (ppo_agent.py:490-520)."""
    from .vec_env import B200BackgammonVecEnv
    device = torch.device(device or net.device)
    env = B200BackgammonVecEnv(num_envs=num_games, device=device, seed=seed, check_every=0)
    env.reset()
    finished = torch.zeros(num_games, dtype=torch.bool, device=device)
    won = torch.zeros(num_games, dtype=torch.bool, device=device)
    points = torch.zeros(num_games, dtype=torch.float32, device=device)
    for t in range(max_steps):
        a_pol, _, _ = net.act(env.boards52, env.players, env.legal_counts, seed=seed, step=t, greedy=greedy)
        a_rnd = env.random_actions(seed + 1, t)
        acts = torch.where(env.players == 0, a_pol, a_rnd).contiguous()
        env.step_device(acts)
        d = env.dones_u8.bool() & ~finished
        won |= d & (env.winner == 0)
        points += torch.where(d, torch.where(env.winner == 0, env.game_score.float(), -env.game_score.float()), torch.zeros_like(points))
        finished |= d
        if t % 64 == 63 and bool(finished.all()):
            break
    env.check_status()
    n = int(finished.sum().item())
    return {"games": n, "win_rate": float(won.sum().item()) / max(1, n), "mean_points": float(points.sum().item()) / max(1, n)}
