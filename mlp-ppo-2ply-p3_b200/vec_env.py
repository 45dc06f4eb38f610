"""B200BackgammonVecEnv -- drop-in for the reference's VectorizedBackgammonEnv
(src/environment/vec_bg_env.py:7-71) with every game resident in HBM.

Same constructor arguments and method set (reset / step / get_action_masks /
get_legal_board_features / close), same return shapes and dtypes, same per-game semantics as
BackgammonEnv (src/environment/backgammon_env.py:38-251, 357-405) including auto-reset on done,
pass turns, invalid-action reward -1 and the first-`max_legal_moves` truncation.  Differences, all
additive: `infos` is a lazy list view over device tensors; legal plays are also available ragged
(`legal_counts`, `legal_starts`, `afterstates`, `afterstate_features`); dice come from a counter-based
Philox stream (or an injected per-game dice list for parity runs) instead of numpy's global MT19937.

All compute is in libbg_b200.so (hand-written sm_100a kernels); torch only owns the memory.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import BgError, EnvState, StepOut, check, lib
from .engine import FEATURES, LD_BF16, from_board52, to_board52, _stream


class StepInfos:
    """Lazy list-of-dicts view of the per-game info tensors (reference: list of dicts, vec_bg_env.py:40).

    Keys as in backgammon_env.py:117,139,149,172: "current_player", "info" (on pass / invalid action),
    "winner" and "game_score" (on a win).  The tensors themselves are the cheap way to read them.
    """

    def __init__(self, current_player, winner, game_score, flags):
        self.current_player, self.winner, self.game_score, self.flags = current_player, winner, game_score, flags
        self._host = None

    def __len__(self):
        return self.current_player.shape[0]

    def _h(self):
        if self._host is None:
            self._host = (self.current_player.cpu().tolist(), self.winner.cpu().tolist(),
                          self.game_score.cpu().tolist(), self.flags.cpu().tolist())
        return self._host

    def __getitem__(self, i):
        cp, w, gs, fl = self._h()
        d = {"current_player": cp[i]}
        if fl[i] & 1:
            d["info"] = "No legal actions, turn passed"
        if fl[i] & 2:
            d["info"] = "Invalid action"
        if w[i] >= 0:
            d["winner"], d["game_score"] = w[i], gs[i]
        return d

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class HostStepBuffers:
    """Pinned host mirrors of what a host-side policy needs from every step -- rewards, dones and the legal-play counts
    of the new positions -- filled by `env.step(actions, host=buffers)` on a copy stream as soon as K2 and K1 have
    finished, i.e. while the encoders (K3: observations, afterstate features) of the same step are still running.
    `wait()` blocks until the three arrays of the latest step are on the host; the device tensors step() returned stay
    asynchronous as usual."""

    def __init__(self, env):
        N, dev = env.num_envs, env.device
        self.rewards = torch.empty(N, dtype=torch.float32).pin_memory()
        self.dones = torch.empty(N, dtype=torch.bool).pin_memory()
        self.legal_counts = torch.empty(N, dtype=torch.int32).pin_memory()
        self.stream = torch.cuda.Stream(device=dev)
        self.stepped = torch.cuda.Event()                  # K2 of the step done: rewards / dones final
        self.legal_ready = torch.cuda.Event()              # K1 of the step done (recorded inside bg_update_legal_plays)
        self.ready = torch.cuda.Event()                    # ... and everything copied to the host
        self.obs_ready = torch.cuda.Event()                # observations of the step encoded (side stream)
        self.actions_up = torch.cuda.Event()               # the step's actions copied to the device (copy stream)
        with torch.cuda.device(dev):
            for e in (self.stepped, self.legal_ready, self.ready, self.obs_ready, self.actions_up):
                e.record()                                 # (handles are created on the first record)

    def wait(self):
        self.ready.synchronize()


class B200BackgammonVecEnv:
    def __init__(self, num_envs=1, match_length=15, max_legal_moves=500, device=None, seed=0x5EED,
                 stream_base=0, rows_per_game=64, dense_budget_bytes=4 << 30, check_every=16, auto_reset=True):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        device = torch.device(device) if device is not None else None
        if device is None or device.type != "cuda":
            raise BgError("B200BackgammonVecEnv needs a CUDA device (there is no CPU fallback)")
        lib()   # fail loudly now if the extension is missing
        self.num_envs, self.match_length, self.max_legal_moves = int(num_envs), int(match_length), int(max_legal_moves)
        self.device, self.seed, self.stream_base = device, int(seed), int(stream_base)
        self.dense_budget_bytes, self.check_every = int(dense_budget_bytes), int(check_every)
        N = self.num_envs
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=device)
        self.boards52, self.players, self.dice = z((N, 52), torch.int8), z((N,), torch.int8), z((N, 2), torch.int8)
        self.scores, self.draws, self.match_over = z((N, 2), torch.int32), z((N,), torch.int32), z((N,), torch.int8)
        # auto_reset=False: BackgammonEnv's terminal behaviour (backgammon_env.py:119-121,156-181) instead of the vec env's
        # (vec_bg_env.py:35-36): finished games keep their terminal board until their next step() resets them (single_env.py)
        self.auto_reset = bool(auto_reset)
        self.game_over = None if self.auto_reset else z((N,), torch.int8)
        self.cap_rows = max(N * int(rows_per_game), 4096)
        self.after52 = z((self.cap_rows, 52), torch.int8)
        self.row_players = z((self.cap_rows,), torch.int8)
        self.legal_starts, self.legal_counts = z((N,), torch.int64), z((N,), torch.int32)
        self.legal_counts_true = z((N,), torch.int32)
        self.status = z((1,), torch.int32)
        self.rewards, self.dones_u8 = z((N,), torch.float32), z((N,), torch.uint8)
        self.info_player, self.winner, self.game_score = z((N,), torch.int8), z((N,), torch.int8), z((N,), torch.int8)
        self.flags = z((N,), torch.uint8)
        self._ws_bytes = int(lib().bg_movegen_workspace_bytes(max(N, 1)))
        self._ws = torch.zeros(self._ws_bytes, dtype=torch.uint8, device=device)
        # the slab row allocator lives in the workspace header (bytes [48, 56)): every K1 call zeroes it together with its own
        # counters, so no separate fill per turn
        self.alloc_rows = self._ws[48:56].view(torch.int64)
        self._ext_dice = None
        self._side = None
        self._st_cache = None
        self._act_dev = None
        self._pinned_ptr = -1
        self._next_out = None
        self._steps = 0
        # reference attributes (vec_bg_env.py:16-18); gym is not a dependency, so plain descriptors
        self.observation_space = {"shape": (FEATURES,), "low": -1.0, "high": 1.0, "dtype": "float32"}
        self.action_space = {"n": self.max_legal_moves}

    # ------------------------------------------------------------------ plumbing
    def _state(self) -> EnvState:
        e = self._ext_dice
        return EnvState(self.num_envs, self.boards52.data_ptr(), self.players.data_ptr(), self.dice.data_ptr(),
                        self.scores.data_ptr(), self.draws.data_ptr(), self.match_over.data_ptr(),
                        self.after52.data_ptr(), self.legal_starts.data_ptr(), self.legal_counts.data_ptr(),
                        self.seed, self.stream_base, e.data_ptr() if e is not None else None,
                        e.shape[1] if e is not None else 0, self.match_length, 0 if self.auto_reset else 1,
                        None if self.auto_reset else self.game_over.data_ptr())

    def _state_cached(self) -> EnvState:
        """The state struct of _state(), built once: every pointer in it is a persistent buffer (set_dice_stream, the only
        thing that swaps one, drops the cache)."""
        if self._st_cache is None:
            self._st_cache = self._state()
        return self._st_cache

    def _refresh_legal_moves(self, with_features: bool = False):
        """update_legal_moves (backgammon_env.py:198-243) for every game: K1 in slab mode.  with_features also
        writes generate_all_board_features (ai/batching.py:10-75) of every legal play as bf16 rows into
        self.after_feats, fused into K1's output stage (no separate encoder launch)."""
        fptr = None
        if with_features:
            if not hasattr(self, "after_feats"):
                self.after_feats = torch.empty((self.cap_rows, LD_BF16), dtype=torch.bfloat16, device=self.device)
            fptr = self.after_feats.data_ptr()
        check(lib().bg_movegen_slab(self.boards52.data_ptr(), self.players.data_ptr(), self.dice.data_ptr(),
                                    self.num_envs, self.max_legal_moves, self.after52.data_ptr(), self.cap_rows,
                                    self.row_players.data_ptr(), fptr, self.legal_counts_true.data_ptr(),
                                    self.legal_counts.data_ptr(), self.legal_starts.data_ptr(),
                                    self.alloc_rows.data_ptr(), self.status.data_ptr(), self._ws.data_ptr(),
                                    self._ws_bytes, _stream()), "bg_movegen_slab")

    def update_legal_plays(self, obs: bool = True, features: bool = False, overlap: bool = True, k1_events=None):
        """update_legal_moves + get_observation (backgammon_env.py:193-243) in one C call (bg_update_legal_plays):
        K1, then K3 into the persistent buffers self.obs_f32 (N,198) f32 and self.after_feats (cap_rows,208) bf16
        (rows [0, alloc_rows) valid).  overlap=True runs the encoders on a second stream beside K1's latency-bound
        overflow tiers (fork/join inside the call; for the caller everything stays ordered on the current stream).
        k1_events: optional pair of torch.cuda.Event (timing enabled, already recorded once) bracketing K1."""
        optr = fptr = None
        if obs:
            if not hasattr(self, "obs_f32"):
                self.obs_f32 = torch.empty((self.num_envs, FEATURES), dtype=torch.float32, device=self.device)
            optr = self.obs_f32.data_ptr()
        if features:
            if not hasattr(self, "after_feats"):
                self.after_feats = torch.empty((self.cap_rows, LD_BF16), dtype=torch.bfloat16, device=self.device)
            fptr = self.after_feats.data_ptr()
        side = None
        if overlap:
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            side = self._side.cuda_stream
        e0 = k1_events[0].cuda_event if k1_events is not None and k1_events[0] is not None else None
        e1 = k1_events[1].cuda_event if k1_events is not None and k1_events[1] is not None else None
        check(lib().bg_update_legal_plays(self.boards52.data_ptr(), self.players.data_ptr(), self.dice.data_ptr(),
                                          self.num_envs, self.max_legal_moves, self.after52.data_ptr(), self.cap_rows,
                                          self.row_players.data_ptr(), self.legal_counts_true.data_ptr(),
                                          self.legal_counts.data_ptr(), self.legal_starts.data_ptr(),
                                          self.alloc_rows.data_ptr(), self.status.data_ptr(), self._ws.data_ptr(),
                                          self._ws_bytes, fptr, LD_BF16, optr, FEATURES, e0, e1, side, _stream()),
              "bg_update_legal_plays")

    def check_status(self):
        s = int(self.status.item())
        if s:
            raise BgError(f"device status {s}: {_lib.status_message(s)}"
                          + (" -- construct the env with a larger rows_per_game" if s & 4 else ""))

    def observations(self, dtype=torch.float32) -> torch.Tensor:
        """get_observation (backgammon_env.py:193-196) of every game: (N,198) f32 (or (N,208) bf16)."""
        from .engine import encode
        return encode(self.boards52, self.players, dtype=dtype)

    # ------------------------------------------------------------------ reference API
    def reset(self):
        """vec_bg_env.py:20-26 -> (N,198) f32"""
        with torch.cuda.device(self.device):
            st = self._state()
            check(lib().bg_env_reset(C.byref(st), None, self.status.data_ptr(), _stream()), "bg_env_reset")
            self.obs_f32 = torch.empty((self.num_envs, FEATURES), dtype=torch.float32, device=self.device)
            self.update_legal_plays(obs=True, features=False)
            obs = self.obs_f32
        self.check_status()
        return obs

    def _alloc_outputs(self, return_obs):
        N, dev = self.num_envs, self.device
        return (return_obs, torch.empty(N, dtype=torch.float32, device=dev), torch.empty(N, dtype=torch.bool, device=dev),   # K2 writes 0/1 bytes
                torch.empty((4, N), dtype=torch.int8, device=dev),
                torch.empty((N, FEATURES), dtype=torch.float32, device=dev) if return_obs else None)

    def step(self, actions, return_obs=True, with_features=False, host: HostStepBuffers | None = None, overlap: bool | None = None):
        """vec_bg_env.py:28-49 -> (obs (N,198) f32, rewards (N,) f32, dones (N,) bool, infos).
        Every returned tensor is a fresh allocation (as in the reference) that the kernels write directly: no
        staging copies.  with_features also refreshes self.after_feats (bf16 features of every legal play).
        host: rewards / dones / new legal-play counts are also copied into these pinned buffers as soon as K2 + K1 are
        done (host.wait()), without waiting for the encoders.  overlap=False runs the encoders after K1 instead of beside
        its overflow tiers: K1 finishes earlier, which is what a host waiting for the counts wants (the default with
        `host`; without it the default is the overlapped form, which is the shorter one for the GPU)."""
        if overlap is None:
            overlap = host is None
        if not isinstance(actions, torch.Tensor):
            import numpy as np
            acts = [0 if a is None else int(a) for a in actions] if not isinstance(actions, np.ndarray) else actions
            actions = torch.as_tensor(acts)
        if actions.shape[0] != self.num_envs:
            raise BgError("step: need one action per env")
        if host is not None and not actions.is_cuda:
            # host actions go up on the copy stream, i.e. beside the encoders of the previous step that are still running on
            # the main stream, not behind them
            # (is_pinned() asks the driver: remember the buffer that passed)
            ptr = actions.data_ptr()
            pinned = ptr == self._pinned_ptr
            if not pinned and actions.dtype == torch.int32 and actions.is_contiguous() and actions.is_pinned():
                self._pinned_ptr, pinned = ptr, True
            if pinned and actions.dtype == torch.int32 and actions.is_contiguous():
                # pinned int32 actions: one cudaMemcpyAsync into a persistent device buffer (no allocation, no stream context)
                if self._act_dev is None:
                    self._act_dev = torch.empty(self.num_envs, dtype=torch.int32, device=self.device)
                host.stream.wait_event(host.stepped)                           # K2 of the previous step has read the buffer
                check(lib().bg_copy_actions_async(self._act_dev.data_ptr(), actions.data_ptr(), self.num_envs,
                                                  host.stream.cuda_stream), "bg_copy_actions_async")
                host.actions_up.record(host.stream)
                actions = self._act_dev
            else:
                with torch.cuda.stream(host.stream):
                    actions = actions.to(device=self.device, dtype=torch.int32, non_blocking=True).contiguous()
                    host.actions_up.record(host.stream)
                actions.record_stream(torch.cuda.current_stream(self.device))
            torch.cuda.current_stream(self.device).wait_event(host.actions_up)
        else:
            actions = actions.to(device=self.device, dtype=torch.int32, non_blocking=True).contiguous()
        N, dev = self.num_envs, self.device
        with torch.cuda.device(dev):
            if host is not None:                                                # K1 rewrites legal_counts: not before the previous
                torch.cuda.current_stream().wait_event(host.ready)              # step's copy has read it (a no-op after host.wait())
            # fresh result tensors (allocated at the end of the previous step, off the launch path); the persistent
            # attributes alias them until the next step
            out = self._next_out if self._next_out is not None and self._next_out[0] == return_obs else self._alloc_outputs(return_obs)
            _, self.rewards, dones, info, obs = out
            if return_obs:
                self.obs_f32 = obs
            # K2 first, from raw pointers (info is (4,N) int8: player, winner, game score, flags) and the cached state struct;
            # the python views of the outputs are made after the launches, off the host's critical path
            ip = info.data_ptr()
            so = StepOut(self.rewards.data_ptr(), dones.data_ptr(), ip, ip + N, ip + 2 * N, ip + 3 * N)
            check(lib().bg_env_step(C.byref(self._state_cached()), actions.data_ptr(), C.byref(so), self.status.data_ptr(), _stream()),
                  "bg_env_step")
            if host is not None:                                                # rewards / dones are final after K2
                host.stepped.record()
            obs_beside_k1 = host is not None and return_obs and not overlap
            if obs_beside_k1:
                # the observations depend on K2 only: encode them on the side stream beside K1's issue-bound tier 0 (20 us of
                # HBM writes that cost K1 nothing) instead of after K1
                if self._side is None:
                    self._side = torch.cuda.Stream(device=dev)
                self._side.wait_event(host.stepped)
                check(lib().bg_encode_f32(self.boards52.data_ptr(), self.players.data_ptr(), 0, N, None, self.obs_f32.data_ptr(),
                                          FEATURES, self._side.cuda_stream), "bg_encode_f32")
                host.obs_ready.record(self._side)
            self.update_legal_plays(obs=return_obs and not obs_beside_k1, features=with_features, overlap=overlap,
                                    k1_events=(None, host.legal_ready) if host is not None else None)
            if obs_beside_k1:
                torch.cuda.current_stream().wait_event(host.obs_ready)
            if host is not None:
                cs = host.stream
                cs.wait_event(host.stepped)
                with torch.cuda.stream(cs):
                    host.rewards.copy_(self.rewards, non_blocking=True)
                    host.dones.copy_(dones, non_blocking=True)
                    cs.wait_event(host.legal_ready)
                    host.legal_counts.copy_(self.legal_counts, non_blocking=True)
                    host.ready.record(cs)
                self.rewards.record_stream(cs); dones.record_stream(cs)
            self.dones_u8 = dones.view(torch.uint8)
            self.info_player, self.winner, self.game_score, self.flags = info[0], info[1], info[2], info[3].view(torch.uint8)
            self._next_out = self._alloc_outputs(return_obs)
        self._steps += 1
        if self.check_every and self._steps % self.check_every == 0:
            self.check_status()
        infos = StepInfos(self.info_player, self.winner, self.game_score, self.flags)
        return (self.obs_f32 if return_obs else None), self.rewards, dones, infos

    def _apply_actions(self, actions_i32: torch.Tensor):
        """K2: step / reward / terminal / auto-reset / dice for every game."""
        st = self._state()
        out = StepOut(self.rewards.data_ptr(), self.dones_u8.data_ptr(), self.info_player.data_ptr(),
                      self.winner.data_ptr(), self.game_score.data_ptr(), self.flags.data_ptr())
        check(lib().bg_env_step(C.byref(st), actions_i32.data_ptr(), C.byref(out), self.status.data_ptr(), _stream()),
              "bg_env_step")

    def step_random_device(self, seed: int, t: int, actions_out: torch.Tensor | None = None):
        """K2 with the uniform-random policy inside (the draw of random_actions(seed, t)): one launch instead of two."""
        st = self._state_cached()
        out = StepOut(self.rewards.data_ptr(), self.dones_u8.data_ptr(), self.info_player.data_ptr(),
                      self.winner.data_ptr(), self.game_score.data_ptr(), self.flags.data_ptr())
        check(lib().bg_env_step_random(C.byref(st), int(seed), int(t), actions_out.data_ptr() if actions_out is not None else None,
                                       C.byref(out), self.status.data_ptr(), _stream()), "bg_env_step_random")

    def step_device(self, actions_i32: torch.Tensor, with_features: bool = False):
        """The hot path only: K2 (step/reward/terminal/reset/dice) + K1 (legal plays of the new positions,
        optionally with their bf16 features).  Results land in self.rewards / dones_u8 / info_* / legal_* /
        after_feats without any host synchronisation."""
        self._apply_actions(actions_i32)
        self._refresh_legal_moves(with_features)

    def encode_resident(self, obs=True, afterstates=True):
        """K3 into persistent buffers, no host sync: self.obs_f32 (N,198) f32 = get_observation of every game,
        self.after_feats (cap_rows,208) bf16 = generate_all_board_features of every game (ragged, rows
        [0, alloc_rows) valid).  What the reference computes eagerly in update_legal_moves/get_observation."""
        from .engine import encode
        if obs:
            if not hasattr(self, "obs_f32"):
                self.obs_f32 = torch.empty((self.num_envs, FEATURES), dtype=torch.float32, device=self.device)
            encode(self.boards52, self.players, dtype=torch.float32, out=self.obs_f32)
        if afterstates:
            if not hasattr(self, "after_feats"):
                self.after_feats = torch.empty((self.cap_rows, LD_BF16), dtype=torch.bfloat16, device=self.device)
            encode(self.after52, self.row_players, dtype=torch.bfloat16, out=self.after_feats, n_rows_dev=self.alloc_rows)

    def random_actions(self, seed: int, t: int, out: torch.Tensor | None = None) -> torch.Tensor:
        """Uniform random policy on the device (Philox, domain "ACT1")."""
        if out is None:
            out = torch.empty(self.num_envs, dtype=torch.int32, device=self.device)
        check(lib().bg_random_actions(self.legal_counts.data_ptr(), self.num_envs, int(seed), self.stream_base, int(t),
                                      out.data_ptr(), _stream()), "bg_random_actions")
        return out

    def get_action_masks(self):
        """vec_bg_env.py:51-56 -> (N, max_legal_moves) f32"""
        ar = torch.arange(self.max_legal_moves, device=self.device, dtype=torch.int32)
        return (ar[None, :] < self.legal_counts[:, None]).to(torch.float32)

    def get_legal_board_features(self):
        """vec_bg_env.py:58-63 -> dense (N, max_legal_moves, 198) f32, zero padded (backgammon_env.py:233-243).
        Guarded: the dense tensor is 396 KB per game; use afterstate_features() at scale."""
        need = self.num_envs * self.max_legal_moves * FEATURES * 4
        if need > self.dense_budget_bytes:
            raise BgError(f"dense legal_board_features would take {need / 2**30:.1f} GiB; use the ragged "
                          "afterstate_features() / legal_starts / legal_counts accessors instead")
        feats, rows_game, rows_idx = self._ragged_features(torch.float32)
        dense = torch.zeros((self.num_envs, self.max_legal_moves, FEATURES), dtype=torch.float32, device=self.device)
        dense[rows_game, rows_idx] = feats
        return dense

    def close(self):
        self.check_status()

    def render(self, index: int = 0, mode: str = "human"):
        """BackgammonEnv.render (backgammon_env.py:253-355) of game `index`; returns the text it prints."""
        if mode != "human":
            raise NotImplementedError("Only 'human' mode is supported")
        from .engine import render_board52
        text = render_board52(self.boards52[index].cpu().numpy())
        print(text)
        return text

    def match_scores(self):
        """(N,2) i32 match points of PLAYER1 / PLAYER2 (backgammon_env.py:42-45,173-181) and (N,) match-over flags."""
        return self.scores, self.match_over.bool()

    # ------------------------------------------------------------------ ragged extras
    def total_rows(self) -> int:
        return int(self.alloc_rows.item())

    def afterstates(self):
        """(rows,52) int8 view of the current legal-play buffer; game g owns rows
        legal_starts[g] .. legal_starts[g]+legal_counts[g], in the reference's legal_moves order."""
        return self.after52[: self.total_rows()]

    def afterstate_features(self, dtype=torch.bfloat16):
        """generate_all_board_features (ai/batching.py:10-75) for every game, ragged: (rows, 208) bf16 or
        (rows,198) f32, row r encoded with its mover's turn flag (batching.py:72-74)."""
        from .engine import encode
        n = self.total_rows()
        return encode(self.after52[:n], self.row_players[:n], dtype=dtype)

    @property
    def legal_offsets(self) -> torch.Tensor:
        """(N+1,) i64 CSR offsets of the legal plays in game order (exclusive prefix sum of legal_counts) -- the layout
        afterstates_csr() returns.  (In the env's own buffer game g's rows start at legal_starts[g]: blocks are in
        completion order, not game order.)"""
        off = torch.zeros(self.num_envs + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(self.legal_counts, 0, out=off[1:])
        return off

    def afterstates_csr(self):
        """-> (offsets (N+1,) i64, afterstates (total,52) i8) with the games' blocks in game order (a gather of the
        env's buffer; rows of a game in the reference's legal_moves order)."""
        off = self.legal_offsets
        total = int(off[-1].item())
        games = torch.repeat_interleave(torch.arange(self.num_envs, device=self.device), self.legal_counts.long(), output_size=total)
        rows = self.legal_starts[games] + (torch.arange(total, device=self.device) - off[games])
        return off, self.after52[rows]

    def _ragged_features(self, dtype):
        from .engine import encode
        counts = self.legal_counts.to(torch.int64)
        total = int(counts.sum().item())
        games = torch.repeat_interleave(torch.arange(self.num_envs, device=self.device), counts)
        offs = torch.cumsum(counts, 0) - counts
        idx = torch.arange(total, device=self.device) - offs[games]
        rows = self.legal_starts[games] + idx
        feats = encode(self.after52[rows], self.players[games], dtype=dtype)
        return feats, games, idx

    # ------------------------------------------------------------------ parity / injection hooks
    def boards(self) -> torch.Tensor:
        """(N,4,24) int8 in the reference layout."""
        return from_board52(self.boards52)

    def set_dice_stream(self, dice: torch.Tensor | None):
        """Inject dice: (N, L, 2) int8, game g consumes dice[g, draw] in order (replaces np.random.randint of
        backgammon_env.py:245-246).  None returns to Philox."""
        if dice is None:
            self._ext_dice = None
        else:
            d = torch.as_tensor(dice).to(device=self.device, dtype=torch.int8).contiguous()
            if d.dim() != 3 or d.shape[0] != self.num_envs or d.shape[2] != 2:
                raise BgError("set_dice_stream: need (N, L, 2)")
            self._ext_dice = d
        self._st_cache = None
        self.draws.zero_()

    def set_dice(self, dice):
        """Overwrite the current roll of every game ((N,2) int, values 1..6) and refresh the legal plays: the
        one-step form of dice injection (the reference: assigning env.dice before update_legal_moves)."""
        d = torch.as_tensor(dice).to(self.device, torch.int8).reshape(-1, 2)
        if d.shape[0] != self.num_envs:
            raise BgError("set_dice: need (N, 2)")
        self.dice.copy_(d)
        with torch.cuda.device(self.device):
            self.update_legal_plays(obs=False, features=False)
        self.check_status()

    def load_positions(self, boards, players, dice):
        """Overwrite every game's position (reference-layout (N,4,24) or packed (N,52)), mover and roll, then
        refresh the legal plays."""
        b = torch.as_tensor(boards).to(self.device)
        self.boards52.copy_(b if b.shape[-1] == 52 and b.dim() == 2 else to_board52(b))
        self.players.copy_(torch.as_tensor(players).to(self.device, torch.int8))
        self.dice.copy_(torch.as_tensor(dice).to(self.device, torch.int8).reshape(-1, 2))
        with torch.cuda.device(self.device):
            self._refresh_legal_moves()
        self.check_status()

    load_boards = load_positions                      # the name SURVEY.md 8(b) uses


# The reference's class name, for callers that do `VectorizedBackgammonEnv(num_envs, ...)` (train.py:128-133)
VectorizedBackgammonEnv = B200BackgammonVecEnv
