"""BackgammonEnv -- drop-in for the reference's single-game class (src/environment/backgammon_env.py:34-251,
357-405) as `train_single.py:16-105` consumes it: `reset()`, `step(action | None)`, and the attributes
`current_player`, `action_mask`, `legal_board_features`, `legal_moves`, `roll_result`, `board`, `game_over`,
`match_over`, `player_scores`, `current_match_winner`, `seed()`, `render()`.

It is one game of the batched engine with `auto_reset=False`, i.e. the reference's own terminal behaviour
(backgammon_env.py:119-121,156-190) instead of VectorizedBackgammonEnv's:
  * the step that wins returns the observation of the TERMINAL board with the winner still to move, sets
    `game_over`, and leaves `roll_result` / `action_mask` / `legal_board_features` / `legal_moves` as they were
    (the reference does not refresh them either);
  * a `step()` on a finished game resets it and returns `(obs, 0.0, True, {"current_player": winner})`.
All computation is K2 / K1 / K3 of libbg_b200.so through B200BackgammonVecEnv; there is no CPU path.

Dice.  The reference draws `np.random.randint(1, 7)` twice per roll from numpy's GLOBAL generator
(backgammon_env.py:245-246) and `seed(s)` seeds it (:357-363).  `dice="numpy"` (default) reproduces that stream
bit for bit: rolls are drawn on the host with the same two calls per roll, `block` rolls ahead, and injected into K2
(so after `env.seed(s)` the trajectory equals the reference's under the same actions; other users of numpy's global
generator see it advanced by up to `block` rolls).  `dice="philox"` uses the engine's counter-based stream.
"""
from __future__ import annotations

from enum import IntEnum

import numpy as np
import torch

from ._lib import BgError
from .engine import FEATURES, encode, from_board52
from .vec_env import B200BackgammonVecEnv


class Player(IntEnum):
    """players/player.py:6-12"""
    PLAYER1 = 0
    PLAYER2 = 1


class BoardView:
    """What callers read from `env.board` (board/immutable_board.py:16-40,171-212): `.tensor` (4,24) int8 in the
    reference layout and `get_board_features(player)`."""

    def __init__(self, board52: torch.Tensor):
        self._b52 = board52.reshape(1, 52)
        self.tensor = from_board52(self._b52)[0]

    def get_board_features(self, current_player) -> torch.Tensor:
        return encode(self._b52, int(current_player), dtype=torch.float32)[0]


class BackgammonEnv:
    metadata = {"render.modes": ["human"]}

    def __init__(self, match_length=15, max_legal_moves=500, device=None, dice: str = "numpy", seed: int = 0x5EED,
                 block: int = 1024):
        if dice not in ("numpy", "philox"):
            raise BgError('dice must be "numpy" or "philox"')
        self._v = B200BackgammonVecEnv(num_envs=1, match_length=match_length, max_legal_moves=max_legal_moves,
                                       device=device, seed=seed, check_every=1, auto_reset=False)
        self.device = self._v.device
        self.match_length, self.max_legal_moves = int(match_length), int(max_legal_moves)
        self.observation_space, self.action_space = self._v.observation_space, self._v.action_space
        self._dice_mode, self._block = dice, int(block)
        self._fed = 0                                        # rolls injected so far (numpy mode)
        self.current_match_winner = None
        self.game_over = False
        self.roll_result = None
        self.action_mask = torch.zeros(self.max_legal_moves, dtype=torch.float32, device=self.device)
        self.legal_board_features = None
        self.legal_moves = []
        self._afterstates = None

    # ------------------------------------------------------------------ dice
    def seed(self, seed=None):
        """backgammon_env.py:357-363"""
        torch.manual_seed(seed)
        if seed is not None:
            np.random.seed(seed)
            if self._dice_mode == "philox":
                self._v.seed = int(seed)
                self._v._st_cache = None
                self._v.draws.zero_()
        if self._dice_mode == "numpy":
            self._fed = 0
            self._v._ext_dice = None
            self._v._st_cache = None
            self._v.draws.zero_()

    def _feed_dice(self, need: int):
        """Keep at least `need` undrawn rolls in the injected stream, drawn exactly as roll_dice does (:245-246)."""
        if self._dice_mode != "numpy":
            return
        used = int(self._v.draws.item())
        if self._v._ext_dice is not None and self._fed - used >= need:
            return
        fresh = np.array([[np.random.randint(1, 7), np.random.randint(1, 7)] for _ in range(max(self._block, need))], np.int8)
        old = self._v._ext_dice[0, used:self._fed].cpu().numpy() if self._v._ext_dice is not None else np.zeros((0, 2), np.int8)
        stream = np.concatenate([old, fresh], 0)
        self._v._ext_dice = torch.as_tensor(stream[None]).to(self.device).contiguous()
        self._v._st_cache = None
        self._v.draws.zero_()
        self._fed = stream.shape[0]

    def unused_dice(self) -> np.ndarray:
        """numpy mode: the rolls drawn ahead and not consumed yet (for callers that interleave their own numpy draws)."""
        used = int(self._v.draws.item())
        return self._v._ext_dice[0, used:self._fed].cpu().numpy() if self._v._ext_dice is not None else np.zeros((0, 2), np.int8)

    # ------------------------------------------------------------------ views
    @property
    def current_player(self) -> Player:
        return Player(int(self._v.players[0].item()))

    @property
    def board(self) -> BoardView:
        return BoardView(self._v.boards52[0])

    @property
    def player_scores(self):
        s = self._v.scores[0].tolist()
        return {Player.PLAYER1: int(s[0]), Player.PLAYER2: int(s[1])}

    @property
    def match_over(self) -> bool:
        return bool(self._v.match_over[0].item())

    def get_observation(self) -> torch.Tensor:
        """backgammon_env.py:193-196"""
        return self._v.observations()[0]

    def _publish_legal(self):
        """update_legal_moves' outputs (backgammon_env.py:198-243) from the engine's ragged buffers."""
        v = self._v
        n = int(v.legal_counts[0].item())
        self.roll_result = [int(x) for x in v.dice[0].tolist()]
        self.action_mask = v.get_action_masks()[0]
        self.legal_board_features = v.get_legal_board_features()[0]
        s = int(v.legal_starts[0].item())
        self._afterstates = v.after52[s:s + n].clone()
        # the reference keeps FullMove objects; here a legal move is its afterstate, (4,24) int8, same order
        self.legal_moves = list(from_board52(self._afterstates)) if n else []

    # ------------------------------------------------------------------ reference API
    def reset(self):
        """backgammon_env.py:78-113 -> (198,) f32"""
        self._feed_dice(64)
        if self.match_over:                                  # :79-82 (the scores are cleared by the kernel's new_game)
            self.current_match_winner = None
        obs = self._v.reset()
        self.game_over = False
        self._publish_legal()
        return obs[0]

    def step(self, action):
        """backgammon_env.py:115-191 -> (obs (198,) f32, reward 0-d f32 tensor, done bool, info dict)"""
        self._feed_dice(64)
        a = 0 if action is None else int(action)
        was_over = self.game_over
        if was_over and self.match_over:
            self.current_match_winner = None
        obs, rew, done, infos = self._v.step(torch.tensor([a], dtype=torch.int32))
        info = dict(infos[0])
        info["current_player"] = Player(info["current_player"])
        if "winner" in info:
            info["winner"] = Player(info["winner"])
        done = bool(done[0].item())
        if was_over:                                         # :119-121: the reset's observation, reward 0, done
            self.game_over = False
            self._publish_legal()
        elif done:                                           # :156-181: terminal board, stale roll / mask / features
            self.game_over = True
            if self.match_over:
                self.current_match_winner = info["winner"]
        elif "Invalid action" in info.get("info", ""):
            print(f"Invalid action selected: {action}. Assigned reward: {rew[0]}")   # :147
        else:
            self._publish_legal()
        return obs[0], rew[0], done, info

    def update_legal_moves(self):
        with torch.cuda.device(self.device):
            self._v.update_legal_plays(obs=False, features=False)
        self._publish_legal()

    def render(self, mode="human"):
        return self._v.render(0, mode)

    def close(self):
        self._v.close()
