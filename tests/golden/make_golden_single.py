"""Golden vector for the single-game BackgammonEnv seam (train_single.py:16-105): the UNMODIFIED reference env
driven exactly as train_single.py drives it, except that finished games are continued with `env.step(...)` (the
game_over branch, backgammon_env.py:119-121) on every second game and with `env.reset()` on the others, so both
ways out of a terminal state are recorded.  Run once in the build container:

    python tests/golden/make_golden_single.py        -> tests/golden/single_env.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def main(seed=4242, n_games=6, match_length=3):
    import torch
    import ref_loader
    ref_loader.load()
    torch.set_num_threads(1)
    from src.environment.backgammon_env import BackgammonEnv
    from src.board.immutable_board import execute_full_move_on_board_copy as exe
    from src.players.player import Player
    from make_golden import pack52
    rng = np.random.RandomState(seed)
    env = BackgammonEnv(match_length=match_length, max_legal_moves=500)
    env.seed(seed)
    rec = {k: [] for k in ("board", "player", "roll", "n_legal", "mask_sum", "action", "reward", "done", "info_player",
                           "passed", "winner", "game_score", "obs_after", "game_over_after", "match_over_after",
                           "scores_after", "player_after", "was_over")}
    feats = []
    obs = env.reset()
    rec_obs0 = obs.numpy().copy()
    g = 0
    while g < n_games:
        n = len(env.legal_moves)
        was_over = bool(env.game_over)
        rec["was_over"].append(was_over)
        rec["board"].append(pack52(env.board.tensor.numpy())[0])
        rec["player"].append(int(env.current_player))
        rec["roll"].append(list(env.roll_result))
        rec["n_legal"].append(n)
        rec["mask_sum"].append(float(env.action_mask.sum()))
        feats.append(env.legal_board_features[:n].numpy().copy())
        a = int(rng.randint(0, n)) if n else None
        obs, reward, done, info = env.step(a)
        rec["action"].append(-1 if a is None else a)
        rec["reward"].append(float(reward))
        rec["done"].append(bool(done))
        rec["info_player"].append(int(info["current_player"]))
        rec["passed"].append("turn passed" in info.get("info", ""))
        rec["winner"].append(int(info["winner"]) if "winner" in info else -1)
        rec["game_score"].append(int(info.get("game_score", 0)))
        rec["obs_after"].append(obs.numpy().copy())
        rec["game_over_after"].append(bool(env.game_over))
        rec["match_over_after"].append(bool(env.match_over))
        rec["scores_after"].append([env.player_scores[Player.PLAYER1], env.player_scores[Player.PLAYER2]])
        rec["player_after"].append(int(env.current_player))
        if done and not was_over:
            g += 1
            if g % 2 == 0:
                obs = env.reset()              # train_single.py:21
            # else: the next loop iteration steps the finished game (backgammon_env.py:119-121)
    out = {k: np.array(v) for k, v in rec.items()}
    out["feat_offsets"] = np.cumsum([0] + [f.shape[0] for f in feats])
    out["feats"] = np.concatenate(feats, 0)
    out["obs0"] = rec_obs0
    out["seed"] = seed
    out["match_length"] = match_length
    np.savez_compressed(os.path.join(HERE, "single_env.npz"), **out)
    print("steps", len(rec["action"]), "games", g, "steps on finished games", int(np.sum(out["was_over"])))


if __name__ == "__main__":
    main()
