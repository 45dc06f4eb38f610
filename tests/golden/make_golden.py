"""Generate golden vectors by running the UNMODIFIED reference (/root/reference).

Run once in the build container (the reference is not on the GPU box):
    python tests/golden/make_golden.py [--procs 8]
Outputs (committed): tests/golden/*.npz.  Afterstates are stored in the
reference's list order, packed to 52 bytes
[P1 points 24][P2 points 24][bar1 bar2 off1 off2].

Reference entry points exercised:
  get_all_possible_moves            src/moves/get_all_moves.py:9-70
  execute_full_move_on_board_copy   src/board/immutable_board.py:224-233
  get_board_features_batch_from_tensors / get_board_features
                                    src/ai/batching.py:78-147, immutable_board.py:171-212
  BackgammonEnv.reset/step          src/environment/backgammon_env.py:78-191
  BackgammonPolicyNetwork.forward   src/agent/policy_network.py:58-75
"""
import argparse
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def _ref():
    import torch
    import ref_loader
    ref_loader.load()
    torch.set_num_threads(1)
    from src.moves.get_all_moves import get_all_possible_moves
    from src.board.immutable_board import ImmutableBoard, execute_full_move_on_board_copy
    from src.players.player import Player
    return torch, get_all_possible_moves, ImmutableBoard, execute_full_move_on_board_copy, Player


def pack52(b96: np.ndarray) -> np.ndarray:
    b = b96.reshape(-1, 4, 24)
    out = np.zeros((b.shape[0], 52), np.int8)
    out[:, 0:24] = b[:, 0]
    out[:, 24:48] = b[:, 1]
    out[:, 48:50] = b[:, 2, 0:2]
    out[:, 50:52] = b[:, 3, 0:2]
    return out


def ref_moves(board96: np.ndarray, player: int, d0: int, d1: int):
    """-> (afterstates (n,4,24) i8 in reference order, n_sub (n,) i32)"""
    torch, gapm, IB, exe, Player = _ref()
    b = IB(torch.tensor(board96.reshape(4, 24), dtype=torch.int8))
    moves = gapm(Player(player), b, [int(d0), int(d1)])
    after = np.zeros((len(moves), 4, 24), np.int8)
    nsub = np.zeros((len(moves),), np.int32)
    for i, m in enumerate(moves):
        after[i] = exe(b, m).tensor.numpy()
        nsub[i] = len(m.sub_move_commands)
    return after, nsub


def _movegen_job(args):
    board96, player, d0, d1 = args
    return ref_moves(board96, player, d0, d1)


def run_cases(pool, boards, players, dice):
    res = pool.map(_movegen_job, [(boards[i], int(players[i]), int(dice[i, 0]), int(dice[i, 1]))
                                  for i in range(len(boards))], chunksize=4)
    counts = np.array([r[0].shape[0] for r in res], np.int32)
    offsets = np.zeros(len(res) + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    after = np.concatenate([pack52(r[0]) for r in res], 0) if len(res) else np.zeros((0, 52), np.int8)
    nsub = np.concatenate([r[1] for r in res], 0) if len(res) else np.zeros((0,), np.int32)
    return counts, offsets, after, nsub


# ----------------------------------------------------------------------------- trajectories

def _trajectory_job(args):
    seed, n_games, max_legal = args
    torch, gapm, IB, exe, Player = _ref()
    from src.environment.backgammon_env import BackgammonEnv
    rng = np.random.RandomState(seed)          # action choice (not the env's dice)
    env = BackgammonEnv(match_length=3, max_legal_moves=max_legal)
    env.seed(seed)
    dice_log = []
    orig_roll = env.roll_dice

    def roll():
        orig_roll()
        dice_log.append(list(env.roll_result))
    env.roll_dice = roll

    rec = {k: [] for k in ("board", "player", "roll", "n_legal", "action", "reward", "done", "passed",
                           "winner", "game_score", "obs_after", "dice_pos", "scores", "game_id")}
    afters = []
    obs = env.reset()
    g = 0
    while g < n_games:
        n = len(env.legal_moves)
        rec["board"].append(env.board.tensor.numpy().copy())
        rec["player"].append(int(env.current_player))
        rec["roll"].append(list(env.roll_result))
        rec["n_legal"].append(n)
        rec["dice_pos"].append(len(dice_log))
        rec["game_id"].append(g)
        afters.append(pack52(np.stack([exe(env.board, m).tensor.numpy() for m in env.legal_moves]))
                      if n else np.zeros((0, 52), np.int8))
        a = int(rng.randint(0, n)) if n else 0
        obs, reward, done, info = env.step(a if n else None)
        rec["action"].append(a)
        rec["reward"].append(float(reward))
        rec["done"].append(bool(done))
        rec["passed"].append("turn passed" in info.get("info", ""))
        rec["winner"].append(int(info["winner"]) if "winner" in info else -1)
        rec["game_score"].append(int(info.get("game_score", 0)))
        rec["obs_after"].append(obs.numpy().copy())
        rec["scores"].append([env.player_scores[Player.PLAYER1], env.player_scores[Player.PLAYER2]])
        if done:
            g += 1
            obs = env.reset()                   # vec_bg_env.py:35-36
    out = {k: np.array(v) for k, v in rec.items()}
    out["dice"] = np.array(dice_log, np.int8)
    out["after"] = np.concatenate(afters, 0)
    out["seed"] = seed
    return out


# ----------------------------------------------------------------------------- synthetic boards

def random_board(rng: np.random.RandomState, kind: str) -> np.ndarray:
    """Random *valid-looking* position: 15 checkers per side, no shared points."""
    b = np.zeros((4, 24), np.int8)
    pts = np.arange(24)
    rng.shuffle(pts)
    n1 = rng.randint(1, 8)
    n2 = rng.randint(1, 8)
    p1_pts, p2_pts = pts[:n1], pts[n1:n1 + n2]
    if kind == "bearoff1":      # P1 all home, P2 anywhere
        p1_pts = np.array([p for p in range(18, 24) if rng.rand() < 0.6] or [23])
        p2_pts = np.array([p for p in pts if p not in set(p1_pts)][:n2])
    if kind == "bearoff2":
        p2_pts = np.array([p for p in range(0, 6) if rng.rand() < 0.6] or [0])
        p1_pts = np.array([p for p in pts if p not in set(p2_pts)][:n1])
    if kind == "bearoff_both":
        p1_pts = np.array([p for p in range(18, 24) if rng.rand() < 0.6] or [23])
        p2_pts = np.array([p for p in range(0, 6) if rng.rand() < 0.6] or [0])
    for side, spts in ((0, p1_pts), (1, p2_pts)):
        left = 15
        off = 0
        bar = 0
        if kind.startswith("bearoff") and ((side == 0 and kind != "bearoff2") or (side == 1 and kind != "bearoff1")):
            off = rng.randint(0, 14)
        elif kind == "bar":
            bar = rng.randint(1, 4) if rng.rand() < 0.8 else 0
            off = 0
        else:
            bar = rng.randint(0, 3) if rng.rand() < 0.3 else 0
            off = rng.randint(0, 6) if rng.rand() < 0.2 else 0
        left -= off + bar
        b[2, side] = bar
        b[3, side] = off
        w = rng.dirichlet(np.ones(len(spts)) * 0.7)
        c = np.floor(w * left).astype(int)
        c[0] += left - c.sum()
        for p, k in zip(spts, c):
            b[side, p] = k
        if left > 0 and b[side].sum() == 0:
            b[side, spts[0]] = left
    return b


def constructed_cases():
    """Hand-built positions from SURVEY.md section 3.3 / 4 (Q1 families, bear-off, blocked bar, partial doubles)."""
    cases = []

    def mk(p1=None, p2=None, bar=(0, 0), off=(0, 0)):
        b = np.zeros((4, 24), np.int8)
        for k, v in (p1 or {}).items():
            b[0, k] = v
        for k, v in (p2 or {}).items():
            b[1, k] = v
        b[2, 0], b[2, 1] = bar
        b[3, 0], b[3, 1] = off
        return b
    # Q1 (i): bar entry, larger die alone suppresses smaller-first play
    cases.append((mk({10: 1}, {6: 2, 11: 2}, bar=(1, 0)), 0, 6, 1))
    cases.append((mk({10: 1}, {6: 2, 11: 2}, bar=(1, 0)), 0, 1, 6))
    # Q1 (ii): last checker comes home
    cases.append((mk({0: 2, 12: 5, 17: 4, 19: 4}, {1: 6, 2: 1, 5: 3, 7: 1}, off=(0, 4)), 1, 6, 5))
    cases.append((mk({0: 2, 12: 5, 17: 4, 19: 4}, {1: 6, 2: 1, 5: 3, 7: 1}, off=(0, 4)), 1, 5, 6))
    # partial doubles
    cases.append((mk({0: 1}, {9: 2}), 0, 3, 3))
    # bear-off P1 / P2 mirror
    cases.append((mk({20: 2, 22: 1}, {0: 15}, off=(12, 0)), 0, 6, 3))
    cases.append((mk({23: 15}, {3: 2, 1: 1}, off=(0, 12)), 1, 6, 3))
    # bar blocked everywhere
    cases.append((mk({12: 14}, {0: 2, 1: 2, 2: 2, 3: 2, 4: 2, 5: 2, 6: 3}, bar=(1, 0)), 0, 3, 5))
    cases.append((mk({12: 14}, {0: 2, 1: 2, 2: 2, 3: 2, 4: 2, 5: 2, 6: 3}, bar=(1, 0)), 0, 4, 4))
    # two on the bar, one entry open
    cases.append((mk({12: 13}, {0: 2, 1: 2, 2: 1, 3: 2, 4: 2, 5: 2, 6: 4}, bar=(2, 0)), 0, 3, 5))
    cases.append((mk({12: 13}, {0: 2, 1: 2, 2: 1, 3: 2, 4: 2, 5: 2, 6: 4}, bar=(2, 0)), 0, 3, 3))
    cases.append((mk({3: 4, 2: 2, 1: 2, 0: 7}, {12: 11, 20: 2}, bar=(0, 2)), 1, 4, 4))
    # game over for the mover
    cases.append((mk({}, {5: 15}, off=(15, 0)), 0, 3, 4))
    # bear-off with contact (opponent blot/point in the home board)
    cases.append((mk({18: 3, 19: 3, 21: 4, 23: 3}, {20: 1, 22: 2, 5: 11}, off=(2, 0), bar=(0, 1)), 0, 2, 1))
    cases.append((mk({18: 3, 19: 3, 21: 4, 23: 3}, {20: 1, 22: 2, 5: 11}, off=(2, 0), bar=(0, 1)), 0, 2, 2))
    # big doubles: spread-out checkers (many distinct plays)
    cases.append((mk({0: 1, 1: 1, 2: 1, 3: 1, 4: 1, 5: 1, 6: 1, 7: 1, 8: 1, 9: 1, 10: 1, 11: 1, 12: 1, 13: 1, 14: 1},
                     {23: 15}), 0, 1, 1))
    cases.append((mk({0: 2, 2: 2, 4: 2, 6: 2, 8: 2, 10: 2, 12: 3}, {23: 2, 21: 2, 19: 11}), 0, 2, 2))
    cases.append((mk({0: 2, 2: 2, 4: 2, 6: 2, 8: 2, 10: 2, 12: 3}, {23: 2, 21: 2, 19: 11}), 0, 3, 3))
    # invalid-count board: P1 has a single checker in its home board (not 15 => NORMAL state, conditions.py:147)
    cases.append((mk({20: 1}, {5: 15}), 0, 6, 5))
    cases.append((mk({20: 1}, {5: 15}), 0, 4, 4))
    return cases


# ----------------------------------------------------------------------------- main

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=8)
    ap.add_argument("--games", type=int, default=6, help="games per trajectory worker")
    ap.add_argument("--allroll-positions", type=int, default=160)
    ap.add_argument("--synthetic", type=int, default=1200)
    args = ap.parse_args()
    t0 = time.time()
    pool = Pool(args.procs)

    # (1) initial-position table: 21 rolls x 2 players
    torch, gapm, IB, exe, Player = _ref()
    init = IB.initial_board(torch.device("cpu")).tensor.numpy()
    boards, players, dice = [], [], []
    for p in (0, 1):
        for a in range(1, 7):
            for b in range(a, 7):
                boards.append(init); players.append(p); dice.append((a, b))
    boards, players, dice = np.array(boards), np.array(players, np.int8), np.array(dice, np.int8)
    c, o, a, ns = run_cases(pool, boards, players, dice)
    np.savez_compressed(os.path.join(HERE, "initial_table.npz"), boards=pack52(boards), players=players, dice=dice,
                        counts=c, offsets=o, after=a, nsub=ns)
    print("initial table", c[:21].tolist(), f"{time.time()-t0:.0f}s", flush=True)

    # (2) trajectories (random vs random through the reference env), one cap-500 set and one tiny-cap set (Q5)
    jobs = [(1000 + i, args.games, 500) for i in range(args.procs)] + [(2000, 2, 6)]
    trajs = pool.map(_trajectory_job, jobs, chunksize=1)
    for t, j in zip(trajs, jobs):
        name = f"traj_seed{j[0]}_cap{j[2]}.npz"
        t = dict(t)
        t["board"] = pack52(t["board"])
        t["obs_after"] = t["obs_after"].astype(np.float32)
        np.savez_compressed(os.path.join(HERE, name), max_legal=j[2], **t)
    nsteps = sum(len(t["action"]) for t in trajs)
    print("trajectories", len(trajs), "steps", nsteps, f"{time.time()-t0:.0f}s", flush=True)

    # (3) all-21-rolls expansion of sampled trajectory positions, both players to move
    rng = np.random.RandomState(7)
    pos = np.concatenate([t["board"] for t in trajs[:-1]], 0)
    sel = rng.choice(len(pos), size=min(args.allroll_positions, len(pos)), replace=False)
    boards, players, dice = [], [], []
    for s in sel:
        mover = int(rng.randint(0, 2))
        for a_ in range(1, 7):
            for b_ in range(a_, 7):
                boards.append(pos[s]); players.append(mover)
                dice.append((a_, b_) if rng.rand() < 0.5 else (b_, a_))
    boards, players, dice = np.array(boards), np.array(players, np.int8), np.array(dice, np.int8)
    c, o, a, ns = run_cases(pool, boards, players, dice)
    np.savez_compressed(os.path.join(HERE, "allrolls.npz"), boards=pack52(boards), players=players, dice=dice,
                        counts=c, offsets=o, after=a, nsub=ns)
    print("allrolls", len(boards), "max n", c.max(), f"{time.time()-t0:.0f}s", flush=True)

    # (4) constructed + synthetic adversarial families
    cons = constructed_cases()
    boards = [x[0] for x in cons]; players = [x[1] for x in cons]; dice = [(x[2], x[3]) for x in cons]
    kinds = ["any", "bar", "bearoff1", "bearoff2", "bearoff_both"]
    for i in range(args.synthetic):
        k = kinds[i % len(kinds)]
        boards.append(random_board(rng, k))
        players.append({"bearoff1": 0, "bearoff2": 1}.get(k, int(rng.randint(0, 2))))
        dd = (int(rng.randint(1, 7)), int(rng.randint(1, 7)))
        if i % 3 == 0:
            dd = (dd[0], dd[0])
        dice.append(dd)
    boards, players, dice = np.array(boards), np.array(players, np.int8), np.array(dice, np.int8)
    c, o, a, ns = run_cases(pool, boards, players, dice)
    np.savez_compressed(os.path.join(HERE, "adversarial.npz"), boards=pack52(boards), players=players, dice=dice,
                        counts=c, offsets=o, after=a, nsub=ns, n_constructed=len(cons))
    print("adversarial", len(boards), "max n", c.max(), "constructed counts", c[:len(cons)].tolist(),
          f"{time.time()-t0:.0f}s", flush=True)

    # (5) encoder vectors: batch encoder and single-board encoder, both flags
    from src.ai.batching import get_board_features_batch_from_tensors
    sample = np.concatenate([pos[rng.choice(len(pos), 300, replace=False)].reshape(-1, 4, 24),
                             np.array(boards[:300]).reshape(-1, 4, 24)], 0)
    tb = torch.tensor(sample, dtype=torch.int8)
    f0 = get_board_features_batch_from_tensors(tb, Player.PLAYER1).numpy()
    f1 = get_board_features_batch_from_tensors(tb, Player.PLAYER2).numpy()
    single = np.stack([IB(tb[i]).get_board_features(Player.PLAYER1).numpy() for i in range(40)])
    assert np.array_equal(single, f0[:40])
    np.savez_compressed(os.path.join(HERE, "encode.npz"), boards=pack52(sample), feat_p1=f0, feat_p2=f1,
                        feat_p1_bf16=torch.tensor(f0).to(torch.bfloat16).view(torch.int16).numpy(),
                        feat_p2_bf16=torch.tensor(f1).to(torch.bfloat16).view(torch.int16).numpy())
    print("encode", sample.shape, f"{time.time()-t0:.0f}s", flush=True)

    # (6) terminal / reward classification through env.step on constructed end positions
    from src.environment.backgammon_env import BackgammonEnv
    ends = []

    def end_case(p1, p2, bar, off, player, roll):
        b = np.zeros((4, 24), np.int8)
        for k, v in p1.items():
            b[0, k] = v
        for k, v in p2.items():
            b[1, k] = v
        b[2, 0], b[2, 1] = bar
        b[3, 0], b[3, 1] = off
        env = BackgammonEnv()
        env.board = IB(torch.tensor(b))
        env.current_player = Player(player)
        env.roll_result = list(roll)
        env.update_legal_moves()
        n = len(env.legal_moves)
        obs, reward, done, info = env.step(0)
        ends.append(dict(board=b, player=player, roll=roll, n=n, reward=float(reward), done=bool(done),
                         winner=int(info.get("winner", -1)), score=int(info.get("game_score", 0)),
                         after=env.board.tensor.numpy().copy(), obs=obs.numpy().copy()))
    # P1 bears off last checker: normal / gammon / backgammon (home) / backgammon (bar)
    end_case({23: 1}, {5: 10}, (0, 0), (14, 5), 0, (6, 5))
    end_case({23: 1}, {5: 15}, (0, 0), (14, 0), 0, (6, 5))
    end_case({23: 1}, {5: 14, 20: 1}, (0, 0), (14, 0), 0, (6, 5))
    end_case({23: 1}, {5: 14}, (0, 1), (14, 0), 0, (6, 5))
    end_case({22: 1, 23: 1}, {5: 15}, (0, 0), (13, 0), 0, (2, 2))
    end_case({22: 1, 23: 1}, {5: 15}, (0, 0), (13, 0), 0, (1, 1))      # both off with 4x1? 22->23->off, 23->off
    # P2 mirror
    end_case({18: 10}, {0: 1}, (0, 0), (5, 14), 1, (6, 5))
    end_case({18: 15}, {0: 1}, (0, 0), (0, 14), 1, (3, 1))
    end_case({18: 14, 3: 1}, {0: 1}, (0, 0), (0, 14), 1, (3, 1))
    end_case({18: 14}, {0: 1}, (1, 0), (0, 14), 1, (3, 1))
    # not terminal
    end_case({22: 1, 23: 1}, {5: 15}, (0, 0), (13, 0), 0, (2, 1))
    keys = ends[0].keys()
    np.savez_compressed(os.path.join(HERE, "env_end.npz"),
                        **{k: np.array([e[k] for e in ends]) for k in keys})
    print("env_end", [(e["reward"], e["done"], e["score"]) for e in ends], flush=True)

    # (7) reset protocol with injected dice (backgammon_env.py:78-113)
    resets = []
    for seq in ([(3, 3), (2, 5), (4, 4), (6, 1)], [(5, 2), (1, 3)], [(1, 1), (2, 2), (6, 6), (1, 2), (3, 3), (4, 3)],
                [(4, 6), (6, 4)]):
        env = BackgammonEnv()
        it = iter(seq)

        def roll(env=env, it=it):
            env.roll_result = list(next(it))
        env.roll_dice = roll
        env.reset()
        used = len(seq) - len(list(it))
        resets.append(dict(seq=np.array(seq + [(0, 0)] * (6 - len(seq)), np.int8), seq_len=len(seq), used=used,
                           player=int(env.current_player), roll=np.array(env.roll_result, np.int8),
                           n=len(env.legal_moves)))
    np.savez_compressed(os.path.join(HERE, "reset.npz"), **{k: np.array([r[k] for r in resets]) for k in resets[0]})
    print("reset", [(r["used"], r["player"], r["roll"].tolist(), r["n"]) for r in resets], flush=True)

    # (8) MLP forward (policy_network.py:44-75) with torch.manual_seed(0) default init
    from src.agent.policy_network import BackgammonPolicyNetwork
    torch.manual_seed(0)
    net = BackgammonPolicyNetwork(198, 128, 500)
    x = torch.tensor(np.concatenate([f0[:200], f1[200:400]], 0))
    with torch.no_grad():
        logits, values = net(x)
    sd = {k: v.numpy() for k, v in net.state_dict().items()}
    np.savez_compressed(os.path.join(HERE, "mlp.npz"), x=x.numpy(), values=values.numpy(),
                        logits_head=logits[:, :8].numpy(),
                        fc1_weight=sd["fc1.weight"], fc1_bias=sd["fc1.bias"],
                        value_weight=sd["value_head.weight"], value_bias=sd["value_head.bias"],
                        action_weight_head=sd["action_head.weight"][:8], action_bias_head=sd["action_head.bias"][:8])
    print("mlp", values[:4].tolist(), f"{time.time()-t0:.0f}s", flush=True)
    pool.close()
    print("done", f"{time.time()-t0:.0f}s")


if __name__ == "__main__":
    main()
