"""2-ply golden vectors from a Python restatement ON THE REFERENCE'S OWN PRIMITIVES
(get_all_possible_moves, generate_all_board_features, BackgammonPolicyNetwork, env reward rules).
The reference's own 2-ply (src/moves/expect_minmax.py) is commented-out code, so this pins the build's
definition, not a reference behaviour.   python tests/golden/make_golden_twoply.py"""
import os
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def unpack52(b52):
    b = np.zeros((4, 24), np.int8)
    b[0], b[1] = b52[0:24], b52[24:48]
    b[2, 0:2], b[3, 0:2] = b52[48:50], b52[50:52]
    return b


def job(args):
    b52, me, d0, d1 = args
    import torch
    import ref_loader
    ref_loader.load()
    torch.set_num_threads(1)
    from src.moves.get_all_moves import get_all_possible_moves
    from src.moves.get_all_dice_rolls import get_all_dice_rolls_tensor
    from src.board.immutable_board import ImmutableBoard, execute_full_move_on_board_copy
    from src.ai.batching import generate_all_board_features, get_board_features_batch_from_tensors
    from src.agent.policy_network import BackgammonPolicyNetwork
    from src.environment.backgammon_env import BackgammonEnv
    from src.players.player import Player
    torch.manual_seed(0)
    net = BackgammonPolicyNetwork(198, 128, 500)
    env = BackgammonEnv()

    def win_reward(board, p):           # backgammon_env.py:156-171 via the env's own predicates
        env.board = board
        if env.check_for_backgammon(p):
            return 2.0
        return 1.5 if env.check_for_gammon(p) else 1.0

    rolls, probs = get_all_dice_rolls_tensor()
    me_p, opp_p = Player(me), Player(1 - me)
    root = ImmutableBoard(torch.tensor(unpack52(b52)))
    A = get_all_possible_moves(me_p, root, [d0, d1])
    scores, leaves = [], 0
    for mv in A:
        Ai = execute_full_move_on_board_copy(root, mv)
        if Ai.tensor[3, me].item() == 15:
            scores.append(win_reward(Ai, me_p)); continue
        acc = torch.tensor(0.0)
        for r in range(21):
            roll = [int(rolls[r, 0]), int(rolls[r, 1])]
            R = get_all_possible_moves(opp_p, Ai, roll)
            if not R:
                x = get_board_features_batch_from_tensors(Ai.tensor[None], opp_p)
                with torch.no_grad():
                    vr = net(x)[1][0]
                leaves += 1
            else:
                feats = generate_all_board_features(Ai, opp_p, R, roll)
                with torch.no_grad():
                    vals = net(feats)[1].clone()
                leaves += len(R)
                for j, rm in enumerate(R):
                    Bj = execute_full_move_on_board_copy(Ai, rm)
                    if Bj.tensor[3, 1 - me].item() == 15:
                        vals[j] = win_reward(Bj, opp_p)
                vr = vals.max()
            acc = acc + probs[r] * vr
        scores.append(float(-acc))
    return np.array(scores, np.float32), leaves


def main():
    d = np.load(os.path.join(HERE, "allrolls.npz"))
    a = np.load(os.path.join(HERE, "adversarial.npz"))
    rng = np.random.RandomState(3)
    cases = []
    idx = [i for i in range(len(d["counts"])) if 2 <= d["counts"][i] <= 5]
    for i in rng.choice(idx, 8, replace=False):
        cases.append((d["boards"][i], int(d["players"][i]), int(d["dice"][i, 0]), int(d["dice"][i, 1])))
    # bear-off endgames (terminal afterstates / terminal replies) from the synthetic families
    idx = [i for i in range(20, len(a["counts"])) if 1 <= a["counts"][i] <= 4 and a["boards"][i][50 + a["players"][i]] >= 11]
    for i in idx[:6]:
        cases.append((a["boards"][i], int(a["players"][i]), int(a["dice"][i, 0]), int(a["dice"][i, 1])))
    with Pool(8) as pool:
        res = pool.map(job, cases, chunksize=1)
    counts = np.array([len(r[0]) for r in res], np.int32)
    np.savez_compressed(os.path.join(HERE, "twoply.npz"), boards=np.array([c[0] for c in cases], np.int8),
                        players=np.array([c[1] for c in cases], np.int8), dice=np.array([c[2:4] for c in cases], np.int8),
                        counts=counts, scores=np.concatenate([r[0] for r in res]), leaves=np.array([r[1] for r in res]))
    print("twoply golden:", counts.tolist(), [r[1] for r in res])


if __name__ == "__main__":
    main()
