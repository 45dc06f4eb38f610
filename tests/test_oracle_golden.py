"""Pin the CPU oracle (oracle/bg_oracle.c) against outputs of the UNMODIFIED reference.

The fixtures under tests/golden/ were produced by tests/golden/make_golden.py, which imports
/root/reference and records get_all_possible_moves / BackgammonEnv / encoder / MLP outputs.
Afterstate lists are compared IN ORDER (the reference's legal_moves order), which is stronger
than the canonically-sorted set equality the north star asks for.
"""
import glob
import os

import numpy as np
import pytest

from oracle import bg_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def _check_movegen(npz):
    d = np.load(npz)
    boards = O.unpack52(d["boards"])
    counts, offsets, after = O.legal_moves_batch(boards, d["players"], d["dice"])
    assert np.array_equal(counts, d["counts"]), npz
    assert np.array_equal(O.pack52(after), d["after"]), npz
    return len(boards), int(counts.max())


@pytest.mark.parametrize("name", ["initial_table", "allrolls", "adversarial"])
def test_movegen_matches_reference_in_order(name):
    n, mx = _check_movegen(os.path.join(G, name + ".npz"))
    assert n > 0


def test_initial_table_known_answers():
    # legal-play counts of the initial position per roll (both players)
    d = np.load(os.path.join(G, "initial_table.npz"))
    expect = [42, 15, 16, 14, 8, 10, 75, 17, 18, 8, 14, 73, 17, 9, 14, 52, 9, 14, 4, 7, 11]
    assert d["counts"][:21].tolist() == expect and d["counts"][21:].tolist() == expect


def test_submove_counts_match_reference():
    d = np.load(os.path.join(G, "adversarial.npz"))
    boards = O.unpack52(d["boards"])
    for i in range(0, len(boards), 7):
        after, nsub, subs = O.legal_moves(boards[i], d["players"][i], d["dice"][i, 0], d["dice"][i, 1], with_moves=True)
        lo, hi = d["offsets"][i], d["offsets"][i + 1]
        assert np.array_equal(nsub, d["nsub"][lo:hi])


def test_encoder_bit_exact():
    d = np.load(os.path.join(G, "encode.npz"))
    boards = O.unpack52(d["boards"])
    assert np.array_equal(O.encode(boards, 0), d["feat_p1"])
    assert np.array_equal(O.encode(boards, 1), d["feat_p2"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(G, "traj_seed*.npz"))))
def test_trajectory_replay(path):
    """Replay the reference env's random games under its recorded dice: every position, roll,
    legal-play list (ordered), reward, done flag, winner, score and observation must match."""
    d = np.load(path)
    env = O.Env(match_length=3, max_legal_moves=int(d["max_legal"]))
    env.set_external_dice(d["dice"])
    env.reset()
    T = len(d["action"])
    off = 0
    for t in range(T):
        s = env.state()
        assert np.array_equal(O.pack52(s["board"])[0], d["board"][t]), (path, t)
        assert s["player"] == d["player"][t] and s["roll"].tolist() == d["roll"][t].tolist()
        assert s["n_legal"] == d["n_legal"][t], (path, t)
        assert env.dice_consumed() == d["dice_pos"][t]
        n = s["n_legal"]
        assert np.array_equal(O.pack52(env.afterstates()), d["after"][off:off + n]), (path, t)
        off += n
        r, done, info = env.step(int(d["action"][t]))
        assert r == d["reward"][t] and done == bool(d["done"][t])
        assert info["passed"] == bool(d["passed"][t])
        assert info["winner"] == d["winner"][t] and info["game_score"] == d["game_score"][t]
        # observation returned by step(): on a win it is the terminal board with the WINNER to move
        assert np.array_equal(env.observation(), d["obs_after"][t])
        assert env.state()["scores"].tolist() == d["scores"][t].tolist()
        if done:
            env.reset()                                     # vec_bg_env.py:35-36


def test_terminal_rewards():
    d = np.load(os.path.join(G, "env_end.npz"))
    for i in range(len(d["reward"])):
        env = O.Env()
        env.set_external_dice(np.array([[1, 2]] * 8, np.int8))
        env.set_position(d["board"][i], int(d["player"][i]), int(d["roll"][i][0]), int(d["roll"][i][1]))
        assert env.state()["n_legal"] == d["n"][i]
        r, done, info = env.step(0)
        assert r == d["reward"][i] and done == bool(d["done"][i])
        assert info["winner"] == d["winner"][i] and info["game_score"] == d["score"][i]
        assert np.array_equal(env.state()["board"], d["after"][i])
        s, rw = O.win_score(d["after"][i], int(d["player"][i]))
        assert s == d["score"][i] and rw == d["reward"][i]


def test_reset_protocol():
    d = np.load(os.path.join(G, "reset.npz"))
    for i in range(len(d["used"])):
        env = O.Env()
        env.set_external_dice(d["seq"][i][: int(d["seq_len"][i])])
        env.reset()
        s = env.state()
        assert env.dice_consumed() == d["used"][i]
        assert s["player"] == d["player"][i] and s["roll"].tolist() == d["roll"][i].tolist()
        assert s["n_legal"] == d["n"][i]
        assert np.array_equal(s["board"], O.initial_board())


def test_mlp_value_matches_torch_reference():
    d = np.load(os.path.join(G, "mlp.npz"))
    v = O.mlp_value(d["x"], d["fc1_weight"], d["fc1_bias"], d["value_weight"], float(d["value_bias"][0]))
    assert np.max(np.abs(v - d["values"])) < 2e-6


def test_pack_roundtrip():
    d = np.load(os.path.join(G, "allrolls.npz"))
    assert np.array_equal(O.pack52(O.unpack52(d["boards"])), d["boards"])


def test_twoply_matches_restatement_on_reference_primitives():
    """bg_twoply (SURVEY 8(c) definition) vs the Python restatement built from the reference's own
    get_all_possible_moves / generate_all_board_features / BackgammonPolicyNetwork (make_golden_twoply.py)."""
    p = os.path.join(G, "twoply.npz")
    if not os.path.exists(p):
        pytest.skip("twoply golden not generated")
    g = np.load(p)
    m = np.load(os.path.join(G, "mlp.npz"))
    w = (m["fc1_weight"], m["fc1_bias"], m["value_weight"], float(m["value_bias"][0]))
    off = 0
    for i in range(len(g["counts"])):
        sc, best, leaves = O.twoply(O.unpack52(g["boards"][i])[0], int(g["players"][i]), int(g["dice"][i, 0]),
                                    int(g["dice"][i, 1]), *w)
        n = int(g["counts"][i])
        assert len(sc) == n and leaves == g["leaves"][i]
        assert np.max(np.abs(sc - g["scores"][off:off + n])) < 5e-6
        off += n
