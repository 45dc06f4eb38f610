"""GPU parity tests: the CUDA path (through the C ABI) against the golden vectors recorded from the
reference and against the CPU oracle on seeded random-play positions.  Integer/byte work: bit-exact,
and afterstate lists are compared IN ORDER (stronger than sorted-set equality)."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def bg():
    import bg_b200
    bg_b200.lib()
    return bg_b200


@pytest.fixture(scope="module")
def O():
    from oracle import bg_oracle
    return bg_oracle


def dev():
    return torch.device("cuda:0")


def _gpu_moves(bg, boards52, players, dice, cap=0):
    c, o, a = bg.legal_moves(torch.as_tensor(boards52).to(dev()), torch.as_tensor(players).to(dev()),
                             torch.as_tensor(dice).to(dev()), max_rows_per_board=cap)
    return c.cpu().numpy(), o.cpu().numpy(), a.cpu().numpy()


@pytest.mark.parametrize("name", ["initial_table", "allrolls", "adversarial"])
def test_movegen_golden_in_order(bg, name):
    d = np.load(os.path.join(G, name + ".npz"))
    c, o, a = _gpu_moves(bg, d["boards"], d["players"], d["dice"])
    assert np.array_equal(c, d["counts"])
    assert np.array_equal(o, d["offsets"])
    assert np.array_equal(a, d["after"])


def _random_play_positions(O, n_games, steps, seed):
    """Positions reached by the oracle's own uniform-random self-play (Philox dice)."""
    boards, players, dice = [], [], []
    for g in range(n_games):
        e = O.Env()
        e.set_philox(seed, g)
        e.reset()
        for t in range(steps):
            s = e.state()
            boards.append(s["board"]); players.append(s["player"]); dice.append(s["roll"])
            n = s["n_legal"]
            a = O.philox_action(seed + 1, g, t, n) if n else 0
            _, done, _ = e.step(a)
            if done:
                e.reset()
    return np.array(boards, np.int8), np.array(players, np.int8), np.array(dice, np.int8)


def test_movegen_vs_oracle_random_play(bg, O):
    boards, players, dice = _random_play_positions(O, 64, 400, seed=11)      # 25,600 positions
    counts, offsets, after = O.legal_moves_batch(boards, players, dice)
    c, o, a = _gpu_moves(bg, O.pack52(boards), players, dice)
    assert np.array_equal(c, counts)
    assert np.array_equal(a, O.pack52(after))
    # sorted-set form (the north star's stated criterion) holds a fortiori; check the cap form too
    c2, o2, a2 = _gpu_moves(bg, O.pack52(boards), players, dice, cap=6)
    assert np.array_equal(c2, counts)
    want = np.concatenate([O.pack52(after[offsets[i]:offsets[i] + min(6, counts[i])]) for i in range(len(counts))])
    assert np.array_equal(a2, want)


def test_movegen_all_rolls_both_dice_orders(bg, O):
    boards, players, _ = _random_play_positions(O, 8, 120, seed=5)
    B = len(boards)
    rolls = np.array([(a, b) for a in range(1, 7) for b in range(1, 7)], np.int8)
    bb = np.repeat(boards, 36, 0); pp = np.repeat(players, 36, 0); dd = np.tile(rolls, (B, 1))
    counts, offsets, after = O.legal_moves_batch(bb, pp, dd)
    c, o, a = _gpu_moves(bg, O.pack52(bb), pp, dd)
    assert np.array_equal(c, counts) and np.array_equal(a, O.pack52(after))


def test_movegen_empty_and_ragged(bg):
    z = torch.zeros((0, 52), dtype=torch.int8, device=dev())
    c, o, a = bg.legal_moves(z, torch.zeros(0, dtype=torch.int8, device=dev()), torch.zeros((0, 2), dtype=torch.int8, device=dev()))
    assert c.numel() == 0 and o.tolist() == [0] and a.shape[0] == 0
    # single position, odd batch sizes
    d = np.load(os.path.join(G, "allrolls.npz"))
    for B in (1, 3, 33, 257):
        c, o, a = _gpu_moves(bg, d["boards"][:B], d["players"][:B], d["dice"][:B])
        assert np.array_equal(c, d["counts"][:B]) and np.array_equal(a, d["after"][: d["offsets"][B]])


def test_movegen_bad_input_is_flagged(bg):
    b = bg.initial_board52(4, dev())
    b[2, 3] = 17                                          # count outside 0..15
    pl = torch.zeros(4, dtype=torch.int8, device=dev())
    dc = torch.tensor([[3, 1]] * 4, dtype=torch.int8, device=dev())
    with pytest.raises(bg.BgError, match="BAD_INPUT"):
        bg.legal_moves(b, pl, dc)
    c, o, a = bg.legal_moves(b, pl, dc, check_status=False)
    assert c.tolist() == [16, 16, -1, 16]


def test_encode_golden(bg):
    d = np.load(os.path.join(G, "encode.npz"))
    b = torch.as_tensor(d["boards"]).to(dev())
    for flag, key in ((0, "feat_p1"), (1, "feat_p2")):
        f = bg.encode(b, flag, dtype=torch.float32)
        assert np.array_equal(f.cpu().numpy(), d[key])
        h = bg.encode(b, flag, dtype=torch.bfloat16)
        assert h.shape[1] == 208
        assert np.array_equal(h[:, :198].contiguous().view(torch.int16).cpu().numpy(), d[key + "_bf16"])
        assert int(h[:, 198:].abs().sum().item()) == 0
        h256 = bg.encode(b, flag, dtype=torch.bfloat16, ld=256)
        assert torch.equal(h256[:, :208], h) and int(h256[:, 208:].abs().sum().item()) == 0
    flags = torch.arange(b.shape[0], device=dev()) % 2
    f = bg.encode(b, flags.to(torch.int8), dtype=torch.float32).cpu().numpy()
    assert np.array_equal(f[0::2], d["feat_p1"][0::2]) and np.array_equal(f[1::2], d["feat_p2"][1::2])


def test_encode_vs_oracle_large(bg, O):
    boards, players, _ = _random_play_positions(O, 16, 300, seed=3)
    f = bg.encode(torch.as_tensor(O.pack52(boards)).to(dev()), torch.as_tensor(players).to(dev()))
    assert np.array_equal(f.cpu().numpy(), O.encode(boards, players))


def _replay(bg, paths, cap):
    ds = [np.load(p) for p in paths]
    N, L = len(ds), max(len(d["dice"]) for d in ds)
    T = max(len(d["action"]) for d in ds)
    dice = np.ones((N, L + 4 * T + 64, 2), np.int8); dice[:, :, 1] = 2      # padding rolls (1,2) past the recording
    for i, d in enumerate(ds):
        dice[i, : len(d["dice"])] = d["dice"]
    env = bg.B200BackgammonVecEnv(num_envs=N, match_length=3, max_legal_moves=cap, device=dev(), check_every=0)
    env.set_dice_stream(torch.as_tensor(dice))
    env.reset()
    offs = [0] * N
    for t in range(T):
        boards, players, rolls = env.boards52.cpu().numpy(), env.players.cpu().numpy(), env.dice.cpu().numpy()
        counts, starts = env.legal_counts.cpu().numpy(), env.legal_starts.cpu().numpy()
        rows = env.after52.cpu().numpy()
        draws = env.draws.cpu().numpy()
        acts = np.zeros(N, np.int32)
        live = []
        for i, d in enumerate(ds):
            if t >= len(d["action"]):
                continue
            live.append(i)
            assert np.array_equal(boards[i], d["board"][t]), (i, t)
            assert players[i] == d["player"][t] and rolls[i].tolist() == d["roll"][t].tolist(), (i, t)
            assert counts[i] == d["n_legal"][t] and draws[i] == d["dice_pos"][t], (i, t)
            n = int(counts[i])
            assert np.array_equal(rows[starts[i]:starts[i] + n], d["after"][offs[i]:offs[i] + n]), (i, t)
            offs[i] += n
            acts[i] = d["action"][t]
        obs, rew, done, infos = env.step(torch.as_tensor(acts))
        rew, done = rew.cpu().numpy(), done.cpu().numpy()
        sc = env.scores.cpu().numpy()
        for i in live:
            d = ds[i]
            assert rew[i] == d["reward"][t] and bool(done[i]) == bool(d["done"][t]), (i, t)
            inf = infos[i]
            assert inf["current_player"] == d["player"][t]
            assert ("turn passed" in inf.get("info", "")) == bool(d["passed"][t])
            assert inf.get("winner", -1) == d["winner"][t] and inf.get("game_score", 0) == d["game_score"][t]
            if not done[i]:
                assert sc[i].tolist() == d["scores"][t].tolist()
    env.check_status()
    return T


def test_env_trajectory_replay_golden(bg):
    paths = sorted(glob.glob(os.path.join(G, "traj_seed*_cap500.npz")))
    assert _replay(bg, paths, 500) > 100


def test_env_trajectory_replay_golden_truncated_cap(bg):
    # Q5: the env keeps the FIRST max_legal_moves plays in the reference's list order
    assert _replay(bg, sorted(glob.glob(os.path.join(G, "traj_seed*_cap6.npz"))), 6) > 50


def test_env_philox_vs_oracle(bg, O):
    N, T, seed, aseed = 384, 160, 0x5EED, 99
    env = bg.B200BackgammonVecEnv(num_envs=N, device=dev(), seed=seed, stream_base=1000, check_every=0)
    obs = env.reset()
    oenvs = []
    for g in range(N):
        e = O.Env()
        e.set_philox(seed, 1000 + g)
        e.reset()
        oenvs.append(e)
    games = 0
    for t in range(T):
        acts = env.random_actions(aseed, t)
        a_host = acts.cpu().numpy()
        if t % 8 == 0 or t == T - 1:
            b, p, dc = env.boards52.cpu().numpy(), env.players.cpu().numpy(), env.dice.cpu().numpy()
            cnt = env.legal_counts_true.cpu().numpy()
            ob = obs.cpu().numpy()
            for g in range(N):
                s = oenvs[g].state()
                assert np.array_equal(O.pack52(s["board"])[0], b[g]) and s["player"] == p[g], (g, t)
                assert s["roll"].tolist() == dc[g].tolist() and s["n_legal_true"] == cnt[g], (g, t)
                assert np.array_equal(oenvs[g].observation(), ob[g])
        obs, rew, done, infos = env.step(acts)
        rew, done = rew.cpu().numpy(), done.cpu().numpy()
        for g in range(N):
            n = oenvs[g].state()["n_legal"]
            a = O.philox_action(aseed, 1000 + g, t, n) if n else 0
            assert a == a_host[g]
            r, dn, info = oenvs[g].step(a)
            assert r == rew[g] and dn == bool(done[g]), (g, t)
            if dn:
                games += 1
                oenvs[g].reset()
    env.check_status()
    assert games > 100


def test_terminal_rewards_golden(bg):
    d = np.load(os.path.join(G, "env_end.npz"))
    N = len(d["reward"])
    env = bg.B200BackgammonVecEnv(num_envs=N, device=dev(), check_every=0)
    env.reset()
    env.load_positions(torch.as_tensor(d["board"]), torch.as_tensor(d["player"]), torch.as_tensor(d["roll"]))
    assert env.legal_counts.cpu().tolist() == d["n"].tolist()
    obs, rew, done, infos = env.step(torch.zeros(N, dtype=torch.int32))
    assert rew.cpu().tolist() == d["reward"].tolist() and done.cpu().tolist() == d["done"].tolist()
    assert env.winner.cpu().tolist() == d["winner"].tolist() and env.game_score.cpu().tolist() == d["score"].tolist()


def test_env_invalid_action_and_masks(bg):
    env = bg.B200BackgammonVecEnv(num_envs=4, device=dev(), seed=1)
    env.reset()
    before = env.boards52.clone()
    counts = env.legal_counts.clone()
    masks = env.get_action_masks()
    assert masks.shape == (4, 500) and torch.equal(masks.sum(1).to(torch.int32), counts)
    feats = env.get_legal_board_features()
    assert feats.shape == (4, 500, 198)
    rag = env.afterstate_features(torch.float32)
    st = env.legal_starts.cpu().tolist()
    for g in range(4):
        n = int(counts[g])
        assert torch.equal(feats[g, :n], rag[st[g]:st[g] + n]) and float(feats[g, n:].abs().sum()) == 0.0
    obs, rew, done, infos = env.step(torch.tensor([499, 0, 499, 0], dtype=torch.int32))
    assert rew.cpu().tolist() == [-1.0, 0.0, -1.0, 0.0] and not done.any()
    assert infos[0]["info"] == "Invalid action" and "info" not in infos[1]
    assert torch.equal(env.boards52[0], before[0]) and torch.equal(env.boards52[2], before[2])
    assert torch.equal(env.legal_counts[[0, 2]], counts[[0, 2]])


def test_c_abi_rejects_bad_arguments(bg):
    L = bg.lib()
    assert L.bg_movegen_count(None, None, None, 4, None, None, None, 0, None) == -1
    assert b"null" in L.bg_last_error()
    assert L.bg_encode_f32(None, None, 0, 4, None, None, 198, None) == -1
    assert L.bg_encode_bf16(None, None, 0, 0, None, None, 100, None) == -1
    assert L.bg_movegen_count(None, None, None, 0, None, None, None, 0, None) == 0      # empty batch is fine


def test_output_overflow_is_flagged(bg):
    env = bg.B200BackgammonVecEnv(num_envs=4096, device=dev(), rows_per_game=1, check_every=0)
    with pytest.raises(bg.BgError, match="OUTPUT_OVERFLOW"):
        env.reset()


def test_fused_afterstate_features_match_encoder(bg):
    """K1's fused feature output (all three tiers) == K3 on the same rows."""
    d = np.load(os.path.join(G, "adversarial.npz"))
    sel = np.argsort(-d["counts"])[:200]                       # the heaviest positions: tiers 1 and 2 included
    sel = np.concatenate([sel, np.arange(0, len(d["counts"]), 5)])
    env = bg.B200BackgammonVecEnv(num_envs=len(sel), device=dev(), rows_per_game=128, max_legal_moves=5000, check_every=0)
    env.boards52.copy_(torch.as_tensor(d["boards"][sel]).to(dev()))
    env.players.copy_(torch.as_tensor(d["players"][sel]).to(dev()))
    env.dice.copy_(torch.as_tensor(d["dice"][sel]).to(dev()))
    env._refresh_legal_moves(with_features=True)
    env.check_status()
    assert env.legal_counts_true.cpu().tolist() == d["counts"][sel].tolist()
    n = env.total_rows()
    want = bg.encode(env.after52[:n], env.row_players[:n], dtype=torch.bfloat16)
    assert torch.equal(env.after_feats[:n].view(torch.int16), want.view(torch.int16))
