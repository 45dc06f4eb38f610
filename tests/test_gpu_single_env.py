"""GPU tests of the single-game BackgammonEnv adapter (the other half of the seam, SURVEY 8(b):
train_single.py:21,34,39-45 over backgammon_env.py:78-191) against golden vectors recorded from the UNMODIFIED
reference (tests/golden/make_golden_single.py, make_golden.py).  Bit-exact: boards, rolls (numpy dice stream
reproduced from env.seed), legal-play counts, features, observations -- including the TERMINAL observation with the
winner to move -- rewards, dones, infos, and the reset-on-step of a finished game."""
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


def _b52(env):
    return env._v.boards52[0].cpu().numpy()


def test_single_env_matches_reference_including_terminal_and_step_after_done(bg):
    d = np.load(os.path.join(G, "single_env.npz"))
    env = bg.BackgammonEnv(match_length=int(d["match_length"]), max_legal_moves=500, device="cuda:0")
    env.seed(int(d["seed"]))
    obs = env.reset()
    assert np.array_equal(obs.cpu().numpy(), d["obs0"])
    T = len(d["action"])
    games = 0
    seen_terminal = seen_step_after_done = 0
    for t in range(T):
        was_over = bool(d["was_over"][t])
        assert env.game_over == was_over, t
        assert np.array_equal(_b52(env), d["board"][t]), t
        assert int(env.current_player) == d["player"][t], t
        assert list(env.roll_result) == d["roll"][t].tolist(), t
        assert len(env.legal_moves) == d["n_legal"][t] and float(env.action_mask.sum()) == d["mask_sum"][t], t
        n = int(d["n_legal"][t])
        f0, f1 = int(d["feat_offsets"][t]), int(d["feat_offsets"][t + 1])
        assert tuple(env.legal_board_features.shape) == (500, 198)
        assert np.array_equal(env.legal_board_features[:n].cpu().numpy(), d["feats"][f0:f1]), t
        assert float(env.legal_board_features[n:].abs().sum()) == 0.0
        a = int(d["action"][t])
        obs, reward, done, info = env.step(None if a < 0 else a)
        assert np.array_equal(obs.cpu().numpy(), d["obs_after"][t]), t        # terminal observation included
        assert float(reward) == d["reward"][t] and bool(done) == bool(d["done"][t]), t
        assert int(info["current_player"]) == d["info_player"][t], t
        assert ("turn passed" in info.get("info", "")) == bool(d["passed"][t]), t
        assert int(info.get("winner", -1)) == d["winner"][t] and info.get("game_score", 0) == d["game_score"][t], t
        assert env.game_over == bool(d["game_over_after"][t]) and env.match_over == bool(d["match_over_after"][t]), t
        sc = env.player_scores
        assert [sc[bg.Player.PLAYER1], sc[bg.Player.PLAYER2]] == d["scores_after"][t].tolist(), t
        assert int(env.current_player) == d["player_after"][t], t
        if was_over:
            seen_step_after_done += 1
            assert float(reward) == 0.0 and done and "winner" not in info
        if done and not was_over:
            seen_terminal += 1
            games += 1
            if games % 2 == 0:
                env.reset()
    assert seen_terminal == 6 and seen_step_after_done == 3
    env.close()


def test_single_env_replays_vec_golden_trajectory(bg):
    """The same recorded reference trajectory the vec env replays (make_golden.py::_trajectory_job), through the
    single-game class with its own numpy dice: env.seed(seed) must reproduce the reference's dice."""
    p = sorted(glob.glob(os.path.join(G, "traj_seed*_cap500.npz")))[0]
    d = np.load(p)
    env = bg.BackgammonEnv(match_length=3, max_legal_moves=500, device="cuda:0")
    env.seed(int(d["seed"]))
    env.reset()
    off = 0
    for t in range(min(250, len(d["action"]))):
        assert np.array_equal(_b52(env), d["board"][t]), t                      # (the trajectory goldens store packed board52 rows)
        assert list(env.roll_result) == d["roll"][t].tolist() and int(env.current_player) == d["player"][t], t
        n = int(d["n_legal"][t])
        assert len(env.legal_moves) == n
        assert np.array_equal(env._afterstates.cpu().numpy(), d["after"][off:off + n]), t
        off += n
        obs, reward, done, info = env.step(int(d["action"][t]) if n else None)
        assert np.array_equal(obs.cpu().numpy(), d["obs_after"][t]), t
        assert float(reward) == d["reward"][t] and bool(done) == bool(d["done"][t]), t
        if done:
            env.reset()


def test_train_single_loop_shape(bg):
    """train_single.py:16-75's loop body, verbatim in structure, against the adapter (a stub agent in place of the PPO
    agent): reads env.current_player / env.action_mask, steps with None on a pass, stops at done."""
    Player = bg.Player
    env = bg.BackgammonEnv(device="cuda:0", dice="philox", seed=7)
    rng = np.random.RandomState(0)

    class Agent:
        memory = []

        def select_action(self, observation, action_mask):
            assert tuple(observation.shape) == (198,) and tuple(action_mask.shape) == (500,)
            n = int(action_mask.sum().item())
            self.memory.append({})
            return [int(rng.randint(0, n))]
    agent = Agent()
    wins = 0
    for episode in range(3):
        observation = env.reset()
        done, t, episode_reward = False, 0, 0.0
        while not done and t < 2000:
            t += 1
            current_player = env.current_player
            assert current_player in (Player.PLAYER1, Player.PLAYER2)
            action_mask = env.action_mask
            if action_mask.sum() == 0:
                observation, reward, done, info = env.step(None)
            else:
                action = agent.select_action(observation, action_mask)[0]
                observation, reward, done, info = env.step(action)
                agent.memory[-1]["reward"], agent.memory[-1]["done"] = reward, done
            episode_reward += reward.item()
            if done and "winner" in info and "game_score" in info:
                wins += 1 if info["winner"] == Player.PLAYER1 else 0
                assert info["game_score"] in (1, 2, 3)
        assert done and episode_reward in (1.0, 1.5, 2.0)
    env.close()


def test_reset_protocol_golden_on_gpu(bg):
    """reset.npz (opening protocol under injected dice, backgammon_env.py:78-113) through bg_env_reset on the device."""
    d = np.load(os.path.join(G, "reset.npz"))
    n = len(d["used"])
    L = int(d["seq"].shape[1])
    dice = np.ones((n, L + 8, 2), np.int8); dice[:, :, 1] = 2
    for i in range(n):
        k = int(d["seq_len"][i])
        dice[i, :k] = d["seq"][i][:k]
    env = bg.B200BackgammonVecEnv(num_envs=n, device="cuda:0", check_every=0)
    env.set_dice_stream(torch.as_tensor(dice))
    env.reset()
    assert np.array_equal(env.draws.cpu().numpy(), d["used"])
    assert np.array_equal(env.players.cpu().numpy(), d["player"])
    assert np.array_equal(env.dice.cpu().numpy(), d["roll"])
    assert np.array_equal(env.legal_counts_true.cpu().numpy(), d["n"])
    init = bg.initial_board52(n, "cuda:0")
    assert bool((env.boards52 == init).all())
    env.check_status()


def test_policy_act_growing_batch(bg):
    """the cached scratch of PolicyValueNet.act must follow bg_policy_workspace_bytes (256 then 400 rows)."""
    dev = torch.device("cuda:0")
    env = bg.B200BackgammonVecEnv(num_envs=400, device=dev, seed=5, check_every=0)
    env.reset()
    net = bg.PolicyValueNet.random_init(dev, seed=0)
    a1, _, v1 = net.act(env.boards52[:256], env.players[:256], env.legal_counts[:256], seed=1, step=0)
    a2, _, v2 = net.act(env.boards52, env.players, env.legal_counts, seed=1, step=0)
    torch.cuda.synchronize()
    assert torch.equal(a1, a2[:256]) and torch.allclose(v1, v2[:256], atol=1e-6)
    a3, _, _ = net.act(env.boards52[:300], env.players[:300], env.legal_counts[:300], seed=1, step=0)
    assert torch.equal(a3, a2[:300])
