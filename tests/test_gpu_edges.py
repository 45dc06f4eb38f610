"""Edge cases of the entry points added after the first parity suite: one-call K1+K3 (serial / overlapped) on the
heaviest fixtures, other feature strides, empty and tiny batches, argument errors, policy kernel without a mask."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def dev():
    return torch.device("cuda:0")


def test_update_legal_plays_heavy_fixtures_serial_and_overlapped():
    """all three K1 tiers + both encoder launches: golden counts/afterstates in order, features == encoder, both modes"""
    import bg_b200 as bg
    d = np.load(os.path.join(G, "adversarial.npz"))
    sel = np.concatenate([np.argsort(-d["counts"])[:150], np.arange(0, len(d["counts"]), 7)])
    for overlap in (False, True):
        env = bg.B200BackgammonVecEnv(num_envs=len(sel), device=dev(), rows_per_game=160, max_legal_moves=5000, check_every=0)
        env.boards52.copy_(torch.as_tensor(d["boards"][sel]).to(dev()))
        env.players.copy_(torch.as_tensor(d["players"][sel]).to(dev()))
        env.dice.copy_(torch.as_tensor(d["dice"][sel]).to(dev()))
        env.update_legal_plays(obs=True, features=True, overlap=overlap)
        torch.cuda.synchronize()
        env.check_status()
        assert env.legal_counts_true.cpu().tolist() == d["counts"][sel].tolist()
        rows = env.after52.cpu().numpy()
        starts = env.legal_starts.cpu().numpy()
        for k, i in enumerate(sel[:60]):
            lo, hi = d["offsets"][i], d["offsets"][i + 1]
            assert np.array_equal(rows[starts[k]:starts[k] + (hi - lo)], d["after"][lo:hi]), (overlap, int(i))
        n = env.total_rows()
        want = bg.encode(env.after52[:n], env.row_players[:n], dtype=torch.bfloat16)
        assert torch.equal(env.after_feats[:n].view(torch.int16), want.view(torch.int16))
        assert torch.equal(env.obs_f32, bg.encode(env.boards52, env.players))


def test_encode_bf16_other_strides_and_tiny_batches():
    import bg_b200 as bg
    from oracle import bg_oracle as O
    d = np.load(os.path.join(G, "encode.npz"))
    b52 = torch.as_tensor(d["boards"]).to(dev())
    flags = (torch.arange(b52.shape[0], device=dev()) % 2).to(torch.int8)
    ref = bg.encode(b52, flags, dtype=torch.float32)
    gold = np.where((flags.cpu().numpy() == 0)[:, None], d["feat_p1"], d["feat_p2"])       # the reference's own vectors
    assert np.array_equal(ref.cpu().numpy(), gold)
    for ld in (200, 208, 256, 512):
        for B in (1, 2, 17, 129, b52.shape[0]):
            out = bg.encode(b52[:B], flags[:B], dtype=torch.bfloat16, ld=ld)
            assert out.shape == (B, ld)
            assert torch.equal(out[:, :198], ref[:B].to(torch.bfloat16)) and bool((out[:, 198:] == 0).all())
    assert bg.encode(b52[:0], 0, dtype=torch.bfloat16).shape[0] == 0


def test_env_single_game_and_odd_sizes():
    import bg_b200 as bg
    from oracle import bg_oracle as O
    for N in (1, 5, 33):
        env = bg.B200BackgammonVecEnv(num_envs=N, device=dev(), seed=21, check_every=1)
        obs = env.reset()
        assert obs.shape == (N, 198)
        for t in range(30):
            obs, rew, done, infos = env.step(env.random_actions(4, t), with_features=True)
            assert obs.shape == (N, 198) and rew.shape == (N,) and done.dtype == torch.bool and len(infos) == N
            b52, pl, dc = env.boards52.cpu().numpy(), env.players.cpu().numpy(), env.dice.cpu().numpy()
            counts, offsets, after = O.legal_moves_batch(O.unpack52(b52), pl, dc)
            assert np.array_equal(env.legal_counts_true.cpu().numpy(), counts)
            assert np.array_equal(obs.cpu().numpy(), O.encode(O.unpack52(b52), pl))
        assert isinstance(infos[0], dict) and "current_player" in infos[0]


def test_new_entry_points_reject_bad_arguments():
    import bg_b200 as bg
    L = bg.lib()
    assert L.bg_update_legal_plays(None, None, None, 4, 0, None, 0, None, None, None, None, None, None, None, 0, None, 208, None, 198,
                                   None, None, None, None) == -1
    assert L.bg_update_legal_plays(None, None, None, 0, 0, None, 0, None, None, None, None, None, None, None, 0, None, 208, None, 198,
                                   None, None, None, None) == 0
    assert L.bg_policy_sample(None, None, 0, 4, None, None, None, None, None, None, 0.0, 0, 0, 0, 0, None, None, None, None, None, 0, None) == -1
    assert L.bg_policy_sample(None, None, 0, 0, None, None, None, None, None, None, 0.0, 0, 0, 0, 0, None, None, None, None, None, 0, None) == 0
    assert L.bg_gae(None, None, None, None, 4, 4, 0.99, 1.0, None, None, None) == -1
    assert L.bg_ppo_loss_grad(None, 1, 499, None, None, None, None, None, None, 4, 0.2, 0.5, 0.01, None, None, None, None, None) == -1
    assert L.bg_twoply_replies_values(None, None, 4, None, 0, None, None, None, None, None, None, 0, None, None, None, 0.0, None, None,
                                      None, None) == -1


def test_policy_without_mask_and_single_row():
    import bg_b200 as bg
    net = bg.PolicyValueNet.random_init(dev(), seed=4)
    b = bg.initial_board52(1, dev())
    a, lp, v, logits = net.act(b, 0, None, greedy=True, want_logits=True)
    assert int(a[0]) == int(logits[0].argmax()) and abs(float(lp[0]) - float(torch.log_softmax(logits[0], -1).max())) < 1e-4
    counts = torch.tensor([3], dtype=torch.int32, device=dev())
    a, lp, v = net.act(b, 0, counts, greedy=True)
    assert int(a[0]) == int(logits[0, :3].argmax())
    # value head agrees with K4 on the same position
    assert abs(float(v[0]) - float(net.values(b, 0)[0])) < 1e-4


def test_host_step_buffers_mirror_the_device_results():
    """step(actions, host=HostStepBuffers): rewards / dones / new legal counts reach pinned host memory behind K1 (while K3
    is still running) and equal the device tensors of the same step; trajectories are unchanged by the early copy."""
    import bg_b200
    dev = torch.device("cuda:0")
    envs = [bg_b200.B200BackgammonVecEnv(num_envs=3000, device=dev, seed=11, check_every=0) for _ in range(2)]
    [e.reset() for e in envs]
    host = bg_b200.HostStepBuffers(envs[0])
    h_acts = torch.empty(3000, dtype=torch.int32).pin_memory()
    rng = np.random.default_rng(5)
    counts = envs[0].legal_counts.cpu().numpy()
    for t in range(150):
        h_acts.numpy()[:] = (rng.random(3000) * counts).astype(np.int32)
        obs, rew, done, _ = envs[0].step(h_acts, with_features=True, host=host)
        obs2, rew2, done2, _ = envs[1].step(h_acts.clone())
        host.wait()
        counts = host.legal_counts.numpy().copy()
        assert np.array_equal(host.rewards.numpy(), rew.cpu().numpy()) and np.array_equal(host.dones.numpy(), done.cpu().numpy())
        assert np.array_equal(counts, envs[0].legal_counts.cpu().numpy())
        assert torch.equal(obs, obs2) and torch.equal(rew, rew2) and torch.equal(done, done2)
    assert float(envs[0].rewards.min()) >= 0.0
    [e.check_status() for e in envs]


def test_reference_layout_inputs_set_dice_and_csr_accessors():
    """legal_moves / encode accept the reference's (B,4,24) tensors; set_dice overrides the roll and refreshes the legal
    plays; afterstates_csr() = the env's blocks gathered in game order == legal_moves() of the same positions."""
    import bg_b200
    dev = torch.device("cuda:0")
    env = bg_b200.B200BackgammonVecEnv(num_envs=700, device=dev, seed=21, check_every=0)
    env.reset()
    for t in range(40):
        env.step(env.random_actions(3, t))
    b424 = env.boards()                                           # (N,4,24) reference layout
    c1, o1, a1 = bg_b200.legal_moves(b424, env.players, env.dice, max_rows_per_board=500)
    c2, o2, a2 = bg_b200.legal_moves(env.boards52, env.players, env.dice, max_rows_per_board=500)
    assert torch.equal(c1, c2) and torch.equal(o1, o2) and torch.equal(a1, a2)
    assert torch.equal(bg_b200.encode(b424, env.players), bg_b200.encode(env.boards52, env.players))
    off, rows = env.afterstates_csr()
    assert torch.equal(off, o2) and torch.equal(rows, a2) and torch.equal(off, env.legal_offsets)
    dice = torch.randint(1, 7, (700, 2), dtype=torch.int8, device=dev)
    env.set_dice(dice)
    c3, o3, a3 = bg_b200.legal_moves(env.boards52, env.players, dice, max_rows_per_board=500)
    off, rows = env.afterstates_csr()
    assert torch.equal(env.legal_counts, c3.clamp(max=500)) and torch.equal(off, o3) and torch.equal(rows, a3)
    with pytest.raises(bg_b200.BgError):
        env.set_dice(dice[:5])


def test_step_random_equals_random_actions_then_step():
    """bg_env_step_random (policy inside K2) == bg_random_actions + bg_env_step, and the slab row counter that lives in K1's
    workspace header is zeroed by K1 itself"""
    import bg_b200 as bg
    a = bg.B200BackgammonVecEnv(num_envs=3000, device=dev(), seed=11, check_every=0)
    b = bg.B200BackgammonVecEnv(num_envs=3000, device=dev(), seed=11, check_every=0)
    a.reset(); b.reset()
    acts = torch.empty(3000, dtype=torch.int32, device=dev())
    for t in range(60):
        a.step_device(a.random_actions(5, t))
        b.step_random_device(5, t, actions_out=acts)
        b._refresh_legal_moves()
    for name in ("boards52", "players", "dice", "scores", "draws", "legal_counts", "rewards", "dones_u8"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert int(a.alloc_rows.item()) == int(a.legal_counts.sum().item()) == int(b.alloc_rows.item())
    a.check_status(); b.check_status()


def test_record_state_copies_boards_players_counts():
    """bg_record_state (the rollout's one-launch record of a step's inputs) == three plain copies"""
    import ctypes as C
    import bg_b200
    from bg_b200._lib import lib, check
    env = bg_b200.B200BackgammonVecEnv(num_envs=1000, device="cuda:0", seed=5, check_every=0)
    env.reset()
    for t in range(9):
        env.step_device(env.random_actions(3, t))
    b = torch.full((1000, 52), -7, dtype=torch.int8, device="cuda:0")
    p = torch.full((1000,), -7, dtype=torch.int8, device="cuda:0")
    c = torch.full((1000,), -7, dtype=torch.int32, device="cuda:0")
    st = env._state()
    check(lib().bg_record_state(C.byref(st), b.data_ptr(), p.data_ptr(), c.data_ptr(), torch.cuda.current_stream().cuda_stream), "bg_record_state")
    torch.cuda.synchronize()
    assert torch.equal(b, env.boards52) and torch.equal(p, env.players) and torch.equal(c, env.legal_counts)
