"""Racecheck substitute for K1's overflow tiers (movegen_team.cu): the cross-warp ordered dedupe (tentative PEND|tid
claim, atomicMin take-over, commit after the barrier) is a data-race-shaped protocol whose RESULT must not depend on
scheduling.  compute-sanitizer is not available on the pool, so the heavy positions (doubles whose levels exceed tier
0's 128 boards -- tiers 1 and 2) are generated 50 times under every team size (tier 1: 128 / 256 / 512 threads per
position, tier 2: 512 / 1024) and in shuffled batch order (different CTA <-> position assignment, different neighbours
on the SM), and every run must reproduce the oracle's afterstate lists bit for bit, in order."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_team_tiers_bit_identical_under_every_team_size_and_order():
    import bg_b200 as bg
    from oracle import bg_oracle as O
    dev = torch.device("cuda:0")
    L = bg.lib()
    # heavy positions: random-play positions x the six doubles, keep those with many plays; plus the heaviest fixtures
    env = bg.B200BackgammonVecEnv(num_envs=4096, device=dev, seed=77, check_every=0)
    env.reset()
    for t in range(70):
        env.step_device(env.random_actions(5, t))
    b = env.boards52.cpu().numpy(); p = env.players.cpu().numpy()
    boards = np.repeat(b, 6, 0); players = np.repeat(p, 6, 0)
    dice = np.tile(np.array([[k, k] for k in range(1, 7)], np.int8), (len(b), 1))
    c, _, _ = bg.legal_moves(torch.as_tensor(boards).to(dev), torch.as_tensor(players).to(dev), torch.as_tensor(dice).to(dev))
    c = c.cpu().numpy()
    heavy = np.argsort(-c)[:700]
    assert int((c[heavy] > 128).sum()) > 50 and c[heavy].max() > 200       # tiers 1 / 2 are really exercised
    d = np.load(os.path.join(G, "adversarial.npz"))
    top = np.argsort(-d["counts"])[:40]
    boards = np.concatenate([boards[heavy], d["boards"][top]]); players = np.concatenate([players[heavy], d["players"][top]])
    dice = np.concatenate([dice[heavy], d["dice"][top]])
    want_c, want_o, want_a = O.legal_moves_batch(O.unpack52(boards), players, dice)
    want_a = O.pack52(want_a)
    n = len(boards)
    rng = np.random.default_rng(0)
    runs = 0
    try:
        for rep in range(9):
            for mid, big in ((128, 512), (256, 512), (512, 1024), (256, 1024), (128, 1024), (0, 0)):
                if runs >= 50:
                    break
                assert L.bg_set_team_threads(mid, big) == 0
                perm = rng.permutation(n)
                gc, go, ga = bg.legal_moves(torch.as_tensor(boards[perm]).to(dev), torch.as_tensor(players[perm]).to(dev),
                                            torch.as_tensor(dice[perm]).to(dev))
                gc, go, ga = gc.cpu().numpy(), go.cpu().numpy(), ga.cpu().numpy()
                assert np.array_equal(gc, want_c[perm]), (rep, mid, big)
                # rows of position perm[k] are ga[go[k]:go[k+1]]: gather them back into the oracle's order and compare everything
                back = np.concatenate([ga[go[k]:go[k + 1]] for k in np.argsort(perm)])
                assert np.array_equal(back, want_a), (rep, mid, big)
                runs += 1
    finally:
        L.bg_set_team_threads(0, 0)
    assert runs == 50
