"""N2 on the GPU: bg_gae against a numpy restatement of compute_returns / GAE and the reference's own returns,
and the self-play trainer end to end (policy kernel -> K2 -> K1 -> returns -> update -> re-packed weights)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ppo.npz")


def _gae_numpy(r, d, v, last, gamma, lam):
    T, N = r.shape
    ret, adv = np.zeros_like(r), np.zeros_like(r)
    nv = last.astype(np.float32).copy()
    a = np.zeros(N, np.float32)
    g, l = np.float32(gamma), np.float32(lam)
    for t in range(T - 1, -1, -1):
        nd = (1 - d[t]).astype(np.float32)
        delta = (r[t] + (g * nd) * nv) + (-v[t])
        a = delta + ((g * l) * nd) * a
        adv[t], ret[t] = a, a + v[t]
        nv = v[t]
    return ret, adv


def test_gae_bit_exact_and_reference_returns():
    from bg_b200.ppo import discounted_returns
    rng = np.random.default_rng(0)
    T, N = 37, 1000
    r = (rng.random((T, N)) < 0.05).astype(np.float32) * rng.choice([1.0, 1.5, 2.0], (T, N)).astype(np.float32)
    d = (r > 0).astype(np.uint8)
    v = rng.standard_normal((T, N)).astype(np.float32)
    last = rng.standard_normal(N).astype(np.float32)
    cu = lambda a: torch.tensor(a).cuda()
    for lam, use_v in ((1.0, False), (0.95, True)):
        ret, adv = discounted_returns(cu(r), cu(d), cu(v) if use_v else None, cu(last) if use_v else None, 0.99, lam)
        wr, wa = _gae_numpy(r, d, v if use_v else np.zeros_like(v), last if use_v else np.zeros(N, np.float32), 0.99, lam)
        assert np.array_equal(ret.cpu().numpy(), wr) and np.array_equal(adv.cpu().numpy(), wa)
    # the reference's compute_returns over its interleaved memory (golden from src/agent/ppo_agent.py)
    g = np.load(GOLDEN)
    M = g["rewards"].shape[0]
    ret, _ = discounted_returns(cu(g["rewards"]).view(M, 1), cu(g["dones"]).view(M, 1), None, None, 0.99, 1.0)
    np.testing.assert_allclose(ret.view(-1).cpu().numpy(), g["returns_interleaved"], rtol=1e-6, atol=1e-6)


def test_policy_kernel_matches_reference_select_action():
    """log-probs and values the reference stored for its own sampled actions (golden) vs the fused kernel's logits/values"""
    import bg_b200
    g = np.load(GOLDEN)
    sd0 = {k[3:]: torch.tensor(g[k]) for k in g.files if k.startswith("w0.")}
    net = bg_b200.PolicyValueNet.from_state_dict(sd0, "cuda:0")
    b52 = bg_b200.to_board52(torch.tensor(g["boards"]).cuda())
    _, _, v, logits = net.act(b52, torch.tensor(g["players"]).cuda(), torch.tensor(g["counts"]).cuda(), want_logits=True)
    counts = torch.tensor(g["counts"]).cuda().long()
    mask = (torch.arange(500, device="cuda")[None, :] < counts[:, None]).float()
    lsm = torch.log_softmax(logits + (mask + 1e-45).log(), -1)
    got = lsm.gather(1, torch.tensor(g["actions"]).cuda().long()[:, None])[:, 0]
    assert (got.cpu() - torch.tensor(g["logp"])).abs().max().item() < 2e-2      # bf16 operands vs the f32 reference
    assert (v.cpu() - torch.tensor(g["values"])).abs().max().item() < 2e-2


def test_trainer_smoke_and_checkpoint_roundtrip(tmp_path):
    import bg_b200
    from bg_b200.ppo import PPOConfig, PPOTrainer, evaluate_vs_random
    dev = torch.device("cuda:0")
    env = bg_b200.B200BackgammonVecEnv(num_envs=512, device=dev, seed=3, check_every=0)
    env.reset()
    net = bg_b200.PolicyValueNet.random_init(dev, seed=0)
    w_before = net.state_dict()
    tr = PPOTrainer(env, net, PPOConfig(t_horizon=24, num_epochs=2), seed=5)
    hist = tr.train(3, log=None)
    assert len(hist) == 3 and all(np.isfinite([h["policy_loss"], h["value_loss"], h["entropy"], h["total_loss"]]).all() for h in hist)
    w_after = net.state_dict()
    assert any(not torch.equal(w_before[k], w_after[k]) for k in w_before)
    # actions the trainer took were legal (no invalid-action penalty) and passes only where there was no legal play
    assert float(tr.buf.rewards.min().item()) >= 0.0
    assert bool(((tr.buf.flags & 1).bool() == (tr.buf.counts == 0)).all())
    # checkpoint in the reference's format (ppo_agent.py:377-403) and back
    path = os.path.join(tmp_path, "ppo_backgammon.pth")
    torch.save(net.state_dict(), path)
    sd = torch.load(path)
    assert set(sd) == {"fc1.weight", "fc1.bias", "action_head.weight", "action_head.bias", "value_head.weight", "value_head.bias"}
    net2 = bg_b200.PolicyValueNet.from_state_dict(sd, dev)
    a1 = net.act(env.boards52, env.players, env.legal_counts, seed=1, step=0)
    a2 = net2.act(env.boards52, env.players, env.legal_counts, seed=1, step=0)
    assert all(torch.equal(x, y) for x, y in zip(a1, a2))
    ev = evaluate_vs_random(net, num_games=256, device=dev)
    assert ev["games"] == 256 and 0.0 <= ev["win_rate"] <= 1.0


def test_fused_loss_kernel_matches_torch_chain():
    """bg_ppo_loss_grad (one pass over the logits) == the torch restatement of ppo_agent.py:271-299, values and gradients"""
    import bg_b200
    from bg_b200.ppo import ppo_loss
    g = np.load(GOLDEN)
    sd0 = {k[3:]: torch.tensor(g[k]).cuda() for k in g.files if k.startswith("w0.")}
    t = lambda k, dt: torch.tensor(g[k]).to(dt).cuda()
    x, counts, actions, logp, values = t("obs", torch.float32), t("counts", torch.int32), t("actions", torch.int32), t("logp", torch.float32), t("values", torch.float32)
    ret = t("returns_interleaved", torch.float32)
    ret = (ret - ret.mean()) / (ret.std() + 1e-5)
    adv = ret - values
    logp = logp + 0.3 * torch.randn_like(logp)                   # push some ratios outside the clip range
    for autocast, tol in ((False, 2e-5), (True, 3e-3)):
        grads = []
        for fused in (False, True):
            p = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
            loss, pl, vl, ent = ppo_loss(p, x, counts, actions, logp, ret, adv, 0.25, 0.5, 0.15, autocast=autocast, fused=fused)
            loss.backward()
            grads.append((loss.item(), pl.item(), vl.item(), ent.item(), {k: v.grad.clone() for k, v in p.items()}))
        a, b = grads
        assert max(abs(a[i] - b[i]) for i in range(4)) < tol, (a[:4], b[:4])
        for k in a[4]:
            scale = a[4][k].abs().max().item() + 1e-8
            assert (a[4][k] - b[4][k]).abs().max().item() < (1e-4 if not autocast else 3e-2) * scale, k
    # want: the reference's own epoch losses (lr = 0 golden) through the fused path as well
    p = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    loss, pl, vl, ent = ppo_loss(p, x, counts, actions, t("logp", torch.float32), ret, adv, 0.25, 0.5, float(g["entropy_coef"]), autocast=True, fused=True)
    got = np.array([pl.item(), vl.item(), ent.item(), loss.item()])
    assert np.abs(got - g["loss_lr0"]).max() < 5e-3, (got, g["loss_lr0"])


def test_manual_update_matches_autograd():
    """ManualUpdate (explicit GEMMs + bg_ppo_loss_grad with the value head as column 500, biases folded) == torch autograd
    over the torch restatement of ppo_agent.py:268-305 under bf16 autocast: losses, every gradient, and a whole update()"""
    from bg_b200.ppo import ManualUpdate, PPOConfig, PPOLearner, ppo_loss
    g = np.load(GOLDEN)
    sd0 = {k[3:]: torch.tensor(g[k]).cuda() for k in g.files if k.startswith("w0.")}
    t = lambda k, dt: torch.tensor(g[k]).to(dt).cuda()
    obs, counts, actions, logp, values = t("obs", torch.float32), t("counts", torch.int32), t("actions", torch.int32), t("logp", torch.float32), t("values", torch.float32)
    B = obs.shape[0]
    x = torch.zeros((B, 208), dtype=torch.bfloat16, device="cuda")
    x[:, :198] = obs
    ret = t("returns_interleaved", torch.float32)
    retn = (ret - ret.mean()) / (ret.std() + 1e-5)
    adv = retn - values
    logp2 = logp + 0.3 * torch.randn_like(logp)
    p = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    loss, pl, vl, ent = ppo_loss(p, x, counts, actions, logp2, retn, adv, 0.25, 0.5, 0.15, autocast=True, fused=False)
    loss.backward()
    want = {k: v.grad.clone() for k, v in p.items()}
    grads = {k: torch.zeros_like(v) for k, v in sd0.items()}
    xm = x.clone(); xm[:, ManualUpdate.ONE_COL] = 1.0
    st = ManualUpdate(torch.device("cuda:0")).epoch(sd0, grads, xm, counts, actions, logp2, adv.contiguous(), retn.contiguous(), 0.25, 0.5, 0.15)
    got = st.cpu().numpy()
    assert np.abs(got - np.array([pl.item(), vl.item(), ent.item(), loss.item()])).max() < 3e-3, got
    for k in want:
        scale = want[k].abs().max().item() + 1e-8
        assert (want[k] - grads[k]).abs().max().item() < 3e-2 * scale, (k, (want[k] - grads[k]).abs().max().item(), scale)
    # a whole four-epoch update through both paths ends at (nearly) the same weights
    outs = []
    for manual in (False, True):
        L = PPOLearner(sd0, "cuda:0", PPOConfig(manual_backward=manual))
        L.update(x.clone(), counts, actions, logp, values, ret)
        outs.append((L.state_dict(), L.last))
    for k in sd0:
        moved = (outs[0][0][k] - sd0[k]).abs().mean().item()          # (Adam steps are sign-like: compare means, not maxima)
        assert (outs[0][0][k] - outs[1][0][k]).abs().mean().item() < 0.25 * moved + 1e-6, k
    for key in ("policy_loss", "value_loss", "entropy", "total_loss"):
        assert abs(outs[0][1][key] - outs[1][1][key]) < 5e-3, (key, outs[0][1], outs[1][1])


def _loss_reference(logits, values, counts, actions, old_logp, adv, returns, eps, vc, ec):
    """ppo_agent.py:271-299 in f64 torch, with autograd for the gradients w.r.t. logits and values"""
    lg = logits.double().clone().requires_grad_(True)
    v = values.double().clone().requires_grad_(True)
    slot = torch.arange(500, device=lg.device)[None, :]
    masked = lg + torch.where(slot < counts[:, None], 0.0, float(np.log(np.float32(1e-45))))
    lp = torch.log_softmax(masked, -1)
    new = lp.gather(1, actions.long()[:, None])[:, 0]
    r = torch.exp(new - old_logp.double())
    pl = -torch.min(r * adv.double(), torch.clamp(r, 1 - eps, 1 + eps) * adv.double()).mean()
    vl = ((v - returns.double()) ** 2).mean()
    ent = -(lp.exp() * lp).sum(-1).mean()
    (pl + vc * vl - ec * ent).backward()
    return pl.item(), vl.item(), ent.item(), lg.grad, v.grad


@pytest.mark.parametrize("ld,bf16", [(500, False), (512, False), (512, True), (504, True)])
def test_loss_kernel_edge_rows(ld, bf16):
    """bg_ppo_loss_grad on synthetic rows: passes (count 0: softmax over all 500 slots), 1 / 128 / 129 / 500 legal slots, a stored
    action outside the legal prefix, ratios on both sides of the clip range; separate value array and value-in-column-500 mode;
    f32 and bf16 logits, ld 500 / 504 / 512 -- against an f64 torch restatement of ppo_agent.py:271-299."""
    from bg_b200._lib import check, lib
    torch.manual_seed(ld + bf16)
    B = 777
    dev = "cuda"
    counts = torch.tensor([0, 1, 2, 17, 127, 128, 129, 256, 257, 384, 385, 499, 500], device=dev, dtype=torch.int32).repeat(B // 13 + 1)[:B].contiguous()
    actions = (torch.rand(B, device=dev) * counts.clamp(min=1)).to(torch.int32)
    actions[5] = 300; counts[5] = 40                                   # stored action outside the legal prefix: literal arithmetic
    actions[0] = 123                                                   # a pass: any slot
    dt = torch.bfloat16 if bf16 else torch.float32
    logits = torch.zeros((B, ld), device=dev, dtype=dt)
    logits[:, :500] = (2.0 * torch.randn(B, 500, device=dev)).to(dt)
    values = torch.randn(B, device=dev)
    if ld > 500:
        logits[:, 500] = values.to(dt)
        values_col = logits[:, 500].float()
    adv, ret = torch.randn(B, device=dev), torch.randn(B, device=dev)
    lsm = torch.log_softmax(logits[:, :500].float() + torch.where(torch.arange(500, device=dev)[None, :] < counts[:, None], 0.0, -103.27893), -1)
    old = (lsm.gather(1, actions.long()[:, None])[:, 0] + 0.4 * torch.randn(B, device=dev)).contiguous()
    eps, vc, ec = 0.25, 0.5, 0.15
    for value_in_col in ([False, True] if ld > 500 else [False]):
        v_used = values_col if value_in_col else values
        pl, vl, ent, dl_want, dv_want = _loss_reference(logits[:, :500].float(), v_used, counts, actions, old, adv, ret, eps, vc, ec)
        dlogits = torch.full((B, ld), 7.0, device=dev, dtype=dt)
        dvalues = torch.zeros(B, device=dev)
        dbias, sums = torch.zeros(512, device=dev), torch.zeros(3, device=dev)
        check(lib().bg_ppo_loss_grad(logits.data_ptr(), int(bf16), ld, None if value_in_col else values.data_ptr(), counts.data_ptr(),
                                     actions.data_ptr(), old.data_ptr(), adv.data_ptr(), ret.data_ptr(), B, eps, vc, ec,
                                     dlogits.data_ptr(), dvalues.data_ptr(), dbias.data_ptr(), sums.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream), "bg_ppo_loss_grad")
        got = (sums / B).cpu().numpy()
        assert np.abs(got - np.array([pl, vl, ent])).max() < 2e-5 * (1 + abs(pl) + vl + ent), (got, pl, vl, ent)
        tol = 1e-2 if bf16 else 1e-5                                   # bf16: the output rounding of dlogits
        scale = dl_want.abs().max().item()
        assert (dlogits[:, :500].double() - dl_want).abs().max().item() < tol * scale
        assert (dvalues.double() - dv_want).abs().max().item() < 1e-6
        if ld > 500:
            pad = dlogits[:, 500:].float()
            if value_in_col:
                assert (pad[:, 0].double() - dv_want).abs().max().item() < (1e-2 if bf16 else 1e-6) * dv_want.abs().max().item() + 1e-9
                assert float(pad[:, 1:].abs().max()) == 0.0
                assert abs(dbias[500].item() - dv_want.sum().item()) < 1e-5
            else:
                assert float(pad.abs().max()) == 0.0
        assert (dbias[:500].double() - dl_want.sum(0)).abs().max().item() < (2e-5 if not bf16 else 2e-5) * B * scale
        # BG_LOSS_DLOGITS_PREZEROED: the same numbers into a buffer the caller has zeroed
        dl2 = torch.zeros((B, ld), device=dev, dtype=dt)
        dv2, db2, s2 = torch.zeros(B, device=dev), torch.zeros(512, device=dev), torch.zeros(3, device=dev)
        check(lib().bg_ppo_loss_grad(logits.data_ptr(), int(bf16) | 2, ld, None if value_in_col else values.data_ptr(), counts.data_ptr(),
                                     actions.data_ptr(), old.data_ptr(), adv.data_ptr(), ret.data_ptr(), B, eps, vc, ec,
                                     dl2.data_ptr(), dv2.data_ptr(), db2.data_ptr(), s2.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream), "bg_ppo_loss_grad")
        assert torch.equal(dl2, dlogits) and torch.equal(dv2, dvalues)
        assert (db2 - dbias).abs().max().item() <= 1e-6 * (1 + dbias.abs().max().item()) and (s2 - sums).abs().max().item() <= 1e-3
