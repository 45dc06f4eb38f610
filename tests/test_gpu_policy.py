"""N1: the fused policy/value kernel (bg_policy_sample) against a torch f32 restatement of
BackgammonPPOAgent.select_action (ppo_agent.py:138-191) on the oracle's features."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
MASK_LOG = float(torch.log(torch.tensor(1e-45, dtype=torch.float32)))


def _positions(n, seed=5):
    import bg_b200
    dev = torch.device("cuda:0")
    env = bg_b200.B200BackgammonVecEnv(num_envs=n, device=dev, seed=seed, check_every=0)
    env.reset()
    for t in range(40):
        env.step_device(env.random_actions(3, t))
    env.check_status()
    return env


def _reference(net, env, bf16_operands):
    """masked log-softmax, values, logits in torch f32 from the oracle's feature encoding"""
    from oracle import bg_oracle as O
    b52, pl = env.boards52.cpu().numpy(), env.players.cpu().numpy()
    x = torch.from_numpy(O.encode(O.unpack52(b52), pl)).cuda()
    p = net.params
    w1, wa = p["fc1.weight"], p["action_head.weight"]
    if bf16_operands:
        w1, wa = w1.bfloat16().float(), wa.bfloat16().float()
    h = torch.relu(x @ w1.T + p["fc1.bias"])
    v = h @ p["value_head.weight"].reshape(-1) + p["value_head.bias"]
    hh = h.bfloat16().float() if bf16_operands else h
    logits = hh @ wa.T + p["action_head.bias"]
    n = env.legal_counts.long()
    mask = (torch.arange(500, device="cuda")[None, :] < n[:, None]).float()
    masked = logits + (mask + 1e-45).log()
    return logits, torch.log_softmax(masked, -1), v, masked


def test_policy_logits_values_logp():
    import bg_b200
    env = _positions(1000)          # not a multiple of 128: ragged last tile
    net = bg_b200.PolicyValueNet.random_init("cuda:0", seed=1)
    # bigger action-head weights so that the softmax is far from uniform
    net.params["action_head.weight"].mul_(6.0); net.sync()
    a, lp, v, logits = net.act(env.boards52, env.players, env.legal_counts, seed=11, step=3, want_logits=True)
    torch.cuda.synchronize()
    ref_logits, ref_lsm, ref_v, _ = _reference(net, env, bf16_operands=True)
    # same bf16 operands, f32 accumulate; a hidden unit that lands on the other side of a bf16 rounding boundary
    # moves a logit by |wa| * ulp_bf16(h) (action-head weights are scaled x6 here)
    assert (logits - ref_logits).abs().max().item() < 1e-2
    assert (v - ref_v).abs().max().item() < 1e-3
    n = env.legal_counts.long()
    al = a.long()
    ok = torch.where(n > 0, al < n, al < 500) & (al >= 0)
    assert bool(ok.all()), "sampled an illegal slot"
    assert (lp - ref_lsm.gather(1, al[:, None])[:, 0]).abs().max().item() < 1e-2
    # against the f32 network (the reference's arithmetic): stated tolerance 2e-2
    f_logits, f_lsm, f_v, _ = _reference(net, env, bf16_operands=False)
    assert (logits - f_logits).abs().max().item() < 5e-2
    assert (v - f_v).abs().max().item() < 2e-2


def test_policy_greedy_is_argmax():
    import bg_b200
    env = _positions(512, seed=9)
    net = bg_b200.PolicyValueNet.random_init("cuda:0", seed=2)
    net.params["action_head.weight"].mul_(4.0); net.sync()
    a, lp, v, logits = net.act(env.boards52, env.players, env.legal_counts, greedy=True, want_logits=True)
    n = env.legal_counts.long()
    mask = (torch.arange(500, device="cuda")[None, :] < n[:, None]).float()
    want = (logits + (mask + 1e-45).log()).argmax(-1)
    assert torch.equal(a.long(), want)


def test_policy_sampling_distribution_and_reproducibility():
    import bg_b200
    env = _positions(128, seed=2)
    net = bg_b200.PolicyValueNet.random_init("cuda:0", seed=3)
    net.params["action_head.weight"].mul_(5.0); net.sync()
    a1, _, _ = net.act(env.boards52, env.players, env.legal_counts, seed=5, stream_base=77, step=9)
    a2, _, _ = net.act(env.boards52, env.players, env.legal_counts, seed=5, stream_base=77, step=9)
    assert torch.equal(a1, a2)
    # sharding invariance: the second half alone with its stream_base gives the same actions
    a3, _, _ = net.act(env.boards52[64:], env.players[64:], env.legal_counts[64:], seed=5, stream_base=77 + 64, step=9)
    assert torch.equal(a1[64:], a3)
    # empirical distribution of 4,000 draws vs softmax probabilities, a few positions (incl. a pass if there is one)
    _, ref_lsm, _, _ = _reference(net, env, bf16_operands=True)
    probs = ref_lsm.exp()
    T = 4000
    draws = torch.stack([net.act(env.boards52, env.players, env.legal_counts, seed=123, step=t)[0] for t in range(T)]).long()
    n = env.legal_counts.cpu().numpy()
    rows = list(range(8)) + [int(i) for i in np.nonzero(n == 0)[0][:2]]
    for r in rows:
        k = int(n[r]) if n[r] > 0 else 500
        emp = torch.bincount(draws[:, r], minlength=500)[:k].float().cpu().numpy()
        exp = probs[r, :k].cpu().numpy() * T
        keep = exp > 5
        if keep.sum() < 2:
            continue
        chi2 = float(((emp[keep] - exp[keep]) ** 2 / exp[keep]).sum())
        dof = int(keep.sum()) - 1
        assert chi2 < dof + 6 * (2 * dof) ** 0.5 + 10, (r, chi2, dof)


def test_policy_class_paths_logp_and_values():
    """The row-class paths (no logits requested: A1 = 1..32 legal slots on 8-slot column warps, A2 = 33..128, B = passes / more):
    the log-prob returned for the sampled action and the value must equal the reference's masked log-softmax / value head for
    every row, the sampled slot must be legal, and greedy must pick the reference's argmax."""
    import bg_b200
    env = _positions(3000, seed=4)
    net = bg_b200.PolicyValueNet.random_init("cuda:0", seed=6)
    net.params["action_head.weight"].mul_(6.0); net.sync()
    n = env.legal_counts.long()
    assert int(((n >= 1) & (n <= 32)).sum()) > 0 and int(((n > 32) & (n <= 128)).sum()) > 0 and int((n == 0).sum()) > 0
    ref_logits, ref_lsm, ref_v, masked = _reference(net, env, bf16_operands=True)
    for step in (0, 1):
        a, lp, v = net.act(env.boards52, env.players, env.legal_counts, seed=21, step=step)
        al = a.long()
        ok = torch.where(n > 0, al < n, al < 500) & (al >= 0)
        assert bool(ok.all()), "sampled an illegal slot"
        assert (lp - ref_lsm.gather(1, al[:, None])[:, 0]).abs().max().item() < 1e-2
        assert (v - ref_v).abs().max().item() < 1e-3
    g, glp, _ = net.act(env.boards52, env.players, env.legal_counts, greedy=True)
    best = masked.max(-1).values
    picked = masked.gather(1, g.long()[:, None])[:, 0]
    assert (best - picked).abs().max().item() < 1e-2                # (ties / bf16 boundary cases: the value, not the index)
    assert (glp - ref_lsm.gather(1, g.long()[:, None])[:, 0]).abs().max().item() < 1e-2


def test_policy_many_class_b_rows_serial_chunks():
    """More class B tiles than the kernel deals out chunk by chunk (kMaxSplitB = 64 tiles): the CTAs then walk the four chunks of
    such a tile one after the other.  9,000 passes (no legal slot: the reference's softmax over all 500 slots) + ordinary rows: same
    checks as above, and the split form (few class B rows) must give the same log-prob for the same row and action."""
    import bg_b200
    env = _positions(3000, seed=9)
    net = bg_b200.PolicyValueNet.random_init("cuda:0", seed=6)
    net.params["action_head.weight"].mul_(4.0); net.sync()
    reps = 4
    boards = env.boards52.repeat(reps, 1).contiguous()
    players = env.players.repeat(reps).contiguous()
    counts = env.legal_counts.to(torch.int32).repeat(reps).contiguous()
    counts[:9000] = 0                                              # 71 class B tiles
    ref_logits, _, ref_v, _ = _reference(net, env, bf16_operands=True)
    ref_logits, ref_v = ref_logits.repeat(reps, 1), ref_v.repeat(reps)
    n = counts.long()
    slot = torch.arange(500, device=ref_logits.device)[None, :]
    masked = ref_logits + torch.where((slot < n[:, None]), 0.0, -103.27893)
    ref_lsm = torch.log_softmax(masked, -1)
    a, lp, v = net.act(boards, players, counts, seed=5, step=2)
    al = a.long()
    assert bool((torch.where(n > 0, al < n, al < 500) & (al >= 0)).all())
    assert (lp - ref_lsm.gather(1, al[:, None])[:, 0]).abs().max().item() < 1e-2
    assert (v - ref_v).abs().max().item() < 1e-3
    # passes sample among all 500 slots: the draws must spread over the chunks
    assert int((al[:9000] >= 128).sum()) > 4000 and int((al[:9000] >= 384).sum()) > 1000


def test_policy_act_with_a_growing_batch():
    """act() on 256 rows, then on 400, then on 3000 with the same net: the cached workspace must follow the C ABI's own requirement
    (bg_policy_workspace_bytes), not a private formula."""
    import bg_b200
    env = _positions(3000, seed=2)
    net = bg_b200.PolicyValueNet.random_init("cuda:0", seed=3)
    ref = net.act(env.boards52, env.players, env.legal_counts, seed=9, step=1)
    for B in (256, 400, 3000):
        a, lp, v = net.act(env.boards52[:B], env.players[:B], env.legal_counts[:B], seed=9, step=1)
        torch.cuda.synchronize()
        assert torch.equal(a, ref[0][:B]) and torch.equal(lp, ref[1][:B]) and torch.equal(v, ref[2][:B])   # keyed by row: independent of the batch
