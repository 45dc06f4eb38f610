"""N2 host logic on the CPU: the PPO loss / update against golden vectors produced by the UNMODIFIED reference
learner (tests/golden/make_golden_ppo.py ran src/agent/ppo_agent.py), and the data-parallel update on two gloo ranks."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "ppo.npz")


def _golden():
    g = np.load(GOLDEN)
    sd0 = {k[3:]: torch.tensor(g[k]) for k in g.files if k.startswith("w0.")}
    sd1 = {k[3:]: torch.tensor(g[k]) for k in g.files if k.startswith("w1.")}
    return g, sd0, sd1


def _returns_interleaved(rewards, dones, gamma=0.99):
    """compute_returns (ppo_agent.py:206-216) over the step-major memory as one sequence, f32 as numpy floats do"""
    out, R = np.zeros_like(rewards), np.float64(0)
    for i in range(len(rewards) - 1, -1, -1):
        if dones[i]:
            R = 0
        R = rewards[i] + gamma * R
        out[i] = R
    return out


def _batch(g):
    t = lambda k, dt: torch.tensor(g[k]).to(dt)
    return (t("obs", torch.float32), t("counts", torch.int32), t("actions", torch.int32), t("logp", torch.float32),
            t("values", torch.float32), t("returns_interleaved", torch.float32))


def test_returns_restate_reference():
    g, _, _ = _golden()
    assert bool(g["mask_is_prefix"].all())                       # the env's masks are prefix masks (backgammon_env.py:228-231)
    np.testing.assert_allclose(_returns_interleaved(g["rewards"], g["dones"]), g["returns_interleaved"], rtol=1e-6, atol=1e-6)


def test_select_action_logp_and_values_match_reference():
    """log-probs / values stored by the reference's select_action == masked log-softmax of the same network"""
    import bg_b200
    from bg_b200.ppo import MASK_LOG, policy_value_forward
    g, sd0, _ = _golden()
    x, counts, actions, logp, values, _ = _batch(g)
    logits, v = policy_value_forward(sd0, x)
    slot = torch.arange(500)[None, :]
    lsm = torch.log_softmax(logits + torch.where(slot < counts[:, None], 0.0, MASK_LOG), -1)
    assert (lsm.gather(1, actions.long()[:, None])[:, 0] - logp).abs().max().item() < 1e-5
    assert (v - values).abs().max().item() < 1e-5


def test_loss_matches_reference_update_epoch():
    import bg_b200
    from bg_b200.ppo import ppo_loss
    g, sd0, _ = _golden()
    x, counts, actions, logp, values, ret = _batch(g)
    ret = (ret - ret.mean()) / (ret.std() + 1e-5)                # ppo_agent.py:256
    adv = ret - values                                           # :259
    want = g["loss_lr0"]                                         # policy, value, entropy, total (lr = 0: one epoch's losses)
    for autocast, tol in ((True, 2e-3), (False, 2e-2)):         # the reference ran under CPU bf16 autocast
        loss, pl, vl, ent = ppo_loss(sd0, x, counts, actions, logp, ret, adv, 0.25, 0.5, float(g["entropy_coef"]), autocast=autocast)
        got = np.array([pl.item(), vl.item(), ent.item(), loss.item()])
        assert np.abs(got - want).max() < tol, (autocast, got, want)


def test_update_moves_weights_like_reference():
    import bg_b200
    from bg_b200.ppo import PPOConfig, PPOLearner
    g, sd0, sd1 = _golden()
    x, counts, actions, logp, values, ret = _batch(g)
    L = PPOLearner(sd0, "cpu", PPOConfig(), host_logic_test=True)
    stats = L.update(x, counts, actions, logp, values, ret)
    want = g["loss_lr1e3"]
    got = np.array([stats["policy_loss"], stats["value_loss"], stats["entropy"], stats["total_loss"]])
    assert np.abs(got - want).max() < 2e-2, (got, want)
    # Adam moves every weight by ~lr per epoch in the direction of its gradient sign: compare the update directions
    new = L.state_dict()
    for k in ("value_head.weight", "value_head.bias", "fc1.bias"):
        d_ref, d_got = (sd1[k] - sd0[k]).flatten(), (new[k] - sd0[k]).flatten()
        cos = float(torch.dot(d_ref, d_got) / (d_ref.norm() * d_got.norm() + 1e-12))
        assert cos > 0.9, (k, cos)
        assert abs(float(d_got.abs().max()) - float(d_ref.abs().max())) < 2e-3


def _dp_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    import bg_b200
    from bg_b200.ppo import PPOConfig, PPOLearner
    g, sd0, _ = _golden()
    x, counts, actions, logp, values, ret = _batch(g)
    n = x.shape[0] // world
    sl = slice(rank * n, (rank + 1) * n)                          # this rank's shard of the samples
    L = PPOLearner(sd0, "cpu", PPOConfig(autocast=False), dist=dist, host_logic_test=True)
    L.update(x[sl], counts[sl], actions[sl], logp[sl], values[sl], ret[sl])
    flat = L.fp.flat.detach().clone()
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        q.put([t.numpy() for t in gathered])
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_update_equals_single_process():
    """two ranks, half of the samples each, one flat-bucket gradient all-reduce per step == one process, all samples"""
    import bg_b200
    from bg_b200.ppo import PPOConfig, PPOLearner
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29811 + (os.getpid() % 150)
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    flats = q.get(timeout=180)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert np.array_equal(flats[0], flats[1])                     # replicas stay bit-identical
    g, sd0, _ = _golden()
    x, counts, actions, logp, values, ret = _batch(g)
    n = (x.shape[0] // world) * world
    torch.set_num_threads(1)
    L = PPOLearner(sd0, "cpu", PPOConfig(autocast=False), host_logic_test=True)
    L.update(x[:n], counts[:n], actions[:n], logp[:n], values[:n], ret[:n])
    single = L.fp.flat.detach().numpy()
    assert L.fp.numel == 90101                                    # 90,101 parameters = one 360 KB bucket
    # Adam normalises the step, so fp reassociation in the gradient sum can flip near-zero-gradient weights by ~lr;
    # the bulk must agree tightly
    diff = np.abs(single - flats[0])
    assert np.quantile(diff, 0.99) < 2e-4 and diff.max() < 1e-2, (np.quantile(diff, 0.99), diff.max())


def test_learner_refuses_cpu_unless_host_logic_test():
    from bg_b200 import BgError
    from bg_b200.ppo import PPOConfig, PPOLearner
    g = np.load(GOLDEN)
    sd0 = {k[3:]: torch.tensor(g[k]) for k in g.files if k.startswith("w0.")}
    with pytest.raises(BgError):
        PPOLearner(sd0, "cpu", PPOConfig())
