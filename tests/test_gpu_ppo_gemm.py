"""The hand-written tcgen05 GEMMs of the PPO update (csrc/ppo_gemm.cu) against plain torch on the same bf16 operands
(f32 accumulation on both sides: tolerance = bf16 rounding of the outputs), the Adam kernel against torch.optim.Adam, and a
whole epoch (TensorCoreUpdate) against torch autograd over the restatement of ppo_agent.py:268-305."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ppo.npz")
DEV = "cuda:0"
OFF = dict(w1=0, b1=128 * 198, wa=128 * 198 + 128, ba=128 * 198 + 128 + 64000, wv=128 * 198 + 128 + 64000 + 500,
           bv=128 * 198 + 128 + 64000 + 500 + 128)


def _setup(seed=0):
    import bg_b200
    from bg_b200._lib import lib, check
    g = torch.Generator(device="cpu").manual_seed(seed)
    flat = (torch.randn(90101, generator=g) * 0.1).to(DEV)
    tc = bg_b200.TensorCoreUpdate(torch.device(DEV))
    L = lib()
    s = torch.cuda.current_stream().cuda_stream
    check(L.bg_ppo_pack_weights(flat.data_ptr(), tc.w1p.data_ptr(), tc.wap_a.data_ptr(), tc.wap_b.data_ptr(), tc.bias_a.data_ptr(),
                                tc.bias_b.data_ptr(), s), "pack")
    W1 = flat[:OFF["b1"]].view(128, 198); b1 = flat[OFF["b1"]:OFF["wa"]]
    Wa = flat[OFF["wa"]:OFF["ba"]].view(500, 128); ba = flat[OFF["ba"]:OFF["wv"]]
    wv = flat[OFF["wv"]:OFF["bv"]]; bv = flat[OFF["bv"]]
    bf = lambda t: t.to(torch.bfloat16).float()
    W1p = torch.zeros(128, 208, device=DEV); W1p[:, :198] = W1; W1p[:, 198] = b1
    WapA = torch.zeros(144, 128, device=DEV); WapA[:128] = Wa[:128]; WapA[128] = wv
    WapB = torch.zeros(512, 128, device=DEV); WapB[:500] = Wa; WapB[500] = wv
    biasA = torch.zeros(144, device=DEV); biasA[:128] = ba[:128]; biasA[128] = bv
    biasB = torch.zeros(512, device=DEV); biasB[:500] = ba; biasB[500] = bv
    return dict(L=L, check=check, s=s, tc=tc, flat=flat, W1p=bf(W1p), WapA=bf(WapA), WapB=bf(WapB), biasA=biasA, biasB=biasB, g=g)


def _rand_bf16(g, shape, scale=1.0):
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).to(DEV)


def _close(got, want, rel=1e-2):
    got, want = got.float(), want.float()
    err = (got - want).abs().max().item()
    scale = want.abs().max().item() + 1e-6
    assert err <= rel * scale, (err, scale)


def _blk(bg, m):
    return bg.TensorCoreUpdate.to_blocked(m)


def _unblk(bg, b, R, C):
    return bg.TensorCoreUpdate.from_blocked(b, R, C)


@pytest.mark.parametrize("TA,TB", [(6, 2), (1, 0), (0, 3), (300, 20)])
def test_forward_gemms_match_torch(TA, TB):
    """HIDDEN / LOGITS_A / LOGITS_B on tile-blocked operands (TA class A tiles, TB class B tiles of 128 rows)"""
    import bg_b200 as bg
    S = _setup()
    L, check, s, tc = S["L"], S["check"], S["s"], S["tc"]
    R, RA, RB = (TA + TB) * 128, TA * 128, TB * 128
    x = _rand_bf16(S["g"], (R, 208)); x[:, 198] = 1.0; x[:, 199:] = 0
    xb = _blk(bg, x)
    hb = torch.full((R * 128,), 7.0, dtype=torch.bfloat16, device=DEV)
    check(L.bg_ppo_gemm_nt(0, xb.data_ptr(), 0, TA + TB, tc.w1p.data_ptr(), None, None, hb.data_ptr(), s), "HIDDEN")
    h = _unblk(bg, hb, R, 128)
    _close(h, torch.relu(x.float() @ S["W1p"].t()))
    la = torch.full((max(RA, 128) * 144,), 7.0, dtype=torch.bfloat16, device=DEV)
    lb = torch.full((max(RB, 128) * 512,), 7.0, dtype=torch.bfloat16, device=DEV)
    check(L.bg_ppo_gemm_nt(1, hb.data_ptr(), 0, TA, tc.wap_a.data_ptr(), tc.bias_a.data_ptr(), None, la.data_ptr(), s), "LOGITS_A")
    check(L.bg_ppo_gemm_nt(2, hb.data_ptr(), TA, TA + TB, tc.wap_b.data_ptr(), tc.bias_b.data_ptr(), None, lb.data_ptr() - TA * 128 * 1024, s), "LOGITS_B")
    torch.cuda.synchronize()
    if TA:
        _close(_unblk(bg, la[:RA * 144], RA, 144)[:, :129], (h[:RA].float() @ S["WapA"].t() + S["biasA"])[:, :129])
    if TB:
        _close(_unblk(bg, lb[:RB * 512], RB, 512)[:, :501], (h[RA:].float() @ S["WapB"].t() + S["biasB"])[:, :501])


@pytest.mark.parametrize("TA,TB", [(6, 2), (3, 0), (1, 1), (200, 12)])
def test_backward_gemms_match_torch(TA, TB):
    import bg_b200 as bg
    S = _setup(seed=1)
    L, check, s, tc, g = S["L"], S["check"], S["s"], S["tc"], S["g"]
    R, RA, RB = (TA + TB) * 128, TA * 128, TB * 128
    x = _rand_bf16(g, (R, 208)); x[:, 198] = 1.0; x[:, 199:] = 0
    h = torch.relu(_rand_bf16(g, (R, 128)))
    dla = _rand_bf16(g, (max(RA, 128), 144), 0.05); dla[:, 129:] = 0
    dlb = _rand_bf16(g, (max(RB, 128), 512), 0.05); dlb[:, 501:] = 0
    xb, hb, dlab, dlbb = _blk(bg, x), _blk(bg, h), _blk(bg, dla), _blk(bg, dlb)
    dpreb = torch.full((R * 128,), 7.0, dtype=torch.bfloat16, device=DEV)
    check(L.bg_ppo_gemm_nt(3, dlab.data_ptr(), 0, TA, tc.wap_a.data_ptr(), None, hb.data_ptr(), dpreb.data_ptr(), s), "DPRE_A")
    check(L.bg_ppo_gemm_nt(4, dlbb.data_ptr() - TA * 128 * 1024, TA, TA + TB, tc.wap_b.data_ptr(), None, hb.data_ptr(), dpreb.data_ptr(), s), "DPRE_B")
    dpre = _unblk(bg, dpreb, R, 128)
    mask = (h.float() > 0).float()
    want = torch.cat([dla[:RA].float() @ S["WapA"], dlb[:RB].float() @ S["WapB"]], 0) * mask
    _close(dpre, want)
    gflat = torch.zeros(90101, device=DEV)
    scratch = torch.empty(199 * 128, device=DEV)
    check(L.bg_ppo_gemm_tn(5, hb.data_ptr(), dlab.data_ptr(), 0, TA, gflat.data_ptr(), None, s), "GRAD_WA_A")
    check(L.bg_ppo_gemm_tn(6, hb.data_ptr(), dlbb.data_ptr() - TA * 128 * 1024, TA, TA + TB, gflat.data_ptr(), None, s), "GRAD_WA_B")
    check(L.bg_ppo_gemm_tn(7, dpreb.data_ptr(), xb.data_ptr(), 0, TA + TB, gflat.data_ptr(), scratch.data_ptr(), s), "GRAD_W1")
    torch.cuda.synchronize()
    gA = dla[:RA].float().t() @ h[:RA].float()                         # (144,128)
    gB = dlb[:RB].float().t() @ h[RA:].float()                         # (512,128)
    want_wa = gB[:500].clone(); want_wa[:128] += gA[:128]
    want_wv = gB[500] + gA[128]
    gW1 = dpre.float().t() @ x.float()                                 # (128,208)
    _close(gflat[OFF["wa"]:OFF["ba"]].view(500, 128), want_wa, 2e-3)
    _close(gflat[OFF["wv"]:OFF["bv"]], want_wv, 2e-3)
    _close(gflat[:OFF["b1"]].view(128, 198), gW1[:, :198], 2e-3)
    _close(gflat[OFF["b1"]:OFF["wa"]], gW1[:, 198], 2e-3)
    assert float(gflat[OFF["ba"]:OFF["wv"]].abs().sum()) == 0.0        # biases are not these kernels' business


def test_gather_block_matches_torch():
    import bg_b200 as bg
    from bg_b200._lib import lib, check
    g = torch.Generator().manual_seed(5)
    x = _rand_bf16(g, (1000, 208))
    perm = torch.randperm(1000, generator=g)[:700].to(torch.int32)
    pp = torch.full((768,), -1, dtype=torch.int32)
    pp[:300] = perm[:300]                                   # class A: 300 rows + 84 padding rows
    pp[384:384 + 380] = perm[300:680]                       # class B: 380 rows + 4 padding rows
    pp = pp.to(DEV)
    out = torch.full((768 * 208,), 3.0, dtype=torch.bfloat16, device=DEV)
    check(lib().bg_ppo_gather_block(x.data_ptr(), 208, pp.data_ptr(), 768, 208, 198, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "gather")
    got = bg.TensorCoreUpdate.from_blocked(out, 768, 208)
    want = torch.zeros((768, 208), dtype=torch.bfloat16, device=DEV)
    valid = pp >= 0
    want[valid] = x[pp[valid].long()]
    want[valid, 198] = 1.0
    assert torch.equal(got, want)


def test_adam_kernel_matches_torch():
    from bg_b200._lib import lib, check
    g = torch.Generator().manual_seed(3)
    p0 = torch.randn(90101, generator=g).to(DEV)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=1e-3)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for step in range(1, 6):
        gr = torch.randn(90101, generator=g).to(DEV) * (0.1 if step % 2 else 3.0)
        p_ref.grad = gr.clone()
        opt.step()
        check(lib().bg_adam_step(p.data_ptr(), (gr * 4).data_ptr(), m.data_ptr(), v.data_ptr(), 90101, 1e-3, 0.9, 0.999, 1e-8, step, 0.25,
                                 torch.cuda.current_stream().cuda_stream), "adam")
        assert (p - p_ref.detach()).abs().max().item() < 2e-6


@pytest.mark.parametrize("impl", ["tcgen05", "tcgen05-unfused-loss", "cublas"])
def test_epoch_matches_autograd(impl):
    """losses and every gradient of one epoch == torch autograd over the torch restatement of ppo_agent.py:268-305 under
    bf16 autocast, on the reference's own rollout (golden) plus synthetic passes / wide rows so that class B is exercised"""
    from bg_b200.ppo import ManualUpdate, TensorCoreUpdate, ppo_loss
    g = np.load(GOLDEN)
    sd0 = {k[3:]: torch.tensor(g[k]).to(DEV) for k in g.files if k.startswith("w0.")}
    t = lambda k, dt: torch.tensor(g[k]).to(dt).to(DEV)
    obs, counts, actions, logp, values = t("obs", torch.float32), t("counts", torch.int32), t("actions", torch.int32), t("logp", torch.float32), t("values", torch.float32)
    # make some rows class B: passes (count 0, any action) and wide rows (count 300, action 250)
    counts = counts.clone(); actions = actions.clone()
    counts[::17] = 0; actions[::17] = 123
    counts[5::29] = 300; actions[5::29] = 250
    B = obs.shape[0]
    x = torch.zeros((B, 208), dtype=torch.bfloat16, device=DEV)
    x[:, :198] = obs
    ret = t("returns_interleaved", torch.float32)
    retn = (ret - ret.mean()) / (ret.std() + 1e-5)
    adv = (retn - values).contiguous()
    logp2 = logp + 0.3 * torch.randn_like(logp)
    p = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    loss, pl, vl, ent = ppo_loss(p, x, counts, actions, logp2, retn, adv, 0.25, 0.5, 0.15, autocast=True, fused=False)
    loss.backward()
    want = {k: v.grad.clone() for k, v in p.items()}
    grads = {k: torch.zeros_like(v) for k, v in sd0.items()}
    xm = x.clone(); xm[:, 198] = 1.0
    upd = (ManualUpdate(torch.device(DEV)) if impl == "cublas" else
           TensorCoreUpdate(torch.device(DEV), fuse_loss=(impl == "tcgen05")))
    st = upd.epoch(sd0, grads, xm, counts, actions, logp2, adv, retn.contiguous(), 0.25, 0.5, 0.15)
    got = st.cpu().numpy()
    assert np.abs(got - np.array([pl.item(), vl.item(), ent.item(), loss.item()])).max() < 3e-3, got
    for k in want:
        scale = want[k].abs().max().item() + 1e-8
        assert (want[k] - grads[k]).abs().max().item() < 3e-2 * scale, (k, (want[k] - grads[k]).abs().max().item(), scale)


def test_update_paths_agree():
    """a whole four-epoch PPOLearner.update through autograd, the cuBLAS path and the tcgen05 path ends at nearly the same weights"""
    from bg_b200.ppo import PPOConfig, PPOLearner
    g = np.load(GOLDEN)
    sd0 = {k[3:]: torch.tensor(g[k]).to(DEV) for k in g.files if k.startswith("w0.")}
    t = lambda k, dt: torch.tensor(g[k]).to(dt).to(DEV)
    obs, counts, actions, logp, values = t("obs", torch.float32), t("counts", torch.int32), t("actions", torch.int32), t("logp", torch.float32), t("values", torch.float32)
    ret = t("returns_interleaved", torch.float32)
    x = torch.zeros((obs.shape[0], 208), dtype=torch.bfloat16, device=DEV)
    x[:, :198] = obs
    outs = []
    for cfg in (PPOConfig(manual_backward=False), PPOConfig(update_impl="cublas"), PPOConfig(update_impl="tcgen05")):
        Lr = PPOLearner(sd0, DEV, cfg)
        Lr.update(x.clone(), counts, actions, logp, values, ret)
        outs.append((Lr.state_dict(), Lr.last))
    for k in sd0:
        moved = (outs[0][0][k] - sd0[k]).abs().mean().item()          # (Adam steps are sign-like: compare means, not maxima)
        for o in outs[1:]:
            assert (outs[0][0][k] - o[0][k]).abs().mean().item() < 0.25 * moved + 1e-6, k
    for key in ("policy_loss", "value_loss", "entropy", "total_loss"):
        for o in outs[1:]:
            assert abs(outs[0][1][key] - o[1][key]) < 5e-3, (key, outs[0][1], o[1])


def test_encode_block_equals_encode_then_gather():
    """bg_ppo_encode_block (features of boards52[perm] straight into the blocked layout) == K3's bf16 encoder followed by bg_ppo_gather_block,
    bit for bit, padding rows and bias column included"""
    import bg_b200
    from bg_b200._lib import lib, check
    from bg_b200.engine import encode
    env = bg_b200.B200BackgammonVecEnv(num_envs=700, device=DEV, seed=11, check_every=0)
    env.reset()
    for t in range(40):
        env.step_device(env.random_actions(5, t))
    rows = 768
    g = torch.Generator(device="cpu").manual_seed(3)
    perm = torch.full((rows,), -1, dtype=torch.int32)
    perm[:700] = torch.randperm(700, generator=g).to(torch.int32)
    perm = perm.to(DEV)
    x = encode(env.boards52, env.players, dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    a = torch.full((rows * 208,), 7.0, dtype=torch.bfloat16, device=DEV)
    b = torch.full((rows * 208,), 9.0, dtype=torch.bfloat16, device=DEV)
    check(lib().bg_ppo_gather_block(x.data_ptr(), x.shape[1], perm.data_ptr(), rows, 208, 198, a.data_ptr(), st), "gather")
    check(lib().bg_ppo_encode_block(env.boards52.data_ptr(), env.players.data_ptr(), perm.data_ptr(), rows, 198, b.data_ptr(), st), "encode_block")
    torch.cuda.synchronize()
    assert torch.equal(a.view(torch.int16), b.view(torch.int16))
