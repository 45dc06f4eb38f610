"""Full-size runs (BASELINE configs[1]: 65,536 games on one GPU) checked through size-independent properties of the
domain, since the CPU oracle cannot replay them in seconds: checker conservation, distinct afterstates, the step
applying exactly the chosen afterstate, GPU-count independence of the trajectories, overlapped == serial encoders,
encoder dtype consistency, and 2-ply invariance to chunking / overlap."""
import pytest
import torch

pytestmark = pytest.mark.gpu
N = 65536


def _row_hash(rows52: torch.Tensor) -> torch.Tensor:
    w = rows52.contiguous().view(torch.int32).view(-1, 13).to(torch.int64)
    h = torch.zeros(w.shape[0], dtype=torch.int64, device=w.device)
    for k in range(13):
        h = h * 1000003 + (w[:, k] & 0xFFFFFFFF) * (2 * k + 1)
    return h


def test_env_invariants_at_65536_games():
    import bg_b200
    dev = torch.device("cuda:0")
    env = bg_b200.B200BackgammonVecEnv(num_envs=N, device=dev, seed=0x5EED, check_every=0)
    env.reset()
    for t in range(150):
        acts = env.random_actions(0xAC7, t)
        check = t % 15 == 14
        if check:
            rows = env.total_rows()
            A = env.after52[:rows].to(torch.int16)
            # 15 men per side on every afterstate (points + bar + off)
            p1 = A[:, 0:24].sum(1) + A[:, 48] + A[:, 50]
            p2 = A[:, 24:48].sum(1) + A[:, 49] + A[:, 51]
            assert bool((p1 == 15).all()) and bool((p2 == 15).all())
            assert int(env.legal_counts.sum().item()) == rows and bool((env.legal_counts == env.legal_counts_true.clamp(max=500)).all())
            # the afterstates of a game are distinct boards (add_unique_board, handle_moves.py:313-341)
            nz = torch.nonzero(env.legal_counts > 0)[:, 0]               # (a game without a legal play owns no rows)
            order = nz[torch.argsort(env.legal_starts[nz])]             # slab blocks are in arbitrary game order
            starts_sorted = env.legal_starts[order]
            cnt_sorted = env.legal_counts[order].long()
            game_of_row = torch.repeat_interleave(order, cnt_sorted)
            assert bool((starts_sorted == torch.cumsum(cnt_sorted, 0) - cnt_sorted).all())   # blocks tile [0, rows) exactly
            key = _row_hash(env.after52[:rows]) * 131071 + game_of_row
            assert torch.unique(key).numel() == rows
            prev_after, prev_starts, prev_counts = env.after52[:rows].clone(), env.legal_starts.clone(), env.legal_counts.clone()
            prev_boards = env.boards52.clone()
        env.step_device(acts)
        if check:
            moved = (prev_counts > 0) & ~env.dones_u8.bool()
            chosen = prev_after[(prev_starts + acts.long()).clamp(max=prev_after.shape[0] - 1)]
            assert bool((env.boards52[moved] == chosen[moved]).all())    # K2 applied exactly the chosen legal play
            passed = (prev_counts == 0)
            assert bool((env.boards52[passed] == prev_boards[passed]).all()) and bool(((env.flags & 1).bool() == passed).all())
    env.check_status()


def test_trajectories_do_not_depend_on_the_shard_count():
    import bg_b200
    dev = torch.device("cuda:0")
    one = bg_b200.B200BackgammonVecEnv(num_envs=N, device=dev, seed=7, check_every=0)
    halves = [bg_b200.B200BackgammonVecEnv(num_envs=N // 2, device=dev, seed=7, stream_base=k * (N // 2), check_every=0) for k in range(2)]
    one.reset(); [h.reset() for h in halves]
    for t in range(40):
        one.step_device(one.random_actions(3, t))
        for h in halves:
            h.step_device(h.random_actions(3, t))
    both = torch.cat([h.boards52 for h in halves])
    assert torch.equal(one.boards52, both) and torch.equal(one.dice, torch.cat([h.dice for h in halves]))
    assert torch.equal(one.legal_counts_true, torch.cat([h.legal_counts_true for h in halves]))


def test_overlapped_encoders_equal_serial_and_dtypes_agree():
    import bg_b200
    dev = torch.device("cuda:0")
    env = bg_b200.B200BackgammonVecEnv(num_envs=N, device=dev, seed=11, check_every=0)
    env.reset()
    for t in range(60):
        env.step_device(env.random_actions(5, t))
    env.update_legal_plays(obs=True, features=True, overlap=False)
    rows = env.total_rows()
    # rows land in arbitrary slab order: compare per game through the starts
    f_serial, o_serial = env.after_feats[:rows].clone(), env.obs_f32.clone()
    st_serial, a_serial = env.legal_starts.clone(), env.after52[:rows].clone()
    env.after_feats.zero_(); env.obs_f32.zero_()
    env.update_legal_plays(obs=True, features=True, overlap=True)
    torch.cuda.synchronize()
    assert env.total_rows() == rows and torch.equal(env.obs_f32, o_serial)
    g = torch.randint(0, N, (4096,), device=dev)
    for gi in g[:256].tolist():
        n = int(env.legal_counts[gi].item())
        s0, s1 = int(st_serial[gi].item()), int(env.legal_starts[gi].item())
        assert torch.equal(a_serial[s0:s0 + n], env.after52[s1:s1 + n]) and torch.equal(f_serial[s0:s0 + n], env.after_feats[s1:s1 + n])
    # every row of the overlapped run: features == encoder applied to the row (bf16), and bf16 == f32 cast (exact except off/15)
    want = bg_b200.encode(env.after52[:rows], env.row_players[:rows], dtype=torch.bfloat16)
    assert torch.equal(env.after_feats[:rows], want)
    f32 = bg_b200.encode(env.after52[:rows], env.row_players[:rows], dtype=torch.float32)
    assert torch.equal(want[:, :198], f32.to(torch.bfloat16)) and bool((want[:, 198:] == 0).all())
    env.check_status()


def test_twoply_invariant_to_chunking_and_overlap():
    import bg_b200
    dev = torch.device("cuda:0")
    env = bg_b200.B200BackgammonVecEnv(num_envs=4096, device=dev, seed=13, check_every=0)
    env.reset()
    for t in range(70):
        env.step_device(env.random_actions(9, t))
    net = bg_b200.ValueNet.random_init(dev, seed=0)
    b, p, d = env.boards52.clone(), env.players.clone(), env.dice.clone()
    ref = None
    for chunk, overlap in ((32768, True), (4096, True), (32768, False)):
        s = bg_b200.TwoPlySearch(net, max_afterstates_per_chunk=chunk, overlap=overlap)
        best, scores, offsets, A = s.search(b, p, d)
        if ref is None:
            ref = (best, scores, offsets, A)
            counts = (offsets[1:] - offsets[:-1])
            has = counts > 0
            assert bool((best[has] >= 0).all()) and bool((best[has].long() < counts[has]).all()) and bool((best[~has] == -1).all())
            seg_max = torch.full((b.shape[0],), -1e30, device=dev).scatter_reduce(0, torch.repeat_interleave(torch.arange(b.shape[0], device=dev), counts), scores, "amax")
            assert torch.equal(scores[(offsets[:-1] + best.long().clamp(min=0))[has]], seg_max[has])
        else:
            assert torch.equal(best, ref[0]) and torch.equal(scores, ref[1]) and torch.equal(A, ref[3])
