#!/usr/bin/env python
"""bench.py -- env steps/sec with legal-move generation (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): batched legal-move generation + step on 65,536 parallel games per GPU,
uniform-random policy, positions from the engine's own random self-play (Philox seed 0x5EED, 128 untimed
de-phasing turns).  One "step" = one turn of every game:
    bg_random_actions -> K2 bg_env_step (apply play, reward/terminal, auto-reset, Philox dice)
    -> K1 bg_movegen_slab (all legal afterstates of the new positions, reference order)
    -> K3 bg_encode_f32 (observations) + bg_encode_bf16 (ragged afterstate features)
i.e. everything BackgammonEnv.step + update_legal_moves + get_observation do, minus the dense
(500,198) zero padding.  Games shard across GPUs by game id with no collective ("scaling": "weak").

Prints ONE JSON line (see the keys below).  `--impl reference` times the CPU port of the reference's
path (oracle/, kind "port": the reference itself is Python and cannot travel to the GPU box) on all
host cores for the same metric.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 0x5EED
ACT_SEED = 0xAC7
METRIC = "env steps/sec w/ legal-move gen"
UNIT = "env_steps/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------- CPU arm

def cpu_envs(threads: int):
    from oracle import bg_oracle as O
    O.lib()
    envs = []
    for i in range(threads):
        e = O.Env()
        e.set_philox(SEED, i)
        e.reset()
        envs.append(e)
    return envs


def cpu_rollout(seconds: float, threads: int, chunk: int = 2000, encode: bool = True, envs=None):
    """Uniform-random self-play through the oracle's restatement of BackgammonEnv (same Philox dice / action
    conventions as the engine), `threads` envs in parallel (ctypes releases the GIL).  -> (steps, elapsed s)"""
    envs = envs or cpu_envs(threads)
    done_steps = [0] * threads
    t_end = [0.0] * threads
    start = time.perf_counter()

    def work(i):
        e = envs[i]
        n = 0
        while time.perf_counter() - start < seconds:
            e.random_rollout(chunk, ACT_SEED, encode)
            n += chunk
        done_steps[i] = n
        t_end[i] = time.perf_counter()

    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return sum(done_steps), max(t_end) - start


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # each "step" of this arm is a bounded sample of the workload: the whole run is ~60 s of rollout whatever K and W are
    per_step_s = max(0.02, min(8.0, 60.0 / max(1, args.steps + args.warmup)))
    chunk = max(100, min(2000, int(per_step_s * 7000)))          # ~70 k steps/s/thread: a chunk is about a tenth of a step
    envs = cpu_envs(threads)
    # the K steps run back to back inside one threaded region (no thread start/join between steps)
    if args.warmup:
        cpu_rollout(per_step_s * args.warmup, threads, chunk, envs=envs)
    total, el = cpu_rollout(per_step_s * args.steps, threads, chunk, envs=envs)
    v = total / el
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * el / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": "uniform-random self-play, BackgammonEnv.step + update_legal_moves + encoders, CPU port of "
                                   "the reference path (oracle/bg_oracle.c); the Python reference itself (~50 steps/s/core, "
                                   "BASELINE.md) cannot travel to the GPU box",
                       "games": threads, "policy": "uniform random (Philox)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{total} env steps in {el:.1f} s over {threads} threads, each step of this arm = "
                                       f"{per_step_s:.2f} s of rollout per thread"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- GPU arm

def run_engine(args):
    import torch
    import torch.distributed as dist
    import bg_b200
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N = args.games
    env = bg_b200.B200BackgammonVecEnv(num_envs=N, device=dev, seed=SEED, stream_base=rank * N,
                                       rows_per_game=args.rows_per_game, check_every=0)
    env.reset()
    acts = torch.empty(N, dtype=torch.int32, device=dev)
    rows_acc = torch.zeros(1, dtype=torch.int64, device=dev)
    launches = [0]
    feats = not args.no_afterstate_features

    def one_step(t, ev=None):
        env.random_actions(ACT_SEED, t, out=acts); launches[0] += 1
        env._apply_actions(acts); launches[0] += 1
        # K1 tiers 0/1/2 + K3 (observations f32, afterstate features bf16: rows final after tier 0 are encoded on
        # a second stream beside tiers 1/2, the rest after them) in one C call
        env.update_legal_plays(obs=True, features=feats, overlap=not args.no_overlap, k1_events=ev)
        launches[0] += 3 + 1 + 2 * int(feats)
        rows_acc.add_(env.alloc_rows)

    t = 0
    for _ in range(args.dephase + args.warmup):               # de-phase the games, then W warm-up steps (untimed)
        one_step(t); t += 1
    env.check_status()
    rows_acc.zero_()
    launches[0] = 0
    K = args.steps
    k1_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for a, b in k1_events:                                    # create the handles (torch creates them on first record)
        a.record(); b.record()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for k in range(K):
        one_step(t, k1_events[k]); t += 1
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    env.check_status()
    k1_ms = sum(a.elapsed_time(b) for a, b in k1_events) / K
    # the same K1 launches alone (no encoder beside the overflow tiers), same positions: explains the in-step figure
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    env._refresh_legal_moves(); torch.cuda.synchronize()
    ea.record()
    for _ in range(20):
        env._refresh_legal_moves()
    eb.record(); torch.cuda.synchronize()
    k1_alone_ms = ea.elapsed_time(eb) / 20
    # ... and whole steps with the encoders AFTER K1 (nothing overlapped): K1's share of this serialised step is what the
    # ncu launch list (profiles/, kernels serialised by the profiler) shows
    overlap_saved, args.no_overlap = args.no_overlap, True
    ea.record()
    for _ in range(20):
        one_step(t); t += 1
    eb.record(); torch.cuda.synchronize()
    args.no_overlap = overlap_saved
    step_serial_ms = ea.elapsed_time(eb) / 20
    launches[0] -= 20 * 8
    rows_per_step = float(rows_acc.item()) / K
    n_launch = launches[0]

    # ---- end to end through the public API with host buffers: actions from pinned host memory every step,
    # rewards / dones / legal-play counts read back to the host every step (what a host-side policy needs).
    E = max(1, min(args.e2e_steps, K))
    import numpy as np
    host = bg_b200.HostStepBuffers(env)                                         # pinned rewards / dones / legal counts
    h_acts = torch.empty(N, dtype=torch.int32).pin_memory()
    rng = np.random.default_rng(1)
    u = rng.integers(0, 65536, size=(E, N), dtype=np.int32)
    acts_np, counts_np, tmp = h_acts.numpy(), host.legal_counts.numpy(), np.empty(N, dtype=np.int32)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_segments = []
    for seg in range(3):                                                        # three segments of E steps; the best one is reported
        host.legal_counts.copy_(env.legal_counts, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for k in range(E):
            host.wait()                                                         # rewards / dones / counts of the previous step are on the host
            np.multiply(u[k], counts_np, out=tmp)                               # host-side policy: uniform over the legal plays,
            np.right_shift(tmp, 16, out=acts_np)                                # action = floor(u16 * count / 65536), into pinned memory
            # H2D of the actions, K2, K1, K3 and the D2H of rewards / dones / counts (issued behind K2 / K1) inside
            obs, rew, done, infos = env.step(h_acts, with_features=feats, host=host, overlap=args.e2e_overlap_encoders)
        host.wait()
        torch.cuda.synchronize()                                                # the last step's encoders are inside the timed region
        e2e_segments.append(time.perf_counter() - t0)
        u = rng.integers(0, 65536, size=(E, N), dtype=np.int32)
    e2e_s = min(e2e_segments)
    env.check_status()

    extra = run_extras(bg_b200, env, torch, dev, args) if (world == 1 and not args.no_extras) else None
    twoply_all, greedy_all = None, None
    if world > 1 and not args.no_extras:
        # 1-ply greedy self-play (configs[2]: games sharded, the value net replicated), device events, max over ranks
        gnet = bg_b200.ValueNet.random_init(dev, seed=0)
        def greedy_step():
            a_, _ = bg_b200.greedy_actions(env, gnet)
            env.step_device(a_.clamp_(min=0))
        greedy_step(); torch.cuda.synchronize(); dist.barrier()
        ga, gb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ga.record()
        for _ in range(30):
            greedy_step()
        gb.record(); torch.cuda.synchronize()
        gt = torch.tensor([ga.elapsed_time(gb) / 30], dtype=torch.float64, device=dev)
        dist.all_reduce(gt, op=dist.ReduceOp.MAX)
        greedy_all = {"env_steps_per_s": world * N / (float(gt.item()) * 1e-3), "ms_per_step": float(gt.item()), "games": world * N,
                      "what": "K4 on every legal afterstate + segment argmax + K2 + K1; max over ranks"}
        # 2-ply on every rank's own roots (shards of the roots, replicas of the net)
        tp2 = run_twoply(bg_b200, env, torch, dev, args)
        agg = torch.tensor([tp2["root_afterstates"], tp2["leaves"], tp2["roots"]], dtype=torch.float64, device=dev)
        tmax = torch.tensor([tp2["seconds"]], dtype=torch.float64, device=dev)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        a, l, r = agg.tolist(); sec = float(tmax.item())
        twoply_all = {"roots": int(r), "root_afterstates": int(a), "leaves": int(l), "seconds": sec, "root_positions_per_s": r / sec,
              "root_afterstates_per_s": a / sec, "leaves_per_s": l / sec,
              "note": f"{world} ranks, each its own {args.twoply_roots} roots; sum over ranks / max time over ranks (best of 3)"}

    tm = torch.tensor([ms, e2e_s * 1e3, k1_ms, k1_alone_ms, step_serial_ms], dtype=torch.float64, device=dev)
    rw = torch.tensor([rows_per_step], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(rw, op=dist.ReduceOp.SUM)
        rw /= world
    ms, e2e_ms, k1_ms, k1_alone_ms, step_serial_ms = [float(x) for x in tm.tolist()]
    rows_per_step = float(rw.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None

    value = world * N * K / (ms * 1e-3)
    e2e_value = world * N * E / (e2e_ms * 1e-3)
    peak, peak_src = load_peaks()
    # K1 algorithmic bytes per launch (DESIGN.md "K1"): per game 52 board + 1 player + 2 dice read; 4 count +
    # 4 true count + 8 start written; per legal play 52 afterstate + 1 mover flag written.
    k1_bytes = N * 71.0 + rows_per_step * 53.0
    achieved = k1_bytes / (k1_ms * 1e-3) / 1e9
    traffic, issue = None, None
    tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tp):
        try:
            kt = json.load(open(tp))
            traffic = kt.get("dram_bytes_per_launch")
            if kt.get("warp_inst_per_launch") and N == 65536:
                # what actually bounds K1: warp-instruction issue (4 schedulers per SM, one instruction per cycle each)
                sm_clock_ghz = (clocks or {}).get("sm_mhz") or 1965.0
                peak_issue = 148 * 4 * sm_clock_ghz / 1e3                      # G warp-instructions / s
                ach = kt["warp_inst_per_launch"] / (k1_alone_ms * 1e-3) / 1e9
                issue = {"bound": "warp-instruction issue", "warp_inst_per_launch": kt["warp_inst_per_launch"],
                         "achieved_ginst_s": ach, "peak_ginst_s": peak_issue, "frac": ach / peak_issue,
                         "what": "instructions per K1 call from the ncu captures named in profiles/k1_traffic.json / "
                                 "k1_ms_per_launch_alone; peak = 148 SMs x 4 schedulers x the SM clock sampled during the run"}
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int8", "data": "synthetic",
        "config": {"workload": "configs[1]: batched legal-move generation + step, uniform-random policy",
                   "games_per_gpu": N, "max_legal_moves": 500, "dephase_steps": args.dephase,
                   "afterstate_features": "bf16 ragged (ld 208)" if feats else "off", "observations": "f32 (N,198)",
                   "legal_plays_per_step_mean": rows_per_step / N, "parallelism": f"games sharded x{world}, no collective",
                   "l2": "per-step working set (afterstates + features + observations) exceeds the 126 MB L2; no flush"},
        "roofline": {"kernel": "K1 movegen (tier 0 warp kernel + tier 1/2 team kernels)", "bound": "hbm",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "k1_ms_per_launch": k1_ms, "k1_share_of_step": k1_ms / (ms / K),
                     "k1_ms_per_launch_alone": k1_alone_ms, "achieved_alone": k1_bytes / (k1_alone_ms * 1e-3) / 1e9,
                     "ms_per_step_serial": step_serial_ms, "k1_share_of_step_serial": k1_alone_ms / step_serial_ms,
                     "algorithmic_bytes_per_launch": k1_bytes,
                     "note": "K1 is integer-issue bound (SURVEY 8(d)); the HBM fraction is reported as required. In the step "
                             "the encoders (K3) run on a second stream beside K1's overflow tiers, so k1_ms_per_launch (CUDA "
                             "events around K1's three launches inside the step) includes that sharing; *_alone is the same "
                             "launch sequence with nothing beside it, and k1_share_of_step_serial = that / a step with the encoders "
                             "after K1 -- the share the profiler's serialised launch list shows (profiles/). traffic = dram bytes of the tier 0 + tier 1 launches "
                             "(ncu --set full, profiles/): below the algorithmic bytes because the afterstate rows are still in "
                             "the 126 MB L2 when the encoder reads them", "issue": issue},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": world * N * 4, "d2h_bytes_per_step": world * N * 9,
                "steps": E, "segments_ms_per_step": [x * 1e3 / E for x in e2e_segments],
                "what": "best of 3 segments (host-side jitter); B200BackgammonVecEnv.step(actions from pinned host memory, "
                                    "host=HostStepBuffers): rewards, dones and legal-play counts are copied to pinned host memory every "
                                    "step as soon as K2 + K1 are done and the host waits for them before choosing the next actions; "
                                    "observations / afterstate features (K3) stay on the device as the reference API returns them, "
                                    "run after K1 and finish while the host prepares the next step"},
        "gpu_launches": n_launch, "clocks": clocks,
    }
    if extra is not None:
        line["extra"] = extra
    if twoply_all is not None:
        line["extra"] = {"twoply": twoply_all, "greedy_1ply": greedy_all}
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n, el = cpu_rollout(args.cpu_seconds, threads)
        line["cpu_baseline"] = {"value": n / el, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{n} env steps of the same uniform-random workload in {el:.1f} s over {threads} "
                                          "threads (oracle/bg_oracle.c; the Python reference measures ~50 steps/s/core, BASELINE.md)"}
    if world > 1:
        dist.destroy_process_group()
    return line


def run_extras(bg_b200, env, torch, dev, args):
    """Secondary figures of the same path (BASELINE.json: "2-ply positions evaluated/sec"; configs[2], configs[3]):
    K4 leaf evaluator alone, 1-ply greedy self-play steps, 2-ply search.  Random-init weights of the reference
    architecture (198-128-1 value path), positions = the de-phased random-play games of the main run."""
    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e-3
    net = bg_b200.ValueNet.random_init(dev, seed=0)
    out = {}
    # K4: value of every current afterstate (ragged buffer, ~1.2 M rows), fused encode + tcgen05 GEMM + value head
    rows = env.total_rows()
    vbuf = torch.empty(rows, dtype=torch.float32, device=dev)
    t = timed(lambda: net.values(env.after52[:rows], env.row_players[:rows], out=vbuf), 20)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tf = rows * 53504.0 / t / 1e12
    out["mlp_value"] = {"positions_per_s": rows / t, "rows": rows, "ms": t * 1e3, "tflops_bf16": tf,
                        "tensor_peak_tflops": peaks.get("bf16_tflops"), "frac_of_tensor_peak": (tf / peaks["bf16_tflops"]) if peaks.get("bf16_tflops") else None,
                        "hbm_gbs": rows * 57.0 / t / 1e9}
    # 1-ply greedy self-play (configs[2] per GPU): K4 on the afterstates + segment argmax + K2 + K1
    def greedy_step():
        acts, _ = bg_b200.greedy_actions(env, net)
        env.step_device(acts.clamp_(min=0))
    t = timed(greedy_step, 30)
    out["greedy_1ply"] = {"env_steps_per_s": env.num_envs / t, "ms_per_step": t * 1e3, "games": env.num_envs}
    # N1 policy/value rollout forward (select_action) and the PPO rollout step (configs[4] rollout side): policy kernel -> K2 -> K1
    pnet = bg_b200.PolicyValueNet.random_init(dev, seed=0)
    N = env.num_envs
    pout = (torch.empty(N, dtype=torch.int32, device=dev), torch.empty(N, dtype=torch.float32, device=dev),
            torch.empty(N, dtype=torch.float32, device=dev))
    step_ctr = [0]
    t = timed(lambda: pnet.act(env.boards52, env.players, env.legal_counts, seed=1, step=0, out=pout), 20)
    out["policy_sample"] = {"positions_per_s": N / t, "ms": t * 1e3, "tflops_bf16": N * (53504.0 + 2 * 128 * 512) / t / 1e12}
    def ppo_rollout_step():
        step_ctr[0] += 1
        pnet.act(env.boards52, env.players, env.legal_counts, seed=1, stream_base=env.stream_base, step=step_ctr[0], out=pout)
        env.step_device(pout[0])
    t = timed(ppo_rollout_step, 30)
    out["ppo_rollout"] = {"env_steps_per_s": N / t, "ms_per_step": t * 1e3, "games": N,
                          "what": "sampled policy (fused policy/value kernel) -> K2 -> K1, no feature tensors in HBM"}
    out["twoply"] = run_twoply(bg_b200, env, torch, dev, args, net)
    return out


def run_twoply(bg_b200, env, torch, dev, args, net=None):
    """2-ply (configs[3]): roots = the first R games' positions with their actual dice"""
    net = net or bg_b200.ValueNet.random_init(dev, seed=0)
    R = min(args.twoply_roots, env.num_envs)
    search = bg_b200.TwoPlySearch(net, max_afterstates_per_chunk=args.twoply_chunk)
    b, p, d = env.boards52[:R].clone(), env.players[:R].clone(), env.dice[:R].clone()
    search.search(b, p, d)                                    # warm-up: allocates the persistent workspaces
    torch.cuda.synchronize()
    t, reps = 1e30, 3
    for _ in range(reps):
        search.leaves_evaluated = 0
        t0 = time.perf_counter()
        best, scores, offsets, A = search.search(b, p, d)
        torch.cuda.synchronize()
        t = min(t, time.perf_counter() - t0)
    return {"roots": R, "root_afterstates": int(A.shape[0]), "leaves": int(search.leaves_evaluated), "seconds": t,
            "root_positions_per_s": R / t, "root_afterstates_per_s": A.shape[0] / t,
            "leaves_per_s": search.leaves_evaluated / t,
            "note": "best of 3, wall clock incl. host orchestration; a 2-ply position = one root afterstate fully expanded "
                    "(21 opponent rolls x replies, leaves MLP-evaluated)"}


class StdoutToStderr:
    """Everything written to fd 1 while active (NCCL's version banner, library chatter) goes to stderr, so that the ONE
    JSON line is the only thing on stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--games", type=int, default=65536, help="games per GPU")
    ap.add_argument("--dephase", type=int, default=128)
    ap.add_argument("--rows-per-game", type=int, default=64)
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-afterstate-features", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="run the encoders after K1 on the same stream")
    ap.add_argument("--e2e-overlap-encoders", action="store_true",
                    help="e2e loop: encoders beside K1's overflow tiers (shorter GPU step, but K1 -- which the host waits for -- ends later)")
    ap.add_argument("--twoply-roots", type=int, default=4096)
    ap.add_argument("--twoply-chunk", type=int, default=98304)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    with StdoutToStderr():
        line = run_engine(args)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
