#!/usr/bin/env python
"""bench.py -- env steps/sec with legal-move generation (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): batched legal-move generation + step on 65,536 parallel games per GPU,
uniform-random policy, positions from the engine's own random self-play (Philox seed 0x5EED, 128 untimed
de-phasing turns).  One "step" = one turn of every game:
    K2 bg_env_step_random (uniform-random action, apply play, reward/terminal, auto-reset, Philox dice)
    -> K1 bg_movegen_slab (all legal afterstates of the new positions, reference order)
    -> K3 bg_encode_f32 (observations) + bg_encode_bf16 (ragged afterstate features)
i.e. everything BackgammonEnv.step + update_legal_moves + get_observation do, minus the dense
(500,198) zero padding.  Games shard across GPUs by game id with no collective ("scaling": "weak").

Prints ONE JSON line.  The timed region is R x K steps (R chosen so that it lasts >= --min-seconds), bracketed once
by barrier + synchronize; `ms_per_step` is its device time / (R K), max over ranks.
`--impl reference` times the reference's path on the host cores: the C port (oracle/, kind "port") for `value`, and
beside it the UNMODIFIED Python reference staged under baseline/_ref (kind "reference-python"), one BackgammonEnv per
process on every core.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
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
REF_STAGED = os.path.join(ROOT, "baseline", "_ref")

# algorithmic bytes per unit; N = games, rows = legal plays of the step
K1_BYTES_PER_GAME, K1_BYTES_PER_ROW = 71.0, 53.0      # board 52 + player 1 + dice 2 read, counts 4 + 4 + start 8 written; row 52 + mover 1
K2_BYTES_PER_GAME = 120.0
K3_OBS_BYTES_PER_GAME = 53.0 + 792.0                  # board + flag read, 198 f32 written
K3_FEAT_BYTES_PER_ROW = 53.0 + 416.0                  # row + flag read, 208 bf16 written


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", d
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)", {}


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
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE,
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


# ----------------------------------------------------------------------------------------- CPU arms

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


def reference_dir():
    """The UNMODIFIED reference tree: the copy build() stages under baseline/_ref (git-ignored, travels to the GPU box
    with the snapshot), else /root/reference when this runs in the build container."""
    for d in (REF_STAGED, "/root/reference"):
        if os.path.isdir(os.path.join(d, "src", "moves")):
            return d
    return None


def _pyref_worker(ref, seconds, seed, barrier, q):
    """One process = one BackgammonEnv of the reference (backgammon_env.py:78-191), uniform-random policy."""
    try:
        sys.stdout = open(os.devnull, "w")
        import numpy as np
        import torch
        torch.set_num_threads(1)
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import ref_loader
        ref_loader.load(ref)
        from src.environment.backgammon_env import BackgammonEnv
        env = BackgammonEnv()
        env.seed(seed)
        rng = np.random.RandomState(seed)
        env.reset()
        for _ in range(5):                                          # warm-up (imports, allocator)
            n = len(env.legal_moves)
            _, _, done, _ = env.step(int(rng.randint(0, n)) if n else None)
            if done:
                env.reset()
        barrier.wait(timeout=600)
        t0 = time.perf_counter()
        steps = 0
        while time.perf_counter() - t0 < seconds:
            n = len(env.legal_moves)
            _, _, done, _ = env.step(int(rng.randint(0, n)) if n else None)
            steps += 1
            if done:
                env.reset()                                         # vec_bg_env.py:35-36
        q.put((steps, time.perf_counter() - t0))
    except Exception as e:                                          # noqa: BLE001
        try:
            barrier.abort()
        except Exception:
            pass
        q.put(("error", repr(e)))


def python_reference_rollout(seconds: float, procs: int):
    """The reference's own Python path on `procs` host processes for `seconds` each -> dict (or {"unavailable": why})."""
    ref = reference_dir()
    if ref is None:
        return {"kind": "reference-python", "unavailable": "reference tree not staged (baseline/_ref missing)"}
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    barrier, q = ctx.Barrier(procs), ctx.Queue()
    ps = [ctx.Process(target=_pyref_worker, args=(ref, seconds, 1000 + i, barrier, q), daemon=True) for i in range(procs)]
    t_spawn = time.perf_counter()
    [p.start() for p in ps]
    res = []
    try:
        for _ in ps:
            res.append(q.get(timeout=600 + seconds))
    except Exception as e:                                          # noqa: BLE001
        return {"kind": "reference-python", "unavailable": f"workers did not report: {e!r}"}
    finally:
        for p in ps:
            p.join(timeout=5)
            if p.is_alive():
                p.terminate()
    bad = [r for r in res if r[0] == "error"]
    if bad:
        return {"kind": "reference-python", "unavailable": bad[0][1][:200]}
    steps, el = sum(r[0] for r in res), max(r[1] for r in res)
    return {"kind": "reference-python", "value": steps / el, "unit": UNIT, "cores": procs,
            "per_process": steps / el / procs,
            "sample": f"{steps} env steps of the UNMODIFIED reference (BackgammonEnv.step + reset on done, uniform-random policy, "
                      f"torch threads 1) in {el:.1f} s on {procs} processes, one env each; staged copy {os.path.relpath(ref, ROOT) if ref.startswith(ROOT) else ref}; "
                      f"start-up {time.perf_counter() - t_spawn - el:.0f} s not counted"}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


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
    py = python_reference_rollout(args.pyref_seconds, threads) if args.pyref_seconds > 0 else None
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * el / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": "uniform-random self-play, BackgammonEnv.step + update_legal_moves + encoders on the host cores: "
                                   "`value` = the C port of the reference path (oracle/bg_oracle.c, the faster baseline); "
                                   "`cpu_baseline.python_reference` = the unmodified Python reference in the same run",
                       "games": threads, "policy": "uniform random (Philox)", "cpu_model": cpu_model()},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{total} env steps in {el:.1f} s over {threads} threads, each step of this arm = "
                                       f"{per_step_s:.2f} s of rollout per thread",
                             "python_reference": py},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- GPU arm

def make_env(bg_b200, dev, games_per_gpu: int, rank: int, world: int, rows_per_game: int):
    """This rank's shard of the world*games_per_gpu games: contiguous global game ids [base, base+count); dice / action
    streams are keyed by the GLOBAL id (stream_base), so the union of trajectories does not depend on `world`."""
    base, count = bg_b200.shard_range(world * games_per_gpu, rank, world)
    env = bg_b200.B200BackgammonVecEnv(num_envs=count, device=dev, seed=SEED, stream_base=base,
                                       rows_per_game=rows_per_game, check_every=0)
    return env


def run_engine(args):
    import numpy as np
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
    ddist = dist if world > 1 else None
    env = make_env(bg_b200, dev, args.games, rank, world, args.rows_per_game)
    N = env.num_envs
    env.reset()
    acts = torch.empty(N, dtype=torch.int32, device=dev)
    feats = not args.no_afterstate_features
    LAUNCHES_PER_STEP = 1 + 3 + 1 + 2 * int(feats)            # K2 (random policy inside), K1 tiers 0/1/2, K3 f32, K3 bf16 (two row ranges)

    def one_step(t, ev=None, overlap=True):
        # K2 with the uniform-random policy inside (bg_env_step_random: the draw of bg_random_actions, one launch less)
        env.step_random_device(ACT_SEED, t)
        # K1 tiers 0/1/2 + K3 (observations f32, afterstate features bf16: rows final after tier 0 are encoded on
        # a second stream beside tiers 1/2, the rest after them) in one C call; the slab row counter lives in K1's workspace
        # header and is zeroed by K1's own memset
        env.update_legal_plays(obs=True, features=feats, overlap=overlap, k1_events=ev)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    t = 0
    for _ in range(args.dephase):                             # de-phase the games (untimed)
        one_step(t); t += 1
    env.check_status()
    K = args.steps
    # W warm-up steps (untimed for the result; timed to choose R so that the timed region lasts >= min_seconds)
    W = max(args.warmup, 1)
    wa, wb = ev(), ev()
    torch.cuda.synchronize(); wa.record()
    for _ in range(W):
        one_step(t, overlap=not args.no_overlap); t += 1
    wb.record(); torch.cuda.synchronize()
    est = torch.tensor([wa.elapsed_time(wb) / W], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(est, op=dist.ReduceOp.MAX)
    R = max(1, min(args.max_reps, math.ceil(args.min_seconds * 1e3 / (float(est.item()) * K))))
    n_timed = R * K
    k1_events = [(ev(), ev()) for _ in range(n_timed)]
    for a, b in k1_events:                                    # create the handles (torch creates them on first record)
        a.record(); b.record()
    e0, e1 = ev(), ev()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for k in range(n_timed):
        one_step(t, k1_events[k], overlap=not args.no_overlap); t += 1
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    env.check_status()
    k1_ms = sum(a.elapsed_time(b) for a, b in k1_events) / n_timed

    # ---- per-kernel device times, each kernel alone on the stream (event pairs, 20 serial steps on the same games)
    names = ("actions", "k2_step", "k1_movegen", "k3_obs_f32", "k3_feats_bf16")
    acc = dict.fromkeys(names, 0.0)
    PK = 20
    pe = [[ev() for _ in range(len(names) + 1)] for _ in range(PK)]
    rows_log = torch.zeros(PK, dtype=torch.int64, device=dev)   # legal plays of these PK steps (counted here, outside the timed region:
    for i_pk, row in enumerate(pe):                             #  the positions are a stationary random-play mix)
        row[0].record()
        env.random_actions(ACT_SEED, t, out=acts); row[1].record()
        env._apply_actions(acts); row[2].record()
        env._refresh_legal_moves(); row[3].record()
        rows_log[i_pk].copy_(env.alloc_rows[0])
        env.encode_resident(obs=True, afterstates=False); row[4].record()
        if feats:
            env.encode_resident(obs=False, afterstates=True)
        row[5].record()
        t += 1
    torch.cuda.synchronize()
    for row in pe:
        for i, nme in enumerate(names):
            acc[nme] += row[i].elapsed_time(row[i + 1]) / PK
    k1_alone_ms = acc["k1_movegen"]
    step_serial_ms = sum(acc.values())
    rows_per_step = float(rows_log.double().mean().item())
    env.check_status()

    # ---- end to end through the public API with host buffers: actions from pinned host memory every step,
    # rewards / dones / legal-play counts read back to the host every step (what a host-side policy needs).
    E = max(1, args.e2e_steps)
    host = bg_b200.HostStepBuffers(env)                                         # pinned rewards / dones / legal counts
    h_acts = torch.empty(N, dtype=torch.int32).pin_memory()
    rng = np.random.default_rng(1)
    acts_np, counts_np, tmp = h_acts.numpy(), host.legal_counts.numpy(), np.empty(N, dtype=np.int32)
    e2e_segments = []
    for seg in range(args.e2e_segments + 1):                                    # segment 0 is an untimed warm-up of the e2e path
        u = rng.integers(0, 65536, size=(E, N), dtype=np.int32)
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
        if seg > 0:
            e2e_segments.append(time.perf_counter() - t0)
    seg_t = torch.tensor(e2e_segments, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(seg_t, op=dist.ReduceOp.MAX)                            # per segment: the slowest rank
    e2e_segments = seg_t.tolist()
    e2e_s = statistics.median(e2e_segments)
    env.check_status()

    extra = run_extras(bg_b200, env, torch, dev, args, ddist, rank, world) if not args.no_extras else None

    tm = torch.tensor([ms, k1_ms, k1_alone_ms, step_serial_ms] + [acc[n] for n in names], dtype=torch.float64, device=dev)
    rw = torch.tensor([rows_per_step], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(rw, op=dist.ReduceOp.SUM)
        rw /= world
    tl = [float(x) for x in tm.tolist()]
    ms, k1_ms, k1_alone_ms, step_serial_ms = tl[:4]
    acc = dict(zip(names, tl[4:]))
    rows_per_step = float(rw.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None

    ms_step = ms / n_timed
    value = world * N / (ms_step * 1e-3)
    e2e_value = world * N * E / e2e_s
    peak, peak_src, peaks = load_peaks()
    k1_bytes = N * K1_BYTES_PER_GAME + rows_per_step * K1_BYTES_PER_ROW
    achieved = k1_bytes / (k1_ms * 1e-3) / 1e9
    step_bytes = (N * (K2_BYTES_PER_GAME + K1_BYTES_PER_GAME + K3_OBS_BYTES_PER_GAME)
                  + rows_per_step * (K1_BYTES_PER_ROW + (K3_FEAT_BYTES_PER_ROW if feats else 0.0)))
    traffic, issue = None, None
    tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tp):
        try:
            kt = json.load(open(tp))
            traffic = kt.get("dram_bytes_per_launch")
            if kt.get("warp_inst_per_launch") and N == 65536:
                # what actually bounds K1: warp-instruction issue (4 schedulers per SM, one instruction per cycle each)
                sm_clock_ghz = ((clocks or {}).get("sm_mhz") or 1965.0) / 1e3
                peak_issue = 148 * 4 * sm_clock_ghz                             # G warp-instructions / s
                ach = kt["warp_inst_per_launch"] / (k1_alone_ms * 1e-3) / 1e9
                issue = {"warp_inst_per_launch": kt["warp_inst_per_launch"], "achieved": ach, "peak": peak_issue,
                         "unit": "G warp-inst/s", "frac": ach / peak_issue,
                         "what": "instructions of one K1 call from the ncu capture named in profiles/k1_traffic.json / "
                                 "k1_ms_per_launch_alone; peak = 148 SMs x 4 schedulers x the SM clock sampled during the run"}
        except Exception:
            traffic = None
    us = lambda x: round(x * 1e3, 2)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int8", "data": "synthetic",
        "config": {"workload": "configs[1]: batched legal-move generation + step, uniform-random policy",
                   "games_per_gpu": N, "max_legal_moves": 500, "dephase_steps": args.dephase,
                   "afterstate_features": "bf16 ragged (ld 208)" if feats else "off", "observations": "f32 (N,198)",
                   "legal_plays_per_step_mean": rows_per_step / N, "parallelism": f"games sharded x{world}, no collective",
                   "timed_region": f"{R} x {K} = {n_timed} steps back to back ({ms * 1e-3:.3f} s), one barrier + synchronize on each side; "
                                   f"R chosen from the warm-up so that the region lasts >= {args.min_seconds} s",
                   "timed_steps": n_timed, "timed_region_s": ms * 1e-3,
                   "l2": "per-step working set (afterstates + features + observations, ~0.6 GB) exceeds the 126 MB L2; no flush"},
        "roofline": {"kernel": "K1 movegen (tier 0 warp kernel + tier 1/2 team kernels), the dominant kernel",
                     "bound": "issue", "bound_note": "K1 is integer-issue bound (SURVEY 8(d)): roofline.issue is the binding figure; "
                     "achieved/peak/frac here are the HBM figure the contract asks for",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": k1_bytes,
                     "algorithmic_bytes_formula": f"{K1_BYTES_PER_GAME:.0f} B x {N} games + {K1_BYTES_PER_ROW:.0f} B x {rows_per_step:.0f} legal plays",
                     "k1_ms_per_launch": k1_ms, "k1_share_of_step": k1_ms / ms_step,
                     "k1_ms_per_launch_alone": k1_alone_ms, "achieved_alone": k1_bytes / (k1_alone_ms * 1e-3) / 1e9,
                     "frac_alone": k1_bytes / (k1_alone_ms * 1e-3) / 1e9 / peak,
                     "ms_per_step_serial": step_serial_ms, "k1_share_of_step_serial": k1_alone_ms / step_serial_ms,
                     "issue": issue,
                     "step_hbm": {"algorithmic_bytes_per_step": step_bytes, "achieved": step_bytes / (ms_step * 1e-3) / 1e9,
                                  "peak": peak, "unit": "GB/s", "frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak,
                                  "what": "whole step: K2 120 B + K1 71 B + K3 observations 845 B per game, K1 53 B + K3 features 469 B per legal play"},
                     "kernels_us": {k: us(v) for k, v in acc.items()},
                     "kernels_hbm_frac": {
                         "k2_step": N * K2_BYTES_PER_GAME / (acc["k2_step"] * 1e-3) / 1e9 / peak,
                         "k1_movegen": k1_bytes / (acc["k1_movegen"] * 1e-3) / 1e9 / peak,
                         "k3_obs_f32": N * K3_OBS_BYTES_PER_GAME / (acc["k3_obs_f32"] * 1e-3) / 1e9 / peak,
                         "k3_feats_bf16": (rows_per_step * K3_FEAT_BYTES_PER_ROW / (acc["k3_feats_bf16"] * 1e-3) / 1e9 / peak) if feats else None},
                     "note": "k1_ms_per_launch: CUDA events around K1's three launches inside the timed steps, where the encoders (K3) run "
                             "on a second stream beside K1's overflow tiers; *_alone / kernels_us: every kernel alone on the stream "
                             "(20 serial steps on the same games) -- the shares the profiler's serialised launch list shows (profiles/). "
                             "traffic = dram bytes of one K1 call (ncu --set full, profiles/k1_traffic.json): below the algorithmic bytes "
                             "because the afterstate rows are still in the 126 MB L2 when the encoder reads them"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": world * N * 4, "d2h_bytes_per_step": world * N * 9,
                "steps": E, "segments_ms_per_step": [x * 1e3 / E for x in e2e_segments],
                "what": f"median of {len(e2e_segments)} segments of {E} steps after one untimed warm segment (per segment: the slowest rank); "
                        "B200BackgammonVecEnv.step(actions from pinned host memory, host=HostStepBuffers): rewards, dones and legal-play "
                        "counts are copied to pinned host memory every step as soon as K2 + K1 are done and the host waits for them before "
                        "choosing the next actions; observations / afterstate features (K3) stay on the device as the reference API returns "
                        "them, run after K1 and finish while the host prepares the next step"},
        "gpu_launches": n_timed * LAUNCHES_PER_STEP, "clocks": clocks,
    }
    if extra is not None:
        line["extra"] = extra
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n, el = cpu_rollout(args.cpu_seconds, threads)
        line["cpu_baseline"] = {"value": n / el, "unit": UNIT, "cores": threads, "kind": "port", "cpu_model": cpu_model(),
                                "sample": f"{n} env steps of the same uniform-random workload in {el:.1f} s over {threads} "
                                          "threads (oracle/bg_oracle.c, the C port of the reference path)",
                                "python_reference": python_reference_rollout(args.pyref_seconds, threads) if args.pyref_seconds > 0 else None}
    if world > 1:
        dist.destroy_process_group()
    return line


def run_extras(bg_b200, env, torch, dev, args, dist, rank, world):
    """Secondary figures of the same path (BASELINE.json: "2-ply positions evaluated/sec"; configs[2..4]), every one timed
    with CUDA events (max over ranks, units summed over ranks): K4 leaf evaluator alone, 1-ply greedy self-play at the main
    run's size and at configs[2]'s 131,072 games per GPU, the policy kernel, the PPO rollout step, a 3-update PPO loop
    (configs[4]) and the 2-ply search (configs[3]).  Random-init weights of the reference architecture."""
    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps * 1e-3], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def total(x):
        v = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
        return float(v.item())
    peaks = load_peaks()[2]
    net = bg_b200.ValueNet.random_init(dev, seed=0)
    N = env.num_envs
    out = {"ranks": world}
    # K4: value of every current afterstate (ragged buffer, ~1.2 M rows per GPU), fused encode + tcgen05 GEMM + value head
    rows = env.total_rows()
    vbuf = torch.empty(rows, dtype=torch.float32, device=dev)
    t = timed(lambda: net.values(env.after52[:rows], env.row_players[:rows], out=vbuf), 20)
    rows_all = total(rows)
    tf = rows_all * 53504.0 / t / 1e12
    out["mlp_value"] = {"positions_per_s": rows_all / t, "rows": int(rows_all), "ms": t * 1e3, "tflops_bf16": tf,
                        "tflops_bf16_per_gpu": tf / world, "useful_tflops_bf16_per_gpu": rows_all * 50944.0 / t / 1e12 / world,
                        "tensor_peak_tflops": peaks.get("bf16_tflops"),
                        "frac_of_tensor_peak": (tf / world / peaks["bf16_tflops"]) if peaks.get("bf16_tflops") else None,
                        "flop_per_position": "53,504 executed (K padded to 208, bias folded into the GEMM); 50,944 useful (SURVEY 8(d))"}

    # 1-ply greedy self-play (configs[2]): K4 on the afterstates + segment argmax + K2 + K1
    def greedy_on(e):
        def f():
            a_, _ = bg_b200.greedy_actions(e, net)
            e.step_device(a_.clamp_(min=0))
        return f
    t = timed(greedy_on(env), 30)
    out["greedy_1ply"] = {"env_steps_per_s": total(N) / t, "ms_per_step": t * 1e3, "games": int(total(N)),
                          "what": "K4 on every legal afterstate + segment argmax + K2 + K1; max over ranks"}
    if args.config2_games > 0:
        # configs[2] as written: 1,048,576 games on 8 GPUs = 131,072 per GPU (the value net replicated, games sharded)
        env2 = make_env(bg_b200, dev, args.config2_games, rank, world, args.rows_per_game)
        env2.reset()
        a2 = torch.empty(env2.num_envs, dtype=torch.int32, device=dev)
        for tt in range(args.dephase):
            env2.random_actions(ACT_SEED, tt, out=a2)
            env2.step_device(a2)
        t = timed(greedy_on(env2), 20)
        env2.check_status()
        out["greedy_1ply_config2"] = {"env_steps_per_s": total(env2.num_envs) / t, "ms_per_step": t * 1e3, "games": int(total(env2.num_envs)),
                                      "games_per_gpu": env2.num_envs,
                                      "what": "configs[2]: MLP value net 1-ply greedy self-play with on-GPU feature encoding (K4 reads boards, "
                                              "features are built in tensor memory), 131,072 games per GPU = 1,048,576 on 8"}
        del env2, a2
    # N1 policy/value rollout forward (select_action) and the PPO rollout step (configs[4] rollout side): policy kernel -> K2 -> K1
    pnet = bg_b200.PolicyValueNet.random_init(dev, seed=0)
    pout = (torch.empty(N, dtype=torch.int32, device=dev), torch.empty(N, dtype=torch.float32, device=dev),
            torch.empty(N, dtype=torch.float32, device=dev))
    step_ctr = [0]
    t = timed(lambda: pnet.act(env.boards52, env.players, env.legal_counts, seed=1, step=0, out=pout), 20)
    # the same call as a CUDA graph (memset + partition + fused kernel), replayed: the device time without the host's launch path
    # (at 65,536 positions the Python call above is launch-bound: ~45 us of host work per call)
    try:
        gstream = torch.cuda.Stream()
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.stream(gstream):
            pnet.act(env.boards52, env.players, env.legal_counts, seed=1, step=0, out=pout)
            torch.cuda.synchronize()
            with torch.cuda.graph(graph, stream=gstream):
                pnet.act(env.boards52, env.players, env.legal_counts, seed=1, step=0, out=pout)
        torch.cuda.synchronize()
        t_dev = timed(graph.replay, 50)
    except Exception as e:                                         # (a capture failure must not take the bench line down)
        sys.stderr.write("policy_sample: CUDA graph capture failed (%s); device_ms = ms\n" % e)
        torch.cuda.synchronize()
        t_dev = t
    # executed tensor-core work of the call: hidden layer 2*128*208 per position (four times for a class B row: its four chunks are
    # four work items that each recompute it), policy GEMM one 128-slot chunk for class A rows (1..128 legal slots), four for class B
    n_legal = env.legal_counts
    n_cls_a = int(((n_legal >= 1) & (n_legal <= 128)).sum().item())
    flop_exec = n_cls_a * (2 * 128 * 208 + 2 * 128 * 128) + (N - n_cls_a) * 4 * (2 * 128 * 208 + 2 * 128 * 128)
    out["policy_sample"] = {"positions_per_s": total(N) / t, "ms": t * 1e3, "device_ms": t_dev * 1e3,
                            "device_positions_per_s": total(N) / t_dev,
                            "executed_tflops_bf16_per_gpu": flop_exec / t_dev / 1e12, "class_a_rows": n_cls_a, "rows": N,
                            "flop_note": "executed MMAs only (chunks actually multiplied, K padded to 208), not the 500-slot head for every row",
                            "what": "fused policy/value kernel: encode + 198->128 + 128->500 on tcgen05, prefix mask, softmax, categorical sample; "
                                    "ms = through PolicyValueNet.act (host launch path included), device_ms = the same call replayed as a CUDA graph"}

    def ppo_rollout_step():
        step_ctr[0] += 1
        pnet.act(env.boards52, env.players, env.legal_counts, seed=1, stream_base=env.stream_base, step=step_ctr[0], out=pout)
        env.step_device(pout[0])
    t = timed(ppo_rollout_step, 30)
    out["ppo_rollout"] = {"env_steps_per_s": total(N) / t, "ms_per_step": t * 1e3, "games": int(total(N)),
                          "what": "sampled policy (fused policy/value kernel) -> K2 -> K1, no feature tensors in HBM"}
    env.check_status()
    if args.ppo_updates > 0:
        out["ppo_loop"] = run_ppo_loop(bg_b200, env, pnet, torch, dev, args, dist, world)
    out["twoply"] = run_twoply(bg_b200, env, torch, dev, args, net, dist, world)
    return out


def run_ppo_loop(bg_b200, env, pnet, torch, dev, args, dist, world):
    """configs[4]: full PPO self-play loop -- rollout (policy kernel -> K2 -> K1, T steps), returns (bg_gae), update
    (num_epochs full-batch epochs, flat-bucket gradient all-reduce over NCCL) -- on the main run's games."""
    from bg_b200.ppo import PPOConfig, PPOTrainer
    cfg = PPOConfig(t_horizon=args.ppo_horizon)
    tr = PPOTrainer(env, pnet, cfg, dist, seed=0)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for _ in range(2):                                                   # warm-up updates: the caching allocator reaches its steady state in the second
        ret = tr.collect(); tr.count_episodes(); tr.update(ret)         # (device_allocs_per_update in the output shows cudaMalloc calls inside the timed ones)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    marks = [[ev(), ev(), ev()] for _ in range(args.ppo_updates)]
    dev_allocs = []                                                       # cudaMalloc calls of the caching allocator during each update
    for m in marks:
        m[0].record()
        ret = tr.collect()
        m[1].record()
        tr.count_episodes()
        n0 = torch.cuda.memory_stats().get("num_device_alloc", 0)
        stats = tr.update(ret)
        dev_allocs.append(int(torch.cuda.memory_stats().get("num_device_alloc", 0) - n0))
        m[2].record()
    torch.cuda.synchronize()
    ar0, ar1 = ev(), ev()
    ar0.record()
    for _ in range(20):
        tr.learner.fp.all_reduce_grads(dist)
    ar1.record(); torch.cuda.synchronize()
    env.check_status()
    t = torch.tensor([sum(m[0].elapsed_time(m[1]) for m in marks), sum(m[1].elapsed_time(m[2]) for m in marks),
                      ar0.elapsed_time(ar1) / 20], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    roll_ms, upd_ms, ar_ms = [float(x) for x in t.tolist()]
    U, T, N = args.ppo_updates, cfg.t_horizon, env.num_envs
    return {"env_steps_per_s": world * N * T * U / ((roll_ms + upd_ms) * 1e-3), "updates": U, "horizon": T, "games_per_gpu": N,
            "samples_per_update_per_gpu": N * T, "epochs": cfg.num_epochs,
            "rollout_ms_per_update": roll_ms / U, "update_ms_per_update": upd_ms / U,
            "update_ms_all_this_rank": [round(m[1].elapsed_time(m[2]), 3) for m in marks],
            "reserved_gib": round(torch.cuda.memory_reserved() / 2**30, 2), "device_allocs_per_update": dev_allocs,
            "allreduce_ms_per_call": ar_ms, "allreduce_calls_per_update": cfg.num_epochs * cfg.num_minibatches,
            "allreduce_bytes": tr.learner.fp.numel * 4, "last_stats": {k: v for k, v in stats.items() if isinstance(v, float)},
            "what": "configs[4]: rollout + GAE + update, one 90,101-float NCCL all-reduce per optimiser step (the only collective); "
                    "CUDA events, max over ranks"}


def run_twoply(bg_b200, env, torch, dev, args, net, dist, world):
    """2-ply (configs[3]): roots = the first R games' positions of every rank with their actual dice"""
    R = min(args.twoply_roots, env.num_envs)
    search = bg_b200.TwoPlySearch(net, max_afterstates_per_chunk=args.twoply_chunk)
    b, p, d = env.boards52[:R].clone(), env.players[:R].clone(), env.dice[:R].clone()
    f = search.search_device(b, p, d)                         # warm-up: allocates the persistent buffers
    torch.cuda.synchronize()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for _ in range(7):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ea.record()
        f = search.search_device(b, p, d)                     # bg_twoply: one C call, no host synchronisation inside
        eb.record()
        torch.cuda.synchronize()
        times.append(ea.elapsed_time(eb) * 1e-3)
    st = int(f["status"].item())
    if st:
        raise RuntimeError(f"2-ply search: device status {st}")
    n_after, n_leaves, ovf_items, ovf_leaves = (int(x) for x in f["stats"].tolist())
    tt = torch.tensor(times, dtype=torch.float64, device=dev)
    agg = torch.tensor([n_after, n_leaves, R, ovf_items, ovf_leaves], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX); dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    t = statistics.median(tt.tolist())
    a, l, r, oi, ol = agg.tolist()
    return {"roots": int(r), "root_afterstates": int(a), "leaves": int(l), "seconds": t, "seconds_all": tt.tolist(),
            "root_positions_per_s": r / t, "root_afterstates_per_s": a / t, "leaves_per_s": l / t,
            "leaf_tflops_bf16_per_gpu": l * 53504.0 / t / 1e12 / world,
            "overflow_items": int(oi), "overflow_items_frac": oi / max(1.0, a * 21), "overflow_leaves_frac": ol / max(1.0, l),
            "note": f"{world} rank(s), each its own {R} roots; bg_twoply (root move generation + fused on-chip expansion / leaf MLP / max + "
                    "scores + argmax in one call); median of 7 searches, CUDA events, per search the slowest rank; a 2-ply position = one "
                    "root afterstate fully expanded (21 opponent rolls x replies, leaves MLP-evaluated)"}


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


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--games", type=int, default=65536, help="games per GPU")
    ap.add_argument("--dephase", type=int, default=128)
    ap.add_argument("--rows-per-game", type=int, default=64)
    ap.add_argument("--min-seconds", type=float, default=0.5, help="minimum length of the timed region (R x K steps)")
    ap.add_argument("--max-reps", type=int, default=400)
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--e2e-segments", type=int, default=7)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--pyref-seconds", type=float, default=20.0, help="seconds of the Python reference per process (0 = skip)")
    ap.add_argument("--no-afterstate-features", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="run the encoders after K1 on the same stream")
    ap.add_argument("--e2e-overlap-encoders", action="store_true",
                    help="e2e loop: encoders beside K1's overflow tiers (shorter GPU step, but K1 -- which the host waits for -- ends later)")
    ap.add_argument("--twoply-roots", type=int, default=4096)
    ap.add_argument("--twoply-chunk", type=int, default=98304)
    ap.add_argument("--config2-games", type=int, default=131072, help="games per GPU of the configs[2] greedy extra (0 = skip)")
    ap.add_argument("--ppo-updates", type=int, default=3, help="timed updates of the configs[4] PPO loop extra (0 = skip)")
    ap.add_argument("--ppo-horizon", type=int, default=64)
    return ap.parse_args(argv)


def main():
    args = parse_args()
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
