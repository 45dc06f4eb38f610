"""Sizing data for K1's per-warp scratch: level sizes of doubles positions met in random self-play (positions from the oracle env,
levels from the host emulation of the kernel's algorithm).  Test infrastructure only (imports oracle/)."""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bg_oracle as O
SRC, SO = os.path.join(ROOT, "tests/host_emul/emul.cu"), os.path.join(ROOT, "tests/host_emul/libemul.so")
if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-Wno-deprecated-gpu-targets",
                           "-I", os.path.join(ROOT, "mlp-ppo-2ply-p3_b200/csrc"), "-o", SO, SRC])
L = C.CDLL(SO)
L.emul_doubles_levels.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
rng = np.random.default_rng(0)
env = O.Env(); env.set_philox(1234, 0); env.reset()
rows, positions = [], 0
for t in range(steps):
    s = env.state()
    positions += 1
    if s["roll"][0] == s["roll"][1]:
        b52 = O.pack52(s["board"][None])[0]
        sz, cd = np.zeros(4, np.int32), np.zeros(4, np.int32)
        L.emul_doubles_levels(b52.ctypes.data, s["player"], int(s["roll"][0]), sz.ctypes.data, cd.ctypes.data)
        rows.append(np.concatenate([sz, cd]))
    n = s["n_legal"]
    _, done, info = env.step(int(rng.integers(n)) if n > 0 else 0)
    if s["match_over"] or info.get("match_over"): env.reset()
R = np.array(rows); sz, cd = R[:, :4], R[:, 4:]
mx = sz.max(1)
pair = np.maximum.reduce([sz[:, 0] + 1, sz[:, 0] + sz[:, 1], sz[:, 1] + sz[:, 2], sz[:, 2] + sz[:, 3]])
print(f"positions {positions}, doubles {len(R)}")
for cap in (96, 128, 160, 192, 256):
    print(f"  max level > {cap}: {np.mean(mx > cap) * len(R) / positions * 100:.3f} % of positions; candidates share {cd[mx > cap].sum() / max(cd.sum(), 1) * 100:.1f} % of doubles candidates")
for tot in (192, 256, 320, 384):
    print(f"  adjacent levels sum > {tot}: {np.mean(pair > tot) * len(R) / positions * 100:.3f} % of positions")
print("  doubles: mean boards per level", sz.mean(0), "mean candidates per level", cd.mean(0))
ov = mx > 128
print("  overflowing (>128): mean level sizes", sz[ov].mean(0), "mean candidates", cd[ov].mean(0), " last level >512:", np.mean(mx > 512) * len(R) / positions * 100, "%")
# routing predictor for the overflow tiers: a position whose THIRD level already exceeds 128 boards ...
early = (sz[:, :3].max(1) > 128)
big = mx > 512
print(f"  level <= 3 already > 128: {early.sum() / positions * 100:.3f} % of positions; of the {big.sum()} positions with a level > 512, {int((early & big).sum())} are among them;"
      f" of the early ones {int((early & ~big).sum())} would have fitted 512")
