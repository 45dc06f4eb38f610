"""CPU-only checks: the C-ABI library loads and exports every symbol include/bg_b200.h declares, the host
wrappers refuse to run without CUDA (no fallback), and layout helpers round-trip."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "bg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import bg_b200
    bg_b200.build()
    L = bg_b200.lib()
    names = _declared_functions()
    assert len(names) >= 10
    from importlib import import_module
    sigs = import_module("bg_b200._lib").SIGNATURES
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/bg_b200.h but not exported"
        assert n in sigs, f"{n} has no ctypes signature"
    assert L.bg_version() >= 100
    assert L.bg_movegen_workspace_bytes(1000) >= 4000


def test_argument_validation_without_gpu():
    import bg_b200
    L = bg_b200.lib()
    assert L.bg_movegen_count(None, None, None, -1, None, None, None, 0, None) == -1
    assert b"negative" in L.bg_last_error()
    assert L.bg_encode_f32(None, None, 0, 1, None, None, 197, None) == -1


def test_no_cpu_fallback():
    import bg_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(bg_b200.BgError):
        bg_b200.B200BackgammonVecEnv(num_envs=2, device="cpu")
    with pytest.raises(bg_b200.BgError):
        bg_b200.legal_moves(torch.zeros((1, 52), dtype=torch.int8), torch.zeros(1, dtype=torch.int8),
                            torch.ones((1, 2), dtype=torch.int8))
    with pytest.raises(bg_b200.BgError):
        bg_b200.encode(torch.zeros((1, 52), dtype=torch.int8), 0)


def test_layout_roundtrip_matches_oracle_packing():
    import bg_b200
    from oracle import bg_oracle as O
    d = np.load(os.path.join(ROOT, "tests", "golden", "allrolls.npz"))
    b52 = torch.as_tensor(d["boards"][:500])
    b96 = bg_b200.from_board52(b52)
    assert np.array_equal(b96.numpy(), O.unpack52(d["boards"][:500]))
    assert torch.equal(bg_b200.to_board52(b96), b52)
    assert np.array_equal(bg_b200.initial_board52(1, "cpu").numpy()[0], O.pack52(O.initial_board())[0])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mlp-ppo-2ply-p3_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                for needle in ("import oracle", "from oracle", "bg_oracle", "libbg_oracle", "oracle/"):
                    assert needle not in txt, f"{f} references the oracle ({needle})"


def test_render_initial_position():
    import numpy as np
    import bg_b200
    from bg_b200.engine import render_board52
    b = np.zeros(52, np.int8)
    for p, c in ((0, 2), (11, 5), (16, 3), (18, 5), (24 + 23, 2), (24 + 12, 5), (24 + 7, 3), (24 + 5, 5)):
        b[p] = c
    b[48], b[51] = 1, 2                                  # one PLAYER1 man on the bar, two PLAYER2 men borne off
    text = render_board52(b)
    lines = text.split("\n")
    assert lines[0].startswith("| 12 | 13 |") and "P=O Home Board" in lines[1]
    top = lines[2]                                       # first row of the top half: points 12..17 | bar | 18..23 | off
    cells = [c.strip() for c in top.strip("|").split("|")]
    assert cells == ["O", "", "", "", "X", "", "", "X", "", "", "", "", "O", "O"]
    assert sum(l.count("X") for l in lines) == 2 + 5 + 3 + 5 + 1 + 1    # 15 men + the bar token + the legend
