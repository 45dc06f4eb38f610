"""K1 against the oracle at scale: 30,000 self-play positions x all 36 ordered rolls = 1.08 M move generations (all three
tiers: ~15,000 of them over 128 plays, ~60 over 512), afterstate lists compared in order.  BG_FUZZ_POSITIONS overrides."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_k1_differential_fuzz_large():
    n = os.environ.get("BG_FUZZ_POSITIONS", "30000")
    r = subprocess.run([sys.executable, os.path.join(HERE, "fuzz_k1.py"), "--positions", n], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout + r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["differences"] == 0 and d["position_roll_pairs"] == int(n) * 36 and d["pairs_over_128"] > 0
