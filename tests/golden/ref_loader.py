"""Import the UNMODIFIED reference (/root/reference) inside the build container.

Used only by tests/golden/make_golden.py to generate golden vectors; the
reference does not exist on the GPU box, so nothing under tests/ that runs
there may import this module's `load()`.

The reference's package __init__ files eagerly import the trainer (boto3,
tensorboardX, `from config import *`), and `gym` is not installed; we register
synthetic parent packages and inert stubs for those.
"""
import os
import sys
import types

REF = os.environ.get("BG_REFERENCE", "/root/reference")


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "src", "moves"))


def load(ref: str = REF):
    if "src.moves" in sys.modules:
        return
    for name, path in (("src", f"{ref}/src"), ("src.agent", f"{ref}/src/agent")):
        pkg = types.ModuleType(name)
        pkg.__path__ = [path]
        sys.modules[name] = pkg
    sys.modules["src"].agent = sys.modules["src.agent"]

    class Env:  # gym.Env stand-in (backgammon_env.py:35)
        def close(self):
            pass

    class Box:
        def __init__(self, low, high, shape, dtype):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    class Discrete:
        def __init__(self, n):
            self.n = n

    spaces = _mod("gym.spaces", Box=Box, Discrete=Discrete)
    _mod("gym", Env=Env, spaces=spaces)

    class _Inert:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, n):
            return lambda *a, **k: None

    _mod("botocore", exceptions=_mod("botocore.exceptions", ClientError=Exception),
         config=_mod("botocore.config", Config=_Inert))
    _mod("boto3", client=lambda *a, **k: _Inert())
    _mod("tensorboardX", SummaryWriter=_Inert,
         record_writer=_mod("tensorboardX.record_writer", RecordWriter=_Inert, S3RecordWriter=_Inert))
    import src.moves  # noqa: F401  (must come first: board<->moves import cycle)
