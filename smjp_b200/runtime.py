"""Process-wide engine handle and the seed stream of the drop-in API.

The reference draws from the global, never-seeded numpy.random state.  The engine is counter based:
every call takes an explicit 64-bit seed.  Drop-in functions that are called without one pull the
next seed from a package-level generator, which `seed(x)` makes reproducible."""
import numpy as np

_ENGINES = {}
_RNG = np.random.default_rng()


def seed(value):
    """Make every subsequent un-seeded drop-in call reproducible."""
    global _RNG
    _RNG = np.random.default_rng(value)


def rng():
    return _RNG


def next_seed():
    return int(_RNG.integers(0, 2 ** 63 - 1))


def get_engine(device=0):
    """Lazily created Engine per device.  Raises if the CUDA library or a GPU is missing: there is
    no CPU path."""
    from .engine import Engine
    if device not in _ENGINES:
        _ENGINES[device] = Engine(device)
    return _ENGINES[device]


def close_engines():
    for e in _ENGINES.values():
        e.close()
    _ENGINES.clear()
