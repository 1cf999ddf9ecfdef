"""networksim_cntfet_b200 -- B200-native (sm_100a) implementation of the data-parallel hot path
of leobrowning92/networksim-cntfet: random CNT stick networks -> junctions -> percolating
clusters -> MNA conductance solve, batched over independent networks.

Host modules mirror the reference's `netsim`, `cnet` and `measure_perc` modules (same entry
points); the compute lives in hand-written CUDA kernels behind the C-ABI of
include/cntnet_b200.h (libcntnet_b200.so).  There is no CPU fallback.
"""
__version__ = "0.1.0"

_engine = None


def get_engine():
    """Process-wide default Engine (one process per GPU).  Raises if there is no CUDA device
    or the CUDA library is not built."""
    global _engine
    if _engine is None:
        from .engine import Engine
        _engine = Engine()
    return _engine


def set_engine(engine):
    global _engine
    _engine = engine
    return engine
