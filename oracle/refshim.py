"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference, usable in the build
container only (`/root/reference` does not exist on the GPU box).

It is used by `tests/golden/make_golden.py` to produce the committed golden fixtures and by
`tests/test_oracle_vs_reference.py` (skipped when the reference is absent) to pin the numpy
restatement in `oracle/cntnet_oracle.py`.  Nothing in the product package imports this.

Compat shims -- none of them touches numerics:
  1. `netsim.py:15` imports matplotlib (unused there; not installed)  -> empty stub module.
  2. `cnet.py:107` calls `nx.to_scipy_sparse_matrix` (removed in networkx 3) -> re-defined on
     top of `nx.to_scipy_sparse_array`.
  3. canonical semantics (hazard H1): `make_G` is overridden to pass
     `nodelist=sorted(graph.nodes())`, the order the reference's own `update_voltages`
     (`cnet.py:137-139`) writes results back in.  Under networkx 2.2 (the author's pin) the
     two orders coincided; under networkx 3 they do not, which silently moves the ground
     node.  The override restores the author-observed behaviour.
  4. `update_voltages` evaluates `sorted(nodes)` inside its loop (`cnet.py:137-138`,
     O(n^2 log n)); `hoist=True` replaces it by a result-identical single sort so that
     20x20 um goldens finish in seconds.
"""
import os
import sys
import types

REFERENCE_DIR = os.environ.get("CNTNET_REFERENCE_DIR", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "netsim.py"))


_loaded = None


def load(hoist=True):
    """Import the reference's `netsim`, `cnet` modules with the compat shims applied."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_DIR)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            sys.modules["matplotlib"] = types.ModuleType("matplotlib")
    import networkx as nx
    import scipy.sparse as sparse
    import scipy.sparse.linalg  # noqa: F401  (reference relies on sparse.linalg being loaded)

    if not hasattr(nx, "to_scipy_sparse_matrix"):
        def to_scipy_sparse_matrix(G, nodelist=None, weight="weight", format="csr"):
            return sparse.lil_matrix(nx.to_scipy_sparse_array(G, nodelist=nodelist, weight=weight))
        nx.to_scipy_sparse_matrix = to_scipy_sparse_matrix
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import cnet as ref_cnet
    import netsim as ref_netsim

    def make_G_canonical(self):
        G = nx.to_scipy_sparse_matrix(self.graph, nodelist=sorted(self.graph.nodes()),
                                      weight="conductance", format="lil")
        dsum = G.sum(axis=0)
        for i in range(self.network_size):
            G[i, i] = -dsum[0, i]
        return G
    ref_cnet.ConductionNetwork.make_G = make_G_canonical

    if hoist:
        import numpy as np

        def update_voltages_hoisted(self, x):
            for i in self.ground_nodes:
                x = np.insert(x, i, 0, axis=0)
            self.source_currents = x[-len(self.voltage_sources):]
            x = x[:-len(self.voltage_sources)]
            nodes = sorted(self.graph.nodes())
            for i in range(len(nodes)):
                self.graph.nodes[nodes[i]]["voltage"] = float(x[i])
        ref_cnet.ConductionNetwork.update_voltages = update_voltages_hoisted
    _loaded = (ref_netsim, ref_cnet)
    return _loaded
