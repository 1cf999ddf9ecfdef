"""Drop-in for the reference's `cnet.py`: conduction elements and `ConductionNetwork`.

Same class names, constructor arguments, attributes and methods as cnet.py:20-183; the
bodies of `update()` and the gate setters run on the GPU (assembly kernel, batched
Jacobi-PCG, write-back kernel) instead of networkx + scipy LIL matrices + SuperLU.

Node order is the canonical one (ascending node label): the
reference's `ground_nodes` / `voltage_sources` are *positions* in that order (cnet.py:121-139).
"""
import numpy as np

from .elements import FermiDiracTransistor, LinExpTransistor, Resistor, StepTransistor  # noqa: F401
from .elements import KIND_CHARS, pair_string


class _EdgeView(dict):
    pass


class ConductionNetwork(object):
    """Solves for the conduction characteristics of a physical network (cnet.py:89-183).

    Two ways to build one:
      * `ConductionNetwork(graph, ground_nodes, voltage_sources)` -- the reference signature:
        `graph` is a networkx graph whose edges carry `component` (an element with
        `gate_voltage` / `get_conductance()`), `kind` and `pos`; exactly one ground node and one
        voltage source are supported (the only configuration the reference ever builds,
        netsim.py:181-182).  Conductances are evaluated per edge on the host through the
        element objects, everything else runs on the GPU.
      * `ConductionNetwork._from_batch(...)` -- used by `netsim.RandomConductingNetwork`: the
        junction table already lives on the GPU and conductances come from a per-kind table.
    """

    def __init__(self, graph, ground_nodes, voltage_sources, engine=None):
        from . import get_engine
        self._engine = engine or get_engine()
        self.ground_nodes = np.array(ground_nodes)
        self.voltage_sources = np.array(voltage_sources)
        self.gate_areas = []
        self.vds = 0.1
        self._graph = graph
        self._graph_dirty = False
        self._table_mode = False
        self._build_from_graph(graph)

    # ------------------------------------------------------------------ construction
    def _build_from_graph(self, graph):
        assert len(self.ground_nodes) == 1 and len(self.voltage_sources) == 1, \
            "GPU path supports one ground node and one voltage source"
        nodes = sorted(graph.nodes())
        self.network_size = len(nodes)
        self._nodes = nodes
        src = nodes[int(self.voltage_sources[0][0])]
        gnd = nodes[int(self.ground_nodes[0])]
        self.vds = float(self.voltage_sources[0][1])
        if abs(self.vds - 0.1) > 0:
            self._vscale = self.vds / 0.1      # kernels are written for Vds = 0.1 (netsim.py:182); the system is linear
        else:
            self._vscale = 1.0
        # relabel: source -> 0, ground -> 1, the rest ascending
        rest = [v for v in nodes if v != src and v != gnd]
        order = [src, gnd] + rest
        rank = {v: k for k, v in enumerate(order)}
        self._order = order
        edges = list(graph.edges())
        a = np.array([min(rank[u], rank[v]) for u, v in edges], dtype=np.int32)
        c = np.array([max(rank[u], rank[v]) for u, v in edges], dtype=np.int32)
        o = np.lexsort((c, a))
        self._edges = [edges[k] for k in o]
        self._st1, self._st2 = a[o], c[o]
        pos = [graph.edges[e].get("pos", [np.nan, np.nan]) for e in self._edges]
        self._jx = np.array([p[0] for p in pos], dtype=np.float64)
        self._jy = np.array([p[1] for p in pos], dtype=np.float64)
        self._components = [graph.edges[e]["component"] for e in self._edges]
        eng = self._engine
        n = len(order)
        self._batch = eng.batch_from_junctions(n, self._st1, self._st2, self._jx, self._jy,
                                               np.zeros(len(a), dtype=np.uint8))
        eng.label_clusters(self._batch)
        eng.assemble_pattern(self._batch)

    @classmethod
    def _from_batch(cls, engine, batch, element, onoffmap, owner=None):
        self = cls.__new__(cls)
        from .engine import GateState
        self._engine = engine
        self._batch = batch
        self._element = element
        self._onoffmap = onoffmap
        self._table_mode = True
        self._owner = owner
        self.ground_nodes = np.array([1])
        self.voltage_sources = np.array([[0, 0.1]])
        self.gate_areas = []
        self.vds = 0.1
        self._vscale = 1.0
        self.network_size = int(batch.h_stats[0][5])       # sticks in the percolating component
        self._gate = GateState(engine, batch)
        self._graph = None
        self._graph_dirty = True
        self._last = None
        return self

    # ------------------------------------------------------------------ gates (cnet.py:159-183)
    def set_global_gate(self, voltage):
        if self._table_mode:
            self._gate.set_global(voltage)
        else:
            for comp in self._components:
                comp.gate_voltage = voltage

    def set_local_gate(self, area, voltage):
        if self._table_mode:
            self._gate.set_local(area, voltage)
        else:
            for k in self._local_edge_ids(area):
                self._components[k].gate_voltage = voltage
        self.gate_areas.append([area, voltage])

    def check_in_area(self, point, area):
        """point is [x, y]; area is [centre x, centre y, x width, y length] (closed rectangle)"""
        half_w, half_h = area[2] / 2, area[3] / 2
        return bool(area[0] - half_w <= point[0] <= area[0] + half_w and
                    area[1] - half_h <= point[1] <= area[1] + half_h)

    def _local_edge_ids(self, area):
        return [k for k in range(len(self._edges)) if self.check_in_area([self._jx[k], self._jy[k]], area)]

    def get_local_edges(self, area):
        g = self.graph
        return [e for e in g.edges if self.check_in_area(g.edges[e]["pos"], area)]

    # ------------------------------------------------------------------ solve (cnet.py:150-156)
    def update_conductivity(self):
        """cnet.py:99-103: refresh per-edge conductances (host element objects in graph mode;
        in table mode this happens inside the assembly kernel)."""
        if not self._table_mode:
            self._g_host = np.array([float(c.get_conductance()) for c in self._components], dtype=np.float64)

    def update(self, show=True, v=False, rtol=1e-15):
        eng, b = self._engine, self._batch
        if self._table_mode:
            res = eng.solve_gates(b, [self._gate], self._onoffmap, self._element, rtol=rtol, want_fields=True)
        else:
            self.update_conductivity()
            res = eng.solve_given_conductances(b, self._g_host, rtol=rtol)
        eng.sync()
        self._last = res
        self.status = int(res["status"][0, 0])
        isrc = res["isrc"][0, 0].cpu().numpy() * self._vscale
        # current through the voltage source (x[-1] of the MNA system, cnet.py:135), power form
        self.source_currents = np.array([isrc[2]])
        self.source_currents_estimators = isrc
        self._graph_dirty = True

    def sheet_conductance(self):
        """G_sq = I_source / Vds of the square sample for the current gate state (derived quantity: the reference
        stops at source_currents, cnet.py:135)"""
        return float(np.sum(self.source_currents)) / self.vds

    def solve_mna(self):
        """cnet.py:147-149: returns [V of the non-ground nodes ..., source current]."""
        self.update()
        V = self.voltages()
        keep = np.ones(len(V), dtype=bool)
        keep[int(self.ground_nodes[0])] = False
        return np.append(V[keep], self.source_currents)

    def voltages(self):
        """node voltages in canonical (ascending label) order"""
        V = self._last["V"][0].cpu().numpy()[:self._batch.S] * self._vscale
        if self._table_mode:
            return V[~np.isnan(V)]
        rank = {v: k for k, v in enumerate(self._order)}
        return np.array([V[rank[v]] for v in self._nodes])

    def edge_currents(self):
        return self._last["I"][0].cpu().numpy()[:self._batch.J] * self._vscale

    def edge_conductances(self):
        return self._last["g"][0].cpu().numpy()[:self._batch.J]

    # ------------------------------------------------------------------ host matrices (cnet.py:104-131)
    def make_G(self):
        """negative Laplacian of the component in canonical node order (scipy LIL), cnet.py:104-111"""
        import scipy.sparse as sp
        g = self.edge_conductances() if self._last is not None else None
        if g is None:
            raise RuntimeError("call update() first")
        b = self._batch
        st1 = b.stick1.cpu().numpy()[:b.J]
        st2 = b.stick2.cpu().numpy()[:b.J]
        lab = b.label.cpu().numpy()[:b.S]
        nodes = np.flatnonzero(lab == 0)
        cid = -np.ones(b.S, dtype=np.int64)
        cid[nodes] = np.arange(len(nodes))
        e = lab[st1] == 0
        r, c, w = cid[st1[e]], cid[st2[e]], g[e]
        W = sp.coo_matrix((np.r_[w, w], (np.r_[r, c], np.r_[c, r])), shape=(len(nodes),) * 2).tocsr()
        return (W - sp.diags(np.asarray(W.sum(axis=0)).ravel())).tolil()

    def make_A(self, G):
        import scipy.sparse as sp
        n = G.shape[0]
        Bm = np.zeros((n, 1))
        Bm[int(self.voltage_sources[0][0]), 0] = 1
        A = sp.bmat([[G, Bm], [Bm.T, np.zeros((1, 1))]], format="csr")
        keep = np.ones(n + 1, dtype=bool)
        keep[int(self.ground_nodes[0])] = False
        return A[keep][:, keep].tocsr()

    def make_z(self):
        return np.append(np.zeros((self.network_size - 1, 1)), [[self.vds]], axis=0)

    # ------------------------------------------------------------------ networkx view
    @property
    def graph(self):
        """networkx view of the conducting component with the reference's attributes:
        node `voltage`, `pos`; edge `conductance`, `resistance`, `current`, `kind`, `pos`, `x`, `y`
        (cnet.py:99-103,137-146; netsim.py:184-187).  Built lazily: ensembles never touch it."""
        if self._graph is None or (self._graph_dirty and self._table_mode):
            self._graph = self._materialise_graph()
            self._graph_dirty = False
        elif self._graph_dirty and self._last is not None:
            V = self._last["V"][0].cpu().numpy() * self._vscale
            I, g = self.edge_currents(), self.edge_conductances()
            rank = {v: k for k, v in enumerate(self._order)}
            for v in self._graph.nodes():
                self._graph.nodes[v]["voltage"] = float(V[rank[v]])
            for k, e in enumerate(self._edges):
                d = self._graph.edges[e]
                d["conductance"] = float(g[k])
                d["resistance"] = 1 / float(g[k])
                d["current"] = float(I[k])
            self._graph_dirty = False
        return self._graph

    @graph.setter
    def graph(self, value):
        self._graph = value

    def _materialise_graph(self):
        import networkx as nx
        b = self._batch
        st1 = b.stick1.cpu().numpy()[:b.J]
        st2 = b.stick2.cpu().numpy()[:b.J]
        jx, jy = b.jx.cpu().numpy()[:b.J], b.jy.cpu().numpy()[:b.J]
        k2 = b.kind2.cpu().numpy()[:b.J]
        lab = b.label.cpu().numpy()[:b.S]
        xc, yc = b.xc.cpu().numpy(), b.yc.cpu().numpy()
        G = nx.Graph()
        have = self._last is not None
        if have:
            V = self._last["V"][0].cpu().numpy()
            I, g = self.edge_currents(), self.edge_conductances()
        lev = self._gate.level.cpu().numpy()
        for v in np.flatnonzero(lab == 0):
            G.add_node(int(v), pos=[float(xc[v]), float(yc[v])])
            if have:
                G.nodes[int(v)]["voltage"] = float(V[v])
        for k in np.flatnonzero(lab[st1] == 0):
            kind = pair_string(int(k2[k]))
            comp = self._element(kind, self._onoffmap)
            comp.gate_voltage = self._gate.levels[int(lev[k])]
            attrs = dict(x=float(jx[k]), y=float(jy[k]), kind=kind, pos=[float(jx[k]), float(jy[k])], component=comp)
            if have:
                attrs.update(conductance=float(g[k]), resistance=1 / float(g[k]), current=float(I[k]))
            G.add_edge(int(st1[k]), int(st2[k]), **attrs)
        return G
