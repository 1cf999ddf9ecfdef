"""Drop-in for the reference's `netsim.py`: `RandomConductingNetwork` / `RandomCNTNetwork`.

Same constructor arguments, attributes (`sticks`, `intersects`, `percolating`, `cnet`, `graph`,
`clustersizes`, `seed`, `fname`) and methods as netsim.py:23-277.  Stick placement, the junction
search, cluster labelling and the electrical solve run on the GPU through
`networksim_cntfet_b200.engine.Engine`; pandas / networkx objects are only materialised for
the caller (lazily where the reference's own callers do not need them).

Differences, all documented in DESIGN.md:
  * `rng='philox'` (default) draws sticks with the counter-based GPU generator;
    `rng='mt19937'` replays the reference's numpy stream on the host (netsim.py:42,105-117) so a
    seed recorded by the reference reproduces the same sticks;
  * sort ties follow the canonical rule (source electrode index 0, drain 1, stable), where the
    reference inherits pandas' unstable quicksort;
  * node order in the solve is canonical.
"""
import os
import sys
import traceback
from datetime import datetime

import numpy as np
import pandas as pd

from .cnet import ConductionNetwork, FermiDiracTransistor, LinExpTransistor, Resistor  # noqa: F401
from .elements import KIND_CHARS, pair_code, pair_string

_KCODE = {"v": 0, "m": 1, "s": 2}


class RandomConductingNetwork(object):
    def __init__(self, n=2, scaling=5, l='exp', pm=0.135, fname='', directory='data', notes='', seed=0,
                 onoffmap=0, element=LinExpTransistor, rng='philox', engine=None):
        from . import get_engine
        self._engine = engine or get_engine()
        self.scaling = scaling
        self.n = n
        self.pm = pm
        self.l = l
        self.notes = notes
        self.directory = directory
        self.percolating = False
        self.onoffmap = onoffmap
        self.element = element
        self.rng = rng
        self._graph = None
        self._batch = None
        self._gen = None
        # seeds are included to ensure proper randomness on distributed computing (netsim.py:37-42)
        if seed:
            self.seed = seed
        else:
            self.seed = np.random.randint(low=0, high=2 ** 32)
        np.random.seed(self.seed)

        if not fname:
            self.sticks, self.intersects = self.make_intersects_kdtree(
                self.make_sticks(n, l=l, pm=pm, scaling=scaling))
            self.make_cnet()
            self.fname = self.make_fname()
        else:
            self.fname = fname
            self.load_system(os.path.join(directory, fname))

    # ------------------------------------------------------------------ reporting (netsim.py:52-73)
    def get_info(self):
        print('=== input parameters ===')
        print('number of sticks: {}'.format(self.n))
        print('stick length : {} ± {} μm'.format(0.66, 0.44))
        print('percentage metallic: {} %'.format(self.pm))
        print('\n=== physical network characteristics ===')
        print('device region: {}x{} μm'.format(self.scaling, self.scaling))
        print('stick density: {} sticks/μm^2'.format(self.n / self.scaling ** 2))
        print('number of clusters: {}'.format(len(self.clustersizes)))
        print('size of stick clusters: {:.2f} ± {:.2f} sticks'.format(self.clustersizes.mean(),
                                                                          self.clustersizes.std()))
        print('maximum cluster size: {} sticks'.format(self.clustersizes.max()))
        print('\n=== electrical characteristics ===')
        print("device conducting: {}".format(self.percolating))
        if self.percolating:
            print('driving voltage: {} V'.format(self.cnet.vds))
            current = sum(self.cnet.source_currents)
            cur = self.cnet.edge_currents()
            cur = pd.Series(cur[~np.isnan(cur)])
            print('device current: {:.2f} A'.format(current))
            print('current variation (std dev across sticks): {:.2e} ± {:.2e} A'.format(cur.mean(), cur.std()))
            return [self.n, self.scaling, self.n / self.scaling ** 2, len(self.clustersizes), self.clustersizes.mean(),
                    self.clustersizes.std(), self.clustersizes.max(), self.percolating, self.cnet.vds, current,
                    cur.mean(), cur.std(), self.fname, self.seed]

    # ------------------------------------------------------------------ geometry helpers
    def check_intersect(self, s1, s2):
        """netsim.py:75-94 for two 2x2 end arrays [[xa,ya],[xb,yb]]; returns [x, y] or False.
        (Host convenience; the batched search uses the CUDA predicate with the same op order.)"""
        s1, s2 = np.asarray(s1, dtype=np.float64), np.asarray(s2, dtype=np.float64)
        with np.errstate(all="ignore"):
            m1 = (s1[0, 1] - s1[1, 1]) / (s1[0, 0] - s1[1, 0])
            m2 = (s2[0, 1] - s2[1, 1]) / (s2[0, 0] - s2[1, 0])
            b1 = s1[0, 1] - m1 * s1[0, 0]
            b2 = s2[0, 1] - m2 * s2[0, 0]
            if m1 == m2:
                return False
            xi = (b2 - b1) / (m1 - m2)
            yi = (b2 * m1 - b1 * m2) / (m1 - m2)
        lo = max(s1[:, 0].min(), s2[:, 0].min())
        hi = min(s1[:, 0].max(), s2[:, 0].max())
        return [xi, yi] if lo < xi < hi else False

    def get_distance(self, p1, p2):
        return np.sqrt((p1[0] - p2[0]) ** 2 + (p1[1] - p2[1]) ** 2)

    def get_ends(self, row):
        xc, yc, angle, length = row[0], row[1], row[2], row[3]
        dx, dy = length / 2 * np.cos(angle), length / 2 * np.sin(angle)
        return np.array([[xc - dx, yc + dy], [xc + dx, yc - dy]])

    # ------------------------------------------------------------------ sticks (netsim.py:105-125)
    def make_stick(self, l=None, kind='s', pm=0, scaling=1):
        """one stick [xc, yc, angle, length, kind, endarray] from the numpy global stream"""
        if np.random.rand() <= pm:
            kind = 'm'
        if type(l) != str:
            stick = [np.random.rand(), np.random.rand(), np.random.rand() * 2 * np.pi, l / scaling, kind]
        elif l == 'exp':
            stick = [np.random.rand(), np.random.rand(), np.random.rand() * 2 * np.pi,
                     abs(np.random.normal(0.66, 0.44)) / scaling, kind]
        else:
            print('invalid L value')
        stick.append(self.get_ends(stick))
        return stick

    def _frame(self, t):
        """dict of arrays -> DataFrame with the reference's columns"""
        ends = np.stack([np.stack([t["xa"], t["ya"]], 1), np.stack([t["xb"], t["yb"]], 1)], 1)
        df = pd.DataFrame({"xc": t["xc"], "yc": t["yc"], "angle": t["angle"], "length": t["length"],
                           "kind": np.array(list(KIND_CHARS))[np.asarray(t["kind"], dtype=np.int64)]})
        df["endarray"] = list(ends)
        return df

    def make_sticks(self, n, **kwargs):
        """source + n random sticks + drain, in generation (pre-sort) order (netsim.py:119-125)"""
        l, pm, scaling = kwargs.get("l", None), kwargs.get("pm", 0), kwargs.get("scaling", 1)
        if self.rng == 'mt19937':
            source = [0.01, 0.5, np.pi / 2 - 1e-6, 100, 'v']
            source.append(self.get_ends(source))
            drain = [.99, 0.5, np.pi / 2 - 1e-6, 100, 'v']
            drain.append(self.get_ends(drain))
            return pd.DataFrame([source] + [self.make_stick(**kwargs) for i in range(n)] + [drain],
                                columns=["xc", "yc", "angle", "length", 'kind', "endarray"])
        eng = self._engine
        b = eng.generate([n], scaling, seeds=[int(self.seed)], pm=pm, l=l)
        t = eng.sticks_to_host(b)[0]
        inv = np.argsort(t["orig"], kind="stable")          # back to generation order
        df = self._frame({k: v[inv] for k, v in t.items()})
        self._gen = (df, b, t)
        return df

    def make_intersects_kdtree(self, sticks):
        """length sort + junction search (netsim.py:131-150); returns (sticks, intersects)"""
        sticks['cluster'] = sticks.index
        gen = self._gen if (self._gen is not None and self._gen[0] is sticks) else None
        sticks.sort_values('length', inplace=True, ascending=False, kind='stable')
        sticks.reset_index(drop=True, inplace=True)
        eng = self._engine
        if gen is not None:
            b = gen[1]                                     # already on the GPU in this very order
        else:
            ends = np.stack(sticks.endarray.values)
            t = dict(xc=sticks.xc.values, yc=sticks.yc.values, length=sticks.length.values,
                     xa=ends[:, 0, 0], ya=ends[:, 0, 1], xb=ends[:, 1, 0], yb=ends[:, 1, 1],
                     kind=np.array([_KCODE[k] for k in sticks.kind.values], dtype=np.uint8),
                     orig=sticks.cluster.values)
            b = eng.batch_from_host([t])
        eng.find_junctions(b)
        self._batch = b
        self._gen = None
        return sticks, self._intersects_frame(b)

    def _intersects_frame(self, b):
        J = b.J
        k2 = b.kind2.cpu().numpy()[:J]
        names = np.array([pair_string(c) for c in range(9)], dtype=object)
        return pd.DataFrame({"stick1": b.stick1.cpu().numpy()[:J].astype(np.int64),
                             "stick2": b.stick2.cpu().numpy()[:J].astype(np.int64),
                             "x": b.jx.cpu().numpy()[:J], "y": b.jy.cpu().numpy()[:J], "kind": names[k2]},
                            columns=["stick1", 'stick2', 'x', 'y', 'kind'])

    def make_trivial_sticks(self):
        """the deterministic 6-stick test network of netsim.py:152-167"""
        rows = [[0.01, 0.5, np.pi / 2 - 1e-6, 1.002, 'm'], [0.3, 0.5, np.pi / 4, 1, 's'],
                [0.7, 0.5, -np.pi / 4, 1, 's'], [0.5, 0.5, -np.pi / 4, 0.1, 's'], [0.5, 0.5, np.pi / 4, 0.1, 's'],
                [.99, 0.5, np.pi / 2 - 1e-6, 1.001, 'm']]
        for r in rows:
            r.append(self.get_ends(r))
        sticks = pd.DataFrame(rows, columns=["xc", "yc", "angle", "length", 'kind', "endarray"])
        self.percolating = False
        self.sticks, self.intersects = self.make_intersects_kdtree(sticks)
        self.make_cnet()

    # ------------------------------------------------------------------ graph / clusters (netsim.py:171-206)
    @property
    def graph(self):
        """networkx graph of ALL junctions (netsim.py:175), built on first access"""
        if self._graph is None:
            import networkx as nx
            self._graph = nx.from_pandas_edgelist(self.intersects, source='stick1', target='stick2', edge_attr=True)
        return self._graph

    @graph.setter
    def graph(self, g):
        self._graph = g

    def make_graph(self):
        """component search + percolation test on the GPU.  Returns the conducting component
        (a lazily materialised networkx view) or (False, False, False) like the reference."""
        eng, b = self._engine, self._batch
        self._graph = None
        eng.label_clusters(b)
        eng.assemble_pattern(b)
        self.percolating = bool(b.h_stats[0][0])
        if self.percolating:
            self.ground_nodes = [1]
            self.voltage_sources = [[0, 0.1]]
            return True
        return False, False, False

    def populate_graph(self, onoffmap):
        for edge in self.graph.edges():
            self.graph.edges[edge]['component'] = self.element(self.graph.edges[edge]['kind'], onoffmap)

    def label_clusters(self):
        """sticks.cluster / clustersizes with the reference's conventions (netsim.py:197-206;
        SURVEY.md H5): sticks that have a junction get their component's ordinal (components in
        order of their smallest stick index), isolated sticks keep the pre-sort index."""
        b = self._batch
        if getattr(b, "label", None) is None:
            self._engine.label_clusters(b)
        lab = b.label.cpu().numpy()[:b.S]
        J = b.J
        has = np.zeros(b.S, dtype=bool)
        has[b.stick1.cpu().numpy()[:J]] = True
        has[b.stick2.cpu().numpy()[:J]] = True
        roots = np.unique(lab[has])
        ordinal = -np.ones(b.S, dtype=np.int64)
        ordinal[roots] = np.arange(len(roots))
        cl = self.sticks['cluster'].values.copy()
        cl[has] = ordinal[lab[has]]
        self.sticks['cluster'] = cl
        sizes = np.bincount(lab, minlength=b.S)
        self.clustersizes = sizes[roots]

    def make_cnet(self):
        try:
            self.make_graph()
            assert self.percolating, "The network is not conducting!"
            self.cnet = ConductionNetwork._from_batch(self._engine, self._batch, self.element, self.onoffmap, self)
            self.cnet.set_global_gate(0)
            self.cnet.update()
        except Exception:
            traceback.print_exc(file=sys.stdout)     # the reference prints and carries on (netsim.py:215-218)

    # ------------------------------------------------------------------ persistence (netsim.py:220-249)
    def timestamp(self):
        return datetime.now().strftime('%y-%m-%d_%H%M%S_%f')

    def make_fname(self):
        self.notes = "{}_{}sticks_{}x{}um_{}L_{}".format(self.seed, self.n, self.scaling, self.scaling, self.l,
                                                          self.notes)
        return os.path.join(self.directory, self.notes)

    def save_system(self, fname=False):
        if not fname:
            fname = self.fname
        self.sticks.to_csv(fname + '_sticks.csv')
        self.intersects.to_csv(fname + '_intersects.csv')

    def load_system(self, fname, network=True):
        """reads `<fname>_sticks.csv` / `<fname>_intersects.csv` (written by either implementation);
        end points are re-derived from xc, yc, angle, length and the saved junction table is used
        as is, exactly like netsim.py:238-249."""
        self.sticks = pd.read_csv(fname + '_sticks.csv', index_col=0, float_precision='round_trip')
        self.sticks.endarray = [self.get_ends(row) for row in self.sticks.values]
        self.intersects = pd.read_csv(fname + '_intersects.csv', index_col=0, float_precision='round_trip')
        st = self.sticks
        ends = np.stack(st.endarray.values)
        t = dict(xc=st.xc.values, yc=st.yc.values, length=st.length.values, xa=ends[:, 0, 0], ya=ends[:, 0, 1],
                 xb=ends[:, 1, 0], yb=ends[:, 1, 1], kind=np.array([_KCODE[k] for k in st.kind.values], dtype=np.uint8),
                 orig=st.cluster.values if 'cluster' in st else np.arange(len(st)))
        eng = self._engine
        b = eng.batch_from_host([t])
        it = self.intersects.sort_values(['stick1', 'stick2'], kind='stable')
        eng.attach_junctions(b, it.stick1.values, it.stick2.values, it.x.values, it.y.values,
                             np.array([pair_code(k) for k in it.kind.values], dtype=np.uint8))
        self._batch = b
        if network:
            self.make_cnet()


class RandomCNTNetwork(RandomConductingNetwork):
    """netsim.py:251-277"""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.gatetype = 'back'
        self.gatevoltage = 0

    def dummy(self):
        print("dummy ran")

    def global_gate(self, vg):
        self.cnet.set_global_gate(vg)
        self.cnet.update()
        return sum(self.cnet.source_currents)

    def local_gate(self, vg, area):
        self.cnet.set_local_gate(area, vg)
        self.cnet.update()
        return sum(self.cnet.source_currents)

    def gate_sweep(self, vgs, gates=('back', 'partial', 'total')):
        """All gate steps of a sweep (measure_perc.add_voltagemeas, measure_perc.py:39-53) as ONE
        batched solve: every (gate, vg) pair is a value slab over the shared CSR pattern.  Returns
        the source currents in the order `for g in gates: for vg in vgs` -- the same numbers
        `gate(vg, g)` gives one by one.  Afterwards the device is in the state of the last step."""
        from .engine import AREA_PARTIAL, AREA_TOTAL, GateState
        eng, b = self._engine, self._batch
        states, steps = [], []
        for g in gates:
            for vg in vgs:
                st = GateState(eng, b).set_global(0)
                if g == 'back':
                    st.set_global(vg)
                elif g == 'partial':
                    st.set_local(AREA_PARTIAL, vg)
                elif g == 'total':
                    st.set_local(AREA_TOTAL, vg)
                states.append(st)
                steps.append((g, vg))
        if not states:
            return []
        res = eng.solve_gates(b, states, self.onoffmap, self.element, want_fields=True)
        eng.sync()
        cur = res["isrc"][:, 0, 2].cpu().numpy()
        self.sweep_status = res["status"][:, 0].cpu().numpy()
        # leave the object as the last gate() call would
        self.gatetype, self.gatevoltage = steps[-1]
        c = self.cnet
        c._gate = states[-1]
        c.gate_areas = [] if steps[-1][0] == 'back' else [[AREA_PARTIAL if steps[-1][0] == 'partial' else AREA_TOTAL,
                                                             steps[-1][1]]]
        c._last = {k: (v[-1:] if hasattr(v, "shape") and v.shape[0] == len(states) else v) for k, v in res.items()
                   if k != "sysm"}
        c.source_currents = np.array([cur[-1]])
        c._graph_dirty = True
        return [float(v) for v in cur]

    def gate(self, vg, gate):
        self.gatetype = gate
        self.gatevoltage = vg
        self.cnet.gate_areas = []
        self.cnet.set_global_gate(0)
        if gate == 'back':
            self.global_gate(vg)
        elif gate == 'partial':
            self.local_gate(vg, [0.5, 0, 0.16, 0.667])
        elif gate == 'total':
            self.local_gate(vg, [0.217, 0.5, 0.167, 1.2])
        return sum(self.cnet.source_currents)
