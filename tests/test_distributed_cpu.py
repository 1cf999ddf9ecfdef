"""CPU, world_size 2, gloo: the N>1 host logic of the ensemble driver
(measure_perc.measure_async: contiguous shard per rank, all-gather of the per-network rows).
The GPU body (`measure_ensemble`) is replaced by a deterministic stub -- here computed with the
oracle on tiny networks -- because this test runs without a GPU; the sharded result must equal
the single-process result row for row."""
import os
import socket
import sys

import numpy as np
import pandas as pd
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_measure_ensemble(ns, scaling, seeds, onoffmap=1, **kw):
    """stand-in for the GPU body: oracle on the same (n, seed) inputs"""
    from networksim_cntfet_b200 import measure_perc
    from oracle import cntnet_oracle as O
    res = {k: [] for k in ("percolating", "nclust", "maxclust", "ion", "ioff_back", "ioff_total", "ioff_partial")}
    for n, seed in zip(ns, seeds):
        s = O.sticks_with_ends(O.canonical_sort(O.make_sticks_mt19937(int(n), "exp", 0.135, scaling, int(seed))))
        m = O.measure_fullnet(s, onoffmap=onoffmap)
        res["percolating"].append(m["percolating"])
        res["nclust"].append(m["nclust_ref"])
        res["maxclust"].append(m["maxclust_ref"])
        for k in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
            res[k].append(m[k])
    res = {k: np.asarray(v) for k, v in res.items()}
    return measure_perc._rows_from_ensemble(res, [int(n) for n in ns], scaling, np.asarray(seeds, dtype=np.uint64), onoffmap)


def _run_async(tmpdir):
    from networksim_cntfet_b200 import measure_perc
    measure_perc.measure_ensemble = _stub_measure_ensemble
    os.chdir(tmpdir)
    seeds = list(range(40, 51))
    return measure_perc.measure_async(2, 150, 25, 11, 5, onoffmap=[0, 1], seeds=seeds)


def _worker(rank, world, port, tmpdir, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from networksim_cntfet_b200 import measure_perc
        lo, hi, r, w = measure_perc._shard(11)
        df = _run_async(os.path.join(tmpdir, "r%d" % rank))
        # default seeds: drawn on rank 0 and broadcast, so every rank works from the same array and the one
        # seeds file rank 0 writes is the record of the whole run
        rnd = measure_perc.measure_async(2, 150, 25, 6, 5, onoffmap=[0])
        q.put((rank, (lo, hi, r, w), df.reset_index(drop=True).to_dict("list"), [int(v) for v in rnd.seed.tolist()]))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(300)
def test_measure_async_world2_gloo(tmp_path):
    for r in range(2):
        os.makedirs(tmp_path / ("r%d" % r))
    os.makedirs(tmp_path / "single")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    got.sort()
    assert got[0][1] == (0, 5, 0, 2) and got[1][1] == (5, 11, 1, 2)        # contiguous, balanced shards
    frames = [pd.DataFrame(g[2]) for g in got]
    assert got[0][3] == got[1][3] and len(set(got[0][3])) > 1              # both ranks used rank 0's random seeds
    files = sorted(n for n in os.listdir(tmp_path / "r0") if n.startswith("seeds_"))
    recorded = [np.loadtxt(tmp_path / "r0" / n, delimiter=",").astype(np.uint64).tolist() for n in files]
    assert sorted(set(got[0][3])) in [sorted(set(r)) for r in recorded]
    cwd = os.getcwd()
    try:
        single = _run_async(str(tmp_path / "single")).reset_index(drop=True)
    finally:
        os.chdir(cwd)
    for f in frames:                                                       # every rank holds the full result
        f = f.reset_index(drop=True)
        assert len(f) == len(single)
        assert f.sticks.tolist() == single.sticks.tolist() and f.gate.tolist() == single.gate.tolist()
        assert f.onoffmap.tolist() == single.onoffmap.tolist() and f.seed.tolist() == single.seed.tolist()
        assert np.allclose(f.ion.values.astype(float), single.ion.values.astype(float), rtol=1e-14, atol=0)
        assert np.allclose(f.ioff.values.astype(float), single.ioff.values.astype(float), rtol=1e-14, atol=0)
    assert set(single.sticks) == {150 + 25 * i for i in range(11)}
    # only rank 0 writes the seeds file (measure_perc.py:226-229)
    assert any(n.startswith("seeds_") for n in os.listdir(tmp_path / "r0"))
    assert not any(n.startswith("seeds_") for n in os.listdir(tmp_path / "r1"))


def test_ensemble_call_is_cut_into_batches_for_two_engines(monkeypatch):
    """host logic of measure_perc.measure_ensemble: batches bounded by BATCH_STICKS (electrodes counted), a large call is
    cut into at least two batches so that two engines can alternate, a small call stays one batch on one engine, every
    network lands in exactly one batch and the order is kept"""
    from networksim_cntfet_b200 import measure_perc as mp

    def covered(batches, n):
        idx = [i for sl in batches for i in range(sl.start, sl.stop)]
        return idx == list(range(n))

    # the bench's end-to-end call: 320 networks of 100k sticks -> two batches of 160
    b, pipe = mp._split_batches([100000] * 320)
    assert pipe and [(s.start, s.stop) for s in b] == [(0, 160), (160, 320)]
    # a long ensemble: batches of <= BATCH_STICKS sticks alternate between the engines
    b, pipe = mp._split_batches([100000] * 1000)
    assert pipe and covered(b, 1000) and all(sum(100002 for _ in range(s.start, s.stop)) <= mp.BATCH_STICKS for s in b)
    assert len(b) == 7 and b[0].stop == 160
    # small calls: one batch, one engine
    b, pipe = mp._split_batches([300] * 5)
    assert not pipe and [(s.start, s.stop) for s in b] == [(0, 5)]
    # one network, however large, cannot be split
    b, pipe = mp._split_batches([10_000_000])
    assert not pipe and len(b) == 1
    # a network larger than the cap is a batch of its own; ragged sizes keep their order
    ns = [4000, 20_000_000, 5000, 6000]
    b, pipe = mp._split_batches(ns)
    assert pipe and covered(b, 4) and any(s.stop - s.start == 1 and ns[s.start] == 20_000_000 for s in b)
    # switched off: the stick bound alone decides
    monkeypatch.setattr(mp, "PIPELINE", False)
    b, pipe = mp._split_batches([100000] * 320)
    assert not pipe and [(s.start, s.stop) for s in b] == [(0, 160), (160, 320)]
    monkeypatch.setattr(mp, "BATCH_STICKS", 9000)
    b, pipe = mp._split_batches([3000, 3600, 4000, 4400])
    assert [(s.start, s.stop) for s in b] == [(0, 2), (2, 4)] and not pipe
    assert mp._split_batches([]) == ([], False)


def _rows_loop(res, ns, scaling, seeds, onoffmap, fnames=None):
    """the row loop of measure_perc.py:190-197 (what _rows_from_ensemble builds as arrays)"""
    import pandas as pd
    from networksim_cntfet_b200.measure_perc import FULLNET_COLUMNS
    rows = []
    for k, n in enumerate(ns):
        fname = fnames[k] if fnames else ''
        base = [n, scaling, n / scaling ** 2, int(res["nclust"][k]), int(res["maxclust"][k])]
        if res["percolating"][k]:
            ion = float(res["ion"][k])
            rows.append(base + [ion, float(res["ioff_back"][k]), 'back', fname, int(seeds[k]), onoffmap])
            rows.append(base + [ion, float(res["ioff_total"][k]), 'total', fname, int(seeds[k]), onoffmap])
            rows.append(base + [ion, float(res["ioff_partial"][k]), 'partial', fname, int(seeds[k]), onoffmap])
        else:
            rows.append(base + [0, 0, 'back', fname, int(seeds[k]), onoffmap])
    return pd.DataFrame(rows, columns=FULLNET_COLUMNS)


def test_rows_of_an_ensemble_equal_the_reference_loop():
    """_rows_from_ensemble (columns as arrays) against the reference's row loop: values, order, dtypes, CSV text"""
    import numpy as np
    from networksim_cntfet_b200 import measure_perc as mp
    rng = np.random.default_rng(5)

    def fake(B, perc):
        return dict(percolating=np.asarray(perc, dtype=bool), ion=rng.random(B) * 1e-4, ioff_back=rng.random(B) * 1e-6,
                    ioff_total=rng.random(B) * 1e-7, ioff_partial=rng.random(B) * 1e-8,
                    nclust=rng.integers(1, 500, B).astype(np.int32), maxclust=rng.integers(1, 90000, B).astype(np.int32),
                    status=np.zeros((4, B), dtype=np.int32))

    cases = [
        ([300, 350, 4000, 100000, 7], 5, [True, False, True, True, False], np.arange(11, 16, dtype=np.uint64), 1, None),
        ([300, 301], 5.0, [False, False], [3, 4], 0, None),                                   # nothing percolates: integer zeros
        ([1000, 2000, 3000], 20, [True, True, True], np.array([2 ** 63 + 5, 1, 2], dtype=np.uint64), 2, ['a', 'b', 'c']),
        ([], 5, [], [], 1, None),
    ]
    for ns, scaling, perc, seeds, onoff, fnames in cases:
        res = fake(len(ns), perc)
        a = _rows_loop(res, ns, scaling, seeds, onoff, fnames)
        b = mp._rows_from_ensemble(res, ns, scaling, seeds, onoff, fnames=fnames)
        assert list(a.columns) == list(b.columns) and len(a) == len(b)
        if len(a):
            assert a.dtypes.astype(str).tolist() == b.dtypes.astype(str).tolist(), (a.dtypes, b.dtypes)
        for col in a.columns:
            assert a[col].tolist() == b[col].tolist(), col
        assert a.to_csv() == b.to_csv()


def _gather_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench
        rg = bench.RowGather(dist, world, torch.device("cpu"))
        rows = lambda step: np.full((3, 7), 100.0 * step + rank) + np.arange(7)
        # the blocking form (end-to-end loop): every rank gets the rows of all ranks in rank order
        full = rg.gather(rows(0))
        # the non-blocking form (device-timed loop): issue one per step, nobody waits; finish() completes them all
        issued = [rg.gather(rows(k), wait=False) for k in range(1, 5)]
        n_pending = len(rg.pending)
        last = rg.finish()
        q.put((rank, full.tolist(), issued, n_pending, last.tolist(), len(rg.pending), rg.finish()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_bench_row_gather_world2_gloo():
    """bench.RowGather, the only collective of the ensemble path: blocking and non-blocking all-gather of the statistics
    rows on two gloo ranks; one rank gathers nothing"""
    import bench
    one = bench.RowGather(None, 1, None)
    r = np.arange(6.0).reshape(2, 3)
    assert one.gather(r) is r and one.gather(r, wait=False) is r and one.finish() is None
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r_, 2, port, q)) for r_ in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    expect = lambda step: np.concatenate([np.full((3, 7), 100.0 * step + rk) + np.arange(7) for rk in range(2)]).tolist()
    for rank, full, issued, n_pending, last, left, again in got:
        assert full == expect(0)
        assert issued == [None] * 4 and n_pending == 4
        assert last == expect(4) and left == 0 and again is None
