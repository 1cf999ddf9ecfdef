#!/usr/bin/env python3
"""Drop-in for the reference's `measure_perc.py` (ensemble driver).

Same functions, arguments and DataFrame / CSV columns as measure_perc.py:39-276.  The
reference fans independent networks out over `multiprocessing.Pool(cores)`
(measure_perc.py:230-235); here the whole ensemble is pushed through the GPU in batches
(`measure_ensemble`), sharded over ranks when torch.distributed is initialised, with one final
all-gather of the per-network rows.  `cores` is accepted and ignored.
"""
import argparse
import os
import sys
import textwrap  # noqa: F401  (kept: the reference module exposes it)
import traceback
import uuid as id
from timeit import default_timer as timer

import numpy as np
import pandas as pd

from . import get_engine, netsim, set_engine  # noqa: F401
from .cnet import FermiDiracTransistor, LinExpTransistor

elements = [FermiDiracTransistor, LinExpTransistor]
VDS = 0.1                                   # netsim.py:182
FULLNET_COLUMNS = ['sticks', 'size', 'density', 'nclust', 'maxclust', 'ion', 'ioff', 'gate', 'fname', 'seed',
                   'onoffmap']
_PINNED = None
PIPELINE = True               # the batches of a call alternate between two engines (two streams, two host threads): the host
                              # work and the latency-bound late rounds of one batch overlap the kernels of the other.
                              # Measured at 100x100 um, device / end to end networks/s: one engine x 224 networks per batch
                              # 3250 / 3140, two engines x 112: 3370 / 3250, x 160: 3450 / 3300, x 224: 3320 / 3200; results
                              # identical (networks are independent and seeded one by one)
PIPELINE_MIN_STICKS = 2_000_000   # ... for calls of at least this many sticks (below: one batch, one engine)
BATCH_STICKS = 16_100_000     # sticks per GPU batch of the ensemble path (160 networks at 100x100 um, ~6 GB)


def _split_batches(ns):
    """(slices of consecutive networks, use two engines?): batches of at most BATCH_STICKS sticks (electrodes included; a
    network larger than that is a batch of its own); a call of at least PIPELINE_MIN_STICKS sticks and two networks is cut
    into at least two batches, which alternate between two engines (Engine.run_pipelined)"""
    total = sum(ns) + 2 * len(ns)
    pipeline = bool(PIPELINE) and len(ns) > 1 and total >= PIPELINE_MIN_STICKS
    cap = BATCH_STICKS
    if pipeline:
        cap = min(cap, max((total + 1) // 2, 1))
    batches, start = [], 0
    while start < len(ns):
        stop, tot = start, 0
        while stop < len(ns) and (stop == start or tot + ns[stop] + 2 <= cap):
            tot += ns[stop] + 2
            stop += 1
        batches.append(slice(start, stop))
        start = stop
    return batches, pipeline and len(batches) > 1


def checkdir(directoryname):
    """create the directory if it does not exist (measure_perc.py:28-38)"""
    if directoryname and not os.path.isdir(directoryname):
        os.makedirs(directoryname, exist_ok=True)


def add_voltagemeas(device, data, vgrange=10, vgnum=3):
    """three gate types x linspace(-vgrange, vgrange, vgnum) solves (measure_perc.py:39-53)"""
    gate, gatevoltage, current = [], [], []
    vgs = np.linspace(-vgrange, vgrange, vgnum)
    if hasattr(device, "gate_sweep"):
        # all 3*vgnum gate steps as one batched GPU solve (same numbers as gate() one by one)
        current = device.gate_sweep(vgs, ['back', 'partial', 'total'])
    for g in ['back', 'partial', 'total']:
        for vg in vgs:
            if not hasattr(device, "gate_sweep"):
                current.append(device.gate(vg, g))
            gate.append(g)
            gatevoltage.append(vg)
    data.gate = gate
    data.gatevoltage = gatevoltage
    data.current = current
    return data


def _true_maxclust(device):
    """measure_perc.py:78: len(max(nx.connected_components(g))) == size of the FIRST component,
    i.e. the one holding the smallest stick index that has a junction."""
    return int(device._batch.h_stats[0][4])


def single_measure(n, scaling, l='exp', dump=False, savedir='test', seed=0, onoffmap=0, v=False,
                   element=LinExpTransistor, vgrange=10, vgnum=3, rng='philox'):
    """one device + gate sweeps (measure_perc.py:54-138); returns (DataFrame, fname)"""
    datacol = ['sticks', 'scaling', 'density', 'current', 'gatevoltage', 'gate', 'nclust', 'maxclust', 'fname',
               'onoffmap', 'seed', 'runtime', 'element']
    checkdir(savedir)
    start = timer()
    d = n / scaling ** 2
    if not seed:
        seed = np.random.randint(low=0, high=2 ** 32)
    fname = os.path.join(savedir, "mnet{:2.2f}_s{}_l{}_om{}_el{}_seed{:010d}".format(
        d, scaling, l, onoffmap, elements.index(element), seed))
    data = pd.DataFrame(columns=datacol)
    if v:
        print("=== measurement start ===\nn{:05d}_d{:2.1f}_seed{:010d}".format(n, d, seed))
    device = netsim.RandomCNTNetwork(n=n, scaling=scaling, notes='run', l=l, seed=seed, onoffmap=onoffmap,
                                     element=element, rng=rng)
    if v:
        print("=== physical device made t = {:0.2}".format(timer() - start))
        print("percolating : {}".format(device.percolating))
    device.label_clusters()
    nclust = len(device.sticks.cluster.drop_duplicates())
    maxclust = _true_maxclust(device)
    if dump:
        try:
            device.save_system(fname)
        except Exception as e:
            if v:
                print("measurement failed: error saving data")
                print("ERROR for {} sticks:\n".format(n), e)
                traceback.print_exc(file=sys.stdout)
    if device.percolating:
        data = add_voltagemeas(device, data, vgrange=vgrange, vgnum=vgnum)
    else:
        data.current = [0]
    data.sticks = n
    data.scaling = scaling
    data.density = d
    data.nclust = nclust
    data.maxclust = maxclust
    data.seed = seed
    data.element = element
    data.onoffmap = onoffmap
    data.fname = fname
    data['runtime'] = timer() - start
    data.to_csv(fname + "_data.csv")
    if v:
        print("=== measurement done ===")
    return data, fname


def _check_status(res):
    """the reference's direct solve either succeeds or raises; an iterative solve that stopped at maxit (status 1)
    or broke down (status 2) must not end up in a CSV unnoticed"""
    st = np.asarray(res.get("status", 0))
    if np.any(st != 0):
        import warnings
        bad = np.argwhere(st != 0)
        warnings.warn("measure_perc: %d of %d solves did not converge (first: configuration %d, network %d, status %d, "
                      "relres %.3g); their currents are unreliable -- re-run with Engine.solver = 'direct' or 'flat'"
                      % (len(bad), st.size, bad[0][0], bad[0][1], int(st[tuple(bad[0])]),
                         float(np.asarray(res["relres"])[tuple(bad[0])])), RuntimeWarning, stacklevel=3)


def _rows_from_ensemble(res, ns, scaling, seeds, onoffmap, fnames=None, sheet=False):
    """per-network statistics -> the reference's measure_fullnet rows (measure_perc.py:190-197); sheet=True appends
    the sheet conductances G_sq = I / Vds of the on state and of the row's gate (not a column of the reference)"""
    _check_status(res)
    # columns built as arrays (a Python loop over the networks was 1.5 ms per 320 networks of every ensemble call); dtypes as
    # pandas infers them from the reference's list of rows: a percolating network gives three rows (back, total, partial),
    # any other one row with integer zeros
    nn = np.asarray([int(n) for n in ns], dtype=np.int64)
    perc = np.asarray(res["percolating"], dtype=bool)[:len(nn)]
    rep = np.where(perc, 3, 1)
    idx = np.repeat(np.arange(len(nn)), rep)
    pos = np.arange(len(idx)) - np.repeat(np.cumsum(rep) - rep, rep)          # 0, 1, 2 inside a network
    cols = {}
    cols['sticks'] = nn[idx]
    cols['size'] = np.full(len(idx), scaling)
    cols['density'] = (nn / scaling ** 2)[idx]
    cols['nclust'] = np.asarray(res["nclust"]).astype(np.int64)[idx]
    cols['maxclust'] = np.asarray(res["maxclust"]).astype(np.int64)[idx]
    if perc.any():
        ioff3 = np.stack([np.asarray(res["ioff_back"], dtype=np.float64), np.asarray(res["ioff_total"], dtype=np.float64),
                          np.asarray(res["ioff_partial"], dtype=np.float64)], 1)
        cols['ion'] = np.where(perc, np.asarray(res["ion"], dtype=np.float64), 0.0)[idx]
        cols['ioff'] = np.where(perc[idx], ioff3[idx, pos], 0.0)
    else:
        cols['ion'] = np.zeros(len(idx), dtype=np.int64)
        cols['ioff'] = np.zeros(len(idx), dtype=np.int64)
    cols['gate'] = np.array(['back', 'total', 'partial'], dtype=object)[pos]
    cols['fname'] = np.array([fnames[k] if fnames else '' for k in range(len(nn))], dtype=object)[idx] if len(nn) else np.array([], dtype=object)
    sd = np.asarray(seeds)[:len(nn)]
    if sd.dtype.kind == 'u' and (len(sd) == 0 or int(sd.max()) <= np.iinfo(np.int64).max):
        sd = sd.astype(np.int64)
    cols['seed'] = sd[idx]
    cols['onoffmap'] = np.full(len(idx), onoffmap)
    df = pd.DataFrame(cols, columns=FULLNET_COLUMNS)
    if sheet:
        df["gsheet_on"] = df["ion"] / VDS
        df["gsheet_off"] = df["ioff"] / VDS
    return df


def measure_ensemble(ns, scaling, seeds, onoffmap=1, element=LinExpTransistor, l='exp', rtol=1e-15,
                     return_io=False, engine=None, sheet=False):
    """GPU body shared by measure_fullnet / measure_async: host arrays (ns, seeds) in, the
    reference's rows out.  ns/seeds are staged in pinned host memory and copied to the device;
    results come back as one statistics row per network."""
    import torch
    eng = engine or get_engine()
    ns = [int(n) for n in ns]
    seeds_np = np.asarray(seeds, dtype=np.uint64)
    # pinned staging buffer, kept between calls (cudaHostAlloc per call costs up to tens of ms); the previous call's
    # copy has completed: every call ends with its results on the host
    global _PINNED
    if _PINNED is None or _PINNED.numel() < 2 * len(ns):
        _PINNED = torch.empty(max(2 * len(ns), 4096), dtype=torch.int64).pin_memory()
    pinned = _PINNED[:2 * len(ns)]
    pinned[:len(ns)] = torch.as_tensor(ns, dtype=torch.int64)
    pinned[len(ns):] = torch.from_numpy(seeds_np.view(np.int64).copy())
    dev_in = pinned.to(eng.device, non_blocking=True)          # H2D of this call's inputs
    batches, pipeline = _split_batches(ns)

    def one(e, sl):
        res = e.measure_ensemble(ns[sl], scaling, seeds_np[sl], onoffmap=onoffmap, element=element, l=l, rtol=rtol,
                                 seeds_dev=dev_in[len(ns) + sl.start:len(ns) + sl.stop])
        nbytes = sum(int(np.asarray(v).nbytes) for k, v in res.items() if isinstance(v, np.ndarray))
        # the rows of a batch are built by the thread that ran it, while the other engine is still computing
        return nbytes, _rows_from_ensemble(res, ns[sl], scaling, seeds_np[sl], onoffmap, sheet=sheet)

    # two engines on two host threads: the host work and the late rounds of one batch overlap the kernels of the other
    # (Engine.run_pipelined)
    results = eng.run_pipelined(batches, one) if pipeline else [one(eng, sl) for sl in batches]
    d2h = sum(r[0] for r in results)
    frames = [r[1] for r in results]
    del dev_in
    df = pd.concat(frames, ignore_index=True) if frames else pd.DataFrame(columns=FULLNET_COLUMNS)
    if return_io:
        return df, {"h2d_bytes": int(pinned.numel() * 8), "d2h_bytes": int(d2h)}
    return df


def measure_fullnet(n, scaling, l='exp', save=False, seed=0, onoffmap=1, v=False, remote=False, rng='philox'):
    """one network: nclust / maxclust / ion / ioff for back, total and partial gates at 10 V
    (measure_perc.py:140-203).  `save` (or rng='mt19937') goes through the object API so the
    stick / junction tables exist on the host; otherwise only statistics leave the GPU."""
    start = timer()
    if not seed:
        seed = np.random.randint(low=0, high=2 ** 32)
    notes = "{}_{}sticks_{}x{}um_{}L_{}".format(seed, n, scaling, scaling, l, 'run')
    fname = os.path.join('data', notes)                        # netsim.py:223-226 with notes='run'
    if save or rng != 'philox':
        collection = netsim.RandomConductingNetwork(n, scaling=scaling, notes='run', l=l, seed=seed,
                                                    onoffmap=onoffmap, rng=rng)
        collection.label_clusters()
        nclust = len(collection.sticks.cluster.drop_duplicates())
        maxclust = _true_maxclust(collection)
        if save:
            try:
                checkdir(os.path.dirname(collection.fname))
                collection.save_system()
            except Exception as e:
                if v:
                    print("measurement failed: error saving data")
                    print("ERROR for {} sticks:\n".format(n), e)
                    traceback.print_exc(file=sys.stdout)
        base = [n, scaling, n / scaling ** 2, nclust, maxclust]
        rows = []
        if collection.percolating:
            c = collection.cnet
            ion = sum(c.source_currents)
            c.set_global_gate(10)
            c.update()
            rows.append(base + [ion, sum(c.source_currents), 'back', fname, seed, onoffmap])
            c.set_global_gate(0)
            c.set_local_gate([0.217, 0.5, 0.167, 1.2], 10)
            c.update()
            rows.append(base + [ion, sum(c.source_currents), 'total', fname, seed, onoffmap])
            c.set_global_gate(0)
            c.set_local_gate([0.5, 0, 0.16, 0.667], 10)
            c.update()
            rows.append(base + [ion, sum(c.source_currents), 'partial', fname, seed, onoffmap])
        else:
            rows.append(base + [0, 0, 'back', fname, seed, onoffmap])
        data = pd.DataFrame(rows, columns=FULLNET_COLUMNS)
    else:
        data = measure_ensemble([n], scaling, [seed], onoffmap=onoffmap, l=l)
        data['fname'] = fname
    data['runtime'] = timer() - start
    if fname and os.path.isdir(os.path.dirname(fname)):
        data.to_csv(fname + "_data.csv")
    return data


def measure_fullnet_distributed(n, scaling, seed, onoffmap=1, element=LinExpTransistor, l='exp', engine=None,
                                mode="gates"):
    """ONE (large) network on all ranks of an initialised torch.distributed job: every rank
    regenerates the network from `seed` (counter-based RNG: identical sticks, junctions, labels and
    CSR pattern everywhere, no exchange).  mode="gates": rank r solves the gate configurations q with
    q % world == r and one small all-reduce sums the per-configuration currents (at most 4 ranks do
    useful work).  mode="rows": the unknown rows of the system -- Z-ordered, i.e. compact spatial
    patches of the device -- are sharded over ALL ranks and the four configurations are solved
    together with per-iteration halo exchange of the search direction and two small all-reduces over
    NCCL (Engine.solve_sharded, slab.py).  mode="slab": the junction search -- the FP64 tests -- is sharded by
    spatial slab over all ranks with one exchange of the junction rows over NCCL, the rest runs replicated with
    the default (direct) solver.  Returns measure_fullnet's rows on every rank.  (The
    reference has no way to spread one network over several workers.)"""
    import torch
    import torch.distributed as dist
    eng = engine or get_engine()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    b = eng.generate([int(n)], scaling, seeds=[int(seed)], l=l)
    if mode == "rows":
        from . import slab
        comm = slab.DistComm() if dist.is_initialized() else slab.LocalComm(1)
        res = eng.measure_fullnet(b, onoffmap=onoffmap, element=element, row_shard=comm)
        df = _rows_from_ensemble(res, [int(n)], scaling, np.asarray([seed], dtype=np.uint64), onoffmap)
        return df, res
    if mode == "slab":
        # front end sharded by spatial slab: every rank tests the home sticks of its slab of the device against
        # their neighbourhood, the junction rows are exchanged over NCCL (Engine._find_junctions_slab); labels,
        # assembly and the direct solve (0.06 s for 10M sticks) run replicated on the identical table
        from . import slab
        comm = slab.DistComm() if dist.is_initialized() else slab.LocalComm(1)
        eng.prepare(b, comm)
        res = eng.measure_fullnet(b, onoffmap=onoffmap, element=element)
        df = _rows_from_ensemble(res, [int(n)], scaling, np.asarray([seed], dtype=np.uint64), onoffmap)
        return df, res
    if mode != "gates":
        raise ValueError("mode must be 'gates', 'rows' or 'slab'")
    res = eng.measure_fullnet(b, onoffmap=onoffmap, element=element, slab_shard=(rank, world))
    cur = torch.as_tensor(np.stack([res["ion"], res["ioff_back"], res["ioff_total"], res["ioff_partial"]]),
                          dtype=torch.float64, device=eng.device)
    its = torch.as_tensor(res["iters"].astype(np.float64), device=eng.device)
    if world > 1:
        dist.all_reduce(cur)
        dist.all_reduce(its)
    cur = cur.cpu().numpy()
    res.update(ion=cur[0], ioff_back=cur[1], ioff_total=cur[2], ioff_partial=cur[3], iters=its.cpu().numpy().astype(np.int32))
    df = _rows_from_ensemble(res, [int(n)], scaling, np.asarray([seed], dtype=np.uint64), onoffmap)
    return df, res


def _shard(number):
    """contiguous shard of range(number) for this rank when torch.distributed is initialised"""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            r, w = dist.get_rank(), dist.get_world_size()
            lo, hi = (number * r) // w, (number * (r + 1)) // w
            return lo, hi, r, w
    except Exception:
        pass
    return 0, number, 0, 1


def measure_async(cores, start, step, number, scaling, save=False, onoffmap=[1], seeds=[]):
    """density sweep n_i = start + i*step, i < number (measure_perc.py:205-242).  Returns the
    concatenated rows with columns
    ['sticks','size','density','nclust','maxclust','ion','ioff','gate','fname','seed','onoffmap'].
    `cores` is ignored: networks are batched on the GPU; under torch.distributed each rank takes
    a contiguous shard and the rows are all-gathered (every rank returns the full frame)."""
    if not os.path.isdir("data"):
        os.makedirs("data", exist_ok=True)
    uuid = id.uuid4()
    starttime = timer()
    nrange = [int(start + i * step) for i in range(number)]
    if not (len(seeds) == number):
        seeds = np.random.randint(low=0, high=2 ** 32, size=number)
    lo, hi, rank, world = _shard(number)
    if world > 1:
        # the seeds file is the reproducibility record (measure_perc.py:226-229): every rank must work from rank 0's
        # seeds and run id, not from its own random draw
        import torch.distributed as dist
        shared = [np.asarray(seeds).tolist(), str(uuid)]
        dist.broadcast_object_list(shared, src=0)
        seeds, uuid = np.asarray(shared[0], dtype=np.uint64), shared[1]
    if rank == 0 and not os.path.isfile("seeds_{}.csv".format(uuid)):
        np.savetxt("seeds_{}.csv".format(uuid), seeds, delimiter=",")
    output = []
    for omap in onoffmap:
        if save:
            part = [measure_fullnet(nrange[i], scaling, 'exp', save, seeds[i], omap) for i in range(lo, hi)]
            part = pd.concat(part) if part else pd.DataFrame(columns=FULLNET_COLUMNS)
        else:
            part = measure_ensemble(nrange[lo:hi], scaling, list(seeds[lo:hi]), onoffmap=omap)
            part['fname'] = [os.path.join('data', "{}_{}sticks_{}x{}um_{}L_{}".format(s, n, scaling, scaling, 'exp', 'run'))
                             for s, n in zip(part.seed, part.sticks)]
        if world > 1:
            import torch.distributed as dist
            gathered = [None] * world
            dist.all_gather_object(gathered, part)
            part = pd.concat(gathered)
        output.append(part)
    print(sum(len(o) for o in output))
    runtime = timer() - starttime
    print('finished with a runtime of {:.0f} '.format(runtime))
    data = pd.concat(output)
    if save and rank == 0:
        data.to_csv('measurement_batch_{}.csv'.format(uuid))
    return data


if __name__ == '__main__':
    parser = argparse.ArgumentParser(formatter_class=argparse.RawDescriptionHelpFormatter, description=__doc__)
    parser.add_argument("function", type=str, choices=["multicore", "singlecore"])
    parser.add_argument("-d", '--directory', type=str, default='')
    parser.add_argument("-t", '--test', action="store_true", default=False)
    parser.add_argument('-s', '--save', action="store_true", default=False)
    parser.add_argument('-v', '--verbose', action="store_true", default=False)
    parser.add_argument("--cores", type=int, default=1)
    parser.add_argument("--start", type=int)
    parser.add_argument("--step", type=int, default=0)
    parser.add_argument('-n', "--number", type=int, default=500)
    parser.add_argument("--scaling", type=int, default=5)
    parser.add_argument("--seed", type=int, default=0)
    parser.add_argument("--onoffmap", type=int, default=0)
    parser.add_argument("--element", type=int, default=0)
    parser.add_argument("--vgrange", type=int, default=10)
    parser.add_argument("--vgnum", type=int, default=3)
    args = parser.parse_args()
    if args.function == "multicore":
        if args.test:
            measure_async(2, 500, 0, 10, 5, save=True)
        else:
            # the reference passes the bare int and crashes iterating it (measure_perc.py:271 vs :232)
            measure_async(args.cores, args.start, args.step, args.number, args.scaling, args.save, [args.onoffmap])
    elif args.function == "singlecore":
        if args.test:
            single_measure(500, 5, v=True)
        else:
            single_measure(args.number, args.scaling, savedir=args.directory, dump=args.save, v=args.verbose,
                           element=elements[args.element], onoffmap=args.onoffmap, seed=args.seed,
                           vgrange=args.vgrange, vgnum=args.vgnum)
