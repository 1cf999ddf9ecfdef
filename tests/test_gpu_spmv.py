"""GPU: the streaming CSR SpMV (cnt_spmv_csr, csrc/spmv_stream.cuh) against scipy -- short and empty rows, a row
longer than the staging buffer, entry counts that are not multiples of the vector width, unaligned value
slices (scalar path), columns beyond n_rows (halo copies behind the own rows), and bit-reproducibility."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp
import torch

pytestmark = pytest.mark.gpu


def _spmv(engine, rowptr, col, val, diag, x, n_rows):
    from networksim_cntfet_b200 import _lib
    lib = _lib.load()
    NS = val.shape[0]
    y = torch.full((NS, n_rows), float("nan"), dtype=torch.float64, device=engine.device)
    vp = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    with engine._work():
        _lib.check(lib.cnt_spmv_csr(NS, n_rows, col.numel(), vp(rowptr), vp(col), vp(val), val.stride(0), vp(diag),
                                    diag.stride(0) if diag is not None else 0, vp(x), x.stride(0), vp(y), y.stride(0),
                                    engine.stream_ptr))
    engine.sync()
    return y


def _random_csr(rng, n, ncols, mean_len, long_row=None):
    lens = rng.poisson(mean_len, n)
    lens[rng.integers(0, n, n // 10)] = 0                      # empty rows
    if long_row is not None:
        lens[long_row[0]] = long_row[1]
    rowptr = np.zeros(n + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(lens)
    nnz = int(rowptr[-1])
    col = rng.integers(0, ncols, nnz)
    return rowptr.astype(np.int32), col.astype(np.int32), nnz


@pytest.mark.parametrize("n,mean_len,long_row,halo,pad", [(1000, 4.7, None, 0, 0), (70001, 4.7, (300, 5000), 0, 0),
                                                           (5000, 2.0, None, 37, 1), (257, 9.0, (256, 2500), 5, 3),
                                                           (3, 1.0, None, 0, 0)])
def test_spmv_matches_scipy(engine, n, mean_len, long_row, halo, pad):
    rng = np.random.default_rng(n)
    NS = 3
    rowptr, col, nnz = _random_csr(rng, n, n + halo, mean_len, long_row)
    stride = nnz + pad                                          # odd strides: slices of q > 0 are not 16-byte aligned
    val = rng.standard_normal((NS, stride))
    diag = rng.uniform(1, 2, (NS, n))
    x = rng.standard_normal((NS, n + halo))
    dev = engine.device
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)
    d_rowptr, d_col = t(rowptr, torch.int32), t(col if nnz else np.zeros(1), torch.int32)
    d_val, d_diag, d_x = t(val if stride else np.zeros((NS, 1)), torch.float64), t(diag, torch.float64), t(x, torch.float64)
    y = _spmv(engine, d_rowptr, d_col[:nnz], d_val, d_diag, d_x, n).cpu().numpy()
    y2 = _spmv(engine, d_rowptr, d_col[:nnz], d_val, d_diag, d_x, n).cpu().numpy()
    assert np.array_equal(y, y2)                                # fixed summation order
    y0 = _spmv(engine, d_rowptr, d_col[:nnz], d_val, None, d_x, n).cpu().numpy()
    for q in range(NS):
        # copies: scipy merges duplicate entries IN PLACE when asked for |A|
        A = sp.csr_matrix((val[q, :nnz].copy(), col.copy(), rowptr.copy()), shape=(n, n + halo))
        want = A @ x[q]
        absA = sp.csr_matrix((np.abs(val[q, :nnz]), col.copy(), rowptr.copy()), shape=(n, n + halo))
        scale = absA @ np.abs(x[q]) + np.abs(diag[q] * x[q, :n]) + 1e-300
        assert np.all(np.abs(y0[q] - want) <= 1e-14 * scale)
        assert np.all(np.abs(y[q] - (want + diag[q] * x[q, :n])) <= 1e-14 * scale)
