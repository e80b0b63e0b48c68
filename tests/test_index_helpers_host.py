"""
CPU: the index helpers the grid and particle kernels inline (ngp_bin, wrap_once, cell_of in csrc/snapshot_kernels.cu; cut_coord,
wrap_idx in csrc/grid_common.cuh), compiled for the HOST (bfg_test_index_helpers_host), against numpy and the oracle port:
NGP cell assignments and cutout index sets bit-exact, cutout coordinates bit-identical to np.linspace.
tools/sass_fingerprint.py shows no kernel changed when these functions became host-compilable.
"""
import numpy as np

from oracle import runners_port as rp


def helper(what, x, L, N, n=None):
    from baryonforge_b200 import _lib
    x = np.ascontiguousarray(np.atleast_1d(x), dtype=np.float64)
    n = x.size if n is None else int(n)
    out_i = np.zeros(max(n, 1), dtype=np.int64)
    out_d = np.zeros(max(n, 1), dtype=np.float64)
    _lib.check(_lib.lib().bfg_test_index_helpers_host(int(what), n, x.ctypes.data, float(L), int(N), out_i.ctypes.data,
                                                      out_d.ctypes.data))
    return out_i[:n], out_d[:n]


def test_ngp_cell_is_histogramdd_bin_for_bin():
    """ParticleSnapshot.make_map == np.histogramdd on np.linspace(0, L, N + 1) (utils/io.py:629-677): interior edges go up,
    x == L goes to the last bin, anything outside is dropped (-1)."""
    rng = np.random.default_rng(3)
    for L, N in ((205.0, 512), (1000.0, 1024), (1.0, 7), (62.5, 3), (33.3, 1000)):
        edges = np.linspace(0, L, N + 1)
        x = np.concatenate([rng.uniform(0, L, 200000), edges, np.nextafter(edges, -np.inf), np.nextafter(edges, np.inf),
                            [-1e-300, -1.0, L * (1 + 1e-15), 2 * L, np.nan, np.inf, -np.inf]])
        got, _ = helper(0, x, L, N)
        want = np.searchsorted(edges, x, side='right') - 1
        want = np.where(x == L, N - 1, want)
        want = np.where((x >= 0) & (x <= L), want, -1)
        assert np.array_equal(got, want), (L, N, np.flatnonzero(got != want)[:5])
        # and through np.histogramdd itself (1-D marginal of what make_map calls)
        ok = got >= 0
        h = np.histogramdd(x[ok][:, None], bins=(edges,))[0]
        assert np.array_equal(h, np.bincount(got[ok], minlength=N))


def test_wrap_once_and_cell_of():
    rng = np.random.default_rng(4)
    L = 205.0
    x = np.concatenate([rng.uniform(-L, 2 * L, 100000), [0.0, L, -0.0, np.nextafter(L, np.inf), np.nextafter(0.0, -np.inf)]])
    _, got = helper(1, x, L, 1)
    want = x.copy()                                                       # SnapshotRunner.py:272-273: one wrap, > L then < 0
    want = np.where(want > L, want - L, want)
    want = np.where(want < 0, want + L, want)
    assert np.array_equal(got, want)
    xin = np.concatenate([rng.uniform(0, L, 100000), [0.0, L, np.nextafter(L, -np.inf)]])
    for nc in (1, 41, 215, 1024):
        got_c, _ = helper(2, xin, L, nc)
        want_c = np.clip(np.floor(xin / L * nc).astype(np.int64), 0, nc - 1)
        assert np.array_equal(got_c, want_c)


def test_cutout_coordinates_and_indices_are_numpy_bit_for_bit():
    """Map2DRunner.py:500-528: x = np.linspace(-Nsize/2, Nsize/2, Nsize) * res ; pick_indices(centre, Nsize // 2, Npix)."""
    rng = np.random.default_rng(5)
    for ns in list(range(2, 130, 2)) + [256, 510, 512]:
        for res in (1.0, 0.9765625, 0.3, 1000.0 / 1024, float(rng.uniform(0.01, 5))):
            N = max(2 * ns, 64)
            cen = int(rng.integers(0, N))
            idx, coord = helper(3, [float(cen)], res, N, n=ns)
            assert np.array_equal(coord, np.linspace(-ns / 2, ns / 2, ns) * res), (ns, res)
            assert np.array_equal(idx, rp._pick_indices(cen, ns // 2, N)), (ns, cen, N)
    # centres at the box edge: the periodic wrap on both sides
    for cen in (0, 1, 63):
        idx, _ = helper(3, [float(cen)], 1.0, 64, n=32)
        assert np.array_equal(idx, rp._pick_indices(cen, 16, 64)) and idx.min() >= 0 and idx.max() < 64


def test_axis_deposit_rebuilds_the_numba_regrid_loops():
    """axis_deposit (csrc/grid_kernels.cu, the kernel's own source on the host) gives, per axis, the two cells that overlap [x, x + 1) and
    their overlap lengths; the products over the axes must reproduce the literal window-scan loops of regrid_pixels_2D / 3D
    (oracle/grid_deposit.c == Map2DRunner.py:13-162) -- periodic wrap, negative and far-away positions, exact integers."""
    from baryonforge_b200 import _lib
    rng = np.random.default_rng(8)

    def axis(pos, N):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        c = np.empty((pos.size, 2), dtype=np.int64)
        w = np.empty((pos.size, 2), dtype=np.float64)
        _lib.check(_lib.lib().bfg_test_axis_deposit_host(pos.size, pos.ctypes.data, int(N), c.ctypes.data, w.ctypes.data))
        return c, w

    for ndim, N, n in ((2, 16, 4000), (2, 37, 4000), (3, 12, 6000), (3, 9, 3000)):
        pos = rng.uniform(-0.5 * N, 1.5 * N, (n, ndim))
        pos[:200] = rng.integers(-N, 2 * N, (200, ndim)).astype(float)          # exact integers: one of the two overlaps is 0
        pos[200:400] = np.round(pos[200:400]) + rng.choice([1e-13, -1e-13], (200, ndim))
        pos[400:420] = rng.uniform(-50 * N, 50 * N, (20, ndim))                   # many periods away
        val = rng.uniform(0.1, 10, n)
        # the oracle applies `pix_offsets[:, k] += grids[k]` to (cell index + offset); hand it absolute positions directly
        grid = np.zeros((N,) * ndim)
        lib = rp._gridlib()
        flat = np.ascontiguousarray(pos)
        if ndim == 2:
            lib.grido_regrid_2d(grid, N, n, flat, np.ascontiguousarray(val))
        else:
            lib.grido_regrid_3d(grid, N, n, flat, np.ascontiguousarray(val))
        cw = [axis(pos[:, k], N) for k in range(ndim)]
        got = np.zeros((N,) * ndim)
        for bits in range(2 ** ndim):
            sel = [(bits >> k) & 1 for k in range(ndim)]
            w = val.copy()
            ok = np.ones(n, dtype=bool)
            wprod = np.ones(n)
            for k in range(ndim):
                wk = cw[k][1][:, sel[k]]
                ok &= wk > 0                                                       # the loops' strict `> 0` test
                wprod = wprod * wk if k else wk.copy()
            # component 0 rides on array axis 1 (x), component 1 on axis 0 (y), component 2 on axis 2   (grid[i = y][j = x][k = z])
            idx = (cw[1][0][:, sel[1]], cw[0][0][:, sel[0]]) + ((cw[2][0][:, sel[2]],) if ndim == 3 else ())
            np.add.at(got, tuple(i[ok] for i in idx), (wprod * val)[ok])
        assert np.allclose(got, grid, rtol=1e-13, atol=1e-13 * grid.max()), (ndim, N, np.abs(got - grid).max())
        assert np.isclose(got.sum(), val.sum(), rtol=1e-12)
