"""
CPU tests of the on-disk side of the shell runners: baryonforge_b200.fits (what `hp.read_map(path)` does at
BaryonForge/utils/io.py:346-347 for the files healpy writes) and LightconeShell(path=...).
"""
import os

import numpy as np
import pytest

from baryonforge_b200 import fits, synth
import baryonforge_b200 as b


@pytest.mark.parametrize("nside", [1, 2, 8, 64])
def test_nest2ring_matches_the_oracle(nside):
    from oracle import hpo
    p = np.arange(12 * nside * nside)
    got = fits.nest2ring(nside, p)
    assert np.array_equal(got, hpo.nest2ring(nside, p))
    assert np.array_equal(np.sort(got), p)                                   # a permutation
    with pytest.raises(ValueError):
        fits.nest2ring(12, p)


@pytest.mark.parametrize("nside,dt", [(4, 'f8'), (32, 'f4'), (64, 'f8'), (16, 'i4')])
def test_write_read_round_trip_both_orderings(tmp_path, nside, dt):
    from oracle import hpo
    rng = np.random.default_rng(nside)
    m = (rng.uniform(-10, 10, 12 * nside * nside) * (100 if dt == 'i4' else 1)).astype(dt)
    n2r = hpo.nest2ring(nside, np.arange(m.size))
    for nest in (False, True):
        path = str(tmp_path / ("m_%d.fits" % nest))
        fits.write_map(path, m, nest=nest)
        assert os.path.getsize(path) % 2880 == 0
        with pytest.raises(OSError):
            fits.write_map(path, m, nest=nest)
        same = fits.read_map(path, nest=nest)
        assert same.dtype == np.dtype(dt) and same.dtype.isnative and np.array_equal(same, m)
        other = fits.read_map(path, nest=not nest)
        if nest:
            assert np.array_equal(other[n2r], m)                             # NESTED file -> RING array
        else:
            assert np.array_equal(other, m[n2r])                             # RING file -> NESTED array


def _cards(*cards):
    txt = ''.join('%-80s' % c for c in cards + ('END',))
    return (txt + ' ' * ((-len(txt)) % 2880)).encode('ascii')


def test_hand_built_two_column_table(tmp_path):
    """A file not written by write_map: two float64 columns with repeat 1, a string card containing '/', D exponents."""
    nside = 2
    n = 12 * nside * nside
    a, c = np.arange(n, dtype='>f8') * 0.5, -np.arange(n, dtype='>f8')
    rows = np.empty(n, dtype=[('a', '>f8'), ('c', '>f8')])
    rows['a'], rows['c'] = a, c
    path = str(tmp_path / "two.fits")
    with open(path, 'wb') as f:
        f.write(_cards("SIMPLE  =                    T / file conforms", "BITPIX  =                    8", "NAXIS   =                    0",
                       "EXTEND  =                    T"))
        f.write(_cards("XTENSION= 'BINTABLE'           / binary table", "BITPIX  =                    8",
                       "NAXIS   =                    2", "NAXIS1  =                   16", "NAXIS2  = %20d" % n,
                       "PCOUNT  =                    0", "GCOUNT  =                    1", "TFIELDS =                    2",
                       "TTYPE1  = 'KAPPA   '", "TFORM1  = 'D       '", "TTYPE2  = 'GAMMA/1 '           / a slash in a string",
                       "TFORM2  = '1D      '", "TSCAL2  =               2.0D0", "PIXTYPE = 'HEALPIX '", "ORDERING= 'RING    '",
                       "NSIDE   = %20d" % nside, "COMMENT no equals sign here", "INDXSCHM= 'IMPLICIT'"))
        raw = rows.tobytes()
        f.write(raw + b'\0' * ((-len(raw)) % 2880))
    assert np.array_equal(fits.read_map(path), a.astype('f8'))
    assert np.array_equal(fits.read_map(path, field=1), 2.0 * c.astype('f8'))
    with pytest.raises(IndexError):
        fits.read_map(path, field=2)


def test_errors(tmp_path):
    p = str(tmp_path / "junk.fits")
    open(p, 'wb').write(b'x' * 100)
    with pytest.raises(ValueError):
        fits.read_map(p)
    with pytest.raises(ValueError):
        fits.write_map(str(tmp_path / "bad.fits"), np.zeros(13))


def test_lightcone_shell_reads_a_path(tmp_path):
    """LightconeShell(path=...) (io.py:346-347) works without healpy."""
    m = synth.shell_map(16, seed=3)
    path = str(tmp_path / "shell.fits")
    fits.write_map(path, m)
    shell = b.LightconeShell(path=path, cosmo=synth.COSMO, redshift=0.3)
    assert shell.NSIDE == 16 and np.array_equal(shell.map, m) and shell.data is shell.map


def test_partial_sky_files_round_trip_like_healpy(tmp_path):
    """INDXSCHM = EXPLICIT (hp.write_map(..., partial=True)): a PIXEL column + the data column, one listed pixel per row;
    hp.read_map returns the full sky with UNSEEN in the pixels the file does not list -- in RING or NESTED order."""
    from baryonforge_b200 import fits
    nside = 16
    npix = 12 * nside * nside
    rng = np.random.default_rng(3)
    m = np.full(npix, fits.UNSEEN)
    seen = np.sort(rng.choice(npix, 500, replace=False))
    m[seen] = rng.uniform(-5, 5, seen.size)
    for dt in (np.float64, np.float32):
        p = str(tmp_path / f"partial_{np.dtype(dt).name}.fits")
        fits.write_map(p, m, dtype=dt, partial=True)
        got = fits.read_map(p)
        assert got.shape == (npix,) and got.dtype == np.dtype(dt)
        assert np.array_equal(got, m.astype(dt))
        assert np.array_equal(np.flatnonzero(got != np.dtype(dt).type(fits.UNSEEN)), seen)
        # the same pixels, asked for in NESTED order
        ring_of_nest = fits.nest2ring(nside, np.arange(npix))
        assert np.array_equal(fits.read_map(p, nest=True), m.astype(dt)[ring_of_nest])
    with pytest.raises(IndexError):
        fits.read_map(p, field=1)                              # one data column only
    shell_map = fits.read_map(p)
    assert np.count_nonzero(shell_map == np.float32(fits.UNSEEN)) == npix - seen.size
