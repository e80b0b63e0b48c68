"""
HEALPix FITS maps on disk -- the data format on the input side of the shell runners.

The reference reads a shell with `hp.read_map(path)` (BaryonForge/utils/io.py:346-347); healpy is a third-party dependency
that is not always installed next to a GPU box, so `read_map` / `write_map` here restate what that call does for the files
healpy itself writes: the first binary-table extension holds the map as big-endian columns (TFORMn = '1024E', 'D', ...),
pixel order from the ORDERING card, full sky (INDXSCHM = IMPLICIT) or partial sky (EXPLICIT, a PIXEL column).  Like hp.read_map's defaults: field 0, RING order
out (a NESTED file is re-ordered), native-endian array of the column's own type.  Host-side I/O only: nothing here is on
the per-halo hot path.
"""
import numpy as np

__all__ = ['read_map', 'write_map', 'nest2ring']

_BLOCK = 2880
_TFORM = {'L': 'u1', 'B': 'u1', 'I': '>i2', 'J': '>i4', 'K': '>i8', 'E': '>f4', 'D': '>f8'}


def _parse_value(raw):
    s = raw.strip()
    if s.startswith("'"):
        end = s.find("'", 1)
        while end != -1 and s[end:end + 2] == "''":          # doubled quote inside a string
            end = s.find("'", end + 2)
        return s[1:end].rstrip()
    s = s.split('/')[0].strip()
    if s in ('T', 'F'):
        return s == 'T'
    try:
        return int(s)
    except ValueError:
        try:
            return float(s.replace('D', 'E'))
        except ValueError:
            return s


def _read_header(f):
    """One header unit: dict of cards (last one wins) read in 2880-byte blocks up to END."""
    cards = {}
    while True:
        block = f.read(_BLOCK)
        if len(block) < _BLOCK:
            raise ValueError("truncated FITS header")
        for i in range(0, _BLOCK, 80):
            card = block[i:i + 80].decode('ascii', errors='replace')
            key = card[:8].strip()
            if key == 'END':
                return cards
            if card[8:10] == '= ':
                cards[key] = _parse_value(card[10:])


def _data_bytes(h):
    naxis = int(h.get('NAXIS', 0))
    if naxis == 0:
        return 0
    n = abs(int(h['BITPIX'])) // 8
    for i in range(1, naxis + 1):
        n *= int(h['NAXIS%d' % i])
    n = (n + int(h.get('PCOUNT', 0))) * int(h.get('GCOUNT', 1))
    return n


def _compress_bits(v):
    """Keep the even bits of v (int64) and pack them: ...b4 b2 b0."""
    v = v & 0x5555555555555555
    v = (v | (v >> 1)) & 0x3333333333333333
    v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0F
    v = (v | (v >> 4)) & 0x00FF00FF00FF00FF
    v = (v | (v >> 8)) & 0x0000FFFF0000FFFF
    v = (v | (v >> 16)) & 0x00000000FFFFFFFF
    return v


_JRLL = np.array([2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4], dtype=np.int64)
_JPLL = np.array([1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7], dtype=np.int64)


def nest2ring(nside, ipnest):
    """RING index of NESTED pixels (nside a power of two), HEALPix's nest2ring: face + (ix, iy) from the interleaved bits,
    ring number jr = jrll[face]*nside - ix - iy - 1, position in ring from jpll[face]."""
    nside = int(nside)
    if nside < 1 or nside & (nside - 1):
        raise ValueError("NESTED ordering needs nside to be a power of two")
    p = np.asarray(ipnest, dtype=np.int64)
    npface = nside * nside
    npix = 12 * npface
    ncap = 2 * nside * (nside - 1)
    face = p // npface
    pf = p - face * npface
    ix = _compress_bits(pf)
    iy = _compress_bits(pf >> 1)
    jr = _JRLL[face] * nside - ix - iy - 1
    north, south = jr < nside, jr > 3 * nside
    nr = np.where(north, jr, np.where(south, 4 * nside - jr, nside))
    n_before = np.where(north, 2 * nr * (nr - 1), np.where(south, npix - 2 * (nr + 1) * nr, ncap + (jr - nside) * 4 * nside))
    kshift = np.where(north | south, 0, (jr - nside) & 1)
    jp = (_JPLL[face] * nr + ix - iy + 1 + kshift) // 2
    jp = np.where(jp > 4 * nr, jp - 4 * nr, jp)
    jp = np.where(jp < 1, jp + 4 * nr, jp)
    return n_before + jp - 1


UNSEEN = -1.6375e30     # healpy's marker of a pixel without data


def read_map(path, field=0, nest=False, hdu=1):
    """
    hp.read_map(path) for full-sky maps (BaryonForge/utils/io.py:347): column `field` of binary-table extension `hdu`,
    flattened row by row, in RING order (nest=False, healpy's default; nest=True returns NESTED order; a file in the other
    order is re-ordered), as a native-endian numpy array of the column's type.  Scaled columns (TSCALn / TZEROn) are
    applied.  Partial-sky files (INDXSCHM = EXPLICIT: a PIXEL column followed by the data columns, as hp.write_map(...,
    partial=True) writes them) come back as a full-sky array with UNSEEN (-1.6375e30) in the pixels the file does not list.
    """
    with open(path, 'rb') as f:
        h = _read_header(f)
        if h.get('SIMPLE') is not True:
            raise ValueError("%s is not a FITS file" % path)
        f.seek(((_data_bytes(h) + _BLOCK - 1) // _BLOCK) * _BLOCK, 1)
        for _ in range(int(hdu)):
            h = _read_header(f)
            start = f.tell()
            size = _data_bytes(h)
            f.seek(start + ((size + _BLOCK - 1) // _BLOCK) * _BLOCK)
        if str(h.get('XTENSION', '')).strip() != 'BINTABLE':
            raise ValueError("HDU %d of %s is not a binary table" % (hdu, path))
        explicit = str(h.get('INDXSCHM', 'IMPLICIT')).strip().upper() == 'EXPLICIT'
        nrow, rowbytes, nfield = int(h['NAXIS2']), int(h['NAXIS1']), int(h['TFIELDS'])
        if explicit:            # partial-sky file: column 1 holds the pixel numbers, `field` counts the data columns after it
            field = field + 1
        if not (1 if explicit else 0) <= field < nfield:
            raise IndexError("field %d out of range (file has %d data columns)" % (field - (1 if explicit else 0),
                                                                                    nfield - (1 if explicit else 0)))
        fields = []
        for i in range(1, nfield + 1):
            tform = str(h['TFORM%d' % i]).strip()
            j = 0
            while j < len(tform) and tform[j].isdigit():
                j += 1
            rep, code = (int(tform[:j]) if j else 1), tform[j:j + 1]
            if code not in _TFORM:
                raise NotImplementedError("TFORM%d = %r is not a numeric column" % (i, tform))
            fields.append(('f%d' % i, _TFORM[code], (rep,)))
        dt = np.dtype(fields)
        if dt.itemsize != rowbytes:
            raise ValueError("row size %d does not match TFORM columns (%d bytes)" % (rowbytes, dt.itemsize))
        f.seek(start)
        if nfield == 1:                                     # the usual healpy layout: one column, rows are plain runs of it
            col = np.fromfile(f, dtype=fields[0][1], count=nrow * fields[0][2][0])
            if col.size != nrow * fields[0][2][0]:
                raise ValueError("truncated FITS table")
        else:
            rows = np.fromfile(f, dtype=dt, count=nrow)
            if rows.size != nrow:
                raise ValueError("truncated FITS table")
            col = rows['f%d' % (field + 1)].reshape(-1)
            pix = rows['f1'].reshape(-1).astype(np.int64) if explicit else None
    if nfield == 1 and col.dtype.byteorder == '>':
        out = col.byteswap(inplace=True).view(col.dtype.newbyteorder('='))     # our own buffer: swap in place (6x faster than astype)
    else:
        out = col.astype(col.dtype.newbyteorder('='))
    scale, zero = h.get('TSCAL%d' % (field + 1), 1), h.get('TZERO%d' % (field + 1), 0)
    if scale != 1 or zero != 0:
        out = out * scale + zero
    if explicit:
        # hp.read_map on a partial-sky file: a full-sky array of UNSEEN with the listed pixels filled in
        nside = int(h['NSIDE'])
        if pix.size != out.size or (pix.size and (pix.min() < 0 or pix.max() >= 12 * nside * nside)):
            raise ValueError("PIXEL column does not match the data column / NSIDE")
        full = np.full(12 * nside * nside, UNSEEN, dtype=out.dtype if out.dtype.kind == 'f' else np.float64)
        full[pix] = out
        out = full
    nside = int(h.get('NSIDE', int(round(np.sqrt(out.size / 12.0)))))
    if 12 * nside * nside != out.size:
        raise ValueError("Wrong pixel number (it is not 12*nside**2)")
    file_nest = str(h.get('ORDERING', 'RING')).strip().upper().startswith('NEST')
    if file_nest != bool(nest):
        ring_of_nest = np.empty(out.size, dtype=np.int64)
        for lo in range(0, out.size, 1 << 20):              # chunks keep nest2ring's temporaries in cache
            hi = min(lo + (1 << 20), out.size)
            ring_of_nest[lo:hi] = nest2ring(nside, np.arange(lo, hi, dtype=np.int64))
        if file_nest:                                       # NESTED file -> RING array
            ring = np.empty_like(out)
            ring[ring_of_nest] = out
            out = ring
        else:                                               # RING file -> NESTED array
            out = out[ring_of_nest]
    return out


def _card(key, value, comment=''):
    if isinstance(value, bool):
        v = '%20s' % ('T' if value else 'F')
    elif isinstance(value, (int, np.integer)):
        v = '%20d' % value
    elif isinstance(value, (float, np.floating)):
        v = '%20s' % repr(float(value)).upper()
    else:
        v = "'%-8s'" % str(value).replace("'", "''")
        v = '%-20s' % v
    s = '%-8s= %s' % (key, v)
    if comment:
        s += ' / ' + comment
    return ('%-80s' % s)[:80]


def _pad(b, fill):
    return b + fill * ((-len(b)) % _BLOCK)


def write_map(path, m, nest=False, dtype=None, column_name='TEMPERATURE', overwrite=False, partial=False):
    """hp.write_map(path, m) for one full-sky map: a primary HDU and one BINTABLE with 1024-pixel rows (one pixel per row
    below nside = 32), big-endian, ORDERING / NSIDE / INDXSCHM cards as healpy writes them.  partial=True writes the
    partial-sky form instead: only the pixels that are not UNSEEN, one per row, behind a PIXEL column (INDXSCHM = EXPLICIT)."""
    import os
    m = np.asarray(m)
    if dtype is not None:
        m = m.astype(dtype)
    nside = int(round(np.sqrt(m.size / 12.0)))
    if m.ndim != 1 or 12 * nside * nside != m.size:
        raise ValueError("Wrong pixel number (it is not 12*nside**2)")
    if partial:
        if m.dtype.str[1:] not in ('f4', 'f8'):
            raise TypeError("partial-sky maps are floating point (UNSEEN marks the missing pixels)")
        if os.path.exists(path) and not overwrite:
            raise OSError("File %s already exists (overwrite=False)" % path)
        pix = np.flatnonzero(m != m.dtype.type(UNSEEN)).astype('>i8')
        rows = np.empty(pix.size, dtype=[('p', '>i8'), ('v', m.dtype.newbyteorder('>'))])
        rows['p'], rows['v'] = pix, m[pix]
        code = {'f4': 'E', 'f8': 'D'}[m.dtype.str[1:]]
        prim = [_card('SIMPLE', True, 'conforms to FITS standard'), _card('BITPIX', 8), _card('NAXIS', 0), _card('EXTEND', True),
                '%-80s' % 'END']
        ext = [_card('XTENSION', 'BINTABLE', 'binary table extension'), _card('BITPIX', 8), _card('NAXIS', 2),
               _card('NAXIS1', rows.dtype.itemsize), _card('NAXIS2', pix.size), _card('PCOUNT', 0), _card('GCOUNT', 1),
               _card('TFIELDS', 2), _card('TTYPE1', 'PIXEL'), _card('TFORM1', '1K'), _card('TTYPE2', column_name),
               _card('TFORM2', '1' + code), _card('PIXTYPE', 'HEALPIX', 'HEALPIX pixelisation'),
               _card('ORDERING', 'NESTED' if nest else 'RING', 'Pixel ordering scheme, either RING or NESTED'),
               _card('NSIDE', nside, 'Resolution parameter of HEALPIX'),
               _card('INDXSCHM', 'EXPLICIT', 'Indexing: IMPLICIT or EXPLICIT'), _card('OBJECT', 'PARTIAL'), _card('GRAIN', 1),
               '%-80s' % 'END']
        with open(path, 'wb') as f:
            f.write(_pad(''.join(prim).encode('ascii'), b' '))
            f.write(_pad(''.join(ext).encode('ascii'), b' '))
            f.write(_pad(rows.tobytes(), b'\0'))
        return
    code = {'f4': 'E', 'f8': 'D', 'i2': 'I', 'i4': 'J', 'i8': 'K', 'u1': 'B'}.get(m.dtype.str[1:])
    if code is None:
        raise TypeError("unsupported map dtype %s" % m.dtype)
    if os.path.exists(path) and not overwrite:
        raise OSError("File %s already exists (overwrite=False)" % path)
    rep = 1024 if m.size % 1024 == 0 else 1
    data = m.astype(m.dtype.newbyteorder('>')).tobytes()
    prim = [_card('SIMPLE', True, 'conforms to FITS standard'), _card('BITPIX', 8), _card('NAXIS', 0), _card('EXTEND', True),
            '%-80s' % 'END']
    ext = [_card('XTENSION', 'BINTABLE', 'binary table extension'), _card('BITPIX', 8), _card('NAXIS', 2),
           _card('NAXIS1', rep * m.dtype.itemsize), _card('NAXIS2', m.size // rep), _card('PCOUNT', 0), _card('GCOUNT', 1),
           _card('TFIELDS', 1), _card('TTYPE1', column_name), _card('TFORM1', '%d%s' % (rep, code)),
           _card('PIXTYPE', 'HEALPIX', 'HEALPIX pixelisation'),
           _card('ORDERING', 'NESTED' if nest else 'RING', 'Pixel ordering scheme, either RING or NESTED'),
           _card('NSIDE', nside, 'Resolution parameter of HEALPIX'), _card('FIRSTPIX', 0), _card('LASTPIX', m.size - 1),
           _card('INDXSCHM', 'IMPLICIT', 'Indexing: IMPLICIT or EXPLICIT'), _card('OBJECT', 'FULLSKY'), '%-80s' % 'END']
    with open(path, 'wb') as f:
        f.write(_pad(''.join(prim).encode('ascii'), b' '))
        f.write(_pad(''.join(ext).encode('ascii'), b' '))
        f.write(_pad(data, b'\0'))
