/*
 * oracle/healpix_ring.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp64 / int64) of the HEALPix RING-scheme geometry that
 * the reference reaches through the un-vendored third-party package `healpy`
 * (version unpinned: /root/reference/pyproject.toml:19).  The reference's call sites:
 *   hp.ang2vec / hp.query_disc / hp.get_interp_weights / hp.pix2vec / hp.vec2ang
 *   -> /root/reference/BaryonForge/Runners/HealpixRunner.py:327,330,334,336,357,358,361,426
 *
 * healpy is absent from this image, so what follows is the PUBLISHED algorithm of
 * HEALPix C++ `T_Healpix_Base<int64>` (healpix_base.cc: pix2loc, loc2pix, ring_above,
 * ring2z, get_ring_info_small, get_ring_info2, query_disc_internal with fact = 0,
 * get_interpol), written out again with the same operation order in fp64 so that
 * floor()/int() decisions fall on the same side.  It is pinned against healpy's own
 * docstring known-answer values in tests/test_oracle_healpix.py and against brute force.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef int64_t i64;

static const double PI = 3.141592653589793238462643383279502884197;
static const double TWOPI = 6.283185307179586476925286766559005768394;
static const double HALFPI = 1.570796326794896619231321691639751442099;
static const double INV_TWOPI = 1.0 / 6.283185307179586476925286766559005768394;
static const double INV_HALFPI = 0.6366197723675813430755350534900574;
static const double TWOTHIRD = 2.0 / 3.0;

typedef struct {
    i64 nside, npix, ncap;
    double fact1, fact2;
} hbase;

static hbase mk(i64 nside) {
    hbase b;
    b.nside = nside;
    b.npix = 12 * nside * nside;
    b.ncap = (nside * (nside - 1)) << 1;
    b.fact2 = 4. / (double)b.npix;
    b.fact1 = (double)(nside << 1) * b.fact2;
    return b;
}

static i64 isqrt64(i64 arg) {
    i64 res = (i64)sqrt((double)arg + 0.5);
    if (arg < ((i64)1 << 50)) return res;
    if (res * res > arg) --res;
    else if ((res + 1) * (res + 1) <= arg) ++res;
    return res;
}

static double fmodulo(double v1, double v2) {
    if (v1 >= 0) return (v1 < v2) ? v1 : fmod(v1, v2);
    double tmp = fmod(v1, v2) + v2;
    return (tmp == v2) ? 0. : tmp;
}

/* healpix_base.cc: T_Healpix_Base::pix2loc, RING branch */
static void pix2loc(const hbase *b, i64 pix, double *z, double *phi, double *sth, int *have_sth) {
    *have_sth = 0;
    if (pix < b->ncap) {
        i64 iring = (1 + isqrt64(1 + 2 * pix)) >> 1;
        i64 iphi = (pix + 1) - 2 * iring * (iring - 1);
        double tmp = (double)(iring * iring) * b->fact2;
        *z = 1.0 - tmp;
        if (*z > 0.99) { *sth = sqrt(tmp * (2. - tmp)); *have_sth = 1; }
        *phi = ((double)iphi - 0.5) * HALFPI / (double)iring;
    } else if (pix < (b->npix - b->ncap)) {
        i64 nl4 = 4 * b->nside;
        i64 ip = pix - b->ncap;
        i64 tmp = ip / nl4;
        i64 iring = tmp + b->nside, iphi = ip - nl4 * tmp + 1;
        double fodd = ((iring + b->nside) & 1) ? 1 : 0.5;
        *z = (double)(2 * b->nside - iring) * b->fact1;
        *phi = ((double)iphi - fodd) * PI * 0.75 * b->fact1;
    } else {
        i64 ip = b->npix - pix;
        i64 iring = (1 + isqrt64(2 * ip - 1)) >> 1;
        i64 iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
        double tmp = (double)(iring * iring) * b->fact2;
        *z = tmp - 1.0;
        if (*z < -0.99) { *sth = sqrt(tmp * (2. - tmp)); *have_sth = 1; }
        *phi = ((double)iphi - 0.5) * HALFPI / (double)iring;
    }
}

/* healpix_base.cc: T_Healpix_Base::loc2pix, RING branch */
static i64 loc2pix(const hbase *b, double z, double phi, double sth, int have_sth) {
    double za = fabs(z);
    double tt = fmodulo(phi * INV_HALFPI, 4.0);
    if (za <= TWOTHIRD) {
        i64 nl4 = 4 * b->nside;
        double temp1 = (double)b->nside * (0.5 + tt);
        double temp2 = (double)b->nside * z * 0.75;
        i64 jp = (i64)(temp1 - temp2);
        i64 jm = (i64)(temp1 + temp2);
        i64 ir = b->nside + 1 + jp - jm;
        i64 kshift = 1 - (ir & 1);
        i64 t1 = jp + jm - b->nside + kshift + 1 + nl4 + nl4;
        i64 ip = (t1 >> 1) % nl4;
        return b->ncap + (ir - 1) * nl4 + ip;
    } else {
        double tp = tt - (double)(i64)tt;
        double tmp = ((za < 0.99) || (!have_sth)) ? (double)b->nside * sqrt(3 * (1 - za))
                                                  : (double)b->nside * sth / sqrt((1. + za) / 3.);
        i64 jp = (i64)(tp * tmp);
        i64 jm = (i64)((1.0 - tp) * tmp);
        i64 ir = jp + jm + 1;
        i64 ip = (i64)(tt * (double)ir);
        return (z > 0) ? 2 * ir * (ir - 1) + ip : b->npix - 2 * ir * (ir + 1) + ip;
    }
}

static i64 ring_above(const hbase *b, double z) {
    double az = fabs(z);
    if (az <= TWOTHIRD) return (i64)((double)b->nside * (2 - 1.5 * z));
    i64 iring = (i64)((double)b->nside * sqrt(3 * (1 - az)));
    return (z > 0) ? iring : 4 * b->nside - iring - 1;
}

static double ring2z(const hbase *b, i64 ring) {
    if (ring < b->nside) return 1 - (double)(ring * ring) * b->fact2;
    if (ring <= 3 * b->nside) return (double)(2 * b->nside - ring) * b->fact1;
    ring = 4 * b->nside - ring;
    return (double)(ring * ring) * b->fact2 - 1;
}

static void ring_info_small(const hbase *b, i64 ring, i64 *startpix, i64 *ringpix, int *shifted) {
    if (ring < b->nside) {
        *shifted = 1; *ringpix = 4 * ring; *startpix = 2 * ring * (ring - 1);
    } else if (ring < 3 * b->nside) {
        *shifted = ((ring - b->nside) & 1) == 0;
        *ringpix = 4 * b->nside;
        *startpix = b->ncap + (ring - b->nside) * (*ringpix);
    } else {
        *shifted = 1;
        i64 nr = 4 * b->nside - ring;
        *ringpix = 4 * nr;
        *startpix = b->npix - 2 * nr * (nr + 1);
    }
}

static void ring_info2(const hbase *b, i64 ring, i64 *startpix, i64 *ringpix, double *theta, int *shifted) {
    i64 northring = (ring > 2 * b->nside) ? 4 * b->nside - ring : ring;
    if (northring < b->nside) {
        double tmp = (double)(northring * northring) * b->fact2;
        double costheta = 1 - tmp;
        double sintheta = sqrt(tmp * (2 - tmp));
        *theta = atan2(sintheta, costheta);
        *ringpix = 4 * northring;
        *shifted = 1;
        *startpix = 2 * northring * (northring - 1);
    } else {
        *theta = acos((double)(2 * b->nside - northring) * b->fact1);
        *ringpix = 4 * b->nside;
        *shifted = ((northring - b->nside) & 1) == 0;
        *startpix = b->ncap + (northring - b->nside) * (*ringpix);
    }
    if (northring != ring) {
        *theta = PI - *theta;
        *startpix = b->npix - *startpix - *ringpix;
    }
}

/* ------------------------------------------------------------------ exported ---- */

void hpo_pix2vec(i64 nside, i64 n, const i64 *pix, double *x, double *y, double *zz) {
    hbase b = mk(nside);
    for (i64 i = 0; i < n; ++i) {
        double z, phi, sth; int have;
        pix2loc(&b, pix[i], &z, &phi, &sth, &have);
        if (!have) sth = sqrt((1. - z) * (1. + z));
        x[i] = sth * cos(phi); y[i] = sth * sin(phi); zz[i] = z;
    }
}

/* pix2vec of the contiguous range [p0, p0+n) (used for whole-map calls) */
void hpo_pix2vec_range(i64 nside, i64 p0, i64 n, double *x, double *y, double *zz) {
    hbase b = mk(nside);
    for (i64 i = 0; i < n; ++i) {
        double z, phi, sth; int have;
        pix2loc(&b, p0 + i, &z, &phi, &sth, &have);
        if (!have) sth = sqrt((1. - z) * (1. + z));
        x[i] = sth * cos(phi); y[i] = sth * sin(phi); zz[i] = z;
    }
}

void hpo_pix2ang(i64 nside, i64 n, const i64 *pix, double *theta, double *phi) {
    hbase b = mk(nside);
    for (i64 i = 0; i < n; ++i) {
        double z, ph, sth; int have;
        pix2loc(&b, pix[i], &z, &ph, &sth, &have);
        theta[i] = have ? atan2(sth, z) : acos(z);
        phi[i] = ph;
    }
}

void hpo_ang2pix(i64 nside, i64 n, const double *theta, const double *phi, i64 *pix) {
    hbase b = mk(nside);
    for (i64 i = 0; i < n; ++i) {
        double th = theta[i];
        pix[i] = ((th < 0.01) || (th > 3.14159 - 0.01)) ? loc2pix(&b, cos(th), phi[i], sin(th), 1)
                                                       : loc2pix(&b, cos(th), phi[i], 0., 0);
    }
}

/* pointing(vec3) + normalize(), as healpy's query_disc wrapper does before calling query_disc */
void hpo_vec2pointing(const double *v, double *theta, double *phi) {
    *theta = atan2(sqrt(v[0] * v[0] + v[1] * v[1]), v[2]);
    double p = ((v[0] == 0.) && (v[1] == 0.)) ? 0. : atan2(v[1], v[0]);
    if (p < 0.) p += TWOPI;
    *phi = fmodulo(p, TWOPI);
}

/*
 * query_disc, RING, non-inclusive (fact = 0).  Writes up to `cap` ascending pixel
 * indices into out (may be NULL to count only) and returns the count.
 */
i64 hpo_query_disc(i64 nside, double theta, double phi, double radius, i64 *out, i64 cap) {
    hbase b = mk(nside);
    i64 cnt = 0;
#define EMIT_RANGE(lo, hi)                                                     \
    do {                                                                       \
        i64 _lo = (lo), _hi = (hi);                                            \
        if (_lo < last_end) _lo = last_end; /* rangeset::append merges overlap */ \
        for (i64 _p = _lo; _p < _hi; ++_p) { if (out && cnt < cap) out[cnt] = _p; ++cnt; } \
        if (_hi > last_end) last_end = _hi;                                    \
    } while (0)
    i64 last_end = 0;
    double rsmall = radius, rbig = radius;
    if (rsmall >= PI) { EMIT_RANGE(0, b.npix); return cnt; }
    rbig = (rbig < PI) ? rbig : PI;
    double cosrbig = cos(rbig);
    double z0 = cos(theta);
    double xa = 1. / sqrt((1 - z0) * (1 + z0));
    double rlat1 = theta - rsmall;
    double zmax = cos(rlat1);
    i64 irmin = ring_above(&b, zmax) + 1;
    if ((rlat1 <= 0) && (irmin > 1)) {
        i64 sp, rp; int dummy;
        ring_info_small(&b, irmin - 1, &sp, &rp, &dummy);
        EMIT_RANGE(0, sp + rp);
    }
    double rlat2 = theta + rsmall;
    double zmin = cos(rlat2);
    i64 irmax = ring_above(&b, zmin);
    for (i64 iz = irmin; iz <= irmax; ++iz) {
        double z = ring2z(&b, iz);
        double x = (cosrbig - z * z0) * xa;
        double ysq = 1 - z * z - x * x;
        double dphi = (ysq <= 0) ? 0 : atan2(sqrt(ysq), x);
        if (dphi > 0) {
            i64 nr, ipix1; int shifted;
            ring_info_small(&b, iz, &ipix1, &nr, &shifted);
            double shift = shifted ? 0.5 : 0.;
            i64 ipix2 = ipix1 + nr - 1;
            i64 ip_lo = (i64)floor((double)nr * INV_TWOPI * (phi - dphi) - shift) + 1;
            i64 ip_hi = (i64)floor((double)nr * INV_TWOPI * (phi + dphi) - shift);
            if (ip_hi >= nr) { ip_lo -= nr; ip_hi -= nr; }
            if (ip_lo < 0) {
                EMIT_RANGE(ipix1, ipix1 + ip_hi + 1);
                EMIT_RANGE(ipix1 + ip_lo + nr, ipix2 + 1);
            } else {
                EMIT_RANGE(ipix1 + ip_lo, ipix1 + ip_hi + 1);
            }
        }
    }
    if ((rlat2 >= PI) && (irmax + 1 < 4 * b.nside)) {
        i64 sp, rp; int dummy;
        ring_info_small(&b, irmax + 1, &sp, &rp, &dummy);
        EMIT_RANGE(sp, b.npix);
    }
#undef EMIT_RANGE
    return cnt;
}

/* get_interpol for n directions; pix/wgt are [4][n] like healpy returns them */
void hpo_get_interpol(i64 nside, i64 n, const double *theta, const double *phi, i64 *pix, double *wgt) {
    hbase b = mk(nside);
    for (i64 k = 0; k < n; ++k) {
        double th = theta[k], ph = phi[k];
        i64 p[4] = {0, 0, 0, 0};
        double w[4] = {0, 0, 0, 0};
        double z = cos(th);
        i64 ir1 = ring_above(&b, z);
        i64 ir2 = ir1 + 1;
        double theta1 = 0, theta2 = 0, w1, tmp, dphi;
        i64 sp, nr, i1, i2; int shift;
        if (ir1 > 0) {
            ring_info2(&b, ir1, &sp, &nr, &theta1, &shift);
            dphi = TWOPI / (double)nr;
            tmp = (ph / dphi - .5 * shift);
            i1 = (tmp < 0) ? (i64)tmp - 1 : (i64)tmp;
            w1 = (ph - ((double)i1 + .5 * shift) * dphi) / dphi;
            i2 = i1 + 1;
            if (i1 < 0) i1 += nr;
            if (i2 >= nr) i2 -= nr;
            p[0] = sp + i1; p[1] = sp + i2;
            w[0] = 1 - w1; w[1] = w1;
        }
        if (ir2 < (4 * b.nside)) {
            ring_info2(&b, ir2, &sp, &nr, &theta2, &shift);
            dphi = TWOPI / (double)nr;
            tmp = (ph / dphi - .5 * shift);
            i1 = (tmp < 0) ? (i64)tmp - 1 : (i64)tmp;
            w1 = (ph - ((double)i1 + .5 * shift) * dphi) / dphi;
            i2 = i1 + 1;
            if (i1 < 0) i1 += nr;
            if (i2 >= nr) i2 -= nr;
            p[2] = sp + i1; p[3] = sp + i2;
            w[2] = 1 - w1; w[3] = w1;
        }
        if (ir1 == 0) {
            double wtheta = th / theta2;
            w[2] *= wtheta; w[3] *= wtheta;
            double fac = (1 - wtheta) * 0.25;
            w[0] = fac; w[1] = fac; w[2] += fac; w[3] += fac;
            p[0] = (p[2] + 2) & 3;
            p[1] = (p[3] + 2) & 3;
        } else if (ir2 == 4 * b.nside) {
            double wtheta = (th - theta1) / (PI - theta1);
            w[0] *= 1 - wtheta; w[1] *= 1 - wtheta;
            double fac = wtheta * 0.25;
            w[0] += fac; w[1] += fac; w[2] = fac; w[3] = fac;
            p[2] = ((p[0] + 2) & 3) + b.npix - 4;
            p[3] = ((p[1] + 2) & 3) + b.npix - 4;
        } else {
            double wtheta = (th - theta1) / (theta2 - theta1);
            w[0] *= 1 - wtheta; w[1] *= 1 - wtheta;
            w[2] *= wtheta; w[3] *= wtheta;
        }
        for (int c = 0; c < 4; ++c) { pix[c * n + k] = p[c]; wgt[c * n + k] = w[c]; }
    }
}

/* ring geometry helpers exported for tests */
i64 hpo_ring_above(i64 nside, double z) { hbase b = mk(nside); return ring_above(&b, z); }
double hpo_ring2z(i64 nside, i64 ring) { hbase b = mk(nside); return ring2z(&b, ring); }
void hpo_ring_info(i64 nside, i64 ring, i64 *startpix, i64 *ringpix, int *shifted) {
    hbase b = mk(nside); ring_info_small(&b, ring, startpix, ringpix, shifted);
}

/*
 * Sequential fp64 scatter-add: restates the numba loop
 * /root/reference/BaryonForge/Runners/HealpixRunner.py:17-71 (regrid_pixels_hpix),
 * child arrays laid out [N][4].
 */
void hpo_regrid_scatter(double *hmap, i64 n, const double *parent_vals, const i64 *child_pix, const double *child_w) {
    for (i64 i = 0; i < n; ++i)
        for (int j = 0; j < 4; ++j) hmap[child_pix[i * 4 + j]] += child_w[i * 4 + j] * parent_vals[i];
}

/* ---- RING <-> NEST (T_Healpix_Base::ring2xyf / xyf2ring / xyf2nest / nest2xyf; nside must be a power of two).
 * The reference itself only uses RING maps (utils/io.py:302); these exist because BASELINE.json's north_star names
 * "ang2pix ring/nest", and are pinned by the hierarchy property of the NESTED scheme (tests/test_oracle_healpix.py). */
static const int JRLL[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
static const int JPLL[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};

static i64 spread(i64 v) {           /* bit k of v -> bit 2k */
    i64 r = 0;
    for (int k = 0; k < 31; ++k) r |= ((v >> k) & 1) << (2 * k);
    return r;
}
static i64 compress(i64 v) {         /* bit 2k of v -> bit k */
    i64 r = 0;
    for (int k = 0; k < 31; ++k) r |= ((v >> (2 * k)) & 1) << k;
    return r;
}

static void ring2xyf(hbase b, i64 pix, i64 *ix, i64 *iy, int *face) {
    i64 n = b.nside, nl2 = 2 * n, iring, iphi, kshift, nr;
    if (pix < b.ncap) {
        iring = (1 + isqrt64(1 + 2 * pix)) >> 1;
        iphi = pix + 1 - 2 * iring * (iring - 1);
        kshift = 0; nr = iring;
        *face = (int)((iphi - 1) / nr);
    } else if (pix < b.npix - b.ncap) {
        i64 ip = pix - b.ncap, tmp = ip / (4 * n);
        iring = tmp + n;
        iphi = ip - tmp * 4 * n + 1;
        kshift = (iring + n) & 1; nr = n;
        i64 ire = tmp + 1, irm = nl2 + 1 - tmp;
        i64 ifm = (iphi - ire / 2 + n - 1) / n, ifp = (iphi - irm / 2 + n - 1) / n;
        *face = (int)((ifp == ifm) ? (ifp | 4) : ((ifp < ifm) ? ifp : (ifm + 8)));
    } else {
        i64 ip = b.npix - pix;
        iring = (1 + isqrt64(2 * ip - 1)) >> 1;
        iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
        kshift = 0; nr = iring;
        iring = 2 * nl2 - iring;
        *face = (int)(8 + (iphi - 1) / nr);
    }
    i64 irt = iring - JRLL[*face] * n + 1;
    i64 ipt = 2 * iphi - JPLL[*face] * nr - kshift - 1;
    if (ipt >= nl2) ipt -= 8 * n;
    *ix = (ipt - irt) >> 1;
    *iy = (-ipt - irt) >> 1;
}

static i64 xyf2ring(hbase b, i64 ix, i64 iy, int face) {
    i64 n = b.nside, nl4 = 4 * n, jr = JRLL[face] * n - ix - iy - 1, nr, n_before, kshift;
    if (jr < n) { nr = jr; n_before = 2 * nr * (nr - 1); kshift = 0; }
    else if (jr > 3 * n) { nr = nl4 - jr; n_before = b.npix - 2 * (nr + 1) * nr; kshift = 0; }
    else { nr = n; n_before = b.ncap + (jr - n) * nl4; kshift = (jr - n) & 1; }
    i64 jp = (JPLL[face] * nr + ix - iy + 1 + kshift) / 2;
    if (jp > nl4) jp -= nl4; else if (jp < 1) jp += nl4;
    return n_before + jp - 1;
}

void hpo_ring2nest(i64 nside, i64 n, const i64 *ring, i64 *nest) {
    hbase b = mk(nside);
    for (i64 i = 0; i < n; ++i) {
        i64 ix, iy; int f;
        ring2xyf(b, ring[i], &ix, &iy, &f);
        nest[i] = (i64)f * nside * nside + spread(ix) + (spread(iy) << 1);
    }
}

void hpo_nest2ring(i64 nside, i64 n, const i64 *nest, i64 *ring) {
    hbase b = mk(nside);
    i64 npface = nside * nside;
    for (i64 i = 0; i < n; ++i) {
        int f = (int)(nest[i] / npface);
        i64 p = nest[i] % npface;
        ring[i] = xyf2ring(b, compress(p), compress(p >> 1), f);
    }
}
