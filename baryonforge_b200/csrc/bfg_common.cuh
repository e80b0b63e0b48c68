// bfg_common.cuh -- shared device code of libbfg_b200.so (sm_100a only).
//
//  * error plumbing for the C ABI
//  * TableView + per-halo row blending + radial read-out: the device form of
//    scipy.interpolate.RegularGridInterpolator(method='linear', bounds_error=False, fill_value=nan) as used at
//    BaryonForge/Profiles/BaryonCorrection.py:322,404-411 and BaryonForge/utils/Tabulate.py:270-271,318-319
//  * HEALPix RING geometry as device functions (replaces healpy at BaryonForge/Runners/HealpixRunner.py:327-361)
#pragma once
#include <cmath>
#include <cstring>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/bfg_b200.h"

typedef long long i64;
#define BFG_LOG2_TAB 128

namespace bfg {

void set_error(const char *fmt, ...);
// Keep stream-ordered scratch (cudaMallocAsync) cached in the device's default pool instead of handing it back to
// the OS at every synchronisation (the default release threshold is 0).  Idempotent per device.
int retain_async_pool();
struct RingTabEntry;
// Per-(device, nside) table of ring colatitudes for regrid_target_fast; *d_tab = nullptr when BFG_REGRID_LITERAL=1 or the
// map is too coarse for the small-angle forms (nside < 32).
int get_ring_table(long long nside, const RingTabEntry **d_tab, void *stream);

#define BFG_CUDA_OK(expr)                                                                  \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            bfg::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return BFG_ERR_CUDA;                                                           \
        }                                                                                  \
    } while (0)

// Stream-ordered scratch that is handed back on EVERY return path of an entry point (BFG_CUDA_OK / BFG_REQUIRE return early).
struct StreamScratch {
    void *p = nullptr;
    cudaStream_t st;
    explicit StreamScratch(cudaStream_t s) : st(s) {}
    StreamScratch(const StreamScratch &) = delete;
    StreamScratch &operator=(const StreamScratch &) = delete;
    ~StreamScratch() { if (p) cudaFreeAsync(p, st); }
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 1, st); }
    template <class T> T *as() const { return (T *)p; }
};

// First statement of every entry point: forget a NON-sticky error another library left behind in this thread.  NCCL probes
// features at its first collective of a kind (8 ranks, first reduce-scatter: an internal call failed with
// cudaErrorOperatingSystem, NCCL fell back -- and the next `cudaGetLastError()` after one of OUR launches reported it).
// Sticky errors (a faulted context) are not cleared by this and still surface in the checks below.
#define BFG_ENTRY() ((void)cudaGetLastError())

#define BFG_REQUIRE(cond, msg)                                    \
    do {                                                          \
        if (!(cond)) {                                            \
            bfg::set_error("%s: %s", __func__, msg);              \
            return BFG_ERR_INVALID;                               \
        }                                                         \
    } while (0)

// ------------------------------------------------------------------------------------------------
// Table
// ------------------------------------------------------------------------------------------------
struct TableView {
    int ndim;              // 3 .. BFG_MAX_TABLE_DIM ; axis 2 is radial
    int flags;
    int uniform_r;         // radial axis uniform to 1e-12*step -> closed-form cell index
    int n[BFG_MAX_TABLE_DIM];
    i64 stride[BFG_MAX_TABLE_DIM];  // in elements
    const double *ax[BFG_MAX_TABLE_DIM];
    const double *v;
    double r0, r1, inv_dr;  // first/last radial node, 1/step
};

// radial axis: uniform in ln r (np.geomspace -> np.log) unlocks the closed-form cell index
inline void describe_radial_axis(const double *ar, int64_t nr, TableView &view) {
    double step = (ar[nr - 1] - ar[0]) / (double)(nr - 1);
    bool uni = true;
    for (int64_t i = 0; i < nr; ++i) {
        double d = ar[i] - (ar[0] + step * (double)i);
        if ((d < 0 ? -d : d) > 1e-12 * (step < 0 ? -step : step)) { uni = false; break; }
    }
    view.uniform_r = uni ? 1 : 0;
    view.r0 = ar[0];
    view.r1 = ar[nr - 1];
    view.inv_dr = 1.0 / step;
}

// entry i: (rc, -log2(rc)) with rc = the rounded reciprocal of the centre of the i-th mantissa interval
inline void fill_log2_table(double2 *h) {
    for (int i = 0; i < BFG_LOG2_TAB; ++i) {
        long double c = 1.0L + ((long double)i + 0.5L) / (long double)BFG_LOG2_TAB;
        double rc = (double)(1.0L / c);
        h[i].x = rc;
        h[i].y = (double)(-log2l((long double)rc));   // consistent with the ROUNDED reciprocal
    }
}

// TableView over HOST arrays (the host test entries run the kernels' read-out source on the CPU)
inline void host_table_view(int ndim, const int64_t *shape, const double *const *h_axes, const double *h_values, int flags,
                            TableView &T) {
    memset(&T, 0, sizeof(T));
    i64 total = 1;
    for (int d = ndim - 1; d >= 0; --d) {
        T.n[d] = (int)shape[d]; T.stride[d] = total; total *= shape[d];
        T.ax[d] = h_axes[d];
    }
    T.ndim = ndim; T.flags = flags; T.v = h_values;
    describe_radial_axis(h_axes[2], shape[2], T);
}

}  // namespace bfg

struct bfg_table {
    bfg::TableView view;
    int device;
    i64 shape[BFG_MAX_TABLE_DIM];
    double *d_axes[BFG_MAX_TABLE_DIM];
    double *d_values;
};

namespace bfg {

// (the read-out pieces below -- axis_cell, RowBlender, row_lookup -- are __host__ __device__: bfg_test_table_readout_host runs the same
// source on the CPU against scipy's RegularGridInterpolator; device-only intrinsics get host stand-ins)
#ifdef __CUDA_ARCH__
#define BFG_LDG(p) __ldg(p)
#define BFG_QNAN CUDART_NAN
#else
#define BFG_LDG(p) (*(p))
#define BFG_QNAN (__builtin_nan(""))
#endif

// Cell index / normalised distance of one coordinate on one axis, scipy find_indices semantics:
// ax[i] <= x < ax[i+1], last interval right-closed, index clipped to [0, n-2].  Returns false when x is outside
// [ax[0], ax[n-1]] (RegularGridInterpolator then yields fill_value = nan).  NaN x is "inside" and propagates.
__host__ __device__ __forceinline__ bool axis_cell(const double *__restrict__ ax, int n, double x, int &i, double &t) {
    bool inside = !(x < ax[0]) && !(x > ax[n - 1]);
    int lo = 0, hi = n - 1;  // invariant: ax[lo] <= x (or lo == 0), x < ax[hi] (or hi == n-1)
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (x >= ax[mid]) lo = mid; else hi = mid;
    }
    i = lo;
    t = (x - ax[i]) / (ax[i + 1] - ax[i]);
    return inside;
}

// Per-halo constants of the read-out: corner offsets/weights over every non-radial axis.
struct HaloCell {
    int ncorner;               // 2^(ndim-1)
    bool valid;                // false -> every read-out is NaN (coordinate outside a non-radial axis)
    i64 base[1 << (BFG_MAX_TABLE_DIM - 1)];
    double w[1 << (BFG_MAX_TABLE_DIM - 1)];
};

// Blend the 2^(ndim-1) (z, M, extras) corner rows into ONE radial row in shared memory:
//   row[k] = sum_c w_c * values[corner_c, k]
// All corners are always multiplied in, like scipy's _evaluate_linear, so 0 * (-inf) = NaN survives.
// RowBlender holds the per-halo part (cells and weights of the non-radial axes); node(k) is one blended value.
struct RowBlender {
    int nd, NR;
    i64 sr;
    bool valid;                            // false -> a non-radial coordinate is outside its axis: every read-out is NaN
    int idx[BFG_MAX_TABLE_DIM];
    double tt[BFG_MAX_TABLE_DIM];
    const double *v00, *v01, *v10, *v11;   // the common (z, M, r) table: corner rows ...
    double w00, w01, w10, w11;             // ... and weights hoisted out of the radial loop

    __host__ __device__ __forceinline__ RowBlender(const TableView &T, double lnz, double lnM, const double *__restrict__ extras) {
        nd = T.ndim; NR = T.n[2]; sr = T.stride[2];
        bool ok = true;
        int e = 0;
        for (int d = 0; d < nd; ++d) {
            if (d == 2) continue;
            double x = (d == 0) ? lnz : (d == 1) ? lnM : extras[e++];
            ok &= axis_cell(T.ax[d], T.n[d], x, idx[d], tt[d]);
        }
        valid = ok;
        v00 = T.v + (i64)idx[0] * T.stride[0] + (i64)idx[1] * T.stride[1];
        v01 = v00 + T.stride[1];
        v10 = v00 + T.stride[0];
        v11 = v10 + T.stride[1];
        // itertools.product order (z, M) = (0,0), (0,1), (1,0), (1,1); weight = w_z * w_M, summed in that order
        w00 = (1.0 - tt[0]) * (1.0 - tt[1]); w01 = (1.0 - tt[0]) * tt[1];
        w10 = tt[0] * (1.0 - tt[1]); w11 = tt[0] * tt[1];
    }

    __host__ __device__ __forceinline__ double node(const TableView &T, int k) const {
        if (nd == 3) {
            const i64 o = (i64)k * sr;
            double acc = 0.0;
            acc = acc + BFG_LDG(v00 + o) * w00;
            acc = acc + BFG_LDG(v01 + o) * w01;
            acc = acc + BFG_LDG(v10 + o) * w10;
            acc = acc + BFG_LDG(v11 + o) * w11;
            return acc;
        }
        const int nc = 1 << (nd - 1);
        double acc = 0.0;
        for (int c = 0; c < nc; ++c) {
            i64 off = (i64)k * sr;
            double w = 1.0;
            int bit = 0;
            for (int d = 0; d < nd; ++d) {
                if (d == 2) continue;
                int up = (c >> (nd - 2 - bit)) & 1;  // first axis = most significant, itertools.product order
                ++bit;
                off += (i64)(idx[d] + up) * T.stride[d];
                w = w * (up ? tt[d] : (1.0 - tt[d]));
            }
            acc = acc + BFG_LDG(T.v + off) * w;
        }
        return acc;
    }
};

// row[k] = blended node k (times `post`).  Called by every thread of the block; caller __syncthreads() afterwards.
__device__ __forceinline__ void blend_row(const TableView &T, double lnz, double lnM, const double *__restrict__ extras,
                                          double *__restrict__ row, bool &valid, const double post = 1.0) {
    const RowBlender B(T, lnz, lnM, extras);
    valid = B.valid;
    for (int k = threadIdx.x; k < B.NR; k += blockDim.x) row[k] = B.node(T, k) * post;
}

// The same, plus the row as (value, step to the next node) pairs for read-outs that fetch a whole cell with one 16-byte load:
// rowp[k] = (row[k], row[k+1] - row[k]), rowp[NR-1] = (row[NR-1], 0).  The neighbour's value comes from the next lane by
// shuffle; only the last lane of a warp blends a second node.  Every thread of the block must call it (blockDim.x % 32 == 0).
__device__ __forceinline__ void blend_row_pairs(const TableView &T, double lnz, double lnM, const double *__restrict__ extras,
                                                double *__restrict__ row, double2 *__restrict__ rowp, bool &valid,
                                                const double post = 1.0) {
    const RowBlender B(T, lnz, lnM, extras);
    valid = B.valid;
    const int lane = threadIdx.x & 31;
    for (int k0 = 0; k0 < B.NR; k0 += blockDim.x) {      // uniform trip count: the shuffle needs the whole warp
        const int k = k0 + threadIdx.x;
        const double v = (k < B.NR) ? B.node(T, k) * post : 0.0;
        double nxt = __shfl_down_sync(0xffffffffu, v, 1);
        if (lane == 31 && k + 1 < B.NR) nxt = B.node(T, k + 1) * post;
        if (k < B.NR) {
            row[k] = v;
            rowp[k] = make_double2(v, (k + 1 < B.NR) ? nxt - v : 0.0);
        }
    }
}

// Radial read-out of a blended row at x = ln r (or ln r/R): NaN outside [r0, r1]; (1-t)*v0 + t*v1 as scipy does.
template <bool UNIFORM>
__host__ __device__ __forceinline__ double row_lookup(const TableView &T, const double *__restrict__ row, double x) {
    const int NR = T.n[2];
    if (!(x >= T.r0) || !(x <= T.r1)) return (x != x) ? x : BFG_QNAN;
    int k;
    double t;
    if (UNIFORM) {
        double u = (x - T.r0) * T.inv_dr;
        k = (int)u;
        k = min(k, NR - 2);
        t = u - (double)k;
    } else {
        const double *__restrict__ ax = T.ax[2];
        int lo = 0, hi = NR - 1;
        // closed-form guess then bounded correction (geomspace axes hit on the first try)
        double u = (x - T.r0) * T.inv_dr;
        int g = min(max((int)u, 0), NR - 2);
        if (x >= BFG_LDG(ax + g) && x < BFG_LDG(ax + g + 1)) {
            lo = g;
        } else {
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (x >= BFG_LDG(ax + mid)) lo = mid; else hi = mid;
            }
        }
        k = lo;
        double a0 = BFG_LDG(ax + k), a1 = BFG_LDG(ax + k + 1);
        t = (x - a0) / (a1 - a0);
    }
    double v0 = row[k], v1 = row[k + 1];
    return (1.0 - t) * v0 + t * v1;
}

// ------------------------------------------------------------------------------------------------
// HEALPix RING geometry (T_Healpix_Base algorithms, fp64/int64, written for the device)
// ------------------------------------------------------------------------------------------------
struct Hpx {
    i64 nside, npix, ncap, nl4;
    double fact1, fact2;
    __host__ __device__ explicit Hpx(i64 ns) {
        nside = ns;
        npix = 12 * ns * ns;
        ncap = 2 * ns * (ns - 1);
        nl4 = 4 * ns;
        fact2 = 4.0 / (double)npix;
        fact1 = (double)(2 * ns) * fact2;
    }
};

#define BFG_PI 3.141592653589793238462643383279502884197
#define BFG_TWOPI 6.283185307179586476925286766559005768394
#define BFG_HALFPI 1.570796326794896619231321691639751442099
#define BFG_INV_TWOPI (1.0 / 6.283185307179586476925286766559005768394)
#define BFG_INV_HALFPI 0.6366197723675813430755350534900574
#define BFG_TWOTHIRD (2.0 / 3.0)

__host__ __device__ __forceinline__ i64 isqrt_i64(i64 v) { return (i64)sqrt((double)v + 0.5); }  // exact for v < 2^50

__host__ __device__ __forceinline__ i64 ring_above(const Hpx &h, double z) {
    double az = fabs(z);
    if (az <= BFG_TWOTHIRD) return (i64)((double)h.nside * (2.0 - 1.5 * z));
    i64 ir = (i64)((double)h.nside * sqrt(3.0 * (1.0 - az)));
    return (z > 0) ? ir : 4 * h.nside - ir - 1;
}

__host__ __device__ __forceinline__ double ring2z(const Hpx &h, i64 ring) {
    if (ring < h.nside) return 1.0 - (double)(ring * ring) * h.fact2;
    if (ring <= 3 * h.nside) return (double)(2 * h.nside - ring) * h.fact1;
    i64 q = 4 * h.nside - ring;
    return (double)(q * q) * h.fact2 - 1.0;
}

// first pixel, pixel count and half-pixel phase of a ring (1 <= ring <= 4 nside - 1)
__host__ __device__ __forceinline__ void ring_info(const Hpx &h, i64 ring, i64 &start, i64 &nr, bool &shifted) {
    if (ring < h.nside) {
        shifted = true; nr = 4 * ring; start = 2 * ring * (ring - 1);
    } else if (ring < 3 * h.nside) {
        shifted = ((ring - h.nside) & 1) == 0; nr = h.nl4; start = h.ncap + (ring - h.nside) * h.nl4;
    } else {
        shifted = true; i64 q = 4 * h.nside - ring; nr = 4 * q; start = h.npix - 2 * q * (q + 1);
    }
}

// colatitude-related quantities of a ring: z and sin(theta) with the polar-cap accurate form
__host__ __device__ __forceinline__ void ring_z_sth(const Hpx &h, i64 ring, double &z, double &sth) {
    if (ring < h.nside) {
        double tmp = (double)(ring * ring) * h.fact2;
        z = 1.0 - tmp;
        sth = (z > 0.99) ? sqrt(tmp * (2.0 - tmp)) : sqrt((1.0 - z) * (1.0 + z));
    } else if (ring <= 3 * h.nside) {
        z = (double)(2 * h.nside - ring) * h.fact1;
        sth = sqrt((1.0 - z) * (1.0 + z));
    } else {
        i64 q = 4 * h.nside - ring;
        double tmp = (double)(q * q) * h.fact2;
        z = tmp - 1.0;
        sth = (z < -0.99) ? sqrt(tmp * (2.0 - tmp)) : sqrt((1.0 - z) * (1.0 + z));
    }
}

// azimuth of pixel `ip` (0-based within its ring)
__host__ __device__ __forceinline__ double ring_phi(const Hpx &h, i64 ring, i64 ip, bool shifted) {
    if (ring < h.nside) return ((double)(ip + 1) - 0.5) * BFG_HALFPI / (double)ring;
    if (ring < 3 * h.nside) return ((double)(ip + 1) - (shifted ? 0.5 : 1.0)) * BFG_PI * 0.75 * h.fact1;
    return ((double)(ip + 1) - 0.5) * BFG_HALFPI / (double)(4 * h.nside - ring);
}

// ring number (1-based from the north pole) and in-ring index of a RING pixel
__host__ __device__ __forceinline__ void pix2ring(const Hpx &h, i64 pix, i64 &ring, i64 &ip) {
    if (pix < h.ncap) {
        ring = (1 + isqrt_i64(1 + 2 * pix)) >> 1;
        ip = pix - 2 * ring * (ring - 1);
    } else if (pix < h.npix - h.ncap) {
        i64 q = pix - h.ncap;
        i64 t = q / h.nl4;
        ring = t + h.nside;
        ip = q - t * h.nl4;
    } else {
        i64 q = h.npix - pix;
        i64 s = (1 + isqrt_i64(2 * q - 1)) >> 1;  // counted from the south pole
        ring = 4 * h.nside - s;
        ip = 4 * s - (q - 2 * s * (s - 1));
    }
}

__host__ __device__ __forceinline__ void pix2vec(const Hpx &h, i64 pix, double &x, double &y, double &z) {
    i64 ring, ip, start, nr;
    bool shifted;
    pix2ring(h, pix, ring, ip);
    ring_info(h, ring, start, nr, shifted);
    double sth;
    ring_z_sth(h, ring, z, sth);
    double s, c;
    sincos(ring_phi(h, ring, ip, shifted), &s, &c);
    x = sth * c;
    y = sth * s;
}

// The rings a non-inclusive disc touches (query_disc_internal with fact = 0).
struct DiscRings {
    i64 irmin, irmax;  // rings that need the per-ring azimuth test
    i64 ra, rb;        // full iteration range; rings in [ra, irmin) and (irmax, rb] are taken whole (pole inside)
    double z0, xa, cosr, phi0;
    bool all_sky;
};

__host__ __device__ __forceinline__ DiscRings disc_rings(const Hpx &h, double theta, double phi, double radius) {
    DiscRings d;
    d.phi0 = phi;
    d.all_sky = false;
    if (radius >= BFG_PI) {
        d.all_sky = true;
        d.ra = 1; d.rb = 4 * h.nside - 1; d.irmin = d.rb + 1; d.irmax = d.rb;
        d.z0 = d.xa = d.cosr = 0;
        return d;
    }
    d.cosr = cos(radius);
    d.z0 = cos(theta);
    d.xa = 1.0 / sqrt((1.0 - d.z0) * (1.0 + d.z0));
    double rlat1 = theta - radius;
    d.irmin = ring_above(h, cos(rlat1)) + 1;
    d.ra = d.irmin;
    if (rlat1 <= 0 && d.irmin > 1) d.ra = 1;
    double rlat2 = theta + radius;
    d.irmax = ring_above(h, cos(rlat2));
    d.rb = d.irmax;
    if (rlat2 >= BFG_PI && d.irmax + 1 < 4 * h.nside) d.rb = 4 * h.nside - 1;
    return d;
}

// Ring-range sharding: can the disc (its rings +-2, which also covers the 4 interpolation neighbours of the <4-pixel
// fallback, HealpixRunner.py:333-334) contain a pixel of [pix_lo, pix_hi)?
__device__ __forceinline__ bool disc_touches_range(const Hpx &h, const DiscRings &d, i64 pix_lo, i64 pix_hi) {
    i64 st0, nr0, st1, nr1;
    bool sh;
    ring_info(h, min(4 * h.nside - 1, max((i64)1, d.ra - 2)), st0, nr0, sh);
    ring_info(h, max((i64)1, min(4 * h.nside - 1, d.rb + 2)), st1, nr1, sh);
    return !(st0 >= pix_hi || st1 + nr1 <= pix_lo);
}

// Pixels of ring `iz` inside the disc: in-ring indices (ip_lo + i) mod nr for i in [0, cnt).
__host__ __device__ __forceinline__ void disc_ring_span(const Hpx &h, const DiscRings &d, i64 iz, i64 &start, i64 &nr,
                                               bool &shifted, i64 &ip_lo, i64 &cnt) {
    ring_info(h, iz, start, nr, shifted);
    if (iz < d.irmin || iz > d.irmax) { ip_lo = 0; cnt = nr; return; }
    double z = ring2z(h, iz);
    double x = (d.cosr - z * d.z0) * d.xa;
    double ysq = 1.0 - z * z - x * x;
    cnt = 0; ip_lo = 0;
    if (ysq <= 0) return;
    double dphi = atan2(sqrt(ysq), x);
    if (!(dphi > 0)) return;
    double shift = shifted ? 0.5 : 0.0;
    i64 lo = (i64)floor((double)nr * BFG_INV_TWOPI * (d.phi0 - dphi) - shift) + 1;
    i64 hi = (i64)floor((double)nr * BFG_INV_TWOPI * (d.phi0 + dphi) - shift);
    if (hi >= nr) { lo -= nr; hi -= nr; }
    i64 c = hi - lo + 1;
    if (c <= 0) return;
    if (c > nr) c = nr;  // rangeset::append would have merged the overlap
    if (lo < 0) lo += nr;
    ip_lo = lo;
    cnt = c;
}

// get_interpol: 4 neighbour pixels + bilinear weights of a direction (theta, phi)
__host__ __device__ __forceinline__ void ring_theta_info(const Hpx &h, i64 ring, i64 &start, i64 &nr, double &theta, bool &shifted) {
    i64 nring = (ring > 2 * h.nside) ? 4 * h.nside - ring : ring;
    if (nring < h.nside) {
        double tmp = (double)(nring * nring) * h.fact2;
        theta = atan2(sqrt(tmp * (2.0 - tmp)), 1.0 - tmp);
        nr = 4 * nring; shifted = true; start = 2 * nring * (nring - 1);
    } else {
        theta = acos((double)(2 * h.nside - nring) * h.fact1);
        nr = h.nl4; shifted = ((nring - h.nside) & 1) == 0; start = h.ncap + (nring - h.nside) * nr;
    }
    if (nring != ring) { theta = BFG_PI - theta; start = h.npix - start - nr; }
}

__host__ __device__ __forceinline__ void ring_pair(i64 nr, bool shifted, i64 start, double phi, i64 &p0, i64 &p1, double &w1) {
    double dphi = BFG_TWOPI / (double)nr;
    double sh = shifted ? 0.5 : 0.0;
    double tmp = phi / dphi - sh;
    i64 i1 = (tmp < 0) ? (i64)tmp - 1 : (i64)tmp;
    w1 = (phi - ((double)i1 + sh) * dphi) / dphi;
    i64 i2 = i1 + 1;
    if (i1 < 0) i1 += nr;
    if (i2 >= nr) i2 -= nr;
    p0 = start + i1;
    p1 = start + i2;
}

__host__ __device__ __forceinline__ void get_interpol(const Hpx &h, double theta, double phi, i64 pix[4], double w[4]) {
    double z = cos(theta);
    i64 ir1 = ring_above(h, z), ir2 = ir1 + 1;
    double th1 = 0, th2 = 0, w1;
    i64 sp, nr;
    bool sh;
    pix[0] = pix[1] = pix[2] = pix[3] = 0;
    w[0] = w[1] = w[2] = w[3] = 0;
    if (ir1 > 0) {
        ring_theta_info(h, ir1, sp, nr, th1, sh);
        ring_pair(nr, sh, sp, phi, pix[0], pix[1], w1);
        w[0] = 1.0 - w1; w[1] = w1;
    }
    if (ir2 < 4 * h.nside) {
        ring_theta_info(h, ir2, sp, nr, th2, sh);
        ring_pair(nr, sh, sp, phi, pix[2], pix[3], w1);
        w[2] = 1.0 - w1; w[3] = w1;
    }
    if (ir1 == 0) {
        double wt = theta / th2;
        w[2] *= wt; w[3] *= wt;
        double fac = (1.0 - wt) * 0.25;
        w[0] = fac; w[1] = fac; w[2] += fac; w[3] += fac;
        pix[0] = (pix[2] + 2) & 3;
        pix[1] = (pix[3] + 2) & 3;
    } else if (ir2 == 4 * h.nside) {
        double wt = (theta - th1) / (BFG_PI - th1);
        w[0] *= 1.0 - wt; w[1] *= 1.0 - wt;
        double fac = wt * 0.25;
        w[0] += fac; w[1] += fac; w[2] = fac; w[3] = fac;
        pix[2] = ((pix[0] + 2) & 3) + h.npix - 4;
        pix[3] = ((pix[1] + 2) & 3) + h.npix - 4;
    } else {
        double wt = (theta - th1) / (th2 - th1);
        w[0] *= 1.0 - wt; w[1] *= 1.0 - wt;
        w[2] *= wt; w[3] *= wt;
    }
}

// ------------------------------------------------------------------------------------------------
// Re-binning target of one displaced pixel without the acos / atan2 / sincos / cos / acos / acos chain of the literal
// restatement (vec2ang -> degrees -> radians -> get_interp_weights, HealpixRunner.py:357-365).  The displaced direction is
// close to the source pixel's, so the two angles the weights need come from exact small-angle forms:
//   azimuth      phi'  = phi_src + atan(cross / dot),   cross = x oy - y ox  (no cancellation),  dot = x x' + y y'
//   colatitude   theta' - theta_1 = asin(sin theta' cos theta_1 - cos theta' sin theta_1)   (ring 1 = the ring above theta')
// with odd series for |t|, |s| <= 0.05 (truncation < 2e-16 relative).  theta_1, theta_2 and their cos / sin come from a
// per-nside ring table built once with the literal ring formulas (ring_theta_info / ring_z_sth), so the denominator
// theta_2 - theta_1 has the reference's bits.  Round-off differs from the literal chain at the 1e-16 level in the angles,
// i.e. ~1e-12 in a weight (the literal chain itself carries ~1e-11 through phi / dphi); a displaced direction within
// round-off of a ring boundary may pick the neighbouring ring pair, where the bilinear weights are continuous.
// Returns false (caller takes the literal path) next to the poles, for large displacements and for coarse maps.
// One 64-byte entry per ring (two 32-byte loads): everything ring_info / ring_theta_info / ring_z_sth would derive, so the fast
// path has no integer division, no ring-type branches and no trigonometry per ring.
struct __align__(32) RingTabEntry {
    double theta, z, sth, inv_dth;            // colatitude (literal ring formula), cos, sin, 1 / (theta of the next ring - theta)
    double two_over_nr, nr_over_2pi, start, nr2s;   // 2 / nr, nr / 2 pi, first pixel, 2 nr + (1 if the ring is shifted by half a pixel)
};

struct RingGeo { double two_over_nr, nr_over_2pi; i64 start, nr; bool shifted; };

// (the re-binning target below is __host__ __device__: bfg_test_regrid_target_host runs the SAME source on the CPU, where the CPU
// suite holds it against the oracle; the two device-only intrinsics have host stand-ins here)
__host__ __device__ __forceinline__ double4 ldg_f64x4(const void *p) {     // 32 bytes through the read-only path (two 16-byte loads)
#ifdef __CUDA_ARCH__
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
#else
    const double *q = reinterpret_cast<const double *>(p);
    return make_double4(q[0], q[1], q[2], q[3]);
#endif
}

__host__ __device__ __forceinline__ double rcp_rn(double x) {
#ifdef __CUDA_ARCH__
    return __drcp_rn(x);
#else
    return 1.0 / x;
#endif
}

// One entry of the per-nside ring table (k_ring_table on the device, bfg_test_regrid_target_host on the CPU)
struct RingTabEntry;
__host__ __device__ __forceinline__ void ring_table_entry(const Hpx &h, i64 r, double out[8]) {
    for (int k = 0; k < 8; ++k) out[k] = 0.0;
    out[1] = 1.0;
    if (r < 1 || r >= 4 * h.nside) return;
    i64 start, nr, s2, n2;
    bool shifted, sh2;
    ring_theta_info(h, r, start, nr, out[0], shifted);        // colatitude by the literal ring formula of get_interpol
    ring_z_sth(h, r, out[1], out[2]);                         // z and sin(theta) by the polar-cap accurate forms
    ring_info(h, r, start, nr, shifted);
    out[4] = 2.0 / (double)nr;
    out[5] = (double)nr * BFG_INV_TWOPI;
    out[6] = (double)start;
    out[7] = (double)(2 * nr + (shifted ? 1 : 0));
    if (r + 1 < 4 * h.nside) {
        double th_next;
        ring_theta_info(h, r + 1, s2, n2, th_next, sh2);
        out[3] = 1.0 / (th_next - out[0]);
    }
}

__host__ __device__ __forceinline__ RingGeo ring_geo(const RingTabEntry *__restrict__ rt, i64 ring) {
    const double4 b = ldg_f64x4(reinterpret_cast<const double4 *>(rt + ring) + 1);
    RingGeo g;
    g.two_over_nr = b.x; g.nr_over_2pi = b.y; g.start = (i64)b.z;
    const i64 k = (i64)b.w;
    g.nr = k >> 1; g.shifted = (k & 1) != 0;
    return g;
}

// ring of a RING pixel without the 64-bit integer division of pix2ring's equatorial branch (~70 instructions): quotient by a
// reciprocal multiplication (exact operands below 2^52) and one correction step
__host__ __device__ __forceinline__ i64 pix2ring_only(const Hpx &h, i64 pix) {
    if (pix < h.ncap) return (1 + isqrt_i64(1 + 2 * pix)) >> 1;
    if (pix < h.npix - h.ncap) {
        const i64 q = pix - h.ncap;
        i64 t = (i64)((double)q * (1.0 / (double)h.nl4));
        if (t * h.nl4 > q) --t;
        else if ((t + 1) * h.nl4 <= q) ++t;
        return t + h.nside;
    }
    const i64 q = h.npix - pix;
    return 4 * h.nside - ((1 + isqrt_i64(2 * q - 1)) >> 1);
}

// in-ring neighbours of azimuth phi and the weight of the second one: the reference's two divisions by dphi folded into ONE
// multiplication -- tmp = phi nr / 2 pi - shift, w1 = (phi - (i1 + shift) dphi) / dphi == tmp - i1 (identical in exact arithmetic;
// ~1e-12 apart in double, like everything that goes through phi / dphi with up to 16384 pixels per ring)
__host__ __device__ __forceinline__ void ring_pair_fast(const RingGeo &g, double phi, i64 &p0, i64 &p1, double &w1) {
    const double tmp = fma(phi, g.nr_over_2pi, g.shifted ? -0.5 : 0.0);
    const double fl = floor(tmp);
    w1 = tmp - fl;
    i64 i1 = (i64)fl, i2 = i1 + 1;
    if (i1 < 0) i1 += g.nr;
    if (i2 >= g.nr) i2 -= g.nr;
    p0 = g.start + i1;
    p1 = g.start + i2;
}

__host__ __device__ __forceinline__ bool regrid_target_fast(const Hpx &h, const RingTabEntry *__restrict__ rt, i64 p, double ox,
                                                   double oy, double oz, i64 pix[4], double w[4]) {
    const i64 ring = pix2ring_only(h, p);
    const RingGeo gs = ring_geo(rt, ring);
    const double4 as = ldg_f64x4(rt + ring);
    const double z = as.y, sth = as.z;
    // azimuth of the source pixel in units of pi (exact half-integers over 2 nr)
    const double a_pi = ((double)(p - gs.start) + (gs.shifted ? 0.5 : 0.0)) * gs.two_over_nr;
    double sn, cs;
    sincospi(a_pi, &sn, &cs);
    const double x = sth * cs, y = sth * sn;
    const double xn = x + ox, yn = y + oy, zn = z + oz;                       // HealpixRunner.py:357 (not re-normalised)
    const double cross = x * oy - y * ox, dot = fma(x, ox, fma(y, oy, sth * sth));
    if (!(dot > 0.0)) return false;
    const double t = cross * rcp_rn(dot);
    if (!(fabs(t) <= 0.05)) return false;
    const double t2 = t * t;
    const double dphi_s = t * fma(t2, fma(t2, fma(t2, fma(t2, 1.0 / 9.0, -1.0 / 7.0), 0.2), -1.0 / 3.0), 1.0);
    double phi = fma(a_pi, BFG_PI, dphi_s);
    if (phi < 0.0) phi += BFG_TWOPI;
    if (phi >= BFG_TWOPI) phi -= BFG_TWOPI;
    const double rho2 = fma(xn, xn, yn * yn);
    const double inv_dn = rsqrt(fma(zn, zn, rho2));
    const double zc = zn * inv_dn, rs = (rho2 * rsqrt(rho2)) * inv_dn;        // cos, sin of the displaced colatitude
    const i64 ir1 = ring_above(h, zc), ir2 = ir1 + 1;
    if (ir1 < 1 || ir2 > 4 * h.nside - 1) return false;                       // polar caps' first / last ring: literal path
    const double4 a1 = ldg_f64x4(rt + ir1);    // theta_1, cos, sin, 1 / (theta_2 - theta_1)
    const double s = fma(rs, a1.y, -(zc * a1.z));                             // sin(theta' - theta_1)
    if (!(fabs(s) <= 0.05)) return false;
    const double s2 = s * s;
    const double dth = s * fma(s2, fma(s2, fma(s2, fma(s2, 35.0 / 1152.0, 15.0 / 336.0), 0.075), 1.0 / 6.0), 1.0);
    const double wt = dth * a1.w;
    double w1;
    ring_pair_fast(ring_geo(rt, ir1), phi, pix[0], pix[1], w1);
    w[0] = (1.0 - w1) * (1.0 - wt); w[1] = w1 * (1.0 - wt);
    ring_pair_fast(ring_geo(rt, ir2), phi, pix[2], pix[3], w1);
    w[2] = (1.0 - w1) * wt; w[3] = w1 * wt;
    return true;
}

// The literal chain (kept as the checked fallback): vec2ang(lonlat=True) :358, get_interp_weights(lonlat=True) :361.
__device__ __forceinline__ void regrid_target_literal(const Hpx &h, i64 p, double ox, double oy, double oz, i64 pix[4],
                                                      double w[4]) {
    double x, y, z;
    pix2vec(h, p, x, y, z);
    x += ox; y += oy; z += oz;                                                // :357 (not re-normalised)
    const double dn = sqrt(x * x + y * y + z * z);
    const double theta = acos(z / dn);
    double phi = atan2(y, x);
    if (phi < 0) phi += BFG_TWOPI;
    const double lon = phi * (180.0 / BFG_PI), lat = 90.0 - theta * (180.0 / BFG_PI);
    const double th2 = BFG_HALFPI - lat * (BFG_PI / 180.0), ph2 = lon * (BFG_PI / 180.0);
    get_interpol(h, th2, ph2, pix, w);
}

__device__ __forceinline__ void regrid_target(const Hpx &h, const RingTabEntry *__restrict__ rt, i64 p, double ox, double oy,
                                              double oz, i64 pix[4], double w[4]) {
    if (rt == nullptr || !regrid_target_fast(h, rt, p, ox, oy, oz, pix, w)) regrid_target_literal(h, p, ox, oy, oz, pix, w);
}

__host__ __device__ __forceinline__ double fmodulo(double v1, double v2) {
    if (v1 >= 0) return (v1 < v2) ? v1 : fmod(v1, v2);
    double tmp = fmod(v1, v2) + v2;
    return (tmp == v2) ? 0.0 : tmp;
}

__host__ __device__ __forceinline__ i64 ang2pix_ring(const Hpx &h, double theta, double phi) {
    double z = cos(theta);
    bool have_sth = (theta < 0.01) || (theta > 3.14159 - 0.01);
    double sth = have_sth ? sin(theta) : 0.0;
    double za = fabs(z);
    double tt = fmodulo(phi * BFG_INV_HALFPI, 4.0);
    if (za <= BFG_TWOTHIRD) {
        double t1 = (double)h.nside * (0.5 + tt), t2 = (double)h.nside * z * 0.75;
        i64 jp = (i64)(t1 - t2), jm = (i64)(t1 + t2);
        i64 ir = h.nside + 1 + jp - jm;
        i64 ks = 1 - (ir & 1);
        i64 q = jp + jm - h.nside + ks + 1 + h.nl4 + h.nl4;
        i64 ip = (q >> 1) % h.nl4;
        return h.ncap + (ir - 1) * h.nl4 + ip;
    }
    double tp = tt - (double)(i64)tt;
    double tmp = ((za < 0.99) || !have_sth) ? (double)h.nside * sqrt(3.0 * (1.0 - za))
                                            : (double)h.nside * sth / sqrt((1.0 + za) / 3.0);
    i64 jp = (i64)(tp * tmp), jm = (i64)((1.0 - tp) * tmp);
    i64 ir = jp + jm + 1;
    i64 ip = (i64)(tt * (double)ir);
    return (z > 0) ? 2 * ir * (ir - 1) + ip : h.npix - 2 * ir * (ir + 1) + ip;
}

// RING <-> NEST (T_Healpix_Base ring2xyf / xyf2ring / xyf2nest / nest2xyf; nside a power of two).  The runners work on
// RING maps like the reference (utils/io.py:302); NEST exists for callers that hold nested maps or shard by base face.
__host__ __device__ __forceinline__ i64 spread_bits(i64 v) {      // bit k -> bit 2k (v < 2^31)
    v = (v | (v << 16)) & 0x0000ffff0000ffffLL;
    v = (v | (v << 8)) & 0x00ff00ff00ff00ffLL;
    v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0fLL;
    v = (v | (v << 2)) & 0x3333333333333333LL;
    v = (v | (v << 1)) & 0x5555555555555555LL;
    return v;
}
__host__ __device__ __forceinline__ i64 compress_bits(i64 v) {    // bit 2k -> bit k
    v &= 0x5555555555555555LL;
    v = (v | (v >> 1)) & 0x3333333333333333LL;
    v = (v | (v >> 2)) & 0x0f0f0f0f0f0f0f0fLL;
    v = (v | (v >> 4)) & 0x00ff00ff00ff00ffLL;
    v = (v | (v >> 8)) & 0x0000ffff0000ffffLL;
    v = (v | (v >> 16)) & 0x00000000ffffffffLL;
    return v;
}

__host__ __device__ __forceinline__ i64 ring2nest(const Hpx &h, i64 pix) {
    const int jrll[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4}, jpll[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};
    const i64 n = h.nside, nl2 = 2 * n;
    i64 iring, iphi, kshift, nr;
    int face;
    if (pix < h.ncap) {
        iring = (1 + isqrt_i64(1 + 2 * pix)) >> 1;
        iphi = pix + 1 - 2 * iring * (iring - 1);
        kshift = 0; nr = iring;
        face = (int)((iphi - 1) / nr);
    } else if (pix < h.npix - h.ncap) {
        const i64 ip = pix - h.ncap, tmp = ip / h.nl4;
        iring = tmp + n;
        iphi = ip - tmp * h.nl4 + 1;
        kshift = (iring + n) & 1; nr = n;
        const i64 ire = tmp + 1, irm = nl2 + 1 - tmp;
        const i64 ifm = (iphi - ire / 2 + n - 1) / n, ifp = (iphi - irm / 2 + n - 1) / n;
        face = (int)((ifp == ifm) ? (ifp | 4) : ((ifp < ifm) ? ifp : (ifm + 8)));
    } else {
        const i64 ip = h.npix - pix;
        iring = (1 + isqrt_i64(2 * ip - 1)) >> 1;
        iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1));
        kshift = 0; nr = iring;
        iring = 2 * nl2 - iring;
        face = (int)(8 + (iphi - 1) / nr);
    }
    const i64 irt = iring - jrll[face] * n + 1;
    i64 ipt = 2 * iphi - jpll[face] * nr - kshift - 1;
    if (ipt >= nl2) ipt -= 8 * n;
    const i64 ix = (ipt - irt) >> 1, iy = (-ipt - irt) >> 1;
    return (i64)face * n * n + spread_bits(ix) + (spread_bits(iy) << 1);
}

__host__ __device__ __forceinline__ i64 nest2ring(const Hpx &h, i64 pix) {
    const int jrll[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4}, jpll[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};
    const i64 n = h.nside, npface = n * n;
    const int face = (int)(pix / npface);
    const i64 p = pix - (i64)face * npface;
    const i64 ix = compress_bits(p), iy = compress_bits(p >> 1);
    const i64 jr = jrll[face] * n - ix - iy - 1;
    i64 nr, n_before, kshift;
    if (jr < n) { nr = jr; n_before = 2 * nr * (nr - 1); kshift = 0; }
    else if (jr > 3 * n) { nr = h.nl4 - jr; n_before = h.npix - 2 * (nr + 1) * nr; kshift = 0; }
    else { nr = n; n_before = h.ncap + (jr - n) * h.nl4; kshift = (jr - n) & 1; }
    i64 jp = (jpll[face] * nr + ix - iy + 1 + kshift) / 2;
    if (jp > h.nl4) jp -= h.nl4; else if (jp < 1) jp += h.nl4;
    return n_before + jp - 1;
}

// ------------------------------------------------------------------------------------------------
// Table-driven log2 for the pixel loops (the CUDA libm log() costs ~60 instructions; this one ~18).
//   x = 2^e * m, m in [1,2);  idx = top 7 mantissa bits;  rc = 1/c_idx (c_idx = bucket centre), f = m*rc - 1,
//   |f| <= 2^-8;  log2(x) = e + lt[idx] + log2(1+f),  lt = -log2(rc) tabulated, log2(1+f) by a degree-5 series.
// Absolute error < 2e-15 (tests/test_gpu_parity.py::test_fast_log2).  Zero / denormal / inf / NaN / negative inputs
// (never a finite, in-table radius) return NaN.  tab = 128 x (rc, lt) in shared memory.
// ------------------------------------------------------------------------------------------------
// host: per-device global copy of the table (allocated and filled once); kernels stage it into shared memory
int get_log2_table(const double2 **d_tab);

__device__ __forceinline__ void load_log2_table(double2 *tab, const double2 *__restrict__ g_tab) {
    for (int i = threadIdx.x; i < BFG_LOG2_TAB; i += blockDim.x) tab[i] = g_tab[i];
}

// bit access for __host__ __device__ code (fast_log2 runs on the CPU in bfg_test_fast_log2_host)
__host__ __device__ __forceinline__ int f64_hi(double x) {
#ifdef __CUDA_ARCH__
    return __double2hiint(x);
#else
    long long b; memcpy(&b, &x, 8); return (int)(b >> 32);
#endif
}
__host__ __device__ __forceinline__ int f64_lo(double x) {
#ifdef __CUDA_ARCH__
    return __double2loint(x);
#else
    long long b; memcpy(&b, &x, 8); return (int)(b & 0xffffffffLL);
#endif
}
__host__ __device__ __forceinline__ double f64_from(int hi, int lo) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(hi, lo);
#else
    long long b = ((long long)hi << 32) | (long long)(unsigned int)lo; double x; memcpy(&x, &b, 8); return x;
#endif
}

__host__ __device__ __forceinline__ double fast_log2(double x, const double2 *__restrict__ tab) {
    const int hi = f64_hi(x);
    // zero / denormal / negative / inf / NaN: every caller turns log(0) = -inf, log(inf) and NaN alike into an
    // out-of-table read-out (NaN -> contribution 0), so one NaN stands for all of them
    if ((unsigned)(hi - 0x00100000) >= (unsigned)(0x7ff00000 - 0x00100000)) return BFG_QNAN;
    const int idx = (hi >> 13) & (BFG_LOG2_TAB - 1);
    const double m = f64_from((hi & 0x000fffff) | 0x3ff00000, f64_lo(x));
    const double2 t = tab[idx];
    const double f = fma(m, t.x, -1.0);
    // log2(1+f) = f * (c1 + f (c2 + f (c3 + f (c4 + f c5)))),  c_k = (-1)^(k+1) / (k ln 2)
    double p = fma(f, 0.28853900817779268, -0.36067376022224085);
    p = fma(f, p, 0.48089834696298783);
    p = fma(f, p, -0.72134752044448170);
    p = fma(f, p, 1.4426950408889634);
    // exponent as a double without I2F: 2^52 + 2^31 + (e + 1023) trick
    const double ed = f64_from(0x43300000, (hi >> 20) ^ 0x80000000) - 4503601774854144.0 - 1023.0;
    return fma(f, p, t.y) + ed;
}

// ------------------------------------------------------------------------------------------------
// Lean building blocks of the pixel / cell / particle loops (the loops sit on the FP64 pipe; everything here exists to
// keep non-FP64 instructions out of them): polynomial constants in the constant bank, shared memory addressed by
// 32-bit shared-window addresses, reciprocal square root without the denormal / inf slow path.
// ------------------------------------------------------------------------------------------------
static __constant__ double c_l2p[5] = {0.28853900817779268, -0.36067376022224085, 0.48089834696298783, -0.72134752044448170,
                                1.4426950408889634};
// (the host shadow of a __constant__ array does not carry its initialiser: host builds of row_at_r2 read this copy)
static const double c_l2p_host[5] = {0.28853900817779268, -0.36067376022224085, 0.48089834696298783, -0.72134752044448170,
                                1.4426950408889634};
#ifdef __CUDA_ARCH__
#define BFG_L2P(i) c_l2p[i]
#else
#define BFG_L2P(i) c_l2p_host[i]
#endif

// (host build: "shared-window addresses" are byte offsets into bfg_host_smem, the arena bfg_test_row_at_r2_host lays out, so
// that row_at_r2 -- the read-out of the default grid / particle / exact shell loops -- runs on the CPU from the same source)
static thread_local const char *bfg_host_smem = nullptr;      // host builds only (never referenced by device code)
__host__ __device__ __forceinline__ double2 lds_f64x2(unsigned addr) {
    double2 v;
#ifdef __CUDA_ARCH__
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
#else
    memcpy(&v, bfg_host_smem + addr, sizeof(v));
#endif
    return v;
}
__host__ __device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
#ifdef __CUDA_ARCH__
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
#else
    memcpy(&v, bfg_host_smem + addr, sizeof(v));
#endif
    return v;
}
// 1/sqrt(x) for positive normal x (full double precision: MUFU seed 2^-22, one cubic step); x = 0 -> NaN, never trapped
__device__ __forceinline__ double rsqrt_pos(double x) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double e = fma(-x, y0 * y0, 1.0);
    const double p = fma(e, 0.375, 0.5);
    return fma(p, y0 * e, y0);
}


// Per-halo constants of a read-out by SQUARED radius from a blended row with a uniform ln r axis:
//   cell coordinate u = log2(r^2) * uA + uB,  uA = 0.5 ln2 / step,  uB = (offset - r0) / step.
struct RowLookup {
    double uA, uB, uMax;   // uMax = NR - 1
    int nrm2;              // NR - 2
    unsigned row_s, l2_s;  // shared-window addresses of the blended row and of the log2 table
};

// Launder values the compiler could re-derive from kernel parameters through a warp shuffle (every lane holds the same
// value): otherwise ptxas rematerialises them INSIDE the inner loop (LDC + I2F + DMUL + the 6-instruction
// generic->shared conversion per element) instead of keeping them in registers.
__device__ __forceinline__ void launder(RowLookup &f) {
    f.nrm2 = __shfl_sync(0xffffffffu, f.nrm2, 0);
    f.row_s = __shfl_sync(0xffffffffu, f.row_s, 0);
    f.l2_s = __shfl_sync(0xffffffffu, f.l2_s, 0);
    f.uMax = __shfl_sync(0xffffffffu, f.uMax, 0);
    f.uA = __shfl_sync(0xffffffffu, f.uA, 0);
}

// Table value at squared radius r2: v0 + t (v1 - v0) in the cell of u (scipy's (1-t) v0 + t v1 up to round-off; a
// non-finite node makes the result non-finite either way).  ok = false outside [r0, r1] (scipy: fill_value = nan);
// r2 = 0, inf, NaN fall outside.  log2 is the table-driven one of fast_log2 without its input test.
__host__ __device__ __forceinline__ double row_at_r2(const RowLookup &f, double r2, bool &ok) {
    const int hi = f64_hi(r2);
    const double m = f64_from((hi & 0x000fffff) | 0x3ff00000, f64_lo(r2));
    const double2 t = lds_f64x2(f.l2_s + (((unsigned)hi >> 9) & 0x7f0u));
    const double fr = fma(m, t.x, -1.0);
    double p = fma(fr, BFG_L2P(0), BFG_L2P(1));
    p = fma(fr, p, BFG_L2P(2));
    p = fma(fr, p, BFG_L2P(3));
    p = fma(fr, p, BFG_L2P(4));
    const double ed = f64_from(0x43300000, (hi >> 20) ^ 0x80000000) - 4503601774855167.0;   // unbiased exponent
    const double l2 = fma(fr, p, t.y) + ed;
    const double uu = fma(l2, f.uA, f.uB);
#ifdef __CUDA_ARCH__
    int k = __double2int_rd(uu);
#else
    int k = (uu >= -2147483648.0 && uu <= 2147483647.0) ? (int)floor(uu) : (int)0x80000000;   // cvt.rmi.s32.f64 saturates; NaN -> 0
    if (uu != uu) k = 0;
    else if (uu > 2147483647.0) k = 2147483647;
#endif
    ok = true;
    if (__builtin_expect((unsigned)k > (unsigned)f.nrm2, 0)) {   // outside the table, or exactly on its last node
        ok = (uu == f.uMax);
        k = f.nrm2;
    }
    const double tt = uu - (double)k;
    const unsigned ra = f.row_s + ((unsigned)k << 3);
    const double v0 = lds_f64(ra);
    return fma(tt, lds_f64(ra + 8) - v0, v0);
}

// fp64 RED (no return value): RED.E.ADD.F64 on sm_100a
__host__ __device__ __forceinline__ void red_add(double *addr, double v) {
#ifdef __CUDA_ARCH__
    atomicAdd(addr, v);
#else
    *addr += v;      // host test entries are single-threaded
#endif
}

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ i64 warp_sum_i64(i64 v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace bfg
