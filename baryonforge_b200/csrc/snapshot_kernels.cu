// snapshot_kernels.cu -- particle-snapshot baryonification on sm_100a.
//
//   bfg_snap_build_cells : periodic cell list + cell-sorted particle copies; replaces the scipy KDTree built in
//                          DefaultRunnerSnapshot.__init__ (BaryonForge/Runners/SnapshotRunner.py:95-100)
//   k_snap_halos         : halo loop of BaryonifySnapshot.process (SnapshotRunner.py:217-260): ball query
//                          (d <= R_q, periodic, inclusive like query_ball_point), minimum-image separation
//                          (:103-158), table read-out, accumulate per-particle offsets
//   k_snap_apply         : add offsets, single wrap (:263-273), un-permute to the caller's particle order
//   k_deposit_ngp        : ParticleSnapshot.make_map == np.histogramdd NGP mass deposit (BaryonForge/utils/io.py:629-677)
//
// Particles are physically re-ordered by cell (counting sort), so that for a fixed (cx, cy) the run of z-cells a
// halo touches is ONE contiguous range of the sorted arrays: lanes read consecutive particles (coalesced 8-byte
// loads) and their REDs land on consecutive addresses of the cell-ordered offset array.
#include <algorithm>
#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "bfg_common.cuh"
#include "grid_common.cuh"

using namespace bfg;

namespace {

constexpr int SNAP_THREADS = 128;

__host__ __device__ __forceinline__ int cell_of(double x, double L, int nc) {
    int c = (int)floor(x / L * (double)nc);
    return min(max(c, 0), nc - 1);
}

// The caller's particles are read through an element stride: 1 for separate coordinate arrays, 4 for the reference's own
// layout -- ONE structured array of 32-byte records (M, x, y, z), utils/io.py:588 -- uploaded as raw bytes.
template <int NDIM>
__global__ void k_cell_count(i64 n, const double *__restrict__ x, const double *__restrict__ y,
                             const double *__restrict__ z, i64 stride, double L, int nc, int *__restrict__ cell_id,
                             unsigned long long *__restrict__ counts) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        int c = cell_of(x[i * stride], L, nc) * nc + cell_of(y[i * stride], L, nc);
        if (NDIM == 3) c = c * nc + cell_of(z[i * stride], L, nc);
        cell_id[i] = c;
        atomicAdd(counts + c, 1ULL);
    }
}

// Scatter stage of the counting sort.  A particle is moved as ONE 32-byte record {x, y, z, original index}: the scattered
// write then covers exactly one full DRAM sector (4 separate 8-byte writes to 4 arrays cost 4 partial sectors and ran at
// a fifth of the HBM rate).  k_unpack_records turns the records back into the SoA arrays with coalesced traffic.
template <int NDIM>
__global__ void k_cell_fill(i64 n, const double *__restrict__ x, const double *__restrict__ y,
                            const double *__restrict__ z, i64 stride, const int *__restrict__ cell_id,
                            const i64 *__restrict__ cell_start, unsigned long long *__restrict__ cursor,
                            double4 *__restrict__ rec) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        int c = cell_id[i];
        i64 slot = cell_start[c] + (i64)atomicAdd(cursor + c, 1ULL);
        double4 r;
        r.x = x[i * stride]; r.y = y[i * stride]; r.z = (NDIM == 3) ? z[i * stride] : 0.0; r.w = __longlong_as_double(i);
        rec[slot] = r;
    }
}

// Two-pass scatter (the default above 65536 cells; measured 26.2 vs 29.6 ms for 2.5e8 particles): with ~10^7 cells a single-pass scatter keeps
// ~10^7 partially written 128-byte lines open, far more than the L2 holds, so DRAM sees isolated 32-byte sector writes.
// Pass A scatters the records into <= 65536 coarse buckets (runs of `group` consecutive cells, bucket start =
// cell_start[bucket * group]): few enough write heads for the L2 to merge the four records of a line before it is evicted.
// Pass B reads the bucket-ordered records coalesced and places them in their cells; everything in flight then lands in a
// window of a few MB.
__global__ void k_cell_coarse(i64 n, const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
                              i64 stride, const int *__restrict__ cell_id, const i64 *__restrict__ cell_start, int group, i64 ncells,
                              unsigned long long *__restrict__ cursor_a, double4 *__restrict__ tmp) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const i64 b = cell_id[i] / group;
        const i64 slot = cell_start[min(b * group, ncells)] + (i64)atomicAdd(cursor_a + b, 1ULL);
        double4 r;
        r.x = x[i * stride]; r.y = y[i * stride]; r.z = z ? z[i * stride] : 0.0; r.w = __longlong_as_double(i);
        tmp[slot] = r;
    }
}

template <int NDIM>
__global__ void k_cell_fine(i64 n, const double4 *__restrict__ tmp, double L, int nc, const i64 *__restrict__ cell_start,
                            unsigned long long *__restrict__ cursor, double4 *__restrict__ rec) {
    for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (i64)gridDim.x * blockDim.x) {
        const double4 r = tmp[p];
        int c = cell_of(r.x, L, nc) * nc + cell_of(r.y, L, nc);                 // the cell k_cell_count assigned
        if (NDIM == 3) c = c * nc + cell_of(r.z, L, nc);
        rec[cell_start[c] + (i64)atomicAdd(cursor + c, 1ULL)] = r;
    }
}

// Index sort (BFG_CELL_SORT=3): the particle INDICES are radix-sorted by cell (8 bytes per particle and pass instead of a
// 32-byte record), then every sorted slot gathers its particle.  Random 32-byte READS run near the sector rate of the HBM,
// random 32-byte WRITES into 10^7 open lines do not; with the caller's particles held as 32-byte records (stride 4) the gather
// is one sector per particle.  The sort is stable, so the order inside a cell is the caller's (deterministic cell lists).
__global__ void k_iota_u32(i64 n, unsigned int *__restrict__ idx) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) idx[i] = (unsigned int)i;
}

template <int NDIM>
__global__ void k_cell_gather(i64 n, const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
                              i64 stride, const unsigned int *__restrict__ sorted_idx, i64 *__restrict__ order,
                              double *__restrict__ xs, double *__restrict__ ys, double *__restrict__ zs) {
    for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (i64)gridDim.x * blockDim.x) {
        const i64 i = (i64)sorted_idx[p];
        xs[p] = x[i * stride]; ys[p] = y[i * stride];
        if (NDIM == 3) zs[p] = z[i * stride];
        order[p] = i;
    }
}

template <int NDIM>
__global__ void k_unpack_records(i64 n, const double4 *__restrict__ rec, i64 *__restrict__ order,
                                 double *__restrict__ xs, double *__restrict__ ys, double *__restrict__ zs) {
    for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (i64)gridDim.x * blockDim.x) {
        const double4 r = rec[p];
        xs[p] = r.x; ys[p] = r.y;
        if (NDIM == 3) zs[p] = r.z;
        if (order) order[p] = __double_as_longlong(r.w);
    }
}

// SnapshotRunner.py:135-158 enforce_periodicity
__device__ __forceinline__ double min_image(double dx, double L) {
    if (dx > 0.5 * L) dx -= L;
    if (dx < -0.5 * L) dx += L;
    return dx;
}

struct HaloSnap {
    double c[3], rq, lnz, lnM, rcut, lnRcom;
};

template <bool UNIFORM, int NDIM>
__global__ void __launch_bounds__(SNAP_THREADS)
k_snap_halos(TableView T, double L, int nc, const double *__restrict__ xs, const double *__restrict__ ys,
             const double *__restrict__ zs, const i64 *__restrict__ cell_start, i64 n_part, i64 n_halo,
             const double *__restrict__ halos, const double *__restrict__ extras, int n_extra,
             double *__restrict__ tot, unsigned long long *npairs, const double2 *__restrict__ g_l2tab) {
    extern __shared__ double row[];
    __shared__ double2 l2tab[BFG_LOG2_TAB];
    load_log2_table(l2tab, g_l2tab);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = SNAP_THREADS / 32;
    const double cell = L / (double)nc;
    i64 done = 0;
    for (i64 h = blockIdx.x; h < n_halo; h += gridDim.x) {
        const double *H = halos + h * BFG_HALO_STRIDE;
        HaloSnap s;
        s.c[0] = __ldg(H + BFG_HB_X); s.c[1] = __ldg(H + BFG_HB_Y); s.c[2] = __ldg(H + BFG_HB_Z);
        s.rq = __ldg(H + BFG_HB_RQ); s.lnz = __ldg(H + BFG_HB_LNZ); s.lnM = __ldg(H + BFG_HB_LNM);
        s.rcut = __ldg(H + BFG_HB_RCUT); s.lnRcom = __ldg(H + BFG_HB_LNRCOM);
        __syncthreads();
        bool valid;
        blend_row(T, s.lnz, s.lnM, extras ? extras + h * n_extra : nullptr, row, valid);
        __syncthreads();
        const double rq2 = s.rq * s.rq, rcut2 = s.rcut * s.rcut;
        RowLookup rl;
        if (UNIFORM) {   // lean read-out: cell coordinate u = log2(d^2) * uA + uB (row_at_r2)
            rl.uA = 0.34657359027997264 * T.inv_dr;
            rl.uB = (((T.flags & BFG_TABLE_RDELTA) ? -s.lnRcom : 0.0) - T.r0) * T.inv_dr;
            rl.uMax = (double)(T.n[2] - 1);
            rl.nrm2 = T.n[2] - 2;
            rl.row_s = (unsigned)__cvta_generic_to_shared(row);
            rl.l2_s = (unsigned)__cvta_generic_to_shared(l2tab);
            launder(rl);
        }
        // cells covering [c - rq, c + rq] per axis (periodic); never more than nc of them
        int lo[3], cnt[3];
        bool wraps = false;       // does the ball (plus one cell of slack) reach across the box boundary?
        for (int d = 0; d < 3; ++d) {
            if (d >= NDIM) { lo[d] = 0; cnt[d] = 1; continue; }
            int a = (int)floor((s.c[d] - s.rq) / cell), b = (int)floor((s.c[d] + s.rq) / cell);
            int c = b - a + 1;
            if (c >= nc) { a = 0; c = nc; }
            lo[d] = a; cnt[d] = c;
            wraps = wraps || !(s.c[d] - s.rq > cell) || !(s.c[d] + s.rq < L - cell);
        }
        const int last = NDIM - 1;                 // fastest-varying cell axis: runs along it are contiguous
        const int nouter = (NDIM == 3) ? cnt[0] * cnt[1] : cnt[0];
        for (int oo = warp; oo < nouter; oo += NW) {
            const int cx = (NDIM == 3) ? oo / cnt[1] : oo;
            const int cy = (NDIM == 3) ? oo - cx * cnt[1] : 0;
            // Column culling: the (cx[, cy]) column of cells is a rectangle in the leading axes; if it lies further than
            // R_q from the halo the column is skipped, otherwise only the chord of the ball along the last axis is walked.
            // (Unwrapped cell coordinates; a relative slack keeps the test conservative against round-off.)
            int zl = lo[last], zc = cnt[last];
            if (cnt[last] < nc) {
                double dmin2 = 0.0;
                for (int d = 0; d < last; ++d) {
                    if (cnt[d] >= nc) continue;                      // the whole axis is covered: no constraint from it
                    const double a0 = (double)(lo[d] + (d == 0 ? cx : cy)) * cell;
                    const double gap = fmax(fmax(a0 - s.c[d], s.c[d] - (a0 + cell)) - 1e-9 * cell, 0.0);
                    dmin2 = fma(gap, gap, dmin2);
                }
                const double rem = rq2 * (1.0 + 1e-12) - dmin2;
                if (rem < 0.0) continue;
                const double rz = sqrt(rem) + 1e-9 * cell;
                const int za = max((int)floor((s.c[last] - rz) / cell), lo[last]);
                const int zb = min((int)floor((s.c[last] + rz) / cell), lo[last] + cnt[last] - 1);
                if (zb < za) continue;
                zl = za; zc = zb - za + 1;
            }
            const int gx = (((lo[0] + cx) % nc) + nc) % nc;
            i64 cbase;
            if (NDIM == 3) {
                const int gy = (((lo[1] + cy) % nc) + nc) % nc;
                cbase = ((i64)gx * nc + gy) * nc;
            } else {
                cbase = (i64)gx * nc;                  // 2-D: the "last axis" is y; runs are over y for fixed x
            }
            // the run along the last axis, split where it wraps around the box
            const int z0 = ((zl % nc) + nc) % nc;
            const int len0 = min(zc, nc - z0), len1 = zc - len0;
            for (int seg = 0; seg < 2; ++seg) {
                const int zlo = seg ? 0 : z0, zlen = seg ? len1 : len0;
                if (zlen <= 0) continue;
                const i64 p0 = cell_start[cbase + zlo], p1 = cell_start[cbase + zlo + zlen];
                for (i64 p = p0 + lane; p < p1; p += 32) {
                    double dx = xs[p] - s.c[0], dy = ys[p] - s.c[1], dz = (NDIM == 3) ? zs[p] - s.c[2] : 0.0;
                    if (wraps) {                                   // SnapshotRunner.py:248-251 / :103-132 minimum image
                        dx = min_image(dx, L); dy = min_image(dy, L);
                        if (NDIM == 3) dz = min_image(dz, L);
                    }
                    const double d2 = (NDIM == 3) ? dx * dx + dy * dy + dz * dz : dx * dx + dy * dy;
                    if (!(d2 <= rq2)) continue;                    // query_ball_point: d <= R_q, inclusive  (:232/:247)
                    ++done;
                    double val;
                    if (UNIFORM) {
                        bool ok;
                        val = row_at_r2(rl, d2, ok);
                        if (!ok) val = CUDART_NAN;
                    } else {
                        double xq = fast_log2(d2, l2tab) * 0.34657359027997264;   // ln d = 0.5 ln2 log2(d^2)
                        if (T.flags & BFG_TABLE_RDELTA) xq -= s.lnRcom;
                        val = row_lookup<false>(T, row, xq);
                    }
                    if (!valid) val = CUDART_NAN;
                    val = (d2 < rcut2) ? val : 0.0;                // BaryonCorrection.py:410-411
                    if (!isfinite(val)) val = 0.0;                 // SnapshotRunner.py:259
                    if (val == 0.0 && d2 > 0.0) continue;          // adds exact zeros
                    const double sc = val * rsqrt_pos(d2);         // d == 0 -> NaN, as in the reference's 0/0 (§10 #11)
                    red_add(tot + p, sc * dx);                     // :260
                    red_add(tot + n_part + p, sc * dy);
                    if (NDIM == 3) red_add(tot + 2 * n_part + p, sc * dz);
                }
            }
        }
    }
    if (npairs) {
        done = warp_sum_i64(done);
        if (lane == 0 && done) atomicAdd(npairs, (unsigned long long)done);
    }
}

__host__ __device__ __forceinline__ double wrap_once(double q, double L) {   // SnapshotRunner.py:272-273
    if (q > L) q -= L;
    if (q < 0) q += L;
    return q;
}

// Displaced position of every particle, scattered back to the caller's order as one full-sector 32-byte record; then
// k_unpack_records writes the three output arrays coalesced.
template <int NDIM>
__global__ void k_snap_apply(i64 n, const double *__restrict__ xs, const double *__restrict__ ys,
                             const double *__restrict__ zs, const double *__restrict__ tot,
                             const i64 *__restrict__ order, double L, double4 *__restrict__ rec) {
    for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (i64)gridDim.x * blockDim.x) {
        double4 r;
        r.x = wrap_once(xs[p] + tot[p], L);
        r.y = wrap_once(ys[p] + tot[n + p], L);
        r.z = (NDIM == 3) ? wrap_once(zs[p] + tot[2 * n + p], L) : 0.0;
        r.w = 0.0;
        rec[order[p]] = r;
    }
}

__device__ __forceinline__ void put_field(double4 &r, int f, double v) {
    r.x = (f == 0) ? v : r.x; r.y = (f == 1) ? v : r.y; r.z = (f == 2) ? v : r.z; r.w = (f == 3) ? v : r.w;
}

// k_snap_apply for callers that hold the particles as 32-byte records (the reference's structured array: M, x, y, z in any
// field order): `new_cat = copy of cat with x, y(, z) replaced` (SnapshotRunner.py:263-273) as ONE full-sector read and ONE
// full-sector write per particle, in place or into a second record array -- no column arrays, no unpack pass.
template <int NDIM>
__global__ void k_snap_apply_records(i64 n, const double *__restrict__ xs, const double *__restrict__ ys,
                                     const double *__restrict__ zs, const double *__restrict__ tot,
                                     const i64 *__restrict__ order, double L, const double4 *rec_in, double4 *rec_out,
                                     int fx, int fy, int fz) {
    for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (i64)gridDim.x * blockDim.x) {
        const i64 j = order[p];
        double4 r = rec_in[j];
        put_field(r, fx, wrap_once(xs[p] + tot[p], L));
        put_field(r, fy, wrap_once(ys[p] + tot[n + p], L));
        if (NDIM == 3) put_field(r, fz, wrap_once(zs[p] + tot[2 * n + p], L));
        rec_out[j] = r;
    }
}

// np.histogramdd bin of x on edges = np.linspace(0, L, N+1): searchsorted(side='right') - 1, x == L -> last bin
__host__ __device__ __forceinline__ i64 ngp_bin(double x, double L, i64 N, double step) {
    if (!(x >= 0.0) || !(x <= L)) return -1;
    if (x == L) return N - 1;
    i64 i = (i64)(x / step);
    if (i > N - 1) i = N - 1;
    // edge(i) = i*step as np.linspace computes it (arange*step), last edge exactly L
    while (i > 0 && x < (double)i * step) --i;
    while (i + 1 < N && x >= (double)(i + 1) * step) ++i;
    return i;
}

template <int NDIM>
__global__ void k_deposit_ngp(i64 n, const double *__restrict__ x, const double *__restrict__ y,
                              const double *__restrict__ z, const double *__restrict__ m, double L, i64 N,
                              double *__restrict__ grid) {
    const double step = L / (double)N;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        i64 bx = ngp_bin(x[i], L, N, step), by = ngp_bin(y[i], L, N, step);
        i64 bz = (NDIM == 3) ? ngp_bin(z[i], L, N, step) : 0;
        if (bx < 0 || by < 0 || bz < 0) continue;
        i64 c = (NDIM == 3) ? (bx * N + by) * N + bz : bx * N + by;
        red_add(grid + c, m[i]);
    }
}

// k_snap_apply + k_deposit_ngp in one pass over the CELL-ORDERED particles, for callers that only need the deposited grid
// (BaryonifySnapshot.process() followed by ParticleSnapshot.make_map, utils/io.py:629-677): the displaced positions are
// never scattered back to the caller's order (the un-permute is a random 32-byte write per particle), and lanes that walk
// the particles of one cell-list cell deposit into a handful of neighbouring grid cells.
template <int NDIM>
__global__ void k_snap_apply_deposit(i64 n, const double *__restrict__ xs, const double *__restrict__ ys,
                                     const double *__restrict__ zs, const double *__restrict__ tot,
                                     const i64 *__restrict__ order, const double *__restrict__ mass, double mass_const,
                                     double L, i64 N, double *__restrict__ grid) {
    const double step = L / (double)N;
    for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (i64)gridDim.x * blockDim.x) {
        const i64 bx = ngp_bin(wrap_once(xs[p] + tot[p], L), L, N, step);
        const i64 by = ngp_bin(wrap_once(ys[p] + tot[n + p], L), L, N, step);
        const i64 bz = (NDIM == 3) ? ngp_bin(wrap_once(zs[p] + tot[2 * n + p], L), L, N, step) : 0;
        if (bx < 0 || by < 0 || bz < 0) continue;
        const i64 c = (NDIM == 3) ? (bx * N + by) * N + bz : bx * N + by;
        red_add(grid + c, mass ? mass[order[p]] : mass_const);
    }
}

int blocks_for(i64 n, int threads) { return (int)std::max<i64>(1, std::min<i64>((n + threads - 1) / threads, 148 * 32)); }

}  // namespace

extern "C" int bfg_snap_build_cells(int ndim, int64_t n_part, const double *d_x, const double *d_y, const double *d_z,
                                    double L, int ncell, int64_t *d_cell_start, int64_t *d_order, double *d_xs,
                                    double *d_ys, double *d_zs, void *stream) {
    BFG_ENTRY();
    return bfg_snap_build_cells_strided(ndim, n_part, d_x, d_y, d_z, 1, L, ncell, d_cell_start, d_order, d_xs, d_ys, d_zs, stream);
}

extern "C" int bfg_snap_build_cells_strided(int ndim, int64_t n_part, const double *d_x, const double *d_y, const double *d_z,
                                            int64_t stride, double L, int ncell, int64_t *d_cell_start, int64_t *d_order,
                                            double *d_xs, double *d_ys, double *d_zs, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(ndim == 2 || ndim == 3, "ndim must be 2 or 3");
    BFG_REQUIRE(stride >= 1, "stride must be >= 1");
    BFG_REQUIRE(d_x && d_y && (ndim == 2 || d_z) && d_cell_start && d_order && d_xs && d_ys && (ndim == 2 || d_zs), "null argument");
    BFG_REQUIRE(ncell >= 1 && (ndim == 2 ? ncell <= 32768 : ncell <= 1024), "ncell out of range");
    BFG_REQUIRE(L > 0, "L must be positive");
    if (int rc = retain_async_pool()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const i64 ncells = (ndim == 3) ? (i64)ncell * ncell * ncell : (i64)ncell * ncell;
    StreamScratch s_cell(st), s_counts(st), s_scan(st), s_rec(st), s_tmp(st), s_cursor(st);   // released on every return path
    size_t scan_bytes = 0;
    BFG_CUDA_OK(s_cell.alloc(sizeof(int) * std::max<i64>(n_part, 1)));
    BFG_CUDA_OK(s_counts.alloc(sizeof(unsigned long long) * (ncells + 1)));
    int *cell_id = s_cell.as<int>();
    unsigned long long *counts = s_counts.as<unsigned long long>();
    BFG_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * (ncells + 1), st));
    if (n_part > 0) {
        if (ndim == 3) k_cell_count<3><<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, d_x, d_y, d_z, stride, L, ncell, cell_id, counts);
        else k_cell_count<2><<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, d_x, d_y, d_z, stride, L, ncell, cell_id, counts);
        BFG_CUDA_OK(cudaGetLastError());
    }
    BFG_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const i64 *)counts, (i64 *)d_cell_start, ncells + 1, st));
    BFG_CUDA_OK(s_scan.alloc(scan_bytes));
    BFG_CUDA_OK(cub::DeviceScan::ExclusiveSum(s_scan.p, scan_bytes, (const i64 *)counts, (i64 *)d_cell_start, ncells + 1, st));
    BFG_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * (ncells + 1), st));
    const char *mode0 = getenv("BFG_CELL_SORT");
    if (n_part > 0 && mode0 && mode0[0] == '3' && n_part < ((i64)1 << 31)) {
        // index sort: (cell id, particle index) pairs through cub::DeviceRadixSort (library call, like bfg_halo_sort), then gather
        StreamScratch s_key2(st), s_idx(st), s_idx2(st), s_sort(st);
        BFG_CUDA_OK(s_key2.alloc(sizeof(unsigned int) * n_part));
        BFG_CUDA_OK(s_idx.alloc(sizeof(unsigned int) * n_part));
        BFG_CUDA_OK(s_idx2.alloc(sizeof(unsigned int) * n_part));
        k_iota_u32<<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, s_idx.as<unsigned int>());
        int end_bit = 1;
        while (end_bit < 32 && ((i64)1 << end_bit) < ncells) ++end_bit;
        size_t sort_bytes = 0;
        BFG_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned int *)cell_id, s_key2.as<unsigned int>(),
                                                    s_idx.as<unsigned int>(), s_idx2.as<unsigned int>(), (int)n_part, 0, end_bit, st));
        BFG_CUDA_OK(s_sort.alloc(sort_bytes));
        BFG_CUDA_OK(cub::DeviceRadixSort::SortPairs(s_sort.p, sort_bytes, (const unsigned int *)cell_id, s_key2.as<unsigned int>(),
                                                    s_idx.as<unsigned int>(), s_idx2.as<unsigned int>(), (int)n_part, 0, end_bit, st));
        if (ndim == 3) k_cell_gather<3><<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, d_x, d_y, d_z, stride, s_idx2.as<unsigned int>(),
                                                                               (i64 *)d_order, d_xs, d_ys, d_zs);
        else k_cell_gather<2><<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, d_x, d_y, d_z, stride, s_idx2.as<unsigned int>(),
                                                                     (i64 *)d_order, d_xs, d_ys, d_zs);
        BFG_CUDA_OK(cudaGetLastError());
        return BFG_OK;
    }
    if (n_part > 0) {
        BFG_CUDA_OK(s_rec.alloc(sizeof(double4) * n_part));   // 32-byte particle records
        double4 *rec = s_rec.as<double4>();
        const char *mode = getenv("BFG_CELL_SORT");      // 1 = single-pass scatter (A/B); default: two passes above 65536 cells
        const bool two_pass = !(mode && mode[0] == '1') && ncells > 65536;
        if (two_pass) {
            const int group = (int)std::max<i64>(ncell, (ncells + 65535) / 65536);
            const i64 n_buckets = (ncells + group - 1) / group;
            BFG_CUDA_OK(s_tmp.alloc(sizeof(double4) * n_part));
            BFG_CUDA_OK(s_cursor.alloc(sizeof(unsigned long long) * n_buckets));
            double4 *tmp = s_tmp.as<double4>();
            unsigned long long *cursor_a = s_cursor.as<unsigned long long>();
            BFG_CUDA_OK(cudaMemsetAsync(cursor_a, 0, sizeof(unsigned long long) * n_buckets, st));
            k_cell_coarse<<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, d_x, d_y, ndim == 3 ? d_z : nullptr, stride, cell_id,
                                                                   (const i64 *)d_cell_start, group, ncells, cursor_a, tmp);
            if (ndim == 3) k_cell_fine<3><<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, tmp, L, ncell, (const i64 *)d_cell_start, counts, rec);
            else k_cell_fine<2><<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, tmp, L, ncell, (const i64 *)d_cell_start, counts, rec);
        } else if (ndim == 3) {
            k_cell_fill<3><<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, d_x, d_y, d_z, stride, cell_id, (const i64 *)d_cell_start, counts, rec);
        } else {
            k_cell_fill<2><<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, d_x, d_y, d_z, stride, cell_id, (const i64 *)d_cell_start, counts, rec);
        }
        if (ndim == 3) k_unpack_records<3><<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, rec, (i64 *)d_order, d_xs, d_ys, d_zs);
        else k_unpack_records<2><<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, rec, (i64 *)d_order, d_xs, d_ys, d_zs);
        BFG_CUDA_OK(cudaGetLastError());
    }
    return BFG_OK;
}

extern "C" int bfg_snap_offsets(const bfg_table *t, int ndim, int64_t n_part, const double *d_xs, const double *d_ys,
                                const double *d_zs, double L, int ncell, const int64_t *d_cell_start, int64_t n_halo,
                                const double *d_halos, const double *d_extras, int n_extra, double *d_tot,
                                int64_t *d_npairs, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(t && d_xs && d_ys && (ndim == 2 || d_zs) && d_cell_start && d_tot && (d_halos || n_halo == 0), "null argument");
    BFG_REQUIRE(ndim == 2 || ndim == 3, "ndim must be 2 or 3");
    BFG_REQUIRE(n_extra == t->view.ndim - 3, "n_extra must equal the table's extra axes");
    BFG_REQUIRE(n_extra == 0 || d_extras, "extras missing");
    BFG_REQUIRE((t->view.flags & BFG_TABLE_LOG_VALUES) == 0, "snapshot baryonification needs a displacement table");
    cudaStream_t st = (cudaStream_t)stream;
    if (d_npairs) BFG_CUDA_OK(cudaMemsetAsync(d_npairs, 0, sizeof(i64), st));
    if (n_halo == 0 || n_part == 0) return BFG_OK;
    size_t smem = sizeof(double) * t->view.n[2];
    BFG_REQUIRE(smem <= 200 * 1024, "radial axis too long for the shared-memory row (max 25600 nodes)");
    int blocks = (int)std::min<i64>(n_halo, (i64)1 << 30);
    const double2 *g_l2tab = nullptr;
    if (int rc = get_log2_table(&g_l2tab)) return rc;
    auto go = [&](auto kern) -> int {
        BFG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, SNAP_THREADS, smem, st>>>(t->view, L, ncell, d_xs, d_ys, d_zs, (const i64 *)d_cell_start, n_part, n_halo,
                                                 d_halos, d_extras, n_extra, d_tot, (unsigned long long *)d_npairs, g_l2tab);
        BFG_CUDA_OK(cudaGetLastError());
        return BFG_OK;
    };
    const bool u = t->view.uniform_r != 0;
    if (ndim == 3) return u ? go(k_snap_halos<true, 3>) : go(k_snap_halos<false, 3>);
    return u ? go(k_snap_halos<true, 2>) : go(k_snap_halos<false, 2>);
}

extern "C" int bfg_snap_apply(int ndim, int64_t n_part, const double *d_xs, const double *d_ys, const double *d_zs,
                              const double *d_tot, const int64_t *d_order, double L, double *d_x_out, double *d_y_out,
                              double *d_z_out, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(ndim == 2 || ndim == 3, "ndim must be 2 or 3");
    BFG_REQUIRE(d_xs && d_ys && d_tot && d_order && d_x_out && d_y_out && (ndim == 2 || (d_zs && d_z_out)), "null argument");
    if (n_part == 0) return BFG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = retain_async_pool()) return rc;
    StreamScratch s_rec(st);
    BFG_CUDA_OK(s_rec.alloc(sizeof(double4) * n_part));
    double4 *rec = s_rec.as<double4>();
    if (ndim == 3) {
        k_snap_apply<3><<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, d_xs, d_ys, d_zs, d_tot, (const i64 *)d_order, L, rec);
        k_unpack_records<3><<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, rec, nullptr, d_x_out, d_y_out, d_z_out);
    } else {
        k_snap_apply<2><<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, d_xs, d_ys, d_zs, d_tot, (const i64 *)d_order, L, rec);
        k_unpack_records<2><<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, rec, nullptr, d_x_out, d_y_out, d_z_out);
    }
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_snap_apply_records(int ndim, int64_t n_part, const double *d_xs, const double *d_ys, const double *d_zs,
                                      const double *d_tot, const int64_t *d_order, double L, const double *d_rec_in,
                                      double *d_rec_out, int fx, int fy, int fz, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(ndim == 2 || ndim == 3, "ndim must be 2 or 3");
    BFG_REQUIRE(d_xs && d_ys && d_tot && d_order && d_rec_in && d_rec_out && (ndim == 2 || d_zs), "null argument");
    BFG_REQUIRE((((uintptr_t)d_rec_in | (uintptr_t)d_rec_out) & 31) == 0, "records must be 32-byte aligned");
    BFG_REQUIRE(fx >= 0 && fx < 4 && fy >= 0 && fy < 4 && fx != fy, "bad field slots");
    BFG_REQUIRE(ndim == 2 || (fz >= 0 && fz < 4 && fz != fx && fz != fy), "bad field slots");
    if (n_part == 0) return BFG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (ndim == 3)
        k_snap_apply_records<3><<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, d_xs, d_ys, d_zs, d_tot, (const i64 *)d_order, L,
                                                                         (const double4 *)d_rec_in, (double4 *)d_rec_out, fx, fy, fz);
    else
        k_snap_apply_records<2><<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, d_xs, d_ys, d_zs, d_tot, (const i64 *)d_order, L,
                                                                         (const double4 *)d_rec_in, (double4 *)d_rec_out, fx, fy, fz);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_snap_deposit_ngp(int ndim, int64_t n_part, const double *d_x, const double *d_y, const double *d_z,
                                    const double *d_mass, double L, int64_t n_grid, double *d_grid, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(ndim == 2 || ndim == 3, "ndim must be 2 or 3");
    BFG_REQUIRE(d_x && d_y && (ndim == 2 || d_z) && d_mass && d_grid, "null argument");
    BFG_REQUIRE(n_grid >= 1 && L > 0, "bad grid");
    if (n_part == 0) return BFG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (ndim == 3) k_deposit_ngp<3><<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, d_x, d_y, d_z, d_mass, L, n_grid, d_grid);
    else k_deposit_ngp<2><<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, d_x, d_y, d_z, d_mass, L, n_grid, d_grid);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_snap_apply_deposit(int ndim, int64_t n_part, const double *d_xs, const double *d_ys, const double *d_zs,
                                      const double *d_tot, const int64_t *d_order, const double *d_mass, double mass_const,
                                      double L, int64_t n_grid, double *d_grid, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(ndim == 2 || ndim == 3, "ndim must be 2 or 3");
    BFG_REQUIRE(n_grid >= 1 && L > 0, "bad grid");
    if (n_part == 0) return BFG_OK;
    BFG_REQUIRE(d_xs && d_ys && (ndim == 2 || d_zs) && d_tot && d_grid, "null argument");
    BFG_REQUIRE(!d_mass || d_order, "per-particle masses are in the caller's order: d_order is needed");
    cudaStream_t st = (cudaStream_t)stream;
    if (ndim == 3)
        k_snap_apply_deposit<3><<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, d_xs, d_ys, d_zs, d_tot, (const i64 *)d_order,
                                                                         d_mass, mass_const, L, n_grid, d_grid);
    else
        k_snap_apply_deposit<2><<<blocks_for(n_part, 256), 256, 0, st>>>(n_part, d_xs, d_ys, d_zs, d_tot, (const i64 *)d_order,
                                                                         d_mass, mass_const, L, n_grid, d_grid);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

// ---------------------------------------------------------------------------------------------------- host test entry
// Pure host, no GPU: the index helpers the grid and particle kernels inline, on the CPU.
//   what = 0  NGP cell of a coordinate (ngp_bin: np.histogramdd on np.linspace(0, L, N + 1) edges, utils/io.py:629-677): h_x [n] -> h_out_i [n]
//   what = 1  wrap_once (SnapshotRunner.py:272-273): h_x [n] -> h_out_d [n]
//   what = 2  cell-list cell of a coordinate (cell_of): h_x [n], N = cells per axis -> h_out_i [n]
//   what = 3  cutout of a halo along one axis (Map2DRunner.py:400-429, :500-528): n = Nsize, L = res, centre = (int)h_x[0], N = cells per
//             axis -> h_out_d [n] = np.linspace(-n/2, n/2, n) * res, h_out_i [n] = pick_indices(centre, n / 2, N)
extern "C" int bfg_test_index_helpers_host(int what, int64_t n, const double *h_x, double L, int64_t N, int64_t *h_out_i,
                                           double *h_out_d) {
    BFG_REQUIRE(n >= 0 && N >= 1 && what >= 0 && what <= 3, "bad argument");
    BFG_REQUIRE(n == 0 || h_x, "null argument");
    if (what == 3) {
        BFG_REQUIRE(n >= 2 && h_out_i && h_out_d && N <= 2147483647LL, "bad cutout");
        const int ns = (int)n, cen = (int)h_x[0];
        for (int i = 0; i < ns; ++i) {
            h_out_d[i] = cut_coord(i, ns, L);
            h_out_i[i] = wrap_idx(cen - ns / 2 + i, (int)N);
        }
        return BFG_OK;
    }
    const double step = L / (double)N;
    for (int64_t i = 0; i < n; ++i) {
        if (what == 0) h_out_i[i] = ngp_bin(h_x[i], L, N, step);
        else if (what == 1) h_out_d[i] = wrap_once(h_x[i], L);
        else h_out_i[i] = cell_of(h_x[i], L, (int)N);
    }
    return BFG_OK;
}
