// grid_kernels.cu -- periodic 2-D / 3-D grid runners on sm_100a.
//
//   k_grid_halos<PAINT>  : halo loop of BaryonifyGrid.process     (BaryonForge/Runners/Map2DRunner.py:482-586)
//                          and PaintProfilesGrid.process          (BaryonForge/Runners/Map2DRunner.py:725-821)
//   k_grid_regrid        : re-binning                             (BaryonForge/Runners/Map2DRunner.py:589-613, :13-162)
//
// One CTA per halo; the cutout is walked with the last array axis fastest so a warp's REDs are contiguous.
// The reference's coordinate conventions are kept literally:
//   * cutout coordinates  x[i] = linspace(-Nsize/2, Nsize/2, Nsize)[i] * res   (stretched by Nsize/(Nsize-1))
//   * meshgrid(indexing='xy'): element (i, j, k) of the cutout -- array axes 0, 1, 2 -- has
//       gx = x[j] + dx,  gy = x[i] + dy,  gz = x[k] + dz        with dx = bins[x_cen] - x_j, ...
//     and component 0 of the offset is gx/r, component 1 is gy/r.
#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include "bfg_common.cuh"
#include "grid_common.cuh"

using namespace bfg;

namespace {

constexpr int GRID_THREADS = 128;

// ELL (2-D only): the last 4 columns of `extras` hold the halo's shear matrix Rmat (Map2DRunner.py:281-350, build_Rmat);
// the radius handed to the table is |(gx, gy) @ Rmat| while the direction stays (gx, gy)/r   (:531-536, :769-774).
constexpr int MODE_BARYONIFY = 0;   // BaryonifyGrid          Map2DRunner.py:482-586
constexpr int MODE_PAINT = 1;       // PaintProfilesGrid      Map2DRunner.py:725-821
constexpr int MODE_ANIS = 2;        // PaintProfilesAnisGrid  Map2DRunner.py:895-1001 (2-D maps only, :847)

struct AnisArgs {
    TableView T2;            // Tracer_model.projected table (log values)
    const double *mtot;      // halo part of Mtot_map, owned planes (Map2DRunner.py:866-871)
    const double *orig;      // GriddedMap.map, owned planes
    double mtot_add;         // dV * drho_m (:888)
};

template <int MODE, bool UNIFORM, int NDIM, bool ELL>
__global__ void __launch_bounds__(GRID_THREADS)
k_grid_halos(TableView T, int N, double res, double scale, i64 n_halo, const double *__restrict__ halos,
             const double *__restrict__ extras, int n_extra, double *__restrict__ out, int plane_lo, int plane_hi,
             unsigned long long *nupd, const double2 *__restrict__ g_l2tab, AnisArgs A) {
    constexpr bool PAINT = (MODE != MODE_BARYONIFY);
    extern __shared__ double row[];
    const double *trow = row + ((MODE == MODE_ANIS) ? T.n[2] : 0);   // anis: tracer row after the paint row
    __shared__ double2 l2tab[BFG_LOG2_TAB];
    load_log2_table(l2tab, g_l2tab);
    const i64 plane = (NDIM == 3) ? (i64)N * N : (i64)N;       // cells per axis-0 plane
    const i64 nloc = (i64)(plane_hi - plane_lo) * plane;
    i64 nloc8 = nloc * 8;
    nloc8 = ((i64)__shfl_sync(0xffffffffu, (int)(nloc8 >> 32), 0) << 32) | (unsigned)__shfl_sync(0xffffffffu, (int)nloc8, 0);
    const double inv_res = 1.0 / res;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = GRID_THREADS / 32;
    i64 done = 0;
    for (i64 h = blockIdx.x; h < n_halo; h += gridDim.x) {
        const HaloBox b = load_box(halos + h * BFG_HALO_STRIDE);
        __syncthreads();
        bool valid;
        blend_row(T, b.lnz, b.lnM, extras ? extras + h * n_extra : nullptr, row, valid);
        bool valid2 = true;
        if (MODE == MODE_ANIS) blend_row(A.T2, b.lnz, b.lnM, extras ? extras + h * n_extra : nullptr, row + T.n[2], valid2);
        __syncthreads();
        const int ns = b.nsize, cw = ns / 2;
        const double cut2 = PAINT ? b.paintcut * b.paintcut : b.rcut * b.rcut;
        // lean read-out (uniform ln r axis): cell coordinate u = log2(r^2) * uA + uB, see row_at_r2
        constexpr bool LEAN = (MODE == MODE_BARYONIFY) && UNIFORM;
        RowLookup rl;
        if (LEAN) {
            rl.uA = 0.34657359027997264 * T.inv_dr;
            rl.uB = (((T.flags & BFG_TABLE_RDELTA) ? -b.lnRcom : 0.0) - T.r0) * T.inv_dr;
            rl.uMax = (double)(T.n[2] - 1);
            rl.nrm2 = T.n[2] - 2;
            rl.row_s = (unsigned)__cvta_generic_to_shared(row);
            rl.l2_s = (unsigned)__cvta_generic_to_shared(l2tab);
            launder(rl);
        }
        double R00 = 1.0, R01 = 0.0, R10 = 0.0, R11 = 1.0;
        if (ELL) {
            const double *e = extras + h * n_extra + (n_extra - 4);
            R00 = __ldg(e); R01 = __ldg(e + 1); R10 = __ldg(e + 2); R11 = __ldg(e + 3);
        }
        // Rows of the cutout = every index but the last array axis; a warp owns a row, lanes walk the last axis, so the
        // REDs of a warp are contiguous in memory (the last axis is the fastest one of the C-order grid).
        const int nrows = (NDIM == 3) ? ns * ns : ns;
        for (int rw = warp; rw < nrows; rw += NW) {
            const int i = (NDIM == 3) ? rw / ns : rw;            // cutout index along array axis 0
            const int j = (NDIM == 3) ? rw - i * ns : 0;         // 3-D: cutout index along array axis 1
            const int c0 = wrap_idx(b.cen[0] - cw + i, N);
            if (c0 < plane_lo || c0 >= plane_hi) continue;
            // element (i, j, k): gx = x[j] + dx, gy = x[i] + dy, gz = x[k] + dz   (Map2DRunner.py:524-528 / :561-566)
            // 2-D: the last axis is j itself, so gx varies along the row and gy is the row constant.
            const double gy = cut_coord(i, ns, res) + b.d[1];
            const double gx_row = (NDIM == 3) ? cut_coord(j, ns, res) + b.d[0] : 0.0;
            const double row2 = (NDIM == 3) ? gx_row * gx_row + gy * gy : 0.0;
            const int c1 = (NDIM == 3) ? wrap_idx(b.cen[1] - cw + j, N) : 0;
            const i64 base = (NDIM == 3) ? ((i64)(c0 - plane_lo) * N + c1) * N : (i64)(c0 - plane_lo) * N;
            const int clast0 = b.cen[NDIM - 1] - cw;
            for (int k = lane; k < ns; k += 32) {
                const double gl = cut_coord(k, ns, res) + b.d[NDIM == 3 ? 2 : 0];   // coordinate along the last axis
                const double gx = (NDIM == 3) ? gx_row : gl;
                const double r2 = (NDIM == 3) ? row2 + gl * gl : gl * gl + gy * gy;
                const i64 cell = base + wrap_idx(clast0 + k, N);
                double rt2 = r2;                                                   // radius^2 the table / cuts see
                if (ELL) {
                    const double ex = gx * R00 + gy * R10, ey = gx * R01 + gy * R11;   // (x, y) @ Rmat
                    rt2 = ex * ex + ey * ey;
                }
                if (LEAN) {
                    // BaryonifyGrid, uniform ln r: FP64-lean update (same arithmetic as the generic branch below up to
                    // round-off; NaN / inf offsets propagate and are cleaned after the loop, Map2DRunner.py:597/:607)
                    bool ok;
                    double val = row_at_r2(rl, rt2, ok);
                    if (!ok || !valid) val = CUDART_NAN;                          // outside the table: fill_value = nan
                    if (!(rt2 < cut2)) val = 0.0;                                 // BaryonCorrection.py:410-411
                    ++done;
                    const double sc = (val * inv_res) * rsqrt_pos(r2);            // offset / res / r   (:540/:583)
                    if (sc == 0.0) continue;                                      // exact zeros add nothing
                    double *q = out + cell;
                    red_add(q, sc * gx);
                    red_add((double *)((char *)q + nloc8), sc * gy);
                    if (NDIM == 3) red_add((double *)((char *)q + 2 * nloc8), sc * gl);
                    continue;
                }
                double xq = fast_log2(rt2, l2tab) * 0.34657359027997264;          // ln r = 0.5 ln2 log2(r^2)
                if (T.flags & BFG_TABLE_RDELTA) xq -= b.lnRcom;
                double val = row_lookup<UNIFORM>(T, row, xq);
                if (!valid) val = CUDART_NAN;
                ++done;
                if (MODE == MODE_ANIS) {
                    const double P = exp(val);                                   // Painting  :981
                    if (!isfinite(P) || !(rt2 < cut2)) continue;                 // mask      :987-992
                    double C = exp(row_lookup<UNIFORM>(A.T2, trow, xq));         // Canvas    :982
                    if (!valid2 || !isfinite(C)) continue;                       // :983 -> 0
                    const double m = A.mtot[cell] + A.mtot_add;                  // Mtot_map[inds]
                    if (!(m > 0.0)) continue;                                    // np.divide(..., where = Mtot > 0)  :984
                    const double add = P * ((C / m) * A.orig[cell]);             // :985, :996
                    if (add != 0.0) red_add(out + cell, add);
                    continue;
                }
                if (PAINT) {
                    val = exp(val);                                  // Tabulate.py:319
                    if (!isfinite(val) || !(rt2 < cut2)) continue;   // Map2DRunner.py:814-818
                    val *= scale;                                    // :825 folded in
                    if (val != 0.0) red_add(out + cell, val);
                } else {
                    val = (rt2 < cut2) ? val : 0.0;                  // BaryonCorrection.py:410-411
                    const double sc = (val * inv_res) * rsqrt(r2);   // offset / res / r   (Map2DRunner.py:540/:583)
                    if (sc == 0.0) continue;                         // adds exact zeros; NaN (r = 0, outside table) goes on
                    red_add(out + cell, sc * gx);                    // NaNs propagate (cleaned after the loop, :597/:607)
                    red_add(out + nloc + cell, sc * gy);
                    if (NDIM == 3) red_add(out + 2 * nloc + cell, sc * gl);
                }
            }
        }
    }
    if (nupd) {
        done = warp_sum_i64(done);
        if (lane == 0 && done) atomicAdd(nupd, (unsigned long long)done);
    }
}

// Python float modulo for a positive modulus
__host__ __device__ __forceinline__ double pymod(double x, double n) {
    double r = fmod(x, n);
    if (r != 0.0 && r < 0.0) r += n;
    return r;
}

// One axis of regrid_pixels_2D/3D: the two cells that overlap [xs, xs+1) and their overlap lengths.
__host__ __device__ __forceinline__ void axis_deposit(double pos, int N, int c[2], double w[2]) {
    double xs = pymod(pos, (double)N);
    int f = (int)xs;                       // xs >= 0
    double fp1 = (double)(f + 1);
    w[0] = fp1 - xs;                       // min(f+1, xs+1) - max(f, xs)
    w[1] = (xs + 1.0) - fp1;               // min(f+2, xs+1) - max(f+1, xs)
    c[0] = f % N;
    c[1] = (f + 1) % N;
}

template <int NDIM>
__global__ void __launch_bounds__(256)
k_grid_regrid(int N, const double *__restrict__ map_in, const double *__restrict__ off, double *__restrict__ map_out,
              int plane_lo, int plane_hi) {
    const i64 plane = (NDIM == 3) ? (i64)N * N : (i64)N;
    const i64 nloc = (i64)(plane_hi - plane_lo) * plane;
    for (i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x; c < nloc; c += (i64)gridDim.x * blockDim.x) {
        double m = map_in[c];
        int a0 = (int)(c / plane) + plane_lo;
        i64 rem = c - (i64)(a0 - plane_lo) * plane;
        int a1 = (NDIM == 3) ? (int)(rem / N) : (int)rem;
        int a2 = (NDIM == 3) ? (int)(rem - (i64)a1 * N) : 0;
        double o0 = off[c], o1 = off[nloc + c], o2 = (NDIM == 3) ? off[2 * nloc + c] : 0.0;
        if (!isfinite(o0)) o0 = 0.0;       // Map2DRunner.py:597/:607, element-wise on the accumulated array
        if (!isfinite(o1)) o1 = 0.0;
        if (!isfinite(o2)) o2 = 0.0;
        // xy-meshgrid: component 0 rides on array axis 1, component 1 on axis 0   (:595-599, :605-610)
        int cx[2], cy[2], cz[2];
        double wx[2], wy[2], wz[2];
        axis_deposit(o0 + (double)a1, N, cx, wx);
        axis_deposit(o1 + (double)a0, N, cy, wy);
        if (NDIM == 3) axis_deposit(o2 + (double)a2, N, cz, wz);
#pragma unroll
        for (int iy = 0; iy < 2; ++iy) {
            if (!(wy[iy] > 0)) continue;
#pragma unroll
            for (int ix = 0; ix < 2; ++ix) {
                if (!(wx[ix] > 0)) continue;
                if (NDIM == 2) {
                    red_add(map_out + (i64)cy[iy] * N + cx[ix], (wx[ix] * wy[iy]) * m);       // grid[i=y, j=x]  :79-82
                } else {
#pragma unroll
                    for (int iz = 0; iz < 2; ++iz) {
                        if (!(wz[iz] > 0)) continue;
                        red_add(map_out + ((i64)cy[iy] * N + cx[ix]) * N + cz[iz], ((wx[ix] * wy[iy]) * wz[iz]) * m);  // :158-162
                    }
                }
            }
        }
    }
}

template <int MODE>
int launch_grid(const bfg_table *t, int ndim, i64 N, double res, double scale, i64 n_halo, const double *d_halos,
                const double *d_extras, int n_extra, int use_ell, double *d_out, i64 plane_lo, i64 plane_hi,
                i64 *d_nupdates, cudaStream_t st, const bfg_table *t2 = nullptr, const double *d_mtot = nullptr,
                const double *d_orig = nullptr, double mtot_add = 0.0) {
    constexpr bool PAINT = (MODE != MODE_BARYONIFY);
    BFG_REQUIRE(t && (d_halos || n_halo == 0) && d_out, "null argument");
    AnisArgs A;
    memset(&A, 0, sizeof(A));
    if (MODE == MODE_ANIS) {
        BFG_REQUIRE(ndim == 2, "Can only paint anisotropic profiles on 2D maps (Map2DRunner.py:847)");
        BFG_REQUIRE(t2 && d_mtot && d_orig, "anisotropic paint needs the tracer table and both maps");
        BFG_REQUIRE(t2->view.ndim == t->view.ndim && (t2->view.flags & BFG_TABLE_LOG_VALUES),
                    "tracer table must be a log-profile table with the paint table's extra axes");
        BFG_REQUIRE(t2->device == t->device, "tables live on different devices");
        A.T2 = t2->view; A.mtot = d_mtot; A.orig = d_orig; A.mtot_add = mtot_add;
    }
    BFG_REQUIRE(ndim == 2 || ndim == 3, "ndim must be 2 or 3");
    BFG_REQUIRE(N >= 4 && N <= 32768, "N out of range (need 4 <= N <= 32768)");
    BFG_REQUIRE(plane_lo >= 0 && plane_hi <= N && plane_lo <= plane_hi, "bad plane range");
    BFG_REQUIRE(!use_ell || ndim == 2, "ellipticity exists for 2-D maps only (Map2DRunner.py:571,801)");
    BFG_REQUIRE(n_extra == t->view.ndim - 3 + (use_ell ? 4 : 0),
                "n_extra must equal the table's extra axes (+4 shear-matrix columns when use_ell)");
    BFG_REQUIRE(n_extra == 0 || d_extras, "extras missing");
    BFG_REQUIRE(PAINT == ((t->view.flags & BFG_TABLE_LOG_VALUES) != 0),
                "paint needs a log-profile table, baryonify a displacement table");
    if (d_nupdates) BFG_CUDA_OK(cudaMemsetAsync(d_nupdates, 0, sizeof(i64), st));
    if (n_halo == 0 || plane_lo == plane_hi) return BFG_OK;
    size_t smem = sizeof(double) * (t->view.n[2] + (MODE == MODE_ANIS ? t2->view.n[2] : 0));
    BFG_REQUIRE(smem <= 200 * 1024, "radial axis too long for the shared-memory row (max 25600 nodes)");
    int blocks = (int)std::min<i64>(n_halo, (i64)1 << 30);
    const double2 *g_l2tab = nullptr;
    if (int rc = get_log2_table(&g_l2tab)) return rc;
    auto go = [&](auto kern) -> int {
        BFG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, GRID_THREADS, smem, st>>>(t->view, (int)N, res, scale, n_halo, d_halos, d_extras, n_extra, d_out,
                                                 (int)plane_lo, (int)plane_hi, (unsigned long long *)d_nupdates, g_l2tab, A);
        BFG_CUDA_OK(cudaGetLastError());
        return BFG_OK;
    };
    const bool u = t->view.uniform_r != 0 && (MODE != MODE_ANIS || t2->view.uniform_r != 0);
    if (MODE != MODE_ANIS && ndim == 3 && u) {
        // 3-D grids: tile-centric gather (grid_tile_kernels.cu) -- each cell is written once instead of once per halo.
        // BFG_GRID_TILES=0 forces the halo-centric scatter kernel (kept for geometries the tiling does not cover).
        const char *env = getenv("BFG_GRID_TILES");
        if (!(env && env[0] == '0')) {
            int rc = launch_grid_tiles(PAINT, t, N, res, scale, n_halo, d_halos, d_extras, n_extra, d_out, plane_lo, plane_hi,
                                       d_nupdates, st);
            if (rc != BFG_ERR_UNSUPPORTED) return rc;
        }
    }
    if (MODE == MODE_ANIS) {   // 2-D only
        if (use_ell) return u ? go(k_grid_halos<MODE_ANIS, true, 2, true>) : go(k_grid_halos<MODE_ANIS, false, 2, true>);
        return u ? go(k_grid_halos<MODE_ANIS, true, 2, false>) : go(k_grid_halos<MODE_ANIS, false, 2, false>);
    }
    constexpr int M = (MODE == MODE_ANIS) ? MODE_PAINT : MODE;   // keeps 3-D anis variants from being instantiated
    if (ndim == 3) return u ? go(k_grid_halos<M, true, 3, false>) : go(k_grid_halos<M, false, 3, false>);
    if (use_ell) return u ? go(k_grid_halos<M, true, 2, true>) : go(k_grid_halos<M, false, 2, true>);
    return u ? go(k_grid_halos<M, true, 2, false>) : go(k_grid_halos<M, false, 2, false>);
}

}  // namespace

extern "C" int bfg_grid_offsets(const bfg_table *t, int ndim, int64_t N, double res, int64_t n_halo,
                                const double *d_halos, const double *d_extras, int n_extra, int use_ell,
                                double *d_offsets, int64_t plane_lo, int64_t plane_hi, int64_t *d_nupdates, void *stream) {
    BFG_ENTRY();
    return launch_grid<MODE_BARYONIFY>(t, ndim, N, res, 1.0, n_halo, d_halos, d_extras, n_extra, use_ell, d_offsets, plane_lo,
                              plane_hi, (i64 *)d_nupdates, (cudaStream_t)stream);
}

extern "C" int bfg_grid_paint(const bfg_table *t, int ndim, int64_t N, double res, double scale, int64_t n_halo,
                              const double *d_halos, const double *d_extras, int n_extra, int use_ell, double *d_map,
                              int64_t plane_lo, int64_t plane_hi, int64_t *d_nupdates, void *stream) {
    BFG_ENTRY();
    return launch_grid<MODE_PAINT>(t, ndim, N, res, scale, n_halo, d_halos, d_extras, n_extra, use_ell, d_map, plane_lo,
                             plane_hi, (i64 *)d_nupdates, (cudaStream_t)stream);
}

extern "C" int bfg_grid_paint_anis(const bfg_table *t_paint, const bfg_table *t_tracer, int64_t N, double res,
                                   int64_t n_halo, const double *d_halos, const double *d_extras, int n_extra, int use_ell,
                                   const double *d_mtot, double mtot_add, const double *d_orig, double *d_map,
                                   int64_t plane_lo, int64_t plane_hi, int64_t *d_nupdates, void *stream) {
    BFG_ENTRY();
    return launch_grid<MODE_ANIS>(t_paint, 2, N, res, 1.0, n_halo, d_halos, d_extras, n_extra, use_ell, d_map, plane_lo,
                                  plane_hi, (i64 *)d_nupdates, (cudaStream_t)stream, t_tracer, d_mtot, d_orig, mtot_add);
}

extern "C" int bfg_grid_regrid(int ndim, int64_t N, const double *d_map_in, const double *d_offsets, double *d_map_out,
                               int64_t plane_lo, int64_t plane_hi, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(d_map_in && d_offsets && d_map_out, "null argument");
    BFG_REQUIRE(ndim == 2 || ndim == 3, "ndim must be 2 or 3");
    BFG_REQUIRE(N >= 4 && N <= 32768, "N out of range (need 4 <= N <= 32768)");
    BFG_REQUIRE(plane_lo >= 0 && plane_hi <= N && plane_lo <= plane_hi, "bad plane range");
    if (plane_lo == plane_hi) return BFG_OK;
    i64 nloc = (plane_hi - plane_lo) * (ndim == 3 ? N * N : N);
    int blocks = (int)std::max<i64>(1, std::min<i64>((nloc + 255) / 256, 148 * 32));
    if (ndim == 3)
        k_grid_regrid<3><<<blocks, 256, 0, (cudaStream_t)stream>>>((int)N, d_map_in, d_offsets, d_map_out, (int)plane_lo, (int)plane_hi);
    else
        k_grid_regrid<2><<<blocks, 256, 0, (cudaStream_t)stream>>>((int)N, d_map_in, d_offsets, d_map_out, (int)plane_lo, (int)plane_hi);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

// ---------------------------------------------------------------------------------------------------- host test entry
// Pure host, no GPU: axis_deposit -- one axis of the re-binning kernel (regrid_pixels_2D/3D, Map2DRunner.py:13-162: the two cells
// that overlap [x, x + 1) after the periodic wrap and their overlap lengths) -- on the CPU: h_c [n][2], h_w [n][2].
extern "C" int bfg_test_axis_deposit_host(int64_t n, const double *h_pos, int64_t N, int64_t *h_c, double *h_w) {
    BFG_REQUIRE(n >= 0 && N >= 1 && N <= 2147483647LL && (n == 0 || (h_pos && h_c && h_w)), "bad argument");
    for (int64_t i = 0; i < n; ++i) {
        int c[2];
        double w[2];
        axis_deposit(h_pos[i], (int)N, c, w);
        h_c[2 * i] = c[0]; h_c[2 * i + 1] = c[1];
        h_w[2 * i] = w[0]; h_w[2 * i + 1] = w[1];
    }
    return BFG_OK;
}
