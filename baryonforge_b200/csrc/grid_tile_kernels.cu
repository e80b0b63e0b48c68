// grid_tile_kernels.cu -- tile-centric (gather) form of the 3-D grid halo loops on sm_100a.
//
//   BaryonifyGrid.process     halo loop   BaryonForge/Runners/Map2DRunner.py:482-586
//   PaintProfilesGrid.process halo loop   BaryonForge/Runners/Map2DRunner.py:725-821 (+ :825)
//
// Why: at 1024^3 cells and 10^6 halos a cell receives ~350 contributions.  The halo-centric scatter kernel
// (k_grid_halos) turns every one of them into fp64 REDs on a 26 GB array that no cache can hold for the ~1200 halos in
// flight: 17.8 TB of read-modify-write traffic, i.e. the kernel sits on the HBM roofline (2.2 s).  Here the loop is
// inverted: a CTA owns a TILE of 8 x 16 x 16 cells, keeps its accumulators in REGISTERS (8 cells x 3 components per
// thread), walks the list of halos whose cutout overlaps the tile, and writes each cell ONCE -- 26 GB of traffic instead
// of 17.8 TB, so the loop runs at the speed of its arithmetic.  The blended radial row of every halo is computed once
// into a global buffer (a tile touches a narrow band of its nodes, which stays in L1/L2), the pair loop has no barrier, and
// cells beyond the model's cut (`where(r < epsilon_max R, d, 0)`, BaryonCorrection.py:410-411 / the paint mask
// Map2DRunner.py:814-815) are skipped before the read-out -- the corners of the cubic cutout are 48 % of its cells.
// Results equal the scatter kernels' up to fp64 summation order (the reference's own loop order is yet another order).
//
// Pipeline: k_tile_count (pairs per tile, update count) -> exclusive scan -> k_tile_fill (halo ids per tile) ->
//           k_blend_rows -> k_tile_gather (persistent CTAs pull tiles from a queue).
#include <algorithm>
#include <cub/device/device_scan.cuh>
#include "grid_common.cuh"

using namespace bfg;

namespace {

constexpr int TI = 8, TA = 16, TB = 16;        // tile extent along array axes 0, 1, 2
constexpr int TILE_THREADS = TA * TB;          // one thread per (a, b) column of TI cells

struct TileGeom {
    int N, ntI, ntA, ntB;                      // grid size; tiles per axis (axis 0 counts owned planes only)
    int plane_lo, plane_hi;
    i64 ntiles;
};

// first cutout cell (wrapped into [0, N)) of a halo along one axis
__device__ __forceinline__ int cut_start(int cen, int cw, int N) {
    int s = cen - cw;
    if (s < 0) s += N;
    return s;
}

// Enumerate the tiles a halo's cutout box overlaps (owned planes only).  f(tile_id) is called once per tile.
template <typename F>
__device__ __forceinline__ void for_each_tile(const TileGeom &g, const HaloBox &b, int lane, F f) {
    const int ns = b.nsize, cw = ns / 2;
    const int s0 = cut_start(b.cen[0], cw, g.N), s1 = cut_start(b.cen[1], cw, g.N), s2 = cut_start(b.cen[2], cw, g.N);
    // unwrapped tile spans [u_lo, u_hi] per axis; N is a multiple of the tile extents, so wrapping is a modulo on tiles
    const int a_lo = s1 / TA, a_hi = (s1 + ns - 1) / TA, b_lo = s2 / TB, b_hi = (s2 + ns - 1) / TB;
    const int i_lo = s0 / TI, i_hi = (s0 + ns - 1) / TI;
    const int nA = a_hi - a_lo + 1, nB = b_hi - b_lo + 1, nI = i_hi - i_lo + 1;
    const int ntA_full = g.N / TA, ntB_full = g.N / TB, ntI_full = g.N / TI;
    const int total = nI * nA * nB;
    for (int q = lane; q < total; q += 32) {
        const int qi = q / (nA * nB), r = q - qi * (nA * nB), qa = r / nB, qb = r - qa * nB;
        const int gi = (i_lo + qi) % ntI_full;                 // global axis-0 tile
        const int first_plane = gi * TI;
        if (first_plane + TI <= g.plane_lo || first_plane >= g.plane_hi) continue;   // not an owned plane range
        const int ti = (first_plane - g.plane_lo) / TI;        // plane_lo is a multiple of TI
        const int ta = (a_lo + qa) % ntA_full, tb = (b_lo + qb) % ntB_full;
        f(((i64)ti * g.ntA + ta) * g.ntB + tb);
    }
}

// pairs per tile + the reference's update count (every cutout cell of an owned plane counts, Map2DRunner.py:510-586)
__global__ void __launch_bounds__(256)
k_tile_count(TileGeom g, i64 n_halo, const double *__restrict__ halos, unsigned int *__restrict__ counts,
             unsigned long long *nupd) {
    const int lane = threadIdx.x & 31;
    const i64 wid = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((i64)gridDim.x * blockDim.x) >> 5;
    unsigned long long upd = 0;
    for (i64 h = wid; h < n_halo; h += nw) {
        const HaloBox b = load_box(halos + h * BFG_HALO_STRIDE);
        for_each_tile(g, b, lane, [&](i64 tile) { atomicAdd(counts + tile, 1u); });
        if (lane == 0) {
            const int ns = b.nsize, s0 = cut_start(b.cen[0], ns / 2, g.N);
            int owned = 0;
            for (int i = 0; i < ns; ++i) {
                int c = s0 + i; if (c >= g.N) c -= g.N;
                owned += (c >= g.plane_lo && c < g.plane_hi) ? 1 : 0;
            }
            upd += (unsigned long long)owned * (unsigned long long)ns * (unsigned long long)ns;
        }
    }
    if (nupd && upd) atomicAdd(nupd, upd);
}

__global__ void __launch_bounds__(256)
k_tile_fill(TileGeom g, i64 n_halo, const double *__restrict__ halos, const i64 *__restrict__ tile_start,
            unsigned int *__restrict__ cursor, unsigned int *__restrict__ pair_halo) {
    const int lane = threadIdx.x & 31;
    const i64 wid = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 h = wid; h < n_halo; h += nw) {
        const HaloBox b = load_box(halos + h * BFG_HALO_STRIDE);
        for_each_tile(g, b, lane, [&](i64 tile) {
            const unsigned slot = atomicAdd(cursor + tile, 1u);
            pair_halo[tile_start[tile] + slot] = (unsigned)h;
        });
    }
}

__global__ void k_widen_counts(i64 n, const unsigned int *__restrict__ c32, i64 *__restrict__ c64) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (i64)gridDim.x * blockDim.x)
        c64[i] = (i < n) ? (i64)c32[i] : 0;
}

// cutout index of global cell c along an axis whose cutout starts at s: (c - s) mod N; inside the cutout iff < ns
__device__ __forceinline__ int cut_index(int c, int s, int N) {
    int i = c - s;
    if (i < 0) i += N;
    return i;
}

// Blended radial rows of ALL halos, once: rows[h][k] = sum_c w_c table[corner_c, k] (blend_row writing to global memory).
// rows_valid[h] = 0 when the halo lies outside a non-radial axis (every read-out is NaN).
__global__ void __launch_bounds__(128)
k_blend_rows(TableView T, i64 n_halo, const double *__restrict__ halos, const double *__restrict__ extras, int n_extra,
             double *__restrict__ rows, unsigned char *__restrict__ rows_valid) {
    for (i64 h = blockIdx.x; h < n_halo; h += gridDim.x) {
        const double *H = halos + h * BFG_HALO_STRIDE;
        bool valid;
        blend_row(T, __ldg(H + BFG_HB_LNZ), __ldg(H + BFG_HB_LNM), extras ? extras + h * n_extra : nullptr,
                  rows + h * T.n[2], valid);
        if (threadIdx.x == 0) rows_valid[h] = valid ? 1 : 0;
    }
}

// row_at_r2 with the blended row in global memory (L1/L2 resident: a tile touches a narrow band of nodes per halo)
__device__ __forceinline__ double grow_at_r2(const double *__restrict__ rowp, double uA, double uB, double uMax, int nrm2,
                                             unsigned l2_s, double r2, bool &ok) {
    const int hi = __double2hiint(r2);
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(r2));
    const double2 t = lds_f64x2(l2_s + (((unsigned)hi >> 9) & 0x7f0u));
    const double fr = fma(m, t.x, -1.0);
    double p = fma(fr, c_l2p[0], c_l2p[1]);
    p = fma(fr, p, c_l2p[2]);
    p = fma(fr, p, c_l2p[3]);
    p = fma(fr, p, c_l2p[4]);
    const double ed = __hiloint2double(0x43300000, (hi >> 20) ^ 0x80000000) - 4503601774855167.0;
    const double uu = fma(fma(fr, p, t.y) + ed, uA, uB);
    int k = __double2int_rd(uu);
    ok = true;
    if (__builtin_expect((unsigned)k > (unsigned)nrm2, 0)) {
        ok = (uu == uMax);
        k = nrm2;
    }
    const double tt = uu - (double)k;
    const double v0 = __ldg(rowp + k);
    return fma(tt, __ldg(rowp + k + 1) - v0, v0);
}

// One CTA per tile (persistent, tile queue); thread (a, b) owns the column of TI cells along axis 0 and keeps their
// accumulators in registers.  The pair loop has no barrier: warps run ahead independently.
template <bool PAINT>
__global__ void __launch_bounds__(TILE_THREADS, 3)
k_tile_gather(TableView T, TileGeom g, double res, double scale, const double *__restrict__ halos,
              const double *__restrict__ rows, const unsigned char *__restrict__ rows_valid, double *__restrict__ out,
              const i64 *__restrict__ tile_start, const unsigned int *__restrict__ pair_halo, unsigned long long *queue,
              const double2 *__restrict__ g_l2tab) {
    __shared__ double2 l2tab[BFG_LOG2_TAB];
    __shared__ i64 s_tile;
    load_log2_table(l2tab, g_l2tab);
    const int t = threadIdx.x;
    const int N = g.N, NR = T.n[2];
    const i64 plane = (i64)N * N;
    const i64 nloc = (i64)(g.plane_hi - g.plane_lo) * plane;
    const double inv_res = 1.0 / res;
    constexpr int NCOMP = PAINT ? 1 : 3;
    // kernel-lifetime constants, laundered through a shuffle so that ptxas keeps them in registers instead of re-deriving
    // them from the parameter bank for every cell (see launder() in bfg_common.cuh)
    RowLookup rl;
    rl.uA = 0.34657359027997264 * T.inv_dr; rl.uB = 0.0; rl.uMax = (double)(NR - 1); rl.nrm2 = NR - 2;
    rl.row_s = 0; rl.l2_s = (unsigned)__cvta_generic_to_shared(l2tab);
    launder(rl);
    const unsigned l2_s = rl.l2_s;
    const double uA = rl.uA, uMax = rl.uMax;
    const int nrm2 = rl.nrm2;
    const bool rdelta = (T.flags & BFG_TABLE_RDELTA) != 0;

    for (;;) {
        __syncthreads();
        if (t == 0) s_tile = (i64)atomicAdd(queue, 1ULL);
        __syncthreads();
        const i64 tile = s_tile;
        if (tile >= g.ntiles) break;
        const int tb = (int)(tile % g.ntB), ta = (int)((tile / g.ntB) % g.ntA), ti = (int)(tile / ((i64)g.ntA * g.ntB));
        const int I0 = g.plane_lo + ti * TI, A = ta * TA + (t >> 4), B = tb * TB + (t & 15);
        double acc[TI][NCOMP];
#pragma unroll
        for (int c = 0; c < TI; ++c)
#pragma unroll
            for (int k = 0; k < NCOMP; ++k) acc[c][k] = 0.0;

        const i64 p_end = tile_start[tile + 1];
        for (i64 p = tile_start[tile]; p < p_end; ++p) {
            const i64 h = pair_halo[p];
            // the halo's blended row (uniform across the CTA); laundered like the constants above, else its 64-bit address
            // arithmetic is redone for every cell
            __syncwarp();
            i64 roff = h * NR;
            roff = ((i64)__shfl_sync(0xffffffffu, (int)(roff >> 32), 0) << 32) | (unsigned)__shfl_sync(0xffffffffu, (int)roff, 0);
            const double *__restrict__ rowp = rows + roff;
            const double *H = halos + h * BFG_HALO_STRIDE;
            const int ns = (int)__ldg(H + BFG_HB_NSIZE), cw = ns / 2;
            const int ia = cut_index(A, cut_start((int)__ldg(H + BFG_HB_CY), cw, N), N);
            const int ib = cut_index(B, cut_start((int)__ldg(H + BFG_HB_CZ), cw, N), N);
            if (ia >= ns || ib >= ns) continue;                                // this column is outside the cutout
            // element (i, j, k) of the cutout: gx = x[j] + dx, gy = x[i] + dy, gz = x[k] + dz   (Map2DRunner.py:561-566)
            const double gx = cut_coord(ia, ns, res) + __ldg(H + BFG_HB_DX);
            const double gz = cut_coord(ib, ns, res) + __ldg(H + BFG_HB_DZ);
            const double gxz2 = fma(gx, gx, gz * gz);
            const double cut = PAINT ? __ldg(H + BFG_HB_PAINTCUT) : __ldg(H + BFG_HB_RCUT);
            const double cut2 = cut * cut;
            if (!(gxz2 < cut2)) continue;                                      // the whole column is beyond the cut: adds 0
            const int s0 = cut_start((int)__ldg(H + BFG_HB_CX), cw, N);
            const double dy = __ldg(H + BFG_HB_DY);
            const double uB = ((rdelta ? -__ldg(H + BFG_HB_LNRCOM) : 0.0) - T.r0) * T.inv_dr;
            const bool valid = rows_valid[h] != 0;
#pragma unroll
            for (int c = 0; c < TI; ++c) {
                const int i0 = cut_index(I0 + c, s0, N);
                if (i0 >= ns || I0 + c >= g.plane_hi) continue;
                const double gy = cut_coord(i0, ns, res) + dy;
                const double r2 = fma(gy, gy, gxz2);
                if (!(r2 < cut2)) continue;                                    // beyond the cut: contributes exactly 0
                bool ok;
                double val = grow_at_r2(rowp, uA, uB, uMax, nrm2, l2_s, r2, ok);
                if (!ok || !valid) val = CUDART_NAN;                           // outside the table: fill_value = nan
                if (PAINT) {
                    val = exp(val);                                            // Tabulate.py:319
                    if (isfinite(val)) acc[c][0] += val * scale;               // Map2DRunner.py:814-821, :825 folded in
                } else {
                    const double sc = (val * inv_res) * rsqrt_pos(r2);         // offset / res / r   (:583); NaN propagates
                    acc[c][0] = fma(sc, gx, acc[c][0]);
                    acc[c][1 % NCOMP] = fma(sc, gy, acc[c][1 % NCOMP]);
                    acc[c][2 % NCOMP] = fma(sc, gz, acc[c][2 % NCOMP]);
                }
            }
        }
        // ---- write the tile once (accumulating into the caller's zeroed array, like the scatter kernels) ------------
        if (A < N && B < N) {
#pragma unroll
            for (int c = 0; c < TI; ++c) {
                if (I0 + c >= g.plane_hi) continue;
                const i64 cell = ((i64)(I0 + c - g.plane_lo) * N + A) * N + B;
#pragma unroll
                for (int k = 0; k < NCOMP; ++k)
                    if (acc[c][k] != 0.0) out[(i64)k * nloc + cell] += acc[c][k];
            }
        }
    }
}

}  // namespace

namespace bfg {

int launch_grid_tiles(bool paint, const bfg_table *t, i64 N, double res, double scale, i64 n_halo, const double *d_halos,
                      const double *d_extras, int n_extra, double *d_out, i64 plane_lo, i64 plane_hi, i64 *d_nupdates,
                      cudaStream_t st) {
    // geometry the tiling needs: tiles must wrap cleanly and a cutout (<= N/2 cells) must meet a tile in ONE run
    if (N % TA != 0 || N % TB != 0 || N % TI != 0 || N < 2 * TA || plane_lo % TI != 0) return BFG_ERR_UNSUPPORTED;
    if (!t->view.uniform_r || n_halo >= ((i64)1 << 32)) return BFG_ERR_UNSUPPORTED;
    if ((double)n_halo * (double)t->view.n[2] * 8.0 > 16e9) return BFG_ERR_UNSUPPORTED;   // blended-row buffer cap
    TileGeom g;
    g.N = (int)N; g.plane_lo = (int)plane_lo; g.plane_hi = (int)plane_hi;
    g.ntI = (int)((plane_hi - plane_lo + TI - 1) / TI); g.ntA = (int)(N / TA); g.ntB = (int)(N / TB);
    g.ntiles = (i64)g.ntI * g.ntA * g.ntB;
    if (g.ntiles >= ((i64)1 << 31)) return BFG_ERR_UNSUPPORTED;
    if (int rc = retain_async_pool()) return rc;
    StreamScratch s_counts(st), s_start(st), s_queue(st), s_scan(st), s_pairs(st), s_rows(st), s_valid(st);   // freed on every return
    size_t scan_bytes = 0;
    BFG_CUDA_OK(s_counts.alloc(sizeof(unsigned int) * (g.ntiles + 1)));
    BFG_CUDA_OK(s_start.alloc(sizeof(i64) * (g.ntiles + 1) * 2));
    BFG_CUDA_OK(s_queue.alloc(sizeof(unsigned long long)));
    unsigned int *counts = s_counts.as<unsigned int>();
    i64 *tile_start = s_start.as<i64>();
    unsigned long long *queue = s_queue.as<unsigned long long>();
    BFG_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(unsigned int) * (g.ntiles + 1), st));
    BFG_CUDA_OK(cudaMemsetAsync(queue, 0, sizeof(unsigned long long), st));
    i64 *counts64 = tile_start + (g.ntiles + 1);
    const int hblocks = (int)std::max<i64>(1, std::min<i64>((n_halo * 32 + 255) / 256, 148 * 16));
    k_tile_count<<<hblocks, 256, 0, st>>>(g, n_halo, d_halos, counts, (unsigned long long *)d_nupdates);
    const int tblocks = (int)std::max<i64>(1, std::min<i64>((g.ntiles + 256) / 256, 148 * 16));
    k_widen_counts<<<tblocks, 256, 0, st>>>(g.ntiles, counts, counts64);
    BFG_CUDA_OK(cudaGetLastError());
    BFG_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, counts64, tile_start, g.ntiles + 1, st));
    BFG_CUDA_OK(s_scan.alloc(scan_bytes));
    BFG_CUDA_OK(cub::DeviceScan::ExclusiveSum(s_scan.p, scan_bytes, counts64, tile_start, g.ntiles + 1, st));
    i64 n_pairs = 0;   // the pair list is sized on the host: one 8-byte read-back per call
    BFG_CUDA_OK(cudaMemcpyAsync(&n_pairs, tile_start + g.ntiles, sizeof(i64), cudaMemcpyDeviceToHost, st));
    BFG_CUDA_OK(cudaStreamSynchronize(st));
    BFG_CUDA_OK(s_pairs.alloc(sizeof(unsigned int) * std::max<i64>(n_pairs, 1)));
    unsigned int *pair_halo = s_pairs.as<unsigned int>();
    BFG_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(unsigned int) * (g.ntiles + 1), st));
    k_tile_fill<<<hblocks, 256, 0, st>>>(g, n_halo, d_halos, tile_start, counts, pair_halo);
    BFG_CUDA_OK(cudaGetLastError());
    const double2 *g_l2tab = nullptr;
    if (int rc = get_log2_table(&g_l2tab)) return rc;
    // blended rows of all halos (n_halo x NR doubles; 4 GB for 10^6 halos x 500 nodes)
    BFG_CUDA_OK(s_rows.alloc(sizeof(double) * n_halo * t->view.n[2]));
    BFG_CUDA_OK(s_valid.alloc((size_t)n_halo));
    double *rows = s_rows.as<double>();
    unsigned char *rows_valid = s_valid.as<unsigned char>();
    k_blend_rows<<<(int)std::min<i64>(n_halo, 148 * 32), 128, 0, st>>>(t->view, n_halo, d_halos, d_extras, n_extra, rows,
                                                                     rows_valid);
    BFG_CUDA_OK(cudaGetLastError());
    int sms = 148;
    BFG_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, t->device));
    const int blocks = (int)std::min<i64>(g.ntiles, (i64)sms * 3);
    if (paint)
        k_tile_gather<true><<<blocks, TILE_THREADS, 0, st>>>(t->view, g, res, scale, d_halos, rows, rows_valid, d_out,
                                                            tile_start, pair_halo, queue, g_l2tab);
    else
        k_tile_gather<false><<<blocks, TILE_THREADS, 0, st>>>(t->view, g, res, scale, d_halos, rows, rows_valid, d_out,
                                                             tile_start, pair_halo, queue, g_l2tab);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

}  // namespace bfg
