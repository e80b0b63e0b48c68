// grid_common.cuh -- per-halo record view and cutout conventions shared by the grid kernels
// (BaryonForge/Runners/Map2DRunner.py:400-429, :500-528).
#pragma once
#include "bfg_common.cuh"

namespace bfg {

struct HaloBox {
    double rq, lnz, lnM, rcut, lnRcom, d[3], paintcut;
    int nsize, cen[3];
};

__device__ __forceinline__ HaloBox load_box(const double *__restrict__ H) {
    HaloBox b;
    b.rq = __ldg(H + BFG_HB_RQ);
    b.nsize = (int)__ldg(H + BFG_HB_NSIZE);
    b.cen[0] = (int)__ldg(H + BFG_HB_CX); b.cen[1] = (int)__ldg(H + BFG_HB_CY); b.cen[2] = (int)__ldg(H + BFG_HB_CZ);
    b.lnz = __ldg(H + BFG_HB_LNZ); b.lnM = __ldg(H + BFG_HB_LNM);
    b.rcut = __ldg(H + BFG_HB_RCUT); b.lnRcom = __ldg(H + BFG_HB_LNRCOM);
    b.d[0] = __ldg(H + BFG_HB_DX); b.d[1] = __ldg(H + BFG_HB_DY); b.d[2] = __ldg(H + BFG_HB_DZ);
    b.paintcut = __ldg(H + BFG_HB_PAINTCUT);
    return b;
}

// np.linspace(-Ns/2, Ns/2, Ns)[i] * res, same operation order as numpy (arange*step + start, last = stop)
// (__host__ __device__: bfg_test_index_helpers_host runs the same source on the CPU against np.linspace / pick_indices)
__host__ __device__ __forceinline__ double cut_coord(int i, int ns, double res) {
    double start = -0.5 * (double)ns, stop = 0.5 * (double)ns;
    double step = (stop - start) / (double)(ns - 1);
#ifdef __CUDA_ARCH__
    double y = (i == ns - 1) ? stop : __dadd_rn(__dmul_rn((double)i, step), start);
#else
    double y = (i == ns - 1) ? stop : ((double)i * step) + start;     // x86-64 baseline: no FMA contraction
#endif
    return y * res;
}

__host__ __device__ __forceinline__ int wrap_idx(int c, int N) {   // pick_indices, Map2DRunner.py:400-429
    if (c < 0) c += N;
    if (c >= N) c -= N;
    return c;
}


// launcher of the tile-centric 3-D kernels (grid_tile_kernels.cu); returns BFG_ERR_UNSUPPORTED when the geometry does not
// fit the tiling (the caller then uses the halo-centric scatter kernels)
int launch_grid_tiles(bool paint, const bfg_table *t, i64 N, double res, double scale, i64 n_halo, const double *d_halos,
                      const double *d_extras, int n_extra, double *d_out, i64 plane_lo, i64 plane_hi, i64 *d_nupdates,
                      cudaStream_t st);

}  // namespace bfg
