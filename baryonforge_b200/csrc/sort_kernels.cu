// sort_kernels.cu -- locality ordering of halo records ("halo batches sorted by sky or box cell", north_star (b)).
//
// The reference walks halos in catalogue order (Runners/HealpixRunner.py:315, Map2DRunner.py:482, SnapshotRunner.py:217);
// the sums it accumulates are order-independent up to fp64 round-off.  On the GPU ~700 halos are in flight at once, so
// putting sky/box neighbours next to each other keeps the pixels/cells/particles they share resident in the 126 MB L2:
// the fp64 REDs then hit L2 instead of forcing an HBM read-modify-write per update.
#include <cub/device/device_radix_sort.cuh>
#include "bfg_common.cuh"

using namespace bfg;

namespace {

// sky: colatitude bands of width `band`, serpentine in azimuth so consecutive bands join up
__global__ void k_keys_sky(i64 n, const double *__restrict__ halos, double band, unsigned long long *__restrict__ keys,
                           unsigned int *__restrict__ idx) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        double theta = halos[i * BFG_HALO_STRIDE + BFG_HS_THETA], phi = halos[i * BFG_HALO_STRIDE + BFG_HS_PHI];
        unsigned long long b = (unsigned long long)fmin(fmax(theta / band, 0.0), 1048575.0);
        double f = fmin(fmax(phi * BFG_INV_TWOPI, 0.0), 0.99999999);
        unsigned long long q = (unsigned long long)(f * 16777216.0);   // 24 bits of azimuth
        if (b & 1ULL) q = 16777215ULL - q;
        keys[i] = (b << 24) | q;
        idx[i] = (unsigned int)i;
    }
}

// sky + ownership (ring-range sharding): as k_keys_sky, but halos whose disc cannot touch [pix_lo, pix_hi) get bit 44 set,
// so they sort behind every owned halo
constexpr int SKIP_BIT = 44;
__global__ void k_keys_sky_owned(i64 n, const double *__restrict__ halos, double band, Hpx h, i64 pix_lo, i64 pix_hi,
                                 unsigned long long *__restrict__ keys, unsigned int *__restrict__ idx) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const double *H = halos + i * BFG_HALO_STRIDE;
        double theta = H[BFG_HS_THETA], phi = H[BFG_HS_PHI];
        unsigned long long b = (unsigned long long)fmin(fmax(theta / band, 0.0), 1048575.0);
        double f = fmin(fmax(phi * BFG_INV_TWOPI, 0.0), 0.99999999);
        unsigned long long q = (unsigned long long)(f * 16777216.0);
        if (b & 1ULL) q = 16777215ULL - q;
        unsigned long long key = (b << 24) | q;
        const DiscRings d = disc_rings(h, theta, phi, H[BFG_HS_RADIUS]);
        if (!disc_touches_range(h, d, pix_lo, pix_hi)) key |= 1ULL << SKIP_BIT;
        keys[i] = key;
        idx[i] = (unsigned int)i;
    }
}

// box: coarse raster cells of side L / nc, serpentine along the last axis
__global__ void k_keys_box(i64 n, const double *__restrict__ halos, double L, int nc, int ndim,
                           unsigned long long *__restrict__ keys, unsigned int *__restrict__ idx) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        unsigned long long key = 0;
        for (int d = 0; d < ndim; ++d) {
            double x = halos[i * BFG_HALO_STRIDE + BFG_HB_X + d];
            int c = (int)floor(x / L * (double)nc);
            c = min(max(c, 0), nc - 1);
            if (d == ndim - 1 && (key & 1ULL)) c = nc - 1 - c;
            key = key * (unsigned long long)nc + (unsigned long long)c;
        }
        keys[i] = key;
        idx[i] = (unsigned int)i;
    }
}

__global__ void k_gather_rows(i64 n, int width, const unsigned int *__restrict__ idx, const double *__restrict__ in,
                              double *__restrict__ out) {
    i64 total = n * width;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
        i64 r = t / width;
        int c = (int)(t - r * width);
        out[t] = in[(i64)idx[r] * width + c];
    }
}

// halo records of other ranks (sorted last): mark them so the halo loop stops there
__global__ void k_mark_skipped(i64 n, const unsigned long long *__restrict__ sorted_keys, double *__restrict__ out) {
    for (i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (i64)gridDim.x * blockDim.x)
        if ((sorted_keys[r] >> SKIP_BIT) & 1ULL) out[r * BFG_HALO_STRIDE + BFG_HS_SKIP] = 1.0;
}

}  // namespace

namespace {
int halo_sort_impl(int mode, int64_t n_halo, const double *d_in, double *d_out, const double *d_extras_in,
                   double *d_extras_out, int n_extra, double p0, double p1, int ndim, int nside, i64 pix_lo, i64 pix_hi,
                   void *stream) {
    BFG_REQUIRE(d_in && d_out && d_in != d_out, "need distinct in/out record buffers");
    BFG_REQUIRE(mode == 0 || mode == 1 || mode == 2, "mode: 0 = sky bands, 1 = box cells, 2 = sky bands + ownership");
    BFG_REQUIRE(n_halo >= 0 && n_halo <= 2147483647LL, "n_halo out of range (the radix sort counts items in an int)");
    BFG_REQUIRE(n_extra == 0 || (d_extras_in && d_extras_out), "extras missing");
    if (n_halo == 0) return BFG_OK;
    if (int rc = retain_async_pool()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // every argument is checked before the first allocation; the stream-ordered scratch is released on every return path
    if (mode == 0 || mode == 2) BFG_REQUIRE(p0 > 0, "band width must be positive");
    if (mode == 2) BFG_REQUIRE(nside >= 1 && nside <= (1 << 24), "nside out of range");
    if (mode == 1) BFG_REQUIRE(p0 > 0 && p1 >= 1 && p1 <= 1024 && (ndim == 2 || ndim == 3), "bad box parameters");
    struct Scratch {
        void *p = nullptr; cudaStream_t st;
        explicit Scratch(cudaStream_t s) : st(s) {}
        ~Scratch() { if (p) cudaFreeAsync(p, st); }
    } s_keys(st), s_idx(st), s_tmp(st);
    size_t tmp_bytes = 0;
    BFG_CUDA_OK(cudaMallocAsync(&s_keys.p, sizeof(unsigned long long) * n_halo * 2, st));
    BFG_CUDA_OK(cudaMallocAsync(&s_idx.p, sizeof(unsigned int) * n_halo * 2, st));
    unsigned long long *keys = (unsigned long long *)s_keys.p, *keys2 = keys + n_halo;
    unsigned int *idx = (unsigned int *)s_idx.p, *idx2 = idx + n_halo;
    int blocks = (int)std::max<i64>(1, std::min<i64>((n_halo + 255) / 256, 148 * 8));
    int end_bit = 64;
    if (mode == 0) {
        k_keys_sky<<<blocks, 256, 0, st>>>(n_halo, d_in, p0, keys, idx);
        end_bit = 24 + 20;
    } else if (mode == 2) {
        k_keys_sky_owned<<<blocks, 256, 0, st>>>(n_halo, d_in, p0, Hpx(nside), pix_lo, pix_hi, keys, idx);
        end_bit = SKIP_BIT + 1;
    } else {
        k_keys_box<<<blocks, 256, 0, st>>>(n_halo, d_in, p0, (int)p1, ndim, keys, idx);
        end_bit = 32;
    }
    BFG_CUDA_OK(cudaGetLastError());
    BFG_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, idx, idx2, (int)n_halo, 0, end_bit, st));
    BFG_CUDA_OK(cudaMallocAsync(&s_tmp.p, tmp_bytes, st));
    BFG_CUDA_OK(cub::DeviceRadixSort::SortPairs(s_tmp.p, tmp_bytes, keys, keys2, idx, idx2, (int)n_halo, 0, end_bit, st));
    int gblocks = (int)std::max<i64>(1, std::min<i64>((n_halo * BFG_HALO_STRIDE + 255) / 256, 148 * 16));
    k_gather_rows<<<gblocks, 256, 0, st>>>(n_halo, BFG_HALO_STRIDE, idx2, d_in, d_out);
    if (n_extra) k_gather_rows<<<gblocks, 256, 0, st>>>(n_halo, n_extra, idx2, d_extras_in, d_extras_out);
    if (mode == 2) k_mark_skipped<<<blocks, 256, 0, st>>>(n_halo, keys2, d_out);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}
}  // namespace

extern "C" int bfg_halo_sort(int mode, int64_t n_halo, const double *d_in, double *d_out, const double *d_extras_in,
                             double *d_extras_out, int n_extra, double p0, double p1, int ndim, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(mode == 0 || mode == 1, "mode: 0 = sky bands, 1 = box cells");
    return halo_sort_impl(mode, n_halo, d_in, d_out, d_extras_in, d_extras_out, n_extra, p0, p1, ndim, 0, 0, 0, stream);
}

extern "C" int bfg_halo_sort_owned(int nside, int64_t pix_lo, int64_t pix_hi, int64_t n_halo, const double *d_in,
                                   double *d_out, const double *d_extras_in, double *d_extras_out, int n_extra,
                                   double band, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(pix_lo >= 0 && pix_lo <= pix_hi, "bad pixel range");
    return halo_sort_impl(2, n_halo, d_in, d_out, d_extras_in, d_extras_out, n_extra, band, 0.0, 3, nside, pix_lo, pix_hi,
                          stream);
}
