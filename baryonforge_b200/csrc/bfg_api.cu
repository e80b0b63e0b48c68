// bfg_api.cu -- library plumbing of the C ABI: errors, device info, tables, small utilities.
#include <stdarg.h>
#include <string.h>
#include <vector>
#include <cmath>
#include <algorithm>
#include <cstdlib>
#include <limits>
#include <mutex>
#include "bfg_common.cuh"

namespace bfg {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace bfg

using namespace bfg;

namespace bfg {
int retain_async_pool() {
    static bool done[64] = {false};
    int dev = 0;
    BFG_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || done[dev]) return BFG_OK;
    cudaMemPool_t pool;
    BFG_CUDA_OK(cudaDeviceGetDefaultMemPool(&pool, dev));
    unsigned long long keep = ~0ULL;
    BFG_CUDA_OK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    done[dev] = true;
    return BFG_OK;
}

int get_log2_table(const double2 **d_tab) {
    static double2 *tabs[64] = {nullptr};
    int dev = 0;
    BFG_CUDA_OK(cudaGetDevice(&dev));
    BFG_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
    if (!tabs[dev]) {
        double2 h[BFG_LOG2_TAB];
        fill_log2_table(h);
        double2 *d = nullptr;
        BFG_CUDA_OK(cudaMalloc(&d, sizeof(h)));
        BFG_CUDA_OK(cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice));
        tabs[dev] = d;
    }
    *d_tab = tabs[dev];
    return BFG_OK;
}
namespace {
__global__ void k_ring_table(Hpx h, RingTabEntry *__restrict__ tab) {
    const i64 n = 4 * h.nside;
    for (i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (i64)gridDim.x * blockDim.x) {
        double e[8];
        ring_table_entry(h, r, e);
        tab[r] = RingTabEntry{e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7]};
    }
}
}  // namespace

int get_ring_table(long long nside, const RingTabEntry **d_tab, void *stream) {
    struct Slot { int dev; long long nside; RingTabEntry *tab; };
    static Slot slots[64];
    static int n_slots = 0;
    static std::mutex mu;
    *d_tab = nullptr;
    const char *lit = getenv("BFG_REGRID_LITERAL");
    if ((lit && lit[0] == '1') || nside < 32) return BFG_OK;      // coarse maps: ring spacing beyond the small-angle series
    int dev = 0;
    BFG_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < n_slots; ++i)
        if (slots[i].dev == dev && slots[i].nside == nside) { *d_tab = slots[i].tab; return BFG_OK; }
    if (n_slots == 64) return BFG_OK;                             // cache full: the literal path is always correct
    RingTabEntry *tab = nullptr;
    BFG_CUDA_OK(cudaMalloc(&tab, sizeof(RingTabEntry) * 4 * nside));
    k_ring_table<<<(int)std::min<long long>((4 * nside + 255) / 256, 148 * 4), 256, 0, (cudaStream_t)stream>>>(Hpx(nside), tab);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);   // other streams may use the table next
    if (e != cudaSuccess) {
        cudaFree(tab);
        set_error("get_ring_table: %s", cudaGetErrorString(e));
        return BFG_ERR_CUDA;
    }
    slots[n_slots++] = Slot{dev, nside, tab};
    *d_tab = tab;
    return BFG_OK;
}
}  // namespace bfg

// test entry (pure host, no GPU): regrid_target_fast -- the source k_shell_regrid runs -- on the CPU, for the CPU parity suite.
// h_off is [3][n] (component-major like the device offsets); h_out_pix / h_out_w are [n][4]; h_fast[i] = 0 where the function
// declined (poles, large displacements) and the kernel would take the literal chain.
extern "C" int bfg_test_regrid_target_host(int nside, int64_t n, const int64_t *h_pix, const double *h_off, int64_t *h_out_pix,
                                           double *h_out_w, int *h_fast) {
    BFG_REQUIRE(nside >= 1 && nside <= (1 << 24) && n >= 0, "bad argument");
    BFG_REQUIRE(n == 0 || (h_pix && h_off && h_out_pix && h_out_w && h_fast), "null argument");
    const Hpx h(nside);
    std::vector<RingTabEntry> tab((size_t)(4 * h.nside + 1));
    for (i64 r = 0; r < 4 * h.nside; ++r) {
        double e[8];
        ring_table_entry(h, r, e);
        tab[(size_t)r] = RingTabEntry{e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7]};
    }
    for (int64_t i = 0; i < n; ++i) {
        i64 pix[4] = {0, 0, 0, 0};
        double w[4] = {0, 0, 0, 0};
        BFG_REQUIRE(h_pix[i] >= 0 && h_pix[i] < h.npix, "pixel out of range");
        const bool ok = regrid_target_fast(h, tab.data(), h_pix[i], h_off[i], h_off[n + i], h_off[2 * n + i], pix, w);
        h_fast[i] = ok ? 1 : 0;
        for (int k = 0; k < 4; ++k) { h_out_pix[4 * i + k] = ok ? pix[k] : -1; h_out_w[4 * i + k] = ok ? w[k] : 0.0; }
    }
    return BFG_OK;
}

// test entries (pure host, no GPU): the HEALPix device functions of bfg_common.cuh -- the SAME source the kernels inline --
// compiled for the CPU, so that the CPU suite can hold them against the oracle over many discs / directions / NSIDE values.
// what = 0: query_disc (a[0] = theta, a[1] = phi, a[2] = radius; out_i [cap] pixels, out_i[cap] = count)
//        1: pix2vec    (idx [n] pixels; out_d [n][3])
//        2: get_interpol / hp.get_interp_weights (a = theta [n], b = phi [n]; out_i [n][4], out_d [n][4])
//        3: ang2pix    (a = theta [n], b = phi [n]; out_i [n])
//        4: ring2nest, 5: nest2ring (idx [n]; out_i [n])
extern "C" int bfg_test_healpix_host(int what, int nside, int64_t n, const int64_t *h_idx, const double *h_a, const double *h_b,
                                     int64_t cap, int64_t *h_out_i, double *h_out_d) {
    BFG_REQUIRE(nside >= 1 && nside <= (1 << 24) && n >= 0, "bad argument");
    const Hpx h(nside);
    if (what == 0) {
        BFG_REQUIRE(h_a && h_out_i && cap >= 0, "null argument");
        const DiscRings d = disc_rings(h, h_a[0], h_a[1], h_a[2]);
        i64 cnt_all = 0;
        for (i64 iz = d.ra; iz <= d.rb; ++iz) {
            i64 start, nr, ip_lo, cnt;
            bool sh;
            disc_ring_span(h, d, iz, start, nr, sh, ip_lo, cnt);
            for (i64 i = 0; i < cnt; ++i) {
                i64 ip = ip_lo + i;
                if (ip >= nr) ip -= nr;
                if (cnt_all < cap) h_out_i[cnt_all] = start + ip;
                ++cnt_all;
            }
        }
        h_out_i[cap] = cnt_all;
        return BFG_OK;
    }
    for (int64_t i = 0; i < n; ++i) {
        if (what == 1) {
            BFG_REQUIRE(h_idx && h_out_d && h_idx[i] >= 0 && h_idx[i] < h.npix, "bad pixel");
            pix2vec(h, h_idx[i], h_out_d[3 * i], h_out_d[3 * i + 1], h_out_d[3 * i + 2]);
        } else if (what == 2) {
            BFG_REQUIRE(h_a && h_b && h_out_i && h_out_d, "null argument");
            i64 pix[4];
            double w[4];
            get_interpol(h, h_a[i], h_b[i], pix, w);
            for (int k = 0; k < 4; ++k) { h_out_i[4 * i + k] = pix[k]; h_out_d[4 * i + k] = w[k]; }
        } else if (what == 3) {
            BFG_REQUIRE(h_a && h_b && h_out_i, "null argument");
            h_out_i[i] = ang2pix_ring(h, h_a[i], h_b[i]);
        } else if (what == 4 || what == 5) {
            BFG_REQUIRE(h_idx && h_out_i && h_idx[i] >= 0 && h_idx[i] < h.npix, "bad pixel");
            BFG_REQUIRE((nside & (nside - 1)) == 0, "NESTED needs a power-of-two nside");
            h_out_i[i] = (what == 4) ? ring2nest(h, h_idx[i]) : nest2ring(h, h_idx[i]);
        } else {
            BFG_REQUIRE(false, "what must be 0..5");
        }
    }
    return BFG_OK;
}

// test entry (pure host, no GPU): row_at_r2 -- the lean read-out of the default grid / particle / exact shell loops: table value at a
// SQUARED radius from the blended row, cell coordinate u = log2(r^2) uA + uB with the table-driven log2 -- from the kernels' own
// source.  offset = ln(1/a) - ln R_com (shells), -ln R_com (R_delta-sampled tables) or 0: what the kernels fold into uB.
// h_out[i] = value (before any exp), h_ok[i] = 0 outside [r0, r1].
extern "C" int bfg_test_row_at_r2_host(int ndim, const int64_t *shape, const double *const *h_axes, const double *h_values, int flags,
                                       double lnz, double lnM, const double *h_extras, double offset, int64_t n, const double *h_r2,
                                       double *h_out, int *h_ok) {
    BFG_REQUIRE(shape && h_axes && h_values && (n == 0 || (h_r2 && h_out && h_ok)), "null argument");
    BFG_REQUIRE(ndim >= 3 && ndim <= BFG_MAX_TABLE_DIM && (ndim == 3 || h_extras), "bad table / extras");
    TableView T;
    host_table_view(ndim, shape, h_axes, h_values, flags, T);
    BFG_REQUIRE(T.uniform_r, "row_at_r2 needs a uniform ln r axis");
    const RowBlender B(T, lnz, lnM, h_extras);
    // the "shared memory" of the host build: [row (NR doubles)][log2 table]
    std::vector<double> arena((size_t)B.NR + 2 * BFG_LOG2_TAB);
    for (int k = 0; k < B.NR; ++k) arena[(size_t)k] = B.node(T, k);
    fill_log2_table(reinterpret_cast<double2 *>(arena.data() + B.NR));
    bfg_host_smem = reinterpret_cast<const char *>(arena.data());
    RowLookup rl;
    rl.uA = 0.34657359027997264 * T.inv_dr;
    rl.uB = (offset - T.r0) * T.inv_dr;
    rl.uMax = (double)(T.n[2] - 1);
    rl.nrm2 = T.n[2] - 2;
    rl.row_s = 0;
    rl.l2_s = (unsigned)(sizeof(double) * (size_t)B.NR);
    for (int64_t i = 0; i < n; ++i) {
        bool ok;
        const double v = row_at_r2(rl, h_r2[i], ok);
        h_out[i] = (ok && B.valid) ? v : std::numeric_limits<double>::quiet_NaN();
        h_ok[i] = (ok && B.valid) ? 1 : 0;
    }
    bfg_host_smem = nullptr;
    return BFG_OK;
}

// test entry (pure host, no GPU): fast_log2 -- table + degree-5 series, the log2 of every non-lean read-out -- on the CPU with the
// table the device gets
extern "C" int bfg_test_fast_log2_host(int64_t n, const double *h_x, double *h_out) {
    BFG_REQUIRE(n >= 0 && (n == 0 || (h_x && h_out)), "bad argument");
    double2 tab[BFG_LOG2_TAB];
    fill_log2_table(tab);
    for (int64_t i = 0; i < n; ++i) h_out[i] = fast_log2(h_x[i], tab);
    return BFG_OK;
}

// test entry: out[i] = fast_log2(x[i])
__global__ void k_fast_log2(i64 n, const double2 *__restrict__ g_tab, const double *__restrict__ x, double *__restrict__ out) {
    __shared__ double2 tab[BFG_LOG2_TAB];
    load_log2_table(tab, g_tab);
    __syncthreads();
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) out[i] = fast_log2(x[i], tab);
}

extern "C" int bfg_test_fast_log2(int64_t n, const double *d_x, double *d_out, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(d_x && d_out, "null argument");
    const double2 *g_tab = nullptr;
    if (int rc = get_log2_table(&g_tab)) return rc;
    if (n == 0) return BFG_OK;
    k_fast_log2<<<(int)std::min<i64>((n + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(n, g_tab, d_x, d_out);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_abi_version(void) { return BFG_ABI_VERSION; }
extern "C" const char *bfg_last_error(void) { return bfg::g_err; }

extern "C" int bfg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int bfg_device_info(int device, int *sm_count, int64_t *mem_total, int64_t *mem_free) {
    BFG_ENTRY();
    cudaDeviceProp p;
    BFG_CUDA_OK(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    int cur = 0;
    BFG_CUDA_OK(cudaGetDevice(&cur));
    BFG_CUDA_OK(cudaSetDevice(device));
    size_t f = 0, t = 0;
    BFG_CUDA_OK(cudaMemGetInfo(&f, &t));
    BFG_CUDA_OK(cudaSetDevice(cur));
    if (mem_total) *mem_total = (int64_t)t;
    if (mem_free) *mem_free = (int64_t)f;
    return BFG_OK;
}

// ------------------------------------------------------------------------------------------------ tables
extern "C" int bfg_table_create(bfg_table **out, int ndim, const int64_t *shape, const double *const *h_axes,
                                const double *h_values, int flags, int device) {
    BFG_ENTRY();
    BFG_REQUIRE(out && shape && h_axes && h_values, "null argument");
    BFG_REQUIRE(ndim >= 3 && ndim <= BFG_MAX_TABLE_DIM, "ndim must be 3..6 (ln(1+z), ln M, ln r, extras...)");
    for (int d = 0; d < ndim; ++d) {
        BFG_REQUIRE(shape[d] >= 2 && shape[d] < (1 << 30), "every axis needs >= 2 nodes");
        for (int64_t i = 1; i < shape[d]; ++i) BFG_REQUIRE(h_axes[d][i] > h_axes[d][i - 1], "axes must be strictly ascending");
    }
    int cur = 0;
    BFG_CUDA_OK(cudaGetDevice(&cur));
    BFG_CUDA_OK(cudaSetDevice(device));
    bfg_table *t = new bfg_table();
    memset(t, 0, sizeof(*t));
    t->device = device;
    i64 total = 1;
    for (int d = ndim - 1; d >= 0; --d) {
        t->shape[d] = shape[d];
        t->view.n[d] = (int)shape[d];
        t->view.stride[d] = total;
        total *= shape[d];
    }
    t->view.ndim = ndim;
    t->view.flags = flags;
    int rc = BFG_OK;
    for (int d = 0; d < ndim && rc == BFG_OK; ++d) {
        if (cudaMalloc(&t->d_axes[d], sizeof(double) * shape[d]) != cudaSuccess ||
            cudaMemcpy(t->d_axes[d], h_axes[d], sizeof(double) * shape[d], cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("bfg_table_create: axis upload failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = BFG_ERR_CUDA;
        }
        t->view.ax[d] = t->d_axes[d];
    }
    if (rc == BFG_OK) {
        if (cudaMalloc(&t->d_values, sizeof(double) * total) != cudaSuccess ||
            cudaMemcpy(t->d_values, h_values, sizeof(double) * total, cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("bfg_table_create: value upload failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = BFG_ERR_CUDA;
        }
        t->view.v = t->d_values;
    }
    describe_radial_axis(h_axes[2], shape[2], t->view);
    cudaSetDevice(cur);
    if (rc != BFG_OK) { bfg_table_destroy(t); return rc; }
    *out = t;
    return BFG_OK;
}

extern "C" int bfg_table_destroy(bfg_table *t) {
    BFG_ENTRY();
    if (!t) return BFG_OK;
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(t->device);
    for (int d = 0; d < BFG_MAX_TABLE_DIM; ++d)
        if (t->d_axes[d]) cudaFree(t->d_axes[d]);
    if (t->d_values) cudaFree(t->d_values);
    cudaSetDevice(cur);
    delete t;
    return BFG_OK;
}

extern "C" int bfg_table_info(const bfg_table *t, int *ndim, int64_t *shape, int *flags, int *device, int *uniform_r) {
    BFG_ENTRY();
    BFG_REQUIRE(t, "null table");
    if (ndim) *ndim = t->view.ndim;
    if (shape) for (int d = 0; d < t->view.ndim; ++d) shape[d] = t->shape[d];
    if (flags) *flags = t->view.flags;
    if (device) *device = t->device;
    if (uniform_r) *uniform_r = t->view.uniform_r;
    return BFG_OK;
}

// read-out test kernel: one block, blends the row for (lnz, lnM, extras) then evaluates n radial points
struct ExtrasArg { double e[BFG_MAX_TABLE_DIM]; };

__global__ void k_table_readout(TableView T, double lnz, double lnM, ExtrasArg ex, i64 n, const double *__restrict__ x,
                                double *__restrict__ out) {
    extern __shared__ double row[];
    bool valid;
    blend_row(T, lnz, lnM, ex.e, row, valid);
    __syncthreads();
    for (i64 i = threadIdx.x; i < n; i += blockDim.x) {
        double v = T.uniform_r ? row_lookup<true>(T, row, x[i]) : row_lookup<false>(T, row, x[i]);
        if (!valid) v = CUDART_NAN;
        if (T.flags & BFG_TABLE_LOG_VALUES) v = exp(v);
        out[i] = v;
    }
}

extern "C" int bfg_table_readout(const bfg_table *t, double lnz, double lnM, const double *h_extras, int64_t n,
                                 const double *d_x, double *d_out, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(t && d_x && d_out, "null argument");
    ExtrasArg ex;
    memset(&ex, 0, sizeof(ex));
    for (int d = 3; d < t->view.ndim; ++d) {
        BFG_REQUIRE(h_extras, "table has extra axes but no extras given");
        ex.e[d - 3] = h_extras[d - 3];
    }
    size_t smem = sizeof(double) * t->view.n[2];
    BFG_CUDA_OK(cudaFuncSetAttribute(k_table_readout, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_table_readout<<<1, 256, smem, (cudaStream_t)stream>>>(t->view, lnz, lnM, ex, n, d_x, d_out);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

// test entry (pure host, no GPU): the read-out every halo-loop kernel performs -- blend the 2^(ndim-1) corner rows of the
// non-radial axes into one radial row (RowBlender), then interpolate along ln r (row_lookup) -- with the SAME source on HOST
// buffers: out[i] = table(lnz, lnM, x[i], extras...) with RegularGridInterpolator(bounds_error=False, fill_value=nan) semantics,
// exp() applied for BFG_TABLE_LOG_VALUES tables (Tabulate.py:319).  force_search != 0 takes the non-uniform radial branch.
extern "C" int bfg_test_table_readout_host(int ndim, const int64_t *shape, const double *const *h_axes, const double *h_values,
                                           int flags, int force_search, double lnz, double lnM, const double *h_extras, int64_t n,
                                           const double *h_x, double *h_out) {
    BFG_REQUIRE(shape && h_axes && h_values && (n == 0 || (h_x && h_out)), "null argument");
    BFG_REQUIRE(ndim >= 3 && ndim <= BFG_MAX_TABLE_DIM, "ndim must be 3..6");
    BFG_REQUIRE(ndim == 3 || h_extras, "table has extra axes but no extras given");
    for (int d = 0; d < ndim; ++d) BFG_REQUIRE(shape[d] >= 2, "every axis needs >= 2 nodes");
    TableView T;
    host_table_view(ndim, shape, h_axes, h_values, flags, T);
    if (force_search) T.uniform_r = 0;
    const RowBlender B(T, lnz, lnM, h_extras);
    std::vector<double> row((size_t)B.NR);
    for (int k = 0; k < B.NR; ++k) row[(size_t)k] = B.node(T, k);
    for (int64_t i = 0; i < n; ++i) {
        double v = T.uniform_r ? row_lookup<true>(T, row.data(), h_x[i]) : row_lookup<false>(T, row.data(), h_x[i]);
        if (!B.valid) v = std::numeric_limits<double>::quiet_NaN();
        if (T.flags & BFG_TABLE_LOG_VALUES) v = exp(v);
        h_out[i] = v;
    }
    return BFG_OK;
}

// ------------------------------------------------------------------------------------------------ utilities
__global__ void k_sum_f64(const double *__restrict__ x, i64 n, double *out) {
    double acc = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) acc += x[i];
    acc = warp_sum(acc);
    __shared__ double part[32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) part[w] = acc;
    __syncthreads();
    if (w == 0) {
        acc = (lane < (blockDim.x >> 5)) ? part[lane] : 0.0;
        acc = warp_sum(acc);
        if (lane == 0) atomicAdd(out, acc);
    }
}

extern "C" int bfg_sum_f64(const double *d_x, int64_t n, double *d_out, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(d_out && (d_x || n == 0), "null argument");
    BFG_CUDA_OK(cudaMemsetAsync(d_out, 0, sizeof(double), (cudaStream_t)stream));
    if (n > 0) {
        int blocks = (int)std::min<i64>((n + 1023) / 1024, 148 * 8);
        k_sum_f64<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_x, n, d_out);
        BFG_CUDA_OK(cudaGetLastError());
    }
    return BFG_OK;
}

__global__ void k_transpose_offsets(const double *__restrict__ in, double *__restrict__ out, i64 n, int nc) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        for (int c = 0; c < nc; ++c) out[i * nc + c] = in[(i64)c * n + i];
}

extern "C" int bfg_transpose_offsets(const double *d_in, double *d_out, int64_t n, int ncomp, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(d_in && d_out && ncomp >= 1 && ncomp <= 64, "bad argument");
    if (n == 0) return BFG_OK;
    int blocks = (int)std::min<i64>((n + 255) / 256, 148 * 16);
    k_transpose_offsets<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_in, d_out, n, ncomp);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

// ------------------------------------------------------------------------------------------------ peer memory (IPC)
extern "C" int bfg_shared_alloc(void **d_ptr, int64_t bytes, int device) {
    BFG_ENTRY();
    BFG_REQUIRE(d_ptr && bytes >= 0, "bad argument");
    int cur = 0;
    BFG_CUDA_OK(cudaGetDevice(&cur));
    BFG_CUDA_OK(cudaSetDevice(device));
    cudaError_t e = cudaMalloc(d_ptr, (size_t)(bytes > 0 ? bytes : 8));
    cudaSetDevice(cur);
    if (e != cudaSuccess) { set_error("bfg_shared_alloc: %s", cudaGetErrorString(e)); return BFG_ERR_NOMEM; }
    return BFG_OK;
}

extern "C" int bfg_shared_free(void *d_ptr) {
    BFG_ENTRY();
    if (d_ptr) BFG_CUDA_OK(cudaFree(d_ptr));
    return BFG_OK;
}

extern "C" int bfg_ipc_export(const void *d_ptr, unsigned char *handle64) {
    BFG_ENTRY();
    BFG_REQUIRE(d_ptr && handle64, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t hdl;
    BFG_CUDA_OK(cudaIpcGetMemHandle(&hdl, const_cast<void *>(d_ptr)));
    memcpy(handle64, &hdl, 64);
    return BFG_OK;
}

extern "C" int bfg_ipc_import(const unsigned char *handle64, void **d_peer_ptr) {
    BFG_ENTRY();
    BFG_REQUIRE(handle64 && d_peer_ptr, "null argument");
    cudaIpcMemHandle_t hdl;
    memcpy(&hdl, handle64, 64);
    BFG_CUDA_OK(cudaIpcOpenMemHandle(d_peer_ptr, hdl, cudaIpcMemLazyEnablePeerAccess));
    return BFG_OK;
}

extern "C" int bfg_ipc_close(void *d_peer_ptr) {
    BFG_ENTRY();
    if (d_peer_ptr) BFG_CUDA_OK(cudaIpcCloseMemHandle(d_peer_ptr));
    return BFG_OK;
}

// ------------------------------------------------------------------------------------------------ shared host maps
// Page-locks a host range that several processes of the box have mapped (memfd / POSIX shared memory), so each rank
// can copy its owned slice of a result map straight into ONE host map at full PCIe speed (parallel.SharedHostMaps).
extern "C" int bfg_host_register(void *h_ptr, int64_t bytes) {
    BFG_ENTRY();
    BFG_REQUIRE(h_ptr && bytes > 0, "bad argument");
    BFG_CUDA_OK(cudaHostRegister(h_ptr, (size_t)bytes, cudaHostRegisterPortable));
    return BFG_OK;
}

extern "C" int bfg_host_unregister(void *h_ptr) {
    BFG_ENTRY();
    if (h_ptr) BFG_CUDA_OK(cudaHostUnregister(h_ptr));
    return BFG_OK;
}

extern "C" int bfg_copy_to_host_async(void *h_dst, const void *d_src, int64_t bytes, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(bytes >= 0 && (bytes == 0 || (h_dst && d_src)), "bad argument");
    if (bytes) BFG_CUDA_OK(cudaMemcpyAsync(h_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return BFG_OK;
}
