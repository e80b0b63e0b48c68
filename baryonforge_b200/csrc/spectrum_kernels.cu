// spectrum_kernels.cu -- the measurement step that follows BaryonifySnapshot.process() in the reference's workflow, kept in
// HBM.  The reference has no library function for it; the algorithm is the cell code of
// examples/10_Reproduce_Schneider_deltaPk.ipynb (cited as nb10:cell):
//
//   k_deposit_folded : `numba_histogram3d(Part % Lbox, bins = Ngrd, 0, Lbox)`, Lbox = L / factor   (nb10:1, nb10:15)
//   k_power_bins     : `(conj(F) * F).real` of the FFT of that grid, summed per k-shell
//                      kinds = floor((|k| - kbins[0]) / (kbins[1] - kbins[0])), bincount(kinds, weights)   (nb10:12, nb10:15)
//
// The FFT itself is a plain library transform (cuFFT D2Z, loaded lazily with dlopen so that libbfg_b200.so carries no
// load-time dependency on it); only the N x N x (N/2+1) half spectrum is computed and the shell sums count every mode of
// the half spectrum whose mirror image lies outside it twice.  |k| and the shell index are evaluated with the notebook's
// own operation order in round-to-nearest fp64 (no FMA contraction), so shell membership is bit-identical to numpy's.
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <dlfcn.h>
#include <cufft.h>
#include "bfg_common.cuh"

using namespace bfg;

namespace {

// np.mod for float64 (floored remainder, sign of the divisor): what `Part % Lbox` evaluates per element
__device__ __forceinline__ double np_mod(double a, double b) {
    double m = fmod(a, b);
    if (m != 0.0) {
        if ((b < 0.0) != (m < 0.0)) m += b;
    } else {
        m = copysign(0.0, b);
    }
    return m;
}

// int((p - 0) / width) of the folded coordinate.  NaN / inf coordinates (a particle exactly on a halo centre comes back as
// NaN, SnapshotRunner.py:253-260) are dropped; a folded value that rounds up to Lfold itself (x = -1e-20) is an out-of-bounds
// write in the notebook's numba loop and goes to the last cell here.
__device__ __forceinline__ i64 folded_cell(double x, double Lfold, double width, i64 N) {
    if (!isfinite(x)) return -1;
    const double r = np_mod(x, Lfold);
    i64 c = (i64)__ddiv_rn(r, width);
    return c > N - 1 ? N - 1 : (c < 0 ? 0 : c);
}

__global__ void __launch_bounds__(256)
k_deposit_folded(i64 n, const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ z,
                 double Lfold, i64 N, double *__restrict__ grid, unsigned long long *__restrict__ n_dropped) {
    const double width = __ddiv_rn(Lfold - 0.0, (double)N);                    // (max_vals - min_vals) / bins
    i64 dropped = 0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const i64 cx = folded_cell(x[i], Lfold, width, N), cy = folded_cell(y[i], Lfold, width, N),
                  cz = folded_cell(z[i], Lfold, width, N);
        if (cx < 0 || cy < 0 || cz < 0) { ++dropped; continue; }
        red_add(grid + (cx * N + cy) * N + cz, 1.0);
    }
    if (n_dropped) {
        dropped = warp_sum_i64(dropped);
        if ((threadIdx.x & 31) == 0 && dropped) atomicAdd(n_dropped, (unsigned long long)dropped);
    }
}

// k_deposit_folded straight from the CELL-ORDERED particles of bfg_snap_build_cells + their accumulated offsets (what
// k_snap_apply_deposit does for the NGP grid): position = wrap_once(xs + tot) exactly as bfg_snap_apply computes it
// (SnapshotRunner.py:263-273), then the folded cell.  Lanes walk the particles of one cell-list cell, so their REDs fall
// into neighbouring grid cells instead of random DRAM sectors (tests/test_gpu_spectrum.py: the same grid, bit for bit).
__global__ void __launch_bounds__(256)
k_apply_deposit_folded(i64 n, const double *__restrict__ xs, const double *__restrict__ ys, const double *__restrict__ zs,
                       const double *__restrict__ tot, double L, double Lfold, i64 N, double *__restrict__ grid,
                       unsigned long long *__restrict__ n_dropped) {
    const double width = __ddiv_rn(Lfold - 0.0, (double)N);
    i64 dropped = 0;
    for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (i64)gridDim.x * blockDim.x) {
        double q[3] = {xs[p] + tot[p], ys[p] + tot[n + p], zs[p] + tot[2 * n + p]};
        i64 c[3];
        bool ok = true;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (q[k] > L) q[k] -= L;                         // the single wrap of SnapshotRunner.py:272-273
            if (q[k] < 0) q[k] += L;
            c[k] = folded_cell(q[k], Lfold, width, N);
            ok = ok && (c[k] >= 0);
        }
        if (!ok) { ++dropped; continue; }
        red_add(grid + (c[0] * N + c[1]) * N + c[2], 1.0);
    }
    if (n_dropped) {
        dropped = warp_sum_i64(dropped);
        if ((threadIdx.x & 31) == 0 && dropped) atomicAdd(n_dropped, (unsigned long long)dropped);
    }
}

// One warp per (a, b) row of the half spectrum, lanes along c (coalesced 16-byte loads).  Along a row |k| never decreases,
// so equal shell indices sit in adjacent lanes: a segmented shuffle reduction leaves one shared-memory atomic per
// (row chunk, shell) instead of one per mode; per-CTA shell sums then go out with one RED per non-empty shell.
constexpr int PB_THREADS = 256;

__global__ void __launch_bounds__(PB_THREADS)
k_power_bins(int N, const double2 *__restrict__ spec, const double *__restrict__ klin, double k0, double dk, int Nk,
             double *__restrict__ pk_sum, double *__restrict__ k_sum, unsigned long long *__restrict__ count) {
    extern __shared__ double s_acc[];                       // [Nk] power, [Nk] k, [Nk] counts
    double *s_pk = s_acc, *s_k = s_acc + Nk;
    unsigned long long *s_n = (unsigned long long *)(s_acc + 2 * (size_t)Nk);
    for (int i = threadIdx.x; i < Nk; i += blockDim.x) { s_pk[i] = 0.0; s_k[i] = 0.0; s_n[i] = 0ULL; }
    __syncthreads();

    const int NH = N / 2 + 1;
    const int lane = threadIdx.x & 31;
    const i64 n_rows = (i64)N * N;
    const i64 warp0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((i64)gridDim.x * blockDim.x) >> 5;
    const unsigned FULL = 0xffffffffu;
    for (i64 row = warp0; row < n_rows; row += n_warps) {
        const int a = (int)(row / N), b = (int)(row - (i64)a * N);
        const double ka = klin[a], kb = klin[b];
        const double ka2 = __dmul_rn(ka, ka), kb2 = __dmul_rn(kb, kb);
        const double2 *__restrict__ src = spec + row * NH;
        for (int c0 = 0; c0 < NH; c0 += 32) {
            const int c = c0 + lane;
            int key = 0x7fffffff;                           // lanes beyond the row end
            double pw = 0.0, kv = 0.0;
            unsigned long long w = 0ULL;
            if (c < NH) {
                const double kc = klin[c];
                // klin[:, None, None]**2 + klin[None, None, :]**2 + klin[None, :, None]**2   (axis 0, axis 2, axis 1)
                const double ksq = __dadd_rn(__dadd_rn(ka2, __dmul_rn(kc, kc)), kb2);
                kv = __dsqrt_rn(ksq);
                const double fb = floor(__ddiv_rn(__dadd_rn(kv, -k0), dk));
                key = (fb < 0.0) ? -1 : (fb >= (double)Nk ? Nk : (int)fb);      // NaN compares false twice -> (int)NaN; masked below
                if (!(fb == fb)) key = Nk;
                const double2 f = spec ? __ldg(src + c) : make_double2(0.0, 0.0);      // NULL spectrum: mode counting only
                const double p1 = __dadd_rn(__dmul_rn(f.x, f.x), __dmul_rn(f.y, f.y));   // (conj(F) * F).real
                const bool twice = (c != 0) && (2 * c != N);                    // the mirror mode is not in the half spectrum
                w = twice ? 2ULL : 1ULL;
                pw = twice ? 2.0 * p1 : p1;
                kv = twice ? 2.0 * kv : kv;
            }
            // segmented reduction over runs of equal keys (keys are sorted along the lanes)
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int key2 = __shfl_down_sync(FULL, key, off);
                const double pw2 = __shfl_down_sync(FULL, pw, off), kv2 = __shfl_down_sync(FULL, kv, off);
                const unsigned long long w2 = __shfl_down_sync(FULL, w, off);
                if (lane + off < 32 && key2 == key) { pw += pw2; kv += kv2; w += w2; }
            }
            const int key_prev = __shfl_up_sync(FULL, key, 1);
            const bool head = (lane == 0) || (key_prev != key);
            if (head && key >= 0 && key < Nk) {
                atomicAdd(s_pk + key, pw);
                atomicAdd(s_k + key, kv);
                atomicAdd(s_n + key, w);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Nk; i += blockDim.x) {
        if (s_n[i]) {
            red_add(pk_sum + i, s_pk[i]);
            red_add(k_sum + i, s_k[i]);
            atomicAdd(count + i, s_n[i]);
        }
    }
}

// ---- cuFFT, resolved at first use -----------------------------------------------------------------------------------
struct CufftApi {
    cufftResult (*Plan3d)(cufftHandle *, int, int, int, cufftType) = nullptr;
    cufftResult (*SetStream)(cufftHandle, cudaStream_t) = nullptr;
    cufftResult (*ExecD2Z)(cufftHandle, cufftDoubleReal *, cufftDoubleComplex *) = nullptr;
    cufftResult (*Destroy)(cufftHandle) = nullptr;
    bool ok = false;
};

const CufftApi &cufft_api() {
    static CufftApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libcufft.so.11", "libcufft.so", "/usr/local/cuda/lib64/libcufft.so.11",
                               "/usr/local/cuda/lib64/libcufft.so", "libcufft.so.12", "libcufft.so.10"};
        void *h = nullptr;
        if (const char *env = getenv("BFG_CUFFT_LIB")) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        for (size_t i = 0; !h && i < sizeof(names) / sizeof(names[0]); ++i) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.Plan3d = (decltype(api.Plan3d))dlsym(h, "cufftPlan3d");
        api.SetStream = (decltype(api.SetStream))dlsym(h, "cufftSetStream");
        api.ExecD2Z = (decltype(api.ExecD2Z))dlsym(h, "cufftExecD2Z");
        api.Destroy = (decltype(api.Destroy))dlsym(h, "cufftDestroy");
        api.ok = api.Plan3d && api.SetStream && api.ExecD2Z && api.Destroy;
    });
    return api;
}

// One cached D2Z plan per (device, N): cufftPlan3d allocates its work area with cudaMalloc, which is not stream-ordered.
struct PlanSlot { int device = -1; int N = 0; cufftHandle plan = 0; };
std::mutex g_plan_mutex;
PlanSlot g_plans[16];

// Caller holds g_plan_mutex from here until the plan's Exec has been issued: a plan is bound to one stream at a time, and a
// full cache evicts slot 0, which no other thread can be using while the lock is held.
int get_plan_locked(int N, cufftHandle *out) {
    const CufftApi &api = cufft_api();
    if (!api.ok) {
        set_error("bfg_grid_power_spectrum: cuFFT (libcufft.so.11) could not be loaded; set BFG_CUFFT_LIB");
        return BFG_ERR_UNSUPPORTED;
    }
    int dev = 0;
    BFG_CUDA_OK(cudaGetDevice(&dev));
    PlanSlot *free_slot = nullptr;
    for (PlanSlot &s : g_plans) {
        if (s.N == N && s.device == dev) { *out = s.plan; return BFG_OK; }
        if (!free_slot && s.N == 0) free_slot = &s;
    }
    if (!free_slot) {                                       // cache full: recycle the first slot
        free_slot = &g_plans[0];
        api.Destroy(free_slot->plan);
        *free_slot = PlanSlot();
    }
    cufftHandle plan = 0;
    const cufftResult r = api.Plan3d(&plan, N, N, N, CUFFT_D2Z);
    if (r != CUFFT_SUCCESS) {
        set_error("bfg_grid_power_spectrum: cufftPlan3d(%d^3, D2Z) failed with cufftResult %d", N, (int)r);
        return r == CUFFT_ALLOC_FAILED ? BFG_ERR_NOMEM : BFG_ERR_CUDA;
    }
    free_slot->device = dev; free_slot->N = N; free_slot->plan = plan;
    *out = plan;
    return BFG_OK;
}

int launch_power_bins(int N, const double2 *spec, const double *d_klin, double k0, double dk, int Nk, double *d_pk_sum,
                      double *d_k_sum, int64_t *d_count, cudaStream_t st) {
    BFG_CUDA_OK(cudaMemsetAsync(d_pk_sum, 0, sizeof(double) * Nk, st));
    BFG_CUDA_OK(cudaMemsetAsync(d_k_sum, 0, sizeof(double) * Nk, st));
    BFG_CUDA_OK(cudaMemsetAsync(d_count, 0, sizeof(int64_t) * Nk, st));
    const size_t smem = (size_t)Nk * 24;
    BFG_CUDA_OK(cudaFuncSetAttribute(k_power_bins, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const i64 rows = (i64)N * N;
    const int warps_per_block = PB_THREADS / 32;
    const int blocks = (int)std::max<i64>(1, std::min<i64>((rows + warps_per_block - 1) / warps_per_block, 148 * 4));
    k_power_bins<<<blocks, PB_THREADS, smem, st>>>(N, spec, d_klin, k0, dk, Nk, d_pk_sum, d_k_sum,
                                                   (unsigned long long *)d_count);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

}  // namespace

extern "C" int bfg_snap_deposit_folded(int64_t n_part, const double *d_x, const double *d_y, const double *d_z, double L_fold,
                                       int64_t n_grid, double *d_grid, int64_t *d_ndropped, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(n_part >= 0 && n_grid >= 1 && n_grid <= 4096 && L_fold > 0, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (d_ndropped) BFG_CUDA_OK(cudaMemsetAsync(d_ndropped, 0, sizeof(int64_t), st));
    if (n_part == 0) return BFG_OK;
    BFG_REQUIRE(d_x && d_y && d_z && d_grid, "null argument");
    const int blocks = (int)std::max<i64>(1, std::min<i64>((n_part + 255) / 256, 148 * 32));
    k_deposit_folded<<<blocks, 256, 0, st>>>(n_part, d_x, d_y, d_z, L_fold, n_grid, d_grid, (unsigned long long *)d_ndropped);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_snap_apply_deposit_folded(int64_t n_part, const double *d_xs, const double *d_ys, const double *d_zs,
                                             const double *d_tot, double L, double L_fold, int64_t n_grid, double *d_grid,
                                             int64_t *d_ndropped, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(n_part >= 0 && n_grid >= 1 && n_grid <= 4096 && L_fold > 0 && L > 0, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (d_ndropped) BFG_CUDA_OK(cudaMemsetAsync(d_ndropped, 0, sizeof(int64_t), st));
    if (n_part == 0) return BFG_OK;
    BFG_REQUIRE(d_xs && d_ys && d_zs && d_tot && d_grid, "null argument");
    const int blocks = (int)std::max<i64>(1, std::min<i64>((n_part + 255) / 256, 148 * 32));
    k_apply_deposit_folded<<<blocks, 256, 0, st>>>(n_part, d_xs, d_ys, d_zs, d_tot, L, L_fold, n_grid, d_grid,
                                                   (unsigned long long *)d_ndropped);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_power_bin_spectrum(int64_t N, const double *d_spec, const double *d_klin, double k0, double dk, int64_t Nk,
                                      double *d_pk_sum, double *d_k_sum, int64_t *d_count, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(N >= 2 && N <= 4096 && Nk >= 1 && Nk <= 8192 && dk > 0, "bad argument");
    BFG_REQUIRE(d_klin && d_pk_sum && d_k_sum && d_count, "null argument");
    return launch_power_bins((int)N, (const double2 *)d_spec, d_klin, k0, dk, (int)Nk, d_pk_sum, d_k_sum, d_count,
                             (cudaStream_t)stream);
}

extern "C" int bfg_grid_power_spectrum(int64_t N, const double *d_grid, const double *d_klin, double k0, double dk, int64_t Nk,
                                       double *d_pk_sum, double *d_k_sum, int64_t *d_count, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(N >= 2 && N <= 4096 && Nk >= 1 && Nk <= 8192 && dk > 0, "bad argument");
    BFG_REQUIRE(d_grid && d_klin && d_pk_sum && d_k_sum && d_count, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = retain_async_pool()) return rc;
    const CufftApi &api = cufft_api();
    StreamScratch s_spec(st);
    int rc = BFG_OK;
    double2 *spec = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_plan_mutex);
        cufftHandle plan = 0;
        if (int prc = get_plan_locked((int)N, &plan)) return prc;
        BFG_CUDA_OK(s_spec.alloc(sizeof(double2) * (size_t)N * N * (N / 2 + 1)));
        spec = s_spec.as<double2>();
        cufftResult r = api.SetStream(plan, st);
        if (r == CUFFT_SUCCESS) r = api.ExecD2Z(plan, const_cast<double *>(d_grid), (cufftDoubleComplex *)spec);
        if (r != CUFFT_SUCCESS) {
            set_error("bfg_grid_power_spectrum: cuFFT D2Z of %lld^3 failed with cufftResult %d", (long long)N, (int)r);
            rc = BFG_ERR_CUDA;
        }
    }
    if (rc == BFG_OK) rc = launch_power_bins((int)N, spec, d_klin, k0, dk, (int)Nk, d_pk_sum, d_k_sum, d_count, st);
    return rc;
}
