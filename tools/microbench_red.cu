// tools/microbench_red.cu -- which path carries the fp64 scatter-add of k_shell_halos fastest?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench_red tools/microbench_red.cu
//   gpurun -- ./tools/microbench_red
// ncu of the v8 / v9 halo loop shows the SM -> L2 request path (l1tex__m_l1tex2xbar_req_cycles_active) 73 % busy while the FP64
// pipe is half idle: the loop is bound by how fast RED.E.ADD.F64 sectors leave the SM.  This benchmark reproduces the access
// pattern without the arithmetic -- spans of `len` consecutive doubles at pseudo-random 8-byte-aligned (RED) or 16-byte-aligned
// (bulk) positions of three component planes -- and times
//   mode 0: RED.E.ADD.F64 from lane groups of GW lanes (what the kernel does today)
//   mode 1: the warp stages the span in shared memory and ONE thread issues cp.reduce.async.bulk (.add.f64) per component.
// Output: one line per (mode, GW / chunk, span length, footprint): Gupdates/s (1 update = 3 doubles) and GB/s.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

// spans handed out in "sky order": consecutive span ids are neighbours in memory (like the sky-sorted halos), with jitter
__device__ __forceinline__ int64_t span_base(uint64_t id, int64_t n, int len, int64_t n_spans, int align = 1) {
    const int64_t stride = (n - len - 64) / n_spans;
    int64_t b = (int64_t)id * stride + (int64_t)(mix(id) % (uint64_t)(8 * len + 1));
    if (b > n - len - 64) b = n - len - 64;
    return b & ~(int64_t)(align - 1);          // align = 1, 4 (32-byte sector) or 16 (128-byte line) pixels
}

template <int GW>
__global__ void __launch_bounds__(128, 7) k_red(double *out, int64_t n, int len, int64_t n_spans, unsigned long long *queue,
                                                 int align) {
    const int lane = threadIdx.x & 31, li = lane & (GW - 1), gi = lane / GW;
    constexpr int NG = 32 / GW;
    for (;;) {
        uint64_t w = 0;
        if (lane == 0) w = atomicAdd(queue, (unsigned long long)NG);
        w = __shfl_sync(0xffffffffu, w, 0);
        if ((int64_t)w >= n_spans) break;
        const uint64_t id = w + gi;
        if ((int64_t)id >= n_spans) continue;
        double *p = out + span_base(id, n, len, n_spans, align);
        const double v = 1e-9 * (double)(lane + 1);
        for (int i = li; i < len; i += GW) {
            atomicAdd(p + i, v);
            atomicAdd(p + n + i, v);
            atomicAdd(p + 2 * n + i, v);
        }
    }
}

// one warp per span: stage 3 x len doubles in shared memory, one bulk reduce per component; two buffers in flight
template <int MAXLEN>
__global__ void __launch_bounds__(128, 7) k_bulk(double *out, int64_t n, int len, int64_t n_spans, unsigned long long *queue) {
    __shared__ __align__(128) double buf[4][2][3][MAXLEN];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int which = 0;
    for (;;) {
        uint64_t w = 0;
        if (lane == 0) w = atomicAdd(queue, 1ULL);
        w = __shfl_sync(0xffffffffu, w, 0);
        if ((int64_t)w >= n_spans) break;
        int64_t b = span_base(w, n, len, n_spans) & ~(int64_t)1;       // 16-byte aligned destination
        const double v = 1e-9 * (double)(lane + 1);
        // the buffer used two spans ago must have been read by the bulk engine
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        double *s = &buf[wid][which][0][0];
        for (int i = lane; i < len; i += 32) { s[i] = v; s[MAXLEN + i] = v; s[2 * MAXLEN + i] = v; }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            const unsigned bytes = (unsigned)len * 8u;
            for (int c = 0; c < 3; ++c) {
                const unsigned sa = (unsigned)__cvta_generic_to_shared(s + c * MAXLEN);
                double *g = out + (int64_t)c * n + b;
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                             :: "l"(g), "r"(sa), "r"(bytes) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        which ^= 1;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
static float time_ms(F launch, unsigned long long *queue) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaMemset(queue, 0, 8));
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best) best = ms;
    }
    return best;
}

int main() {
    const int64_t n = 201326592;                 // pixels of an NSIDE = 4096 map; 3 planes = 4.8 GB
    double *out; unsigned long long *queue;
    CK(cudaMalloc(&out, 3 * n * sizeof(double)));
    CK(cudaMemset(out, 0, 3 * n * sizeof(double)));
    CK(cudaMalloc(&queue, 8));
    const int grid = 148 * 7;
    const int lens[] = {48, 112, 256};
    const int64_t total_updates = 4000000000LL;  // ~ a quarter of the headline step
    for (int len : lens) {
        const int64_t n_spans = total_updates / len;
        float ms;
        for (int align : {1, 4, 16}) {
            ms = time_ms([&] { k_red<8><<<grid, 128>>>(out, n, len, n_spans, queue, align); }, queue);
            printf("RED  GW=8   align=%2d len=%3d  %7.2f ms  %6.1f Gupd/s  %7.1f GB/s payload\n", align, len, ms, n_spans * len / ms * 1e-6, n_spans * len * 24.0 / ms * 1e-6);
            ms = time_ms([&] { k_red<16><<<grid, 128>>>(out, n, len, n_spans, queue, align); }, queue);
            printf("RED  GW=16  align=%2d len=%3d  %7.2f ms  %6.1f Gupd/s  %7.1f GB/s payload\n", align, len, ms, n_spans * len / ms * 1e-6, n_spans * len * 24.0 / ms * 1e-6);
            ms = time_ms([&] { k_red<32><<<grid, 128>>>(out, n, len, n_spans, queue, align); }, queue);
            printf("RED  GW=32  align=%2d len=%3d  %7.2f ms  %6.1f Gupd/s  %7.1f GB/s payload\n", align, len, ms, n_spans * len / ms * 1e-6, n_spans * len * 24.0 / ms * 1e-6);
        }
        ms = time_ms([&] { k_bulk<256><<<grid, 128>>>(out, n, len, n_spans, queue); }, queue);
        printf("BULK        len=%3d  %7.2f ms  %6.1f Gupd/s  %7.1f GB/s payload\n", len, ms, n_spans * len / ms * 1e-6, n_spans * len * 24.0 / ms * 1e-6);
        fflush(stdout);
    }
    // correctness of the bulk path: every element of a small region must equal the RED result
    {
        const int64_t m = 1 << 20;
        double *a, *b;
        CK(cudaMalloc(&a, 3 * m * 8)); CK(cudaMalloc(&b, 3 * m * 8));
        CK(cudaMemset(a, 0, 3 * m * 8)); CK(cudaMemset(b, 0, 3 * m * 8));
        const int len = 104; const int64_t ns = 40000;
        CK(cudaMemset(queue, 0, 8));
        k_red<32><<<grid, 128>>>(a, m, len, ns, queue, 1);
        CK(cudaMemset(queue, 0, 8));
        k_bulk<256><<<grid, 128>>>(b, m, len, ns, queue);
        CK(cudaDeviceSynchronize());
        double *ha = (double *)malloc(3 * m * 8), *hb = (double *)malloc(3 * m * 8);
        CK(cudaMemcpy(ha, a, 3 * m * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hb, b, 3 * m * 8, cudaMemcpyDeviceToHost));
        // the RED kernel uses 8-byte-aligned bases, the bulk kernel rounds them down to 16 bytes: compare plane sums instead
        double sa = 0, sb = 0;
        for (int64_t i = 0; i < 3 * m; ++i) { sa += ha[i]; sb += hb[i]; }
        printf("check: sum RED %.12e  sum BULK %.12e  rel diff %.3e\n", sa, sb, (sa - sb) / sa);
    }
    return 0;
}
