// sht_kernels.cu -- spherical-harmonic transforms on the HEALPix RING grid for the C_l measurement that follows BaryonifyShell in
// the reference's workflow: `hp.anafast(map)` (examples/04_Baryonify_Density_Shell.ipynb cell 18).
//
// First run on a B200 in round 2 (tests/test_gpu_harmonics.py green, timings in profiles/r2_anafast.json).  The algorithm is the
// one of oracle/anafast_rings.py (checked on the CPU against the dense definition oracle/anafast_port.py); parity is UNPINNED at
// the healpy boundary (healpy absent).  First version: correctness before speed --
//   k_sht_ring_analysis  : F_m(r) = sum_j f(r, j) exp(-i m phi_j) for every ring r and 0 <= m <= lmax, as a direct sum with exact
//                          seeds for the twiddle recurrence (m phi_j / pi is a rational number with denominator 4 n_r / 4);
//   k_sht_leg_analysis   : a_lm += w sum_r lambda_lm(cos theta_r) F_m(r), one warp per (m, 32 ring pairs), north/south rings
//                          paired through lambda_lm(-x) = (-1)^(l+m) lambda_lm(x), a power-of-two exponent per lane keeps
//                          sin^m(theta) from underflowing, warp-shuffle sum then one RED per (l, warp);
//   k_sht_leg_synthesis  : b_m(r) = sum_l a_lm lambda_lm(cos theta_r), same recursion, accumulators in registers;
//   k_sht_ring_synthesis : f(r, j) = b_0 + 2 Re sum_{m>0} b_m exp(i m phi_j);
//   k_sht_alm2cl         : C_l = (|a_l0|^2 + 2 sum_{m>0} |a_lm|^2) / (2l + 1).
// A ring transform costs n_r (lmax + 1) complex multiply-adds per ring (2.5e12 at NSIDE = 4096); replacing it by batched
// cuFFT for the 2 NSIDE + 1 equatorial rings is the obvious next step once this version is parity-green.
#include <algorithm>
#include "bfg_common.cuh"

using namespace bfg;

namespace {

constexpr int SHT_SCALE_BITS = 256;

// ---- ring geometry -------------------------------------------------------------------------------------------------------
// The per-ring sums are written as __host__ __device__ functions so that the phase conventions can be checked on the CPU
// (bfg_test_sht_ring_host) against the oracle; the kernels only distribute them over threads.
struct RingPhase {
    i64 n;             // pixels in the ring; phi_j / pi = (2 j + odd) / n
    int odd;           // caps: phi_j = (2j + 1) pi / (4 i), n = 4 i;  belt: phi_j = (2j + s) pi / (4 nside), n = 4 nside
};

__host__ __device__ __forceinline__ void sincospi_hd(double x, double *s, double *c) {
#ifdef __CUDA_ARCH__
    sincospi(x, s, c);
#else
    *s = sin(BFG_PI * x);
    *c = cos(BFG_PI * x);
#endif
}

// exp(+i m phi_j) with an exact argument reduction: m (2j + odd) / n is taken modulo 2 in integers
__host__ __device__ __forceinline__ void twiddle(i64 m, i64 j, const RingPhase &g, double &c, double &s) {
    const i64 num = (m * (2 * j + g.odd)) % (2 * g.n);
    sincospi_hd((double)num / (double)g.n, &s, &c);
}

constexpr int RING_THREADS = 256;
constexpr int RESEED = 64;          // twiddle recurrence steps between exact seeds
constexpr int B_CHUNK = 1024;       // b_m staged through shared memory in chunks of this many m

// F_m = sum_j f_j exp(-i m phi_j) of one ring
__host__ __device__ __forceinline__ double2 ring_dft_m(int m, const RingPhase &g, const double *ring) {
    double rc, rs;                                               // one step in j: exp(-i m 2 pi / n)
    {
        const i64 num = (2 * (i64)m) % (2 * g.n);
        sincospi_hd(-(double)num / (double)g.n, &rs, &rc);
    }
    double accr = 0.0, acci = 0.0;
    for (i64 j0 = 0; j0 < g.n; j0 += RESEED) {
        double c, s;
        twiddle(m, j0, g, c, s);
        s = -s;                                                  // exp(-i m phi_j0)
        const i64 j1 = (j0 + RESEED < g.n) ? j0 + RESEED : g.n;
        for (i64 j = j0; j < j1; ++j) {
            const double f = ring[j];
            accr = fma(f, c, accr);
            acci = fma(f, s, acci);
            const double c2 = c * rc - s * rs;
            s = fma(s, rc, c * rs);
            c = c2;
        }
    }
    double2 out; out.x = accr; out.y = acci;
    return out;
}

// acc += sum_{m0 <= m < m1} c_m Re(b_m exp(i m phi_j)), c_0 = 1, c_{m>0} = 2; b is indexed by m - m0; (pc, ps) = exp(i phi_j)
__host__ __device__ __forceinline__ void ring_synth_chunk(i64 j, const RingPhase &g, int m0, int m1, const double2 *b, double pc,
                                                          double ps, double &acc) {
    for (int ms = m0; ms < m1; ms += RESEED) {
        double c, s;
        twiddle(ms, j, g, c, s);                                 // exp(i ms phi_j), exact seed
        const int me = (ms + RESEED < m1) ? ms + RESEED : m1;
        for (int m = ms; m < me; ++m) {
            const double2 bm = b[m - m0];
            const double w = (m == 0) ? 1.0 : 2.0;
            acc = fma(w, fma(bm.x, c, -bm.y * s), acc);
            const double c2 = c * pc - s * ps;
            s = fma(s, pc, c * ps);
            c = c2;
        }
    }
}

__device__ __forceinline__ RingPhase ring_phase(const Hpx &h, i64 ring, i64 &start) {
    RingPhase g;
    bool shifted;
    ring_info(h, ring, start, g.n, shifted);
    g.odd = shifted ? 1 : 0;
    return g;
}

// F[m][ring-1] = sum_j f_j exp(-i m phi_j).  One CTA per ring; the ring sits in shared memory; thread t takes m = t, t + T, ...
__global__ void __launch_bounds__(RING_THREADS)
k_sht_ring_analysis(Hpx h, int lmax, const double *__restrict__ map, double2 *__restrict__ F, i64 ring_stride) {
    extern __shared__ double s_ring[];
    const i64 ring = (i64)blockIdx.x + 1;
    i64 start;
    const RingPhase g = ring_phase(h, ring, start);
    for (i64 j = threadIdx.x; j < g.n; j += blockDim.x) s_ring[j] = map[start + j];
    __syncthreads();
    for (int m = threadIdx.x; m <= lmax; m += blockDim.x) F[(i64)m * ring_stride + (ring - 1)] = ring_dft_m(m, g, s_ring);
}

// f_j = b_0 + 2 Re sum_{m>0} b_m exp(i m phi_j).  One CTA per ring; thread t takes pixels j = t, t + T, ...
__global__ void __launch_bounds__(RING_THREADS)
k_sht_ring_synthesis(Hpx h, int lmax, const double2 *__restrict__ B, i64 ring_stride, double *__restrict__ map) {
    __shared__ double2 s_b[B_CHUNK];
    const i64 ring = (i64)blockIdx.x + 1;
    i64 start;
    const RingPhase g = ring_phase(h, ring, start);
    const i64 n_pass = (g.n + blockDim.x - 1) / blockDim.x;
    for (i64 pass = 0; pass < n_pass; ++pass) {
        const i64 j = pass * blockDim.x + threadIdx.x;
        const bool live = j < g.n;
        double acc = 0.0, pc = 1.0, ps = 0.0;
        if (live) twiddle(1, j, g, pc, ps);                      // exp(i phi_j): one step in m
        for (int m0 = 0; m0 <= lmax; m0 += B_CHUNK) {
            const int m1 = min(m0 + B_CHUNK, lmax + 1);
            __syncthreads();
            for (int m = m0 + threadIdx.x; m < m1; m += blockDim.x) s_b[m - m0] = B[(i64)m * ring_stride + (ring - 1)];
            __syncthreads();
            if (live) ring_synth_chunk(j, g, m0, m1, s_b, pc, ps, acc);
        }
        if (live) map[start + j] = acc;
    }
}

// ---- Legendre recursion --------------------------------------------------------------------------------------------------
// lambda_mm and the recursion coefficients follow oracle/anafast_rings.py.  State per lane: (prev, cur) mantissas sharing one
// power-of-two exponent `expo` (a multiple of SHT_SCALE_BITS, <= 0); true value = mantissa * 2^expo.
struct LegState {
    double prev, cur, sf;    // sf = 2^expo (0 when that underflows: the contribution is negligible)
    int expo;
};

__host__ __device__ __forceinline__ double pow2_or_zero(int e) { return (e < -1000) ? 0.0 : ldexp(1.0, e); }

__host__ __device__ __forceinline__ LegState leg_start(int m, double ln_mm, double sin2) {
    LegState st;
    st.prev = 0.0;
    if (m > 0 && !(sin2 > 0.0)) {            // exactly on a pole: lambda_lm = 0 for every m > 0
        st.cur = 0.0; st.expo = 0; st.sf = 1.0;
        return st;
    }
    const double ln = (m == 0) ? ln_mm : fma(0.5 * (double)m, log(sin2), ln_mm);
    const double log2v = ln * 1.4426950408889634074;
    double e = floor(log2v / (double)SHT_SCALE_BITS) * (double)SHT_SCALE_BITS;
    if (e > 0.0) e = 0.0;
    if (e < -1.0e9) e = -1.0e9;
    st.expo = (int)e;
    st.cur = exp2(log2v - e) * ((m & 1) ? -1.0 : 1.0);
    st.sf = pow2_or_zero(st.expo);
    return st;
}

// advance l - 1 -> l (l > m): a_l (x cur - c_{l-1} prev), rescale when the mantissa outgrows 2^SHT_SCALE_BITS
__host__ __device__ __forceinline__ void leg_step(LegState &st, int l, int m, double x, double &c_prev) {
    const double l2 = (double)l * (double)l, m2 = (double)m * (double)m;
    const double a = sqrt((4.0 * l2 - 1.0) / (l2 - m2));
    const double nxt = a * (x * st.cur - c_prev * st.prev);
    st.prev = st.cur;
    st.cur = nxt;
    if (fabs(st.cur) > 1.157920892373162e77) {          // 2^256
        st.cur *= 8.636168555094445e-78;                // 2^-256
        st.prev *= 8.636168555094445e-78;
        st.expo += SHT_SCALE_BITS;
        st.sf = pow2_or_zero(st.expo);
    }
    c_prev = sqrt((l2 - m2) / (4.0 * l2 - 1.0));
}

// pair p of rings: north ring p + 1 with its mirror 4 nside - (p + 1); p = 2 nside - 1 is the equator alone (rs = -1)
__host__ __device__ __forceinline__ void pair_rings(const Hpx &h, i64 p, i64 &rn, i64 &rs) {
    rn = p + 1;
    rs = (rn == 2 * h.nside) ? -1 : 4 * h.nside - rn;
}

// What one lane of the Legendre kernels holds for its ring pair.  Written as __host__ __device__ so that pairing, parity and
// packing can be checked on the CPU (bfg_test_sht_legendre_host); the kernels add the thread mapping and the warp sum.
struct PairLane {
    double x;                  // cos(theta) of the northern ring
    double Pr, Pi, Qr, Qi;     // F_north +- F_south: multiply lambda_lm when l + m is even / odd
    LegState st;
    i64 rn, rs;
};

__host__ __device__ __forceinline__ PairLane pair_lane(const Hpx &h, i64 p, int m, double ln_mm_m, const double2 *F_m /* [ring-1] */) {
    PairLane q;
    pair_rings(h, p, q.rn, q.rs);
    double sth;
    ring_z_sth(h, q.rn, q.x, sth);
    q.Pr = q.Pi = q.Qr = q.Qi = 0.0;
    if (F_m) {
        const double2 fn = F_m[q.rn - 1];
        double2 fs; fs.x = 0.0; fs.y = 0.0;
        if (q.rs > 0) fs = F_m[q.rs - 1];
        q.Pr = fn.x + fs.x; q.Pi = fn.y + fs.y;
        q.Qr = fn.x - fs.x; q.Qi = fn.y - fs.y;
    }
    q.st = leg_start(m, ln_mm_m, sth * sth);
    return q;
}

// healpy packing of the m >= 0 coefficients: idx(l, m) = alm_base(m) + l
__host__ __device__ __forceinline__ i64 alm_base(int lmax, int m) { return (i64)m * (2 * (i64)lmax + 1 - m) / 2; }

// this pair's contribution to a_lm (before the sum over pairs): weight * lambda_lm(x) * (P or Q)
__host__ __device__ __forceinline__ void pair_contribution(const PairLane &q, int l, int m, double weight, double &vr, double &vi) {
    const double lam = q.st.cur * q.st.sf * weight;
    const bool even = ((l + m) & 1) == 0;
    vr = lam * (even ? q.Pr : q.Qr);
    vi = lam * (even ? q.Pi : q.Qi);
}

// b_m of the pair's two rings from a_lm (l = m .. lmax): lambda_lm(-x) = (-1)^(l+m) lambda_lm(x)
__host__ __device__ __forceinline__ void pair_synthesis(PairLane &q, int m, int lmax, const double2 *alm_m /* indexed by l */,
                                                        double2 &bn, double2 &bs) {
    double er = 0.0, ei = 0.0, orr = 0.0, oi = 0.0, c_prev = 0.0;
    for (int l = m; l <= lmax; ++l) {
        if (l > m) leg_step(q.st, l, m, q.x, c_prev);
        const double lam = q.st.cur * q.st.sf;
        const double2 a = alm_m[l];
        if (((l + m) & 1) == 0) { er = fma(a.x, lam, er); ei = fma(a.y, lam, ei); }
        else { orr = fma(a.x, lam, orr); oi = fma(a.y, lam, oi); }
    }
    bn.x = er + orr; bn.y = ei + oi;
    bs.x = er - orr; bs.y = ei - oi;
}

__global__ void __launch_bounds__(256)
k_sht_leg_analysis(Hpx h, int lmax, const double *__restrict__ ln_mm, const double2 *__restrict__ F, i64 ring_stride,
                   double weight, double *__restrict__ alm) {
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 n_pairs = 2 * h.nside, chunks = (n_pairs + 31) / 32;
    const i64 m64 = warp / chunks;
    if (m64 > lmax) return;                                       // whole warp leaves together
    const int m = (int)m64;
    const i64 p = (warp - m64 * chunks) * 32 + lane;
    PairLane q;
    if (p < n_pairs) {
        q = pair_lane(h, p, m, ln_mm[m], F + (i64)m * ring_stride);
    } else {                                                      // idle lane: contributes zeros to the warp sums
        q.x = 0.0; q.Pr = q.Pi = q.Qr = q.Qi = 0.0;
        q.st.prev = 0.0; q.st.cur = 0.0; q.st.sf = 0.0; q.st.expo = 0;
        q.rn = 1; q.rs = -1;
    }
    const i64 base = alm_base(lmax, m);
    double c_prev = 0.0;
    for (int l = m; l <= lmax; ++l) {
        if (l > m) leg_step(q.st, l, m, q.x, c_prev);
        double vr, vi;
        pair_contribution(q, l, m, weight, vr, vi);
        vr = warp_sum(vr);
        vi = warp_sum(vi);
        if (lane == 0 && (vr != 0.0 || vi != 0.0)) {
            red_add(alm + 2 * (base + l), vr);
            red_add(alm + 2 * (base + l) + 1, vi);
        }
    }
}

__global__ void __launch_bounds__(256)
k_sht_leg_synthesis(Hpx h, int lmax, const double *__restrict__ ln_mm, const double2 *__restrict__ alm, double2 *__restrict__ B,
                    i64 ring_stride) {
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 n_pairs = 2 * h.nside, chunks = (n_pairs + 31) / 32;
    const i64 m64 = warp / chunks;
    if (m64 > lmax) return;
    const int m = (int)m64;
    const i64 p = (warp - m64 * chunks) * 32 + (threadIdx.x & 31);
    if (p >= n_pairs) return;                                     // no warp-wide operation below
    PairLane q = pair_lane(h, p, m, ln_mm[m], nullptr);
    double2 bn, bs;
    pair_synthesis(q, m, lmax, alm + alm_base(lmax, m), bn, bs);
    B[(i64)m * ring_stride + (q.rn - 1)] = bn;
    if (q.rs > 0) B[(i64)m * ring_stride + (q.rs - 1)] = bs;
}

__global__ void k_sht_alm2cl(int lmax, const double2 *__restrict__ alm, double *__restrict__ cl) {
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l <= lmax; l += gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int m = 0; m <= l; ++m) {
            const double2 a = alm[(i64)m * (2 * (i64)lmax + 1 - m) / 2 + l];
            acc += ((m == 0) ? 1.0 : 2.0) * fma(a.x, a.x, a.y * a.y);
        }
        cl[l] = acc / (double)(2 * l + 1);
    }
}

int check_sht_args(int nside, int lmax) {
    BFG_REQUIRE(nside >= 1 && nside <= 8192, "nside out of range");
    BFG_REQUIRE(lmax >= 0 && lmax <= 4 * nside, "lmax out of range (0 <= lmax <= 4 nside)");
    return BFG_OK;
}

}  // namespace

extern "C" int64_t bfg_sht_workspace_elems(int nside, int lmax) {
    // complex128 elements of the [lmax + 1][4 nside - 1] ring-coefficient array F_m(r) / b_m(r)
    if (nside < 1 || lmax < 0) return 0;
    return (int64_t)(lmax + 1) * (4 * (int64_t)nside - 1);
}

extern "C" int bfg_sht_map2alm_pass(int nside, int lmax, const double *d_map, const double *d_ln_mm, double *d_work,
                                    double *d_alm, void *stream) {
    BFG_ENTRY();
    if (int rc = check_sht_args(nside, lmax)) return rc;
    BFG_REQUIRE(d_map && d_ln_mm && d_work && d_alm, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const Hpx h(nside);
    const i64 n_rings = 4 * (i64)nside - 1;
    const size_t smem = sizeof(double) * 4 * (size_t)nside;
    BFG_REQUIRE(smem <= 200 * 1024, "ring too long for shared memory (nside <= 6400)");
    BFG_CUDA_OK(cudaFuncSetAttribute(k_sht_ring_analysis, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_sht_ring_analysis<<<(unsigned)n_rings, RING_THREADS, smem, st>>>(h, lmax, d_map, (double2 *)d_work, n_rings);
    BFG_CUDA_OK(cudaGetLastError());
    const i64 chunks = (2 * (i64)nside + 31) / 32, warps = (i64)(lmax + 1) * chunks;
    const i64 blocks = (warps + 7) / 8;
    BFG_REQUIRE(blocks <= 0x7fffffff, "transform too large for one launch");
    k_sht_leg_analysis<<<(unsigned)blocks, 256, 0, st>>>(h, lmax, d_ln_mm, (const double2 *)d_work, n_rings,
                                                          4.0 * BFG_PI / (double)h.npix, d_alm);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_sht_alm2map(int nside, int lmax, const double *d_alm, const double *d_ln_mm, double *d_work, double *d_map,
                               void *stream) {
    BFG_ENTRY();
    if (int rc = check_sht_args(nside, lmax)) return rc;
    BFG_REQUIRE(d_map && d_ln_mm && d_work && d_alm, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const Hpx h(nside);
    const i64 n_rings = 4 * (i64)nside - 1;
    const i64 chunks = (2 * (i64)nside + 31) / 32, warps = (i64)(lmax + 1) * chunks;
    const i64 blocks = (warps + 7) / 8;
    BFG_REQUIRE(blocks <= 0x7fffffff, "transform too large for one launch");
    k_sht_leg_synthesis<<<(unsigned)blocks, 256, 0, st>>>(h, lmax, d_ln_mm, (const double2 *)d_alm, (double2 *)d_work, n_rings);
    BFG_CUDA_OK(cudaGetLastError());
    k_sht_ring_synthesis<<<(unsigned)n_rings, RING_THREADS, 0, st>>>(h, lmax, (const double2 *)d_work, n_rings, d_map);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_sht_alm2cl(int lmax, const double *d_alm, double *d_cl, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(lmax >= 0 && d_alm && d_cl, "bad argument");
    k_sht_alm2cl<<<(lmax + 256) / 256, 256, 0, (cudaStream_t)stream>>>(lmax, (const double2 *)d_alm, d_cl);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

// Unit-test entry, HOST side: the scaled recursion exactly as the kernels run it (same source, compiled for the host), so that
// its underflow handling can be checked without a GPU.  h_out[l - m] = lambda_lm(x) for l = m .. lmax.
extern "C" int bfg_test_sht_lambda_host(int m, int lmax, double ln_mm, double x, double sin2, double *h_out) {
    BFG_REQUIRE(m >= 0 && lmax >= m && h_out, "bad argument");
    LegState st = leg_start(m, ln_mm, sin2);
    double c_prev = 0.0;
    for (int l = m; l <= lmax; ++l) {
        if (l > m) leg_step(st, l, m, x, c_prev);
        h_out[l - m] = st.cur * st.sf;
    }
    return BFG_OK;
}

// Unit-test entry, HOST side: the per-ring sums exactly as the kernels run them (same functions, compiled for the host), to
// pin the phase conventions without a GPU.  One ring of n pixels with phi_j = (2 j + odd) pi / n:
//   h_F   [lmax + 1][2]  <- sum_j h_ring[j] exp(-i m phi_j)                       (k_sht_ring_analysis)
//   h_out [n]            <- sum_m c_m Re(h_b[m] exp(i m phi_j)), c_0 = 1, c_m = 2  (k_sht_ring_synthesis, chunked the same way)
extern "C" int bfg_test_sht_ring_host(int64_t n, int odd, int lmax, const double *h_ring, double *h_F, const double *h_b,
                                      double *h_out) {
    BFG_REQUIRE(n >= 1 && lmax >= 0 && (odd == 0 || odd == 1) && h_ring && h_F && h_b && h_out, "bad argument");
    RingPhase g; g.n = n; g.odd = odd;
    for (int m = 0; m <= lmax; ++m) {
        const double2 f = ring_dft_m(m, g, h_ring);
        h_F[2 * m] = f.x; h_F[2 * m + 1] = f.y;
    }
    const double2 *b = (const double2 *)h_b;
    for (i64 j = 0; j < n; ++j) {
        double acc = 0.0, pc, ps;
        twiddle(1, j, g, pc, ps);
        for (int m0 = 0; m0 <= lmax; m0 += B_CHUNK) {
            const int m1 = (m0 + B_CHUNK < lmax + 1) ? m0 + B_CHUNK : lmax + 1;
            ring_synth_chunk(j, g, m0, m1, b + m0, pc, ps, acc);
        }
        h_out[j] = acc;
    }
    return BFG_OK;
}

// Unit-test entry, HOST side: the Legendre stage of one m over ALL ring pairs with the lane functions of the two kernels --
// pairing of north/south rings, parity, healpy packing -- so that only the thread mapping and the warp sum remain for the GPU.
//   h_F_m   [4 nside - 1][2]  F_m(r) of every ring            ->  h_alm_m [lmax + 1][2] (entries l >= m) = w sum_pairs ...
//   h_alm_in [lmax + 1][2]    a_lm for this m, indexed by l   ->  h_B_m [4 nside - 1][2] = b_m(r)
extern "C" int bfg_test_sht_legendre_host(int nside, int lmax, int m, const double *h_ln_mm, const double *h_F_m, double *h_alm_m,
                                          const double *h_alm_in, double *h_B_m) {
    if (int rc = check_sht_args(nside, lmax)) return rc;
    BFG_REQUIRE(m >= 0 && m <= lmax && h_ln_mm && h_F_m && h_alm_m && h_alm_in && h_B_m, "bad argument");
    const Hpx h(nside);
    const double weight = 4.0 * BFG_PI / (double)h.npix;
    for (int l = 0; l <= lmax; ++l) { h_alm_m[2 * l] = 0.0; h_alm_m[2 * l + 1] = 0.0; }
    for (i64 p = 0; p < 2 * (i64)nside; ++p) {
        PairLane q = pair_lane(h, p, m, h_ln_mm[m], (const double2 *)h_F_m);
        double c_prev = 0.0;
        for (int l = m; l <= lmax; ++l) {
            if (l > m) leg_step(q.st, l, m, q.x, c_prev);
            double vr, vi;
            pair_contribution(q, l, m, weight, vr, vi);
            h_alm_m[2 * l] += vr; h_alm_m[2 * l + 1] += vi;
        }
        PairLane q2 = pair_lane(h, p, m, h_ln_mm[m], nullptr);
        double2 bn, bs;
        pair_synthesis(q2, m, lmax, (const double2 *)h_alm_in, bn, bs);
        h_B_m[2 * (q2.rn - 1)] = bn.x; h_B_m[2 * (q2.rn - 1) + 1] = bn.y;
        if (q2.rs > 0) { h_B_m[2 * (q2.rs - 1)] = bs.x; h_B_m[2 * (q2.rs - 1) + 1] = bs.y; }
    }
    return BFG_OK;
}
