// shell_kernels.cu -- HEALPix-shell runners on sm_100a.
//
//   k_shell_halos<PAINT>  : the per-halo loop of BaryonifyShell.process  (BaryonForge/Runners/HealpixRunner.py:315-355)
//                           and PaintProfilesShell.process              (BaryonForge/Runners/HealpixRunner.py:449-481)
//   k_shell_regrid        : the re-binning step                         (BaryonForge/Runners/HealpixRunner.py:357-365, :17-71)
//
// One CTA per halo: the CTA blends the halo's 2^(ndim-1) table rows into one radial row in shared memory, derives
// the disc's ring range on the device (query_disc, no host round trip), and its warps walk (halo, ring) segments
// with lanes over consecutive pixels of the ring, so the fp64 REDs of a warp hit consecutive addresses of the
// component-major offsets array.
#include <algorithm>
#include "bfg_common.cuh"

using namespace bfg;

namespace {

constexpr int SHELL_THREADS = 128;

struct HaloSph {
    double vx, vy, vz, theta, phi, D, a, radius, lnz, lnM, rcut, lnRcom, scale, theta_ll, phi_ll;
};

__device__ __forceinline__ HaloSph load_halo(const double *__restrict__ H) {
    HaloSph s;
    s.vx = __ldg(H + BFG_HS_VX); s.vy = __ldg(H + BFG_HS_VY); s.vz = __ldg(H + BFG_HS_VZ);
    s.theta = __ldg(H + BFG_HS_THETA); s.phi = __ldg(H + BFG_HS_PHI);
    s.D = __ldg(H + BFG_HS_D); s.a = __ldg(H + BFG_HS_A); s.radius = __ldg(H + BFG_HS_RADIUS);
    s.lnz = __ldg(H + BFG_HS_LNZ); s.lnM = __ldg(H + BFG_HS_LNM); s.rcut = __ldg(H + BFG_HS_RCUT);
    s.lnRcom = __ldg(H + BFG_HS_LNRCOM); s.scale = __ldg(H + BFG_HS_SCALE);
    s.theta_ll = __ldg(H + BFG_HS_THETA_LL); s.phi_ll = __ldg(H + BFG_HS_PHI_LL);
    return s;
}

// One (halo, pixel) update.  (x, y, z) is the pixel's unit vector.
template <bool PAINT, bool UNIFORM>
__device__ __forceinline__ void shell_update(const TableView &T, const double *__restrict__ row, bool valid,
                                             const HaloSph &s, double x, double y, double z, double *__restrict__ out,
                                             i64 nloc, i64 lp) {
    // HealpixRunner.py:337-341  pos = vec*D ; diff = pos - pos_j ; r_sep = sqrt(sum(diff^2))
    double px = x * s.D, py = y * s.D, pz = z * s.D;
    double dx = px - s.vx * s.D, dy = py - s.vy * s.D, dz = pz - s.vz * s.D;
    double r_sep = sqrt(dx * dx + dy * dy + dz * dz);
    double rc = r_sep / s.a;                       // :345 / :472 comoving radius handed to the model
    double xq = log(rc);
    if (T.flags & BFG_TABLE_RDELTA) xq -= s.lnRcom;
    double val = row_lookup<UNIFORM>(T, row, xq);
    if (!valid) val = CUDART_NAN;
    if (PAINT) {
        val = exp(val);                            // Tabulate.py:319
        if (!isfinite(val)) return;                // HealpixRunner.py:473 (adds 0)
        val *= s.scale;                            // :478
        if (val != 0.0) red_add(out + lp, val);    // :481
    } else {
        val = (rc < s.rcut) ? val : 0.0;           // BaryonCorrection.py:410-411
        double off = val * s.a;                    // HealpixRunner.py:345
        double ox = off * (dx / r_sep), oy = off * (dy / r_sep), oz = off * (dz / r_sep);  // :346
        if (!isfinite(ox)) ox = 0.0;               // :347, element-wise
        if (!isfinite(oy)) oy = 0.0;
        if (!isfinite(oz)) oz = 0.0;
        if (ox == 0.0 && oy == 0.0 && oz == 0.0) return;   // delta would be round-off only
        double nx = px + ox, ny = py + oy, nz = pz + oz;   // :350
        double nn = sqrt(nx * nx + ny * ny + nz * nz);
        red_add(out + lp, nx / nn - x);                    // :351-355
        red_add(out + nloc + lp, ny / nn - y);
        red_add(out + 2 * nloc + lp, nz / nn - z);
    }
}

template <bool PAINT, bool UNIFORM>
__global__ void __launch_bounds__(SHELL_THREADS)
k_shell_halos(TableView T, Hpx h, i64 n_halo, const double *__restrict__ halos, const double *__restrict__ extras,
              int n_extra, double *__restrict__ out, i64 pix_lo, i64 pix_hi, unsigned long long *nupd) {
    extern __shared__ double row[];
    __shared__ i64 s_cnt[SHELL_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = SHELL_THREADS / 32;
    const i64 nloc = pix_hi - pix_lo;
    i64 done = 0;

    for (i64 j = blockIdx.x; j < n_halo; j += gridDim.x) {
        const HaloSph s = load_halo(halos + j * BFG_HALO_STRIDE);
        __syncthreads();  // previous halo's row no longer in use
        bool valid;
        blend_row(T, s.lnz, s.lnM, extras ? extras + j * n_extra : nullptr, row, valid);
        const DiscRings d = disc_rings(h, s.theta, s.phi, s.radius);

        bool fallback = false;
        if (!PAINT) {
            // `if pixind.size < 4` (HealpixRunner.py:333): only discs of a few pixels can get there
            double expect = s.radius * s.radius * (double)h.npix * 0.25;
            if (expect < 64.0) {
                i64 c = 0;
                for (i64 iz = d.ra + threadIdx.x; iz <= d.rb; iz += SHELL_THREADS) {
                    i64 start, nr, ip_lo, cnt; bool sh;
                    disc_ring_span(h, d, iz, start, nr, sh, ip_lo, cnt);
                    c += cnt;
                }
                c = warp_sum_i64(c);
                if (lane == 0) s_cnt[warp] = c;
                __syncthreads();
                i64 tot = 0;
                for (int w = 0; w < NW; ++w) tot += s_cnt[w];
                fallback = tot < 4;
            }
        }
        __syncthreads();  // row ready

        if (fallback) {
            if (threadIdx.x < 4) {
                i64 pix[4]; double w[4];
                get_interpol(h, s.theta_ll, s.phi_ll, pix, w);   // HealpixRunner.py:334
                i64 p = pix[threadIdx.x];
                if (p >= pix_lo && p < pix_hi) {
                    double x, y, z;
                    pix2vec(h, p, x, y, z);
                    shell_update<PAINT, UNIFORM>(T, row, valid, s, x, y, z, out, nloc, p - pix_lo);
                    ++done;
                }
            }
            continue;
        }

        for (i64 iz = d.ra + warp; iz <= d.rb; iz += NW) {
            i64 start, nr, ip_lo, cnt; bool sh;
            disc_ring_span(h, d, iz, start, nr, sh, ip_lo, cnt);
            if (cnt == 0 || start >= pix_hi || start + nr <= pix_lo) continue;
            double z, sth;
            ring_z_sth(h, iz, z, sth);
            for (i64 i = lane; i < cnt; i += 32) {
                i64 ip = ip_lo + i;
                if (ip >= nr) ip -= nr;
                i64 p = start + ip;
                if (p < pix_lo || p >= pix_hi) continue;
                double sn, cs;
                sincos(ring_phi(h, iz, ip, sh), &sn, &cs);
                shell_update<PAINT, UNIFORM>(T, row, valid, s, sth * cs, sth * sn, z, out, nloc, p - pix_lo);
                ++done;
            }
        }
    }
    if (nupd) {
        done = warp_sum_i64(done);
        if (lane == 0 && done) atomicAdd(nupd, (unsigned long long)done);
    }
}

// Re-binning: one thread per source pixel.
__global__ void __launch_bounds__(256)
k_shell_regrid(Hpx h, const double *__restrict__ map_in, const double *__restrict__ off, double *__restrict__ map_out,
               i64 pix_lo, i64 pix_hi) {
    const i64 nloc = pix_hi - pix_lo;
    for (i64 lp = (i64)blockIdx.x * blockDim.x + threadIdx.x; lp < nloc; lp += (i64)gridDim.x * blockDim.x) {
        double m = map_in[lp];
        if (m == 0.0) continue;                                  // HealpixRunner.py:359
        double x, y, z;
        pix2vec(h, pix_lo + lp, x, y, z);
        x += off[lp]; y += off[nloc + lp]; z += off[2 * nloc + lp];   // :357 (not re-normalised)
        // hp.vec2ang(lonlat=True)  :358
        double dn = sqrt(x * x + y * y + z * z);
        double theta = acos(z / dn);
        double phi = atan2(y, x);
        if (phi < 0) phi += BFG_TWOPI;
        double lon = phi * (180.0 / BFG_PI), lat = 90.0 - theta * (180.0 / BFG_PI);
        // hp.get_interp_weights(lonlat=True)  :361
        double th2 = BFG_HALFPI - lat * (BFG_PI / 180.0), ph2 = lon * (BFG_PI / 180.0);
        i64 pix[4]; double w[4];
        get_interpol(h, th2, ph2, pix, w);
#pragma unroll
        for (int k = 0; k < 4; ++k) red_add(map_out + pix[k], w[k] * m);   // :17-71
    }
}

__global__ void k_disc_counts(Hpx h, i64 n_halo, const double *__restrict__ halos, i64 *__restrict__ npix) {
    const int lane = threadIdx.x & 31;
    i64 wid = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    i64 nw = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 j = wid; j < n_halo; j += nw) {
        const double *H = halos + j * BFG_HALO_STRIDE;
        DiscRings d = disc_rings(h, __ldg(H + BFG_HS_THETA), __ldg(H + BFG_HS_PHI), __ldg(H + BFG_HS_RADIUS));
        i64 c = 0;
        for (i64 iz = d.ra + lane; iz <= d.rb; iz += 32) {
            i64 start, nr, ip_lo, cnt; bool sh;
            disc_ring_span(h, d, iz, start, nr, sh, ip_lo, cnt);
            c += cnt;
        }
        c = warp_sum_i64(c);
        if (lane == 0) npix[j] = c;
    }
}

// test helper: pixel list of one halo (ring by ring; ascending inside each emitted span)
__global__ void k_query_disc(Hpx h, const double *__restrict__ H, i64 *__restrict__ pix, i64 cap, i64 *count) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    DiscRings d = disc_rings(h, H[BFG_HS_THETA], H[BFG_HS_PHI], H[BFG_HS_RADIUS]);
    i64 n = 0;
    for (i64 iz = d.ra; iz <= d.rb; ++iz) {
        i64 start, nr, ip_lo, cnt; bool sh;
        disc_ring_span(h, d, iz, start, nr, sh, ip_lo, cnt);
        for (i64 i = 0; i < cnt; ++i) {
            i64 ip = ip_lo + i;
            if (ip >= nr) ip -= nr;
            if (n < cap) pix[n] = start + ip;
            ++n;
        }
    }
    *count = n;
}

__global__ void k_pix2vec(Hpx h, i64 pix_lo, i64 pix_hi, double *__restrict__ xyz) {
    i64 n = pix_hi - pix_lo;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        double x, y, z;
        pix2vec(h, pix_lo + i, x, y, z);
        xyz[i] = x; xyz[n + i] = y; xyz[2 * n + i] = z;
    }
}

__global__ void k_interp_weights(Hpx h, i64 n, const double *__restrict__ th, const double *__restrict__ ph,
                                 i64 *__restrict__ pix, double *__restrict__ w) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        i64 p[4]; double ww[4];
        get_interpol(h, th[i], ph[i], p, ww);
        for (int k = 0; k < 4; ++k) { pix[k * n + i] = p[k]; w[k * n + i] = ww[k]; }
    }
}

__global__ void k_ang2pix(Hpx h, i64 n, const double *__restrict__ th, const double *__restrict__ ph, i64 *__restrict__ pix) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        pix[i] = ang2pix_ring(h, th[i], ph[i]);
}

int check_nside(int nside) {
    if (nside < 1 || nside > (1 << 24)) { set_error("nside out of range"); return BFG_ERR_INVALID; }
    return BFG_OK;
}

int grid_for(i64 n, int threads, int max_blocks = 148 * 32) {
    return (int)std::max<i64>(1, std::min<i64>((n + threads - 1) / threads, max_blocks));
}

template <bool PAINT>
int launch_shell(const bfg_table *t, int nside, i64 n_halo, const double *d_halos, const double *d_extras, int n_extra,
                 double *d_out, i64 pix_lo, i64 pix_hi, i64 *d_nupdates, cudaStream_t st) {
    BFG_REQUIRE(t && d_halos && d_out, "null argument");
    if (int rc = check_nside(nside)) return rc;
    Hpx h(nside);
    BFG_REQUIRE(pix_lo >= 0 && pix_hi <= h.npix && pix_lo <= pix_hi, "bad pixel range");
    BFG_REQUIRE(n_extra == t->view.ndim - 3, "n_extra must equal the table's extra axes");
    BFG_REQUIRE(n_extra == 0 || d_extras, "extras missing");
    BFG_REQUIRE(PAINT == ((t->view.flags & BFG_TABLE_LOG_VALUES) != 0),
                "paint needs a log-profile table, baryonify a displacement table");
    if (d_nupdates) BFG_CUDA_OK(cudaMemsetAsync(d_nupdates, 0, sizeof(i64), st));
    if (n_halo == 0 || pix_lo == pix_hi) return BFG_OK;
    size_t smem = sizeof(double) * t->view.n[2];
    BFG_REQUIRE(smem <= 200 * 1024, "radial axis too long for the shared-memory row (max 25600 nodes)");
    int blocks = (int)std::min<i64>(n_halo, (i64)1 << 30);
    auto go = [&](auto kern) -> int {
        BFG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, SHELL_THREADS, smem, st>>>(t->view, h, n_halo, d_halos, d_extras, n_extra, d_out, pix_lo, pix_hi,
                                                  (unsigned long long *)d_nupdates);
        BFG_CUDA_OK(cudaGetLastError());
        return BFG_OK;
    };
    return t->view.uniform_r ? go(k_shell_halos<PAINT, true>) : go(k_shell_halos<PAINT, false>);
}

}  // namespace

extern "C" int bfg_shell_offsets(const bfg_table *t, int nside, int64_t n_halo, const double *d_halos,
                                 const double *d_extras, int n_extra, double *d_offsets, int64_t pix_lo, int64_t pix_hi,
                                 int64_t *d_nupdates, void *stream) {
    return launch_shell<false>(t, nside, n_halo, d_halos, d_extras, n_extra, d_offsets, pix_lo, pix_hi,
                               (i64 *)d_nupdates, (cudaStream_t)stream);
}

extern "C" int bfg_shell_paint(const bfg_table *t, int nside, int64_t n_halo, const double *d_halos,
                               const double *d_extras, int n_extra, double *d_map, int64_t pix_lo, int64_t pix_hi,
                               int64_t *d_nupdates, void *stream) {
    return launch_shell<true>(t, nside, n_halo, d_halos, d_extras, n_extra, d_map, pix_lo, pix_hi, (i64 *)d_nupdates,
                              (cudaStream_t)stream);
}

extern "C" int bfg_shell_regrid(int nside, const double *d_map_in, const double *d_offsets, double *d_map_out,
                                int64_t pix_lo, int64_t pix_hi, void *stream) {
    BFG_REQUIRE(d_map_in && d_offsets && d_map_out, "null argument");
    if (int rc = check_nside(nside)) return rc;
    Hpx h(nside);
    BFG_REQUIRE(pix_lo >= 0 && pix_hi <= h.npix && pix_lo <= pix_hi, "bad pixel range");
    if (pix_lo == pix_hi) return BFG_OK;
    k_shell_regrid<<<grid_for(pix_hi - pix_lo, 256), 256, 0, (cudaStream_t)stream>>>(h, d_map_in, d_offsets, d_map_out,
                                                                                    pix_lo, pix_hi);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_healpix_disc_counts(int nside, int64_t n_halo, const double *d_halos, int64_t *d_npix, void *stream) {
    BFG_REQUIRE(d_halos && d_npix, "null argument");
    if (int rc = check_nside(nside)) return rc;
    if (n_halo == 0) return BFG_OK;
    k_disc_counts<<<grid_for(n_halo * 32, 256), 256, 0, (cudaStream_t)stream>>>(Hpx(nside), n_halo, d_halos, (i64 *)d_npix);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_healpix_query_disc(int nside, const double *d_halo, int64_t *d_pix, int64_t cap, int64_t *d_count,
                                      void *stream) {
    BFG_REQUIRE(d_halo && d_count && (d_pix || cap == 0), "null argument");
    if (int rc = check_nside(nside)) return rc;
    k_query_disc<<<1, 32, 0, (cudaStream_t)stream>>>(Hpx(nside), d_halo, (i64 *)d_pix, cap, (i64 *)d_count);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_healpix_pix2vec(int nside, int64_t pix_lo, int64_t pix_hi, double *d_xyz, void *stream) {
    BFG_REQUIRE(d_xyz, "null argument");
    if (int rc = check_nside(nside)) return rc;
    Hpx h(nside);
    BFG_REQUIRE(pix_lo >= 0 && pix_hi <= h.npix && pix_lo <= pix_hi, "bad pixel range");
    if (pix_lo == pix_hi) return BFG_OK;
    k_pix2vec<<<grid_for(pix_hi - pix_lo, 256), 256, 0, (cudaStream_t)stream>>>(h, pix_lo, pix_hi, d_xyz);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_healpix_interp_weights(int nside, int64_t n, const double *d_theta, const double *d_phi,
                                          int64_t *d_pix, double *d_w, void *stream) {
    BFG_REQUIRE(d_theta && d_phi && d_pix && d_w, "null argument");
    if (int rc = check_nside(nside)) return rc;
    if (n == 0) return BFG_OK;
    k_interp_weights<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(Hpx(nside), n, d_theta, d_phi, (i64 *)d_pix, d_w);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_healpix_ang2pix(int nside, int64_t n, const double *d_theta, const double *d_phi, int64_t *d_pix,
                                   void *stream) {
    BFG_REQUIRE(d_theta && d_phi && d_pix, "null argument");
    if (int rc = check_nside(nside)) return rc;
    if (n == 0) return BFG_OK;
    k_ang2pix<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(Hpx(nside), n, d_theta, d_phi, (i64 *)d_pix);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

// ------------------------------------------------------------------------------------------------ host-buffer calls
namespace {
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) {
        if (cudaMalloc(&p, bytes ? bytes : 8) != cudaSuccess) {
            set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
            return BFG_ERR_NOMEM;
        }
        return BFG_OK;
    }
};
}  // namespace

extern "C" int bfg_shell_baryonify_host(const bfg_table *t, int nside, int64_t n_halo, const double *h_halos,
                                        const double *h_extras, int n_extra, const double *h_map_in, double *h_map_out,
                                        int64_t *h_nupdates, double *h_sums) {
    BFG_REQUIRE(t && h_map_in && h_map_out && (h_halos || n_halo == 0), "null argument");
    if (int rc = check_nside(nside)) return rc;
    BFG_CUDA_OK(cudaSetDevice(t->device));
    Hpx h(nside);
    DevBuf halos, extras, map_in, map_out, off, scal;
    int rc;
    if ((rc = halos.alloc(sizeof(double) * BFG_HALO_STRIDE * n_halo))) return rc;
    if ((rc = extras.alloc(sizeof(double) * n_extra * n_halo))) return rc;
    if ((rc = map_in.alloc(sizeof(double) * h.npix))) return rc;
    if ((rc = map_out.alloc(sizeof(double) * h.npix))) return rc;
    if ((rc = off.alloc(sizeof(double) * 3 * h.npix))) return rc;
    if ((rc = scal.alloc(32))) return rc;
    cudaStream_t st = 0;
    BFG_CUDA_OK(cudaMemcpyAsync(halos.p, h_halos, sizeof(double) * BFG_HALO_STRIDE * n_halo, cudaMemcpyHostToDevice, st));
    if (n_extra) BFG_CUDA_OK(cudaMemcpyAsync(extras.p, h_extras, sizeof(double) * n_extra * n_halo, cudaMemcpyHostToDevice, st));
    BFG_CUDA_OK(cudaMemcpyAsync(map_in.p, h_map_in, sizeof(double) * h.npix, cudaMemcpyHostToDevice, st));
    BFG_CUDA_OK(cudaMemsetAsync(off.p, 0, sizeof(double) * 3 * h.npix, st));
    BFG_CUDA_OK(cudaMemsetAsync(map_out.p, 0, sizeof(double) * h.npix, st));
    i64 *d_n = (i64 *)scal.p;
    double *d_s = (double *)scal.p + 1;
    if ((rc = bfg_shell_offsets(t, nside, n_halo, (double *)halos.p, n_extra ? (double *)extras.p : nullptr, n_extra,
                                (double *)off.p, 0, h.npix, (int64_t *)d_n, st))) return rc;
    if ((rc = bfg_shell_regrid(nside, (double *)map_in.p, (double *)off.p, (double *)map_out.p, 0, h.npix, st))) return rc;
    if ((rc = bfg_sum_f64((double *)map_out.p, h.npix, d_s, st))) return rc;
    if ((rc = bfg_sum_f64((double *)map_in.p, h.npix, d_s + 1, st))) return rc;
    BFG_CUDA_OK(cudaMemcpyAsync(h_map_out, map_out.p, sizeof(double) * h.npix, cudaMemcpyDeviceToHost, st));
    i64 n_up = 0;
    double sums[2];
    BFG_CUDA_OK(cudaMemcpyAsync(&n_up, d_n, sizeof(i64), cudaMemcpyDeviceToHost, st));
    BFG_CUDA_OK(cudaMemcpyAsync(sums, d_s, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    BFG_CUDA_OK(cudaStreamSynchronize(st));
    if (h_nupdates) *h_nupdates = n_up;
    if (h_sums) { h_sums[0] = sums[0]; h_sums[1] = sums[1]; }
    return BFG_OK;
}

extern "C" int bfg_shell_paint_host(const bfg_table *t, int nside, int64_t n_halo, const double *h_halos,
                                    const double *h_extras, int n_extra, double *h_map_out, int64_t *h_nupdates) {
    BFG_REQUIRE(t && h_map_out && (h_halos || n_halo == 0), "null argument");
    if (int rc = check_nside(nside)) return rc;
    BFG_CUDA_OK(cudaSetDevice(t->device));
    Hpx h(nside);
    DevBuf halos, extras, map_out, scal;
    int rc;
    if ((rc = halos.alloc(sizeof(double) * BFG_HALO_STRIDE * n_halo))) return rc;
    if ((rc = extras.alloc(sizeof(double) * n_extra * n_halo))) return rc;
    if ((rc = map_out.alloc(sizeof(double) * h.npix))) return rc;
    if ((rc = scal.alloc(8))) return rc;
    cudaStream_t st = 0;
    BFG_CUDA_OK(cudaMemcpyAsync(halos.p, h_halos, sizeof(double) * BFG_HALO_STRIDE * n_halo, cudaMemcpyHostToDevice, st));
    if (n_extra) BFG_CUDA_OK(cudaMemcpyAsync(extras.p, h_extras, sizeof(double) * n_extra * n_halo, cudaMemcpyHostToDevice, st));
    BFG_CUDA_OK(cudaMemsetAsync(map_out.p, 0, sizeof(double) * h.npix, st));
    if ((rc = bfg_shell_paint(t, nside, n_halo, (double *)halos.p, n_extra ? (double *)extras.p : nullptr, n_extra,
                              (double *)map_out.p, 0, h.npix, (int64_t *)scal.p, st))) return rc;
    BFG_CUDA_OK(cudaMemcpyAsync(h_map_out, map_out.p, sizeof(double) * h.npix, cudaMemcpyDeviceToHost, st));
    i64 n_up = 0;
    BFG_CUDA_OK(cudaMemcpyAsync(&n_up, scal.p, sizeof(i64), cudaMemcpyDeviceToHost, st));
    BFG_CUDA_OK(cudaStreamSynchronize(st));
    if (h_nupdates) *h_nupdates = n_up;
    return BFG_OK;
}
