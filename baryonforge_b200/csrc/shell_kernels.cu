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
constexpr int SHELL_MIN_CTAS = 6;            // caps registers at 85/thread: 24 warps/SM hide the fp64 latency
constexpr int RING_CHUNK = SHELL_THREADS;   // ring segments staged in shared memory per pass (one per thread)

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

// Per-halo constants of the pixel update, hoisted out of the pixel loop.
struct HaloUpd {
    double D, a, pjx, pjy, pjz;   // pos_j = vec_j * D                            HealpixRunner.py:337
    double ln_inv_a;              // ln(1/a): ln(r_sep/a) = 0.5 ln(r^2) + ln(1/a)  :345
    double rcut2;                 // (model eps * R_com * a)^2 : r_com < rcut  <=>  r_sep^2 < rcut2   BaryonCorrection.py:410
    double lnRcom, scale;
};

__device__ __forceinline__ HaloUpd make_upd(const HaloSph &s) {
    HaloUpd u;
    u.D = s.D; u.a = s.a;
    u.pjx = s.vx * s.D; u.pjy = s.vy * s.D; u.pjz = s.vz * s.D;
    u.ln_inv_a = s.lnz;           // the record's ln(1/a) is exactly this quantity
    double rc = s.rcut * s.a;
    u.rcut2 = rc * rc;
    u.lnRcom = s.lnRcom; u.scale = s.scale;
    return u;
}

// One (halo, pixel) update.  (x, y, z) = pixel unit vector; (px, py, pz) = (x, y, z) * D.
template <bool PAINT, bool UNIFORM>
__device__ __forceinline__ void shell_update(const TableView &T, const double *__restrict__ row, bool valid,
                                             const HaloUpd &u, double x, double y, double z, double px, double py,
                                             double pz, double *__restrict__ out, i64 nloc, i64 lp,
                                             const double2 *__restrict__ l2tab) {
    // HealpixRunner.py:338-341  diff = pos - pos_j ; r_sep^2 = sum(diff^2)
    double dx = px - u.pjx, dy = py - u.pjy, dz = pz - u.pjz;
    double r2 = dx * dx + dy * dy + dz * dz;
    // ln(r_sep / a) = 0.5 ln2 log2(r^2) + ln(1/a): no square root, no division, table-driven log2  (:345 / :472)
    double xq = fma(fast_log2(r2, l2tab), 0.34657359027997264, u.ln_inv_a);
    if (T.flags & BFG_TABLE_RDELTA) xq -= u.lnRcom;
    double val = row_lookup<UNIFORM>(T, row, xq);
    if (!valid) val = CUDART_NAN;
    if (PAINT) {
        val = exp(val);                            // Tabulate.py:319
        if (!isfinite(val)) return;                // HealpixRunner.py:473 (adds 0)
        val *= u.scale;                            // :478
        if (val != 0.0) red_add(out + lp, val);    // :481
    } else {
        val = (r2 < u.rcut2) ? val : 0.0;          // BaryonCorrection.py:410-411
        double sc = (val * u.a) * rsqrt(r2);       // offset / r_sep                 HealpixRunner.py:345-346
        // :347 non-finite -> 0 (r_sep = 0, NaN/inf table value, outside the table); exact zeros add nothing
        if (!isfinite(sc) || sc == 0.0) return;
        double nx = px + sc * dx, ny = py + sc * dy, nz = pz + sc * dz;   // :350 nw_pos = pos + offset
        double ninv = rsqrt(nx * nx + ny * ny + nz * nz);
        red_add(out + lp, nx * ninv - x);                                 // :351-355
        red_add(out + nloc + lp, ny * ninv - y);
        red_add(out + 2 * nloc + lp, nz * ninv - z);
    }
}

// A (halo, ring) segment staged in shared memory by the thread that derived it.
struct RingSeg {
    i64 lbase;             // ring's first pixel - pix_lo
    int nr, ip_lo, cnt, shifted;
    double z, sth;         // ring z, sin(theta)
    double rotS, rotC;     // sin / cos of 32 pixel steps in azimuth
};

template <bool PAINT, bool UNIFORM>
__global__ void __launch_bounds__(SHELL_THREADS, SHELL_MIN_CTAS)
k_shell_halos(TableView T, Hpx h, i64 n_halo, const double *__restrict__ halos, const double *__restrict__ extras,
              int n_extra, double *__restrict__ out, i64 pix_lo, i64 pix_hi, unsigned long long *nupd,
              const double2 *__restrict__ g_l2tab) {
    extern __shared__ double row[];
    __shared__ RingSeg segs[RING_CHUNK];
    __shared__ double2 l2tab[BFG_LOG2_TAB];
    __shared__ int s_next;
    load_log2_table(l2tab, g_l2tab);   // visible after the first __syncthreads() below
    const int lane = threadIdx.x & 31;
    const i64 nloc = pix_hi - pix_lo;
    i64 done = 0;

    for (i64 j = blockIdx.x; j < n_halo; j += gridDim.x) {
        const HaloSph s = load_halo(halos + j * BFG_HALO_STRIDE);
        __syncthreads();  // previous halo's row / segments no longer in use
        bool valid;
        blend_row(T, s.lnz, s.lnM, extras ? extras + j * n_extra : nullptr, row, valid);
        const DiscRings d = disc_rings(h, s.theta, s.phi, s.radius);
        const HaloUpd u = make_upd(s);
        // `if pixind.size < 4` (HealpixRunner.py:333) can only trigger for discs of a few pixels (<= ~12 rings)
        const bool tiny = !PAINT && (s.radius * s.radius * (double)h.npix * 0.25 < 64.0);

        if (tiny) {   // count the disc (block-wide); discs with no pixel centre at all land here too
            __shared__ int s_cnt[SHELL_THREADS / 32];
            int c = 0;
            for (i64 iz = d.ra + threadIdx.x; iz <= d.rb; iz += SHELL_THREADS) {
                i64 start, nr, ip_lo, cnt; bool sh;
                disc_ring_span(h, d, iz, start, nr, sh, ip_lo, cnt);
                c += (int)cnt;
            }
            c = (int)warp_sum_i64(c);
            if (lane == 0) s_cnt[threadIdx.x >> 5] = c;
            __syncthreads();   // also: row ready
            int tot = 0;
            for (int w = 0; w < SHELL_THREADS / 32; ++w) tot += s_cnt[w];
            if (tot < 4) {
                if (threadIdx.x < 4) {
                    i64 pix[4]; double w[4];
                    get_interpol(h, s.theta_ll, s.phi_ll, pix, w);   // HealpixRunner.py:334
                    i64 p = pix[threadIdx.x];
                    if (p >= pix_lo && p < pix_hi) {
                        double x, y, z;
                        pix2vec(h, p, x, y, z);
                        shell_update<PAINT, UNIFORM>(T, row, valid, u, x, y, z, x * u.D, y * u.D, z * u.D, out, nloc,
                                                     p - pix_lo, l2tab);
                        ++done;
                    }
                }
                continue;   // uniform across the block
            }
        }

        for (i64 base = d.ra; base <= d.rb; base += RING_CHUNK) {
            // ---- stage up to RING_CHUNK ring segments: one ring per thread ------------------------------
            {
                i64 iz = base + threadIdx.x;
                RingSeg g;
                g.cnt = 0; g.nr = 0;
                if (iz <= d.rb) {
                    i64 start, nr, ip_lo, cnt; bool sh;
                    disc_ring_span(h, d, iz, start, nr, sh, ip_lo, cnt);
                    g.cnt = (int)cnt;     // counted even when the ring is outside the owned range (fallback test)
                    g.nr = 0;             // nr == 0 marks "nothing to do here"
                    if (cnt > 0 && start < pix_hi && start + nr > pix_lo) {
                        g.lbase = start - pix_lo;
                        g.nr = (int)nr; g.ip_lo = (int)ip_lo; g.shifted = sh ? 1 : 0;
                        ring_z_sth(h, iz, g.z, g.sth);
                        sincospi(64.0 / (double)nr, &g.rotS, &g.rotC);   // 32 pixels * (2/nr) half-turns
                    }
                }
                segs[threadIdx.x] = g;
                if (threadIdx.x == 0) s_next = 0;
            }
            __syncthreads();  // segments + row ready
            const int nseg = (int)min((i64)RING_CHUNK, d.rb - base + 1);

            // ---- warps pull ring segments; lanes walk consecutive pixels -------------------------------
            for (;;) {
                int r = 0;
                if (lane == 0) r = atomicAdd(&s_next, 1);
                r = __shfl_sync(0xffffffffu, r, 0);
                if (r >= nseg) break;
                const int cnt = segs[r].cnt;
                const int nr = segs[r].nr;
                if (cnt == 0 || nr == 0) continue;
                const i64 lbase = segs[r].lbase;
                const double z = segs[r].z, sth = segs[r].sth;
                const double pz = z * u.D, sD = sth * u.D;
                int ip = segs[r].ip_lo + lane;
                if (ip >= nr) ip -= nr;
                double sn, cs;
                sincospi(((double)ip + (segs[r].shifted ? 0.5 : 0.0)) * (2.0 / (double)nr), &sn, &cs);
                const double rotS = segs[r].rotS, rotC = segs[r].rotC;
                for (int i = lane; i < cnt; i += 32) {
                    i64 lp = lbase + ip;
                    if ((unsigned long long)lp < (unsigned long long)nloc) {
                        shell_update<PAINT, UNIFORM>(T, row, valid, u, sth * cs, sth * sn, z, sD * cs, sD * sn, pz, out,
                                                     nloc, lp, l2tab);
                        ++done;
                    }
                    ip += 32;
                    if (ip >= nr) ip -= nr;
                    double c2 = cs * rotC - sn * rotS;   // advance the azimuth by 32 pixels
                    sn = sn * rotC + cs * rotS;
                    cs = c2;
                }
            }
            __syncthreads();  // before the next chunk overwrites the segments
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
    BFG_REQUIRE(t && (d_halos || n_halo == 0) && (d_out || pix_lo == pix_hi), "null argument");
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
    const double2 *g_l2tab = nullptr;
    if (int rc = get_log2_table(&g_l2tab)) return rc;
    auto go = [&](auto kern) -> int {
        BFG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, SHELL_THREADS, smem, st>>>(t->view, h, n_halo, d_halos, d_extras, n_extra, d_out, pix_lo, pix_hi,
                                                  (unsigned long long *)d_nupdates, g_l2tab);
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
