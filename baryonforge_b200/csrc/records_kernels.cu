// records_kernels.cu -- per-halo scalar prep of the shell runners on the device.
//
// The reference computes these scalars at the top of every loop iteration with two pyccl calls, a CubicSpline call and
// hp.ang2vec (BaryonForge/Runners/HealpixRunner.py:317-329, :451-462; Profiles/BaryonCorrection.py:371,398-399,410).
// Here the raw catalogue columns go to HBM once and ONE thread per halo writes the 128-byte record the halo-loop
// kernels read:
//   D_A(z)      the reference's own CubicSpline (HealpixRunner.py:297-299), evaluated from its PPoly coefficients in
//               scipy's operation order (no FMA), so D_j is bit-identical to D_a(z_j);
//   R_200c      R = cbrt(M) * g(ln(1+z)), g = (4 pi/3 Delta rho(a))^(-1/3) tabulated by the host from whatever cosmology
//               object is in use (pyccl or cosmology.Background) on 2048 nodes and splined (error < 1e-13 relative);
//   ln(1+z), ln M are taken from the host (numpy's own log) so that halos sitting exactly on a table edge fall on the
//               same side as in the reference (SURVEY.md §8a "shared semantics").
#include <algorithm>
#include <limits>
#include <vector>
#include "bfg_common.cuh"

using namespace bfg;

namespace {

// The per-halo bodies below are __host__ __device__: bfg_test_records_host runs the same source on the CPU (tests/
// test_records_host.py).  Device-only intrinsics get host stand-ins; on the device the expansion is the original text.
#ifdef __CUDA_ARCH__
#define REC_LD(p) __ldg(p)
#define REC_SUB(a, b) __dsub_rn(a, b)
#define REC_ADD(a, b) __dadd_rn(a, b)
#define REC_MUL(a, b) __dmul_rn(a, b)
#define REC_NAN CUDART_NAN
#define REC_INF CUDART_INF
#else
#define REC_LD(p) (*(p))
#define REC_SUB(a, b) ((a) - (b))
#define REC_ADD(a, b) ((a) + (b))
#define REC_MUL(a, b) ((a) * (b))
#define REC_NAN (std::numeric_limits<double>::quiet_NaN())
#define REC_INF (std::numeric_limits<double>::infinity())
#endif

struct PPolyView {
    int n;              // breakpoints
    const double *x;    // [n] ascending
    const double *c;    // [4][n-1], scipy PPoly.c layout: c[k][i] multiplies (x - x[i])^(3-k)
};

// scipy.interpolate._ppoly.evaluate (extrapolate=True): interval x[i] <= xv < x[i+1] (last one right-closed, outside
// values use the end intervals), then res = c3 + c2*s + c1*(s*s) + c0*((s*s)*s) summed in that order.
__host__ __device__ __forceinline__ double ppoly_eval(const PPolyView &p, double xv) {
    if (xv != xv) return xv;
    const int nint = p.n - 1;
    const double x0 = REC_LD(p.x), x1 = REC_LD(p.x + nint);
    int i;
    if (xv < x0) i = 0;
    else if (xv >= x1) i = nint - 1;
    else {
        i = (int)((xv - x0) / (x1 - x0) * (double)nint);
        i = min(max(i, 0), nint - 1);
        while (i > 0 && xv < REC_LD(p.x + i)) --i;
        while (i < nint - 1 && xv >= REC_LD(p.x + i + 1)) ++i;
    }
    const double s = REC_SUB(xv, REC_LD(p.x + i));
    double res = REC_LD(p.c + 3 * (i64)nint + i);
    double zp = s;
    res = REC_ADD(res, REC_MUL(REC_LD(p.c + 2 * (i64)nint + i), zp));
    zp = REC_MUL(zp, s);
    res = REC_ADD(res, REC_MUL(REC_LD(p.c + 1 * (i64)nint + i), zp));
    zp = REC_MUL(zp, s);
    res = REC_ADD(res, REC_MUL(REC_LD(p.c + i), zp));
    return res;
}

struct RecParams {
    int paint;
    double eps_run, eps_model, pixarea;   // pixarea > 0: paint scale = pixarea * D^2 (HealpixRunner.py:478), else 1
    PPolyView DA, g_run, g_mod;
};

__host__ __device__ __forceinline__ void shell_record_one(i64 n, i64 j, const double *__restrict__ cols, const RecParams &P,
                                                          double *__restrict__ halos, double *__restrict__ aux) {
    const double M = cols[j], z = cols[n + j], ra = cols[2 * n + j], dec = cols[3 * n + j];
    const double lnz = cols[4 * n + j], lnM = cols[5 * n + j];
    const double a = 1.0 / (1.0 + z);                                 // :319
    const double cm = cbrt(M);
    const double R = cm * ppoly_eval(P.g_run, lnz);                   // :320 physical Mpc
    const double D = ppoly_eval(P.DA, z);                             // :321
    // hp.ang2vec(ra, dec, lonlat=True)  :327
    const double theta_ll = BFG_HALFPI - dec * (BFG_PI / 180.0), phi_ll = ra * (BFG_PI / 180.0);
    double st, ct, sp, cp;
    sincos(theta_ll, &st, &ct);
    sincos(phi_ll, &sp, &cp);
    const double vx = st * cp, vy = st * sp, vz = ct;
    double *H = halos + j * BFG_HALO_STRIDE;
    H[BFG_HS_VX] = vx; H[BFG_HS_VY] = vy; H[BFG_HS_VZ] = vz;
    // pointing(vec), as healpy's query_disc wrapper rebuilds it  :330
    H[BFG_HS_THETA] = atan2(sqrt(vx * vx + vy * vy), vz);
    double phi = atan2(vy, vx);
    if (phi < 0) phi += BFG_TWOPI;
    H[BFG_HS_PHI] = phi;
    H[BFG_HS_D] = D;
    H[BFG_HS_A] = a;
    H[BFG_HS_RADIUS] = R * P.eps_run / D;                             // :329
    H[BFG_HS_LNZ] = lnz;
    H[BFG_HS_LNM] = lnM;
    double Rcom = REC_NAN;
    if (P.paint) {
        H[BFG_HS_RCUT] = REC_INF;
        H[BFG_HS_LNRCOM] = 0.0;
        H[BFG_HS_SCALE] = (P.pixarea > 0) ? P.pixarea * (D * D) : 1.0;  // :478
    } else {
        Rcom = cm * ppoly_eval(P.g_mod, lnz) / a;                     // BaryonCorrection.py:399
        H[BFG_HS_RCUT] = P.eps_model * Rcom;                          // :410
        H[BFG_HS_LNRCOM] = log(Rcom);                                 // :408
        H[BFG_HS_SCALE] = 1.0;
    }
    H[BFG_HS_THETA_LL] = theta_ll; H[BFG_HS_PHI_LL] = phi_ll;
    H[BFG_HS_SKIP] = 0.0;
    if (aux) { aux[j] = R; aux[n + j] = D; aux[2 * n + j] = Rcom; }
}

__global__ void __launch_bounds__(256)
k_shell_records(i64 n, const double *__restrict__ cols, RecParams P, double *__restrict__ halos, double *__restrict__ aux) {
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (i64)gridDim.x * blockDim.x)
        shell_record_one(n, j, cols, P, halos, aux);
}

}  // namespace

extern "C" int bfg_shell_records(int64_t n_halo, const double *d_cols, int paint, double eps_run, double eps_model,
                                 double pixarea, int n_DA, const double *d_DA_x, const double *d_DA_c, int n_g,
                                 const double *d_g_x, const double *d_g_run_c, const double *d_g_mod_c, double *d_halos,
                                 double *d_aux, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(n_halo >= 0, "negative halo count");
    if (n_halo == 0) return BFG_OK;
    BFG_REQUIRE(d_cols && d_halos && d_DA_x && d_DA_c && d_g_x && d_g_run_c, "null argument");
    BFG_REQUIRE(n_DA >= 2 && n_g >= 2, "splines need >= 2 breakpoints");
    BFG_REQUIRE(paint || d_g_mod_c, "baryonify records need the model-cosmology radius spline");
    RecParams P;
    P.paint = paint ? 1 : 0;
    P.eps_run = eps_run; P.eps_model = eps_model; P.pixarea = pixarea;
    P.DA = PPolyView{n_DA, d_DA_x, d_DA_c};
    P.g_run = PPolyView{n_g, d_g_x, d_g_run_c};
    P.g_mod = PPolyView{n_g, d_g_x, d_g_mod_c ? d_g_mod_c : d_g_run_c};
    int blocks = (int)std::max<i64>(1, std::min<i64>((n_halo + 255) / 256, 148 * 8));
    k_shell_records<<<blocks, 256, 0, (cudaStream_t)stream>>>(n_halo, d_cols, P, d_halos, d_aux);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

// ------------------------------------------------------------------------------------------------ box runners
// Per-halo scalars of the grid / snapshot runners on the device (Map2DRunner.py:484-520, :727-760; SnapshotRunner.py:219-228;
// BaryonCorrection.py:398-399,410).  All halos of a HaloNDCatalog share one redshift, so the cosmology enters through
// two scalars: g_run, g_mod with R_delta(M, a) = cbrt(M) * g (physical Mpc).  ln M comes from the host because the reference
// evaluates it in float32.
namespace {

struct BoxParams {
    int ndim, paint, grid;          // grid = 1: BaryonifyGrid / PaintProfilesGrid records, 0: BaryonifySnapshot records
    double a, lnz, g_run, g_mod, eps_run, eps_mod, res, rq_clip;
    int N;
    const double *bins;             // [N] cell centres (grid runners)
};

// np.argmin(np.abs(bins - x)): first minimum among the cells around round((x - bins[0]) / res)  (Map2DRunner.py:512-513)
__device__ __forceinline__ int nearest_bin(const double *__restrict__ bins, int N, double res, double x) {
    double cf = floor((x - bins[0]) / res + 0.5);
    cf = fmin(fmax(cf, 0.0), (double)(N - 1));
    const int c = (int)cf;
    int best = min(max(c - 1, 0), N - 1);
    double dbest = fabs(bins[best] - x);
    for (int o = 0; o <= 1; ++o) {          // ascending index order + strict '<' == argmin's first-minimum rule
        const int cand = min(max(c + o, 0), N - 1);
        const double d = fabs(bins[cand] - x);
        if (d < dbest) { best = cand; dbest = d; }
    }
    return best;
}

__global__ void __launch_bounds__(256)
k_box_records(i64 n, const double *__restrict__ cols, BoxParams P, double *__restrict__ halos, double *__restrict__ aux) {
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (i64)gridDim.x * blockDim.x) {
        const double M = cols[j], lnM = cols[4 * n + j];
        double *H = halos + j * BFG_HALO_STRIDE;
        const double cm = cbrt(M);
        const double R_phys = cm * P.g_run;                                   // get_radius(cosmo, M_j, a_j)  :491 / :734 / :226
        double R_mod = CUDART_NAN;
        double Nf;
        if (P.paint) {
            const double R_com = R_phys / P.a;                                // :734
            Nf = 2.0 * P.eps_run * R_com / P.res;                             // :740
            H[BFG_HB_PAINTCUT] = R_com * P.eps_run;                           // :815
            H[BFG_HB_RQ] = R_com * P.eps_run;
            H[BFG_HB_RCUT] = CUDART_INF;
            H[BFG_HB_LNRCOM] = 0.0;
        } else {
            const double R_q = fmin(fmax(P.eps_run * R_phys / P.a, 0.0), P.rq_clip);   // :492-493 / SnapshotRunner.py:227-228
            Nf = 2.0 * R_q / P.res;                                           // :500
            H[BFG_HB_RQ] = R_q;
            R_mod = cm * P.g_mod / P.a;                                       // BaryonCorrection.py:399
            H[BFG_HB_RCUT] = P.eps_mod * R_mod;                               // :410
            H[BFG_HB_LNRCOM] = log(R_mod);                                    // :408
            H[BFG_HB_PAINTCUT] = 0.0;
        }
        H[BFG_HB_LNZ] = P.lnz;
        H[BFG_HB_LNM] = lnM;
        for (int k = 0; k < 3; ++k) {
            if (k < P.ndim) {
                const double x = cols[(1 + k) * n + j];
                H[BFG_HB_X + k] = x;
                if (P.grid) {
                    const int cen = nearest_bin(P.bins, P.N, P.res, x);
                    H[BFG_HB_CX + k] = (double)cen;
                    H[BFG_HB_DX + k] = P.bins[cen] - x;                       // :519-520
                } else {
                    H[BFG_HB_CX + k] = 0.0; H[BFG_HB_DX + k] = 0.0;
                }
            } else {
                H[BFG_HB_X + k] = 0.0; H[BFG_HB_CX + k] = 0.0; H[BFG_HB_DX + k] = 0.0;
            }
        }
        if (P.grid) {
            double ns = floor(Nf / 2.0) * 2.0;                                // int(Nsize // 2) * 2   :501
            ns = fmin(fmax(ns, 2.0), (double)(P.N / 2));                      // np.clip(Nsize, 2, bins.size // 2)  :503
            H[BFG_HB_NSIZE] = ns;
        } else {
            H[BFG_HB_NSIZE] = 0.0;
        }
        if (aux) { aux[j] = R_phys; aux[n + j] = R_mod; }
    }
}

}  // namespace

extern "C" int bfg_box_records(int64_t n_halo, const double *d_cols, int ndim, int grid, int paint, double a, double lnz,
                               double g_run, double g_mod, double eps_run, double eps_mod, double res, double rq_clip,
                               int64_t N, const double *d_bins, double *d_halos, double *d_aux, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(n_halo >= 0 && (ndim == 2 || ndim == 3), "bad argument");
    if (n_halo == 0) return BFG_OK;
    BFG_REQUIRE(d_cols && d_halos, "null argument");
    BFG_REQUIRE(!grid || (d_bins && N >= 2 && N <= 32768 && res > 0), "grid records need the cell centres");
    BoxParams P;
    P.ndim = ndim; P.paint = paint ? 1 : 0; P.grid = grid ? 1 : 0;
    P.a = a; P.lnz = lnz; P.g_run = g_run; P.g_mod = g_mod; P.eps_run = eps_run; P.eps_mod = eps_mod;
    P.res = grid ? res : 1.0; P.rq_clip = rq_clip; P.N = (int)N; P.bins = d_bins;
    int blocks = (int)std::max<i64>(1, std::min<i64>((n_halo + 255) / 256, 148 * 8));
    k_box_records<<<blocks, 256, 0, (cudaStream_t)stream>>>(n_halo, d_cols, P, d_halos, d_aux);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

// ---------------------------------------------------------------------------------------------------- host test entries
// Pure host, no GPU: the per-halo record body above (shell_record_one -- the source k_shell_records runs) on HOST buffers, same
// argument meaning as bfg_shell_records.  (k_box_records stays as written: factoring its body out changed its SASS -- plain instead of
// read-only loads of the cell centres -- and a changed kernel needs a GPU run.)
extern "C" int bfg_test_shell_records_host(int64_t n_halo, const double *h_cols, int paint, double eps_run, double eps_model,
                                           double pixarea, int n_DA, const double *h_DA_x, const double *h_DA_c, int n_g,
                                           const double *h_g_x, const double *h_g_run_c, const double *h_g_mod_c,
                                           double *h_halos, double *h_aux) {
    BFG_REQUIRE(n_halo >= 0, "negative halo count");
    if (n_halo == 0) return BFG_OK;
    BFG_REQUIRE(h_cols && h_halos && h_DA_x && h_DA_c && h_g_x && h_g_run_c, "null argument");
    BFG_REQUIRE(n_DA >= 2 && n_g >= 2, "splines need >= 2 breakpoints");
    BFG_REQUIRE(paint || h_g_mod_c, "baryonify records need the model-cosmology radius spline");
    RecParams P;
    P.paint = paint ? 1 : 0;
    P.eps_run = eps_run; P.eps_model = eps_model; P.pixarea = pixarea;
    P.DA = PPolyView{n_DA, h_DA_x, h_DA_c};
    P.g_run = PPolyView{n_g, h_g_x, h_g_run_c};
    P.g_mod = PPolyView{n_g, h_g_x, h_g_mod_c ? h_g_mod_c : h_g_run_c};
    for (i64 j = 0; j < n_halo; ++j) shell_record_one(n_halo, j, h_cols, P, h_halos, h_aux);
    return BFG_OK;
}
