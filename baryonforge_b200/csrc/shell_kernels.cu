// shell_kernels.cu -- HEALPix-shell runners on sm_100a.
//
//   k_shell_halos<MODE>   : the per-halo loop of BaryonifyShell.process        (BaryonForge/Runners/HealpixRunner.py:315-355),
//                           PaintProfilesShell.process                        (BaryonForge/Runners/HealpixRunner.py:449-481)
//                           and PaintProfilesAnisShell.process                (BaryonForge/Runners/HealpixRunner.py:589-631)
//   k_shell_regrid[_range|_p2p] : the re-binning step                         (BaryonForge/Runners/HealpixRunner.py:357-365, :17-71)
//
// Persistent CTAs (148 SMs x 7) pull halos from a global queue.  Per halo the CTA blends the 2^(ndim-1) table rows into one
// radial row in shared memory, derives the disc's ring range on the device (query_disc, no host round trip), stages the
// per-ring constants, and its warps walk the rings -- 8 or 16 lanes per ring over consecutive pixels, so the fp64 REDs of a
// lane group hit consecutive addresses of the component-major offsets array.
#include <algorithm>
#include <string.h>
#include <vector>
#include "bfg_common.cuh"

using namespace bfg;

namespace {

constexpr int SHELL_THREADS = 128;
#ifndef BFG_SHELL_MIN_CTAS
#define BFG_SHELL_MIN_CTAS 7
#endif
// resident CTAs per SM = size of the persistent grid per SM.  7 -> 72 registers/thread, 28 warps/SM: measured best for the v8
// kernel (flat catalogue 99.7 / 98.6 / 100.0 ms with 8 / 7 / 6; profiles/README.md has the sweeps of the earlier kernels)
constexpr int SHELL_MIN_CTAS = BFG_SHELL_MIN_CTAS;
constexpr int RING_CHUNK = SHELL_THREADS;   // ring segments staged in shared memory per pass (one per thread)

struct HaloSph {
    double vx, vy, vz, theta, phi, D, a, radius, lnz, lnM, rcut, lnRcom, scale, theta_ll, phi_ll, skip;
};

__host__ __device__ __forceinline__ HaloSph load_halo(const double *__restrict__ H) {
    HaloSph s;
    s.vx = BFG_LDG(H + BFG_HS_VX); s.vy = BFG_LDG(H + BFG_HS_VY); s.vz = BFG_LDG(H + BFG_HS_VZ);
    s.theta = BFG_LDG(H + BFG_HS_THETA); s.phi = BFG_LDG(H + BFG_HS_PHI);
    s.D = BFG_LDG(H + BFG_HS_D); s.a = BFG_LDG(H + BFG_HS_A); s.radius = BFG_LDG(H + BFG_HS_RADIUS);
    s.lnz = BFG_LDG(H + BFG_HS_LNZ); s.lnM = BFG_LDG(H + BFG_HS_LNM); s.rcut = BFG_LDG(H + BFG_HS_RCUT);
    s.lnRcom = BFG_LDG(H + BFG_HS_LNRCOM); s.scale = BFG_LDG(H + BFG_HS_SCALE);
    s.theta_ll = BFG_LDG(H + BFG_HS_THETA_LL); s.phi_ll = BFG_LDG(H + BFG_HS_PHI_LL);
    s.skip = BFG_LDG(H + BFG_HS_SKIP);
    return s;
}

// Per-halo constants of the pixel update, hoisted out of the pixel loop.
struct HaloUpd {
    double D, a, pjx, pjy, pjz;   // pos_j = vec_j * D                            HealpixRunner.py:337
    double rcut2;                 // (model eps * R_com * a)^2 : r_com < rcut  <=>  r_sep^2 < rcut2   BaryonCorrection.py:410
    double scale;
    double xq0;                   // ln(r_sep/a) [- ln R_com] = 0.5 ln2 log2(r^2) + xq0              :345, BaryonCorrection.py:408
    double uA, uB, uMax;          // uniform ln r axis: cell coordinate u = log2(r^2) * uA + uB in [0, NR-1]
};

__host__ __device__ __forceinline__ HaloUpd make_upd(const TableView &T, const HaloSph &s) {
    HaloUpd u;
    u.D = s.D; u.a = s.a;
    u.pjx = s.vx * s.D; u.pjy = s.vy * s.D; u.pjz = s.vz * s.D;
    double rc = s.rcut * s.a;
    u.rcut2 = rc * rc;
    u.scale = s.scale;
    u.xq0 = s.lnz - ((T.flags & BFG_TABLE_RDELTA) ? s.lnRcom : 0.0);   // the record's ln(1/a) is s.lnz
    u.uA = 0.34657359027997264 * T.inv_dr;
    u.uB = (u.xq0 - T.r0) * T.inv_dr;
    u.uMax = (double)(T.n[2] - 1);
    return u;
}

// Table value at squared separation r2 (NaN when outside the table / not a positive normal number).
template <bool UNIFORM>
__host__ __device__ __forceinline__ double table_at_l2(const TableView &T, const double *__restrict__ row, const HaloUpd &u,
                                              double l2) {
    if (UNIFORM) {
        const int NR = T.n[2];
        const double uu = fma(l2, u.uA, u.uB);                      // (ln r - r0) / step
        if (!(uu >= 0.0) || !(uu <= u.uMax)) return BFG_QNAN;
        const int k = min((int)uu, NR - 2);
        const double t = uu - (double)k;
        return fma(t, row[k + 1], (1.0 - t) * row[k]);              // (1-t) v0 + t v1, as scipy
    }
    return row_lookup<false>(T, row, fma(l2, 0.34657359027997264, u.xq0));
}

template <bool UNIFORM>
__host__ __device__ __forceinline__ double table_at(const TableView &T, const double *__restrict__ row, const HaloUpd &u,
                                           double r2, const double2 *__restrict__ l2tab) {
    return table_at_l2<UNIFORM>(T, row, u, fast_log2(r2, l2tab));
}

// MODE of the halo loop: the three shell runners share geometry and differ in the per-pixel update
constexpr int MODE_BARYONIFY = 0;   // BaryonifyShell          HealpixRunner.py:315-355
constexpr int MODE_PAINT = 1;       // PaintProfilesShell      HealpixRunner.py:449-481
constexpr int MODE_ANIS = 2;        // PaintProfilesAnisShell  HealpixRunner.py:589-631

// Extra inputs of the anisotropic painter: the tracer ("canvas") table and two map-shaped gathers.
struct AnisArgs {
    TableView T2;            // Tracer_model.projected table (log values)
    const double *mtot;      // halo part of the total-mass map, owned range (HealpixRunner.py:565-570)
    const double *orig;      // LightconeShell.map, owned range
    double mtot_add;         // dV * drho_m: the uniform background added to every pixel of Mtot_map (:582)
};

// One (halo, pixel) update.  (x, y, z) = pixel unit vector; (px, py, pz) = (x, y, z) * D; p0/p1/p2 = the pixel's slots
// in the three offset components (paint: p0 only).
template <int MODE, bool UNIFORM>
__host__ __device__ __forceinline__ void shell_update(const TableView &T, const double *__restrict__ row, const HaloUpd &u,
                                             double x, double y, double z, double px, double py, double pz,
                                             double *__restrict__ p0, double *__restrict__ p1, double *__restrict__ p2,
                                             const double2 *__restrict__ l2tab, const AnisArgs &A,
                                             const double *__restrict__ row2, const HaloUpd &u2) {
    // HealpixRunner.py:338-341  diff = pos - pos_j ; r_sep^2 = sum(diff^2)
    const double dx = px - u.pjx, dy = py - u.pjy, dz = pz - u.pjz;
    const double r2 = dx * dx + dy * dy + dz * dz;
    if (MODE == MODE_ANIS) {
        // p1 / p2 point at this pixel's Mtot_map / orig_map entries
        const double l2 = fast_log2(r2, l2tab);
        const double P = exp(table_at_l2<UNIFORM>(T, row, u, l2));        // Painting  :610
        if (!isfinite(P)) return;                                         // :611 -> 0
        const double C = exp(table_at_l2<UNIFORM>(A.T2, row2, u2, l2));   // Canvas    :612
        if (!isfinite(C)) return;                                         // :613 -> 0
        const double m = *p1 + A.mtot_add;                                // Mtot_map[pixind] (halos + background)
        if (!(m > 0.0)) return;                                           // np.divide(..., where = Mtot > 0)  :614
        const double add = (P * u.scale) * ((C / m) * *p2);               // :615-623
        if (add != 0.0) red_add(p0, add);
        return;
    }
    double val = table_at<UNIFORM>(T, row, u, r2, l2tab);          // :345 / :472 via ln(r_sep/a), no sqrt, no division
    if (MODE == MODE_PAINT) {
        val = exp(val);                            // Tabulate.py:319
        if (!isfinite(val)) return;                // HealpixRunner.py:473 (adds 0)
        val *= u.scale;                            // :478
        if (val != 0.0) red_add(p0, val);          // :481
    } else {
        if (!(r2 < u.rcut2)) return;               // BaryonCorrection.py:410-411: zero beyond the model's cut
        const double sc = (val * u.a) * rsqrt(r2); // offset / r_sep                 HealpixRunner.py:345-346
        // :347 non-finite -> 0 (r_sep = 0, NaN/inf table value, outside the table); exact zeros add nothing
        if (!isfinite(sc) || sc == 0.0) return;
        const double nx = fma(sc, dx, px), ny = fma(sc, dy, py), nz = fma(sc, dz, pz);   // :350 nw_pos = pos + offset
        const double ninv = rsqrt(nx * nx + ny * ny + nz * nz);
        red_add(p0, fma(nx, ninv, -x));            // :351-355
        red_add(p1, fma(ny, ninv, -y));
        red_add(p2, fma(nz, ninv, -z));
    }
}

// A (halo, ring) segment staged in shared memory by the thread that derived it (the warp loop then pays LDS only).
struct __align__(16) RingSeg {
    i64 lbase;             // ring's first pixel - pix_lo
    int nr, ip_lo;
    int cnt, flags;        // flags: 1 = ring straddles the owned pixel range, 2 = equatorial ring (nr = 4 nside), 4 = shifted
    int active, pad;
    double z, sth;         // ring z, sin(theta)
    double pz, sD;         // z * D, sin(theta) * D
    double phase0, inv2nr; // azimuth of pixel ip_lo in half-turns, 2 / nr
    double c0, s0;         // cos / sin of phase0 (equatorial rings)
    double rotC, rotS;     // cos / sin of the azimuth step of one loop iteration (32 pixels; the lane-group width in the fast path)
    double dz, dz2;        // fast path: z - vz of the halo and its square (unit-sphere chord, see span_pixels_fast)
};

template <int MODE, bool UNIFORM, bool CHECK>
__device__ __forceinline__ void ring_pixels(const TableView &T, const double *__restrict__ row, const HaloUpd &u,
                                            const RingSeg &g, double cs, double sn, int lane, double *__restrict__ out,
                                            i64 nloc, const double2 *__restrict__ l2tab, const AnisArgs &A,
                                            const double *__restrict__ row2, const HaloUpd &u2) {
    const int cnt = g.cnt, nr = g.nr;
    const double z = g.z, sth = g.sth, pz = g.pz, sD = g.sD, rotC = g.rotC, rotS = g.rotS;
    double *__restrict__ b0 = out + g.lbase;
    int ip = g.ip_lo + lane;
    if (ip >= nr) ip -= nr;
    for (int i = lane; i < cnt; i += 32) {
        if (!CHECK || (unsigned long long)(g.lbase + ip) < (unsigned long long)nloc) {
            if (MODE == MODE_ANIS)
                shell_update<MODE, UNIFORM>(T, row, u, sth * cs, sth * sn, z, sD * cs, sD * sn, pz, b0 + ip,
                                            const_cast<double *>(A.mtot) + g.lbase + ip,
                                            const_cast<double *>(A.orig) + g.lbase + ip, l2tab, A, row2, u2);
            else
                shell_update<MODE, UNIFORM>(T, row, u, sth * cs, sth * sn, z, sD * cs, sD * sn, pz, b0 + ip,
                                            b0 + nloc + ip, b0 + 2 * nloc + ip, l2tab, A, row2, u2);
        }
        ip += 32;
        if (ip >= nr) ip -= nr;
        const double c2 = cs * rotC - sn * rotS;   // advance the azimuth by 32 pixels
        sn = fma(sn, rotC, cs * rotS);
        cs = c2;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Fast pixel loop of the headline case (BaryonifyShell, uniform ln r axis, ring wholly inside the owned range).
// Same arithmetic as shell_update<MODE_BARYONIFY>, reorganised so that the loop is bounded by the FP64 pipe instead of
// by instruction issue:
//   * everything is done on the UNIT sphere: with pos = D vec, the chord is diff = D (vec - vec_j) =: D d, so
//       ln r_sep = ln D + 0.5 ln |d|^2            (ln D folded into the per-halo table offset)
//       nw_vec   = normalise(vec + sc d),  sc = displacement a / r_sep = (val a / D) rsqrt(|d|^2)
//     D drops out of the new direction, and dz = z - vz is a per-ring constant;
//   * shared memory is addressed with 32-bit shared-window addresses (ld.shared), polynomial constants sit in the
//     constant bank, rsqrt is MUFU.RSQ64H + one cubic iteration without the denormal/inf slow path (|d|^2 = 0 or a
//     non-finite value ends up non-finite and is dropped by the same test as before), the invalid-input test of
//     fast_log2 is subsumed by the table-range test, and a span never wraps (a wrapping ring is walked as two spans).
// ------------------------------------------------------------------------------------------------------------------
struct FastHalo {
    double vx, vy;        // halo unit vector (vz enters through RingSeg.dz)
    double rcut2;         // (model eps * R_com * a / D)^2  on the unit sphere
    double aD;            // a / D   (paint: the record's SCALE, pixarea * D^2 or 1)
    RowLookup t;          // cell coordinate u = log2(|d|^2) * uA + uB   (uB includes ln D and ln(1/a) [- ln R_com])
    unsigned et_s;        // lean loop: shared-window address of the per-halo exponent table E[i] = uB + uA (i + V9_EMIN)
    unsigned rowp_s;      // ... and of the row as (v_k, v_{k+1} - v_k) pairs
    int lean0;            // lean loop allowed as far as the table's range goes (no in-table radius below r^2 = 2^-61)
    double lean_thr;      // ... and every finite node of the (a / D)-scaled row must be smaller than this (|eps| < 2^-6)
};

// ------------------------------------------------------------------------------------------------------------------
// Lean pixel loop (round 2, "v11").  What round 2's experiments said about the v8 loop (profiles/README.md "Round 2"):
// with its REDs compiled out the kernel takes 84-90 ms instead of 94, i.e. the arithmetic alone is as slow as the whole
// kernel, while WITH the REDs the L2 atomic unit is 82-84 % busy (lts__d_atomic_input_cycles_active) -- two ceilings of equal
// height, each hiding the other.  The arithmetic side: ~100 SASS instructions per 32 updates, 47 of them on the FP64 pipe
// (2 issue cycles each), 4 on the XU pipe, 6 branches, ~21 shared-memory wavefronts (half of them bank-conflict replays of
// the 128-entry log2 table), issue slots 66 % busy.  This loop trims all of those:
//   * (x, y) = sin(theta) (cos phi, sin phi) is rotated directly (a rotation is linear: no sth * cs, sth * sn per pixel);
//   * log2 of the mantissa from a 32-entry table held in the lanes' registers (two warp shuffles, no shared-memory bank
//     conflicts) + a degree-4 series (|f| <= 2^-6: abs. error 2.7e-10 in log2 r^2, 1e-10 in ln r -- four orders below
//     the 1e-6 bar);
//   * the exponent's share of the cell coordinate, uB + uA e, comes from a 64-entry per-halo table in shared memory
//     (r^2 in [2^-61, 8) on the unit sphere; anything else reads a NaN entry and lands outside the table);
//   * floor(u) with ONE round-down add of 2^52 + 2^51 (DADD.RM): the cell index is the low word of the sum, its double is
//     the sum minus the constant -- no F2I / I2F (quarter-rate XU instructions);
//   * the table row is pre-multiplied by a / D at blend time and stored as (value, step to the next node) pairs: one
//     16-byte read per update;
//   * 1/sqrt(r^2) = MUFU seed + one Newton step (rel. error 1e-13);
//   * the re-normalisation nw_vec - vec = normalise(vec + sc d) - vec is expanded in eps = |vec + sc d|^2 - 1.  Both vectors
//     are unit vectors, so vec . d = |d|^2 / 2 and eps = sc r^2 (1 + sc) needs no dot product; with
//     dl = 1/sqrt(1 + eps) - 1 = eps (-1/2 + 3/8 eps - 5/16 eps^2) the result is sc (1 + dl) d + dl vec.  dl vec is
//     <~ 1 % of the result, so the truncation (0.27 eps^4) is < 1e-8 of it for |eps| < 2^-6.  No cancellation, unlike
//     fma(nx, ninv, -x);
//   * spans start on a 32-byte sector boundary, so a lane group's RED covers whole sectors (the L2 atomic unit is 82 % busy);
//   * no branch in the body except the one around the three REDs.
// 36 FP64-pipe + 1 XU + ~50 other instructions per 32 updates (profiles/r2_shell_halos_v11_loop.sass), 9 shared-memory
// wavefronts.  A halo takes this loop only if its disc is large (16 lanes per ring) and it is provably safe for it
// (FastHalo.lean, decided per halo): every finite node of its row is small enough that |eps| < 2^-6 anywhere in the disc
// and its table does not reach below r^2 = 2^-61.  Other halos (displacements comparable to their distance, exotic tables)
// take the exact loop (span_pixels_fast).  Deviations from it, all documented boundary ties: a pixel EXACTLY on the table's
// last radial node counts as outside; cell coordinates differ by <= 4e-9 cells, so a pixel within that of a node or of the
// table's edge may land in the neighbouring cell (same value: the interpolant is continuous) or just outside.
// ------------------------------------------------------------------------------------------------------------------
#ifdef BFG_SHELL_V8
constexpr bool SHELL_LEAN = false;     // A/B build: every halo takes the exact loop (span_pixels_fast)
#else
constexpr bool SHELL_LEAN = true;
#endif
constexpr bool SHELL_PRESCALED = true; // the baryonify row holds displacement * a / D
constexpr int SHELL_MAX_PAIR_NODES = 8192;   // radial axes up to this length get the (value, step) pair copy of the row (lean loop)
constexpr int V9_EMIN = -61, V9_NE = 64;          // exponent table: entries 0 .. 63 = uB + uA (i + V9_EMIN), entry 64 = NaN
constexpr double LEAN_EPS_MAX = 0.015625;         // 2^-6

// Per-lane constants of the lean loop's log2: lane l holds entry l of a 32-entry table (bucket centre c_l = 1 + (l + 1/2) / 32 of
// the mantissa): rc = 1 / c_l cut to 17 significant bits and lt = -log2(rc).  The loop fetches its entry from the lane whose
// number is the top 5 mantissa bits with two warp shuffles -- no shared-memory traffic and no bank conflicts (the 128-entry
// shared-memory table of the exact loop costs ~10 wavefronts per warp read because the indices are scattered).  rc's 16
// mantissa bits ride in the low 16 bits of lt's low word (they perturb lt by < 2^-36).
struct Log2Lane { int lt_hi, lt_lo; };

__device__ __forceinline__ Log2Lane make_log2_lane() {
    const int lane = threadIdx.x & 31;
    const double c = 1.0 + ((double)lane + 0.5) * 0.03125;
    const int m16 = (__double2hiint(1.0 / c) >> 4) & 0xffff;                       // 1/c in (1/2, 1): exponent word 0x3fe
    const double rc = __hiloint2double(0x3fe00000 | (m16 << 4), 0);
    const double lt = -log2(rc);
    Log2Lane t;
    t.lt_hi = __double2hiint(lt);
    t.lt_lo = (__double2loint(lt) & 0xffff0000) | m16;
    return t;
}

// One lean pixel loop over this lane's pixels p0 + it * GW, it_lo <= it < n, of a ring span ((cs, sn) = azimuth of *p0; it_lo = 1
// when *p0 itself lies before the disc -- spans start on a sector boundary).  CONVERGENT: all 32 lanes of the warp must call it
// together (lanes without work pass n = 0) -- the trip count is the warp's maximum, because the log2 table lives in the lanes'
// registers.
template <bool CHECK, int GW>
__device__ __forceinline__ void span_pixels_lean(const FastHalo &f, const RingSeg &g, const Log2Lane &L2, double cs, double sn,
                                                 double *__restrict__ p0, int it_lo, int n, i64 nloc8,
                                                 const double *own_lo = nullptr, const double *own_hi = nullptr) {
    const double z = g.z, dz = g.dz, dz2 = g.dz2, rotC = g.rotC, rotS = g.rotS;
    const double vx = f.vx, vy = f.vy, rcut2 = f.rcut2, uA = f.t.uA;
    const unsigned rowp_s = f.rowp_s, nrm2 = (unsigned)f.t.nrm2, et_s = f.et_s;
    double x = g.sth * cs, y = g.sth * sn;
    const int n_warp = __reduce_max_sync(0xffffffffu, n);
    for (int it = 0; it < n_warp; ++it) {
        const double dx = x - vx, dy = y - vy;
        const double r2 = fma(dx, dx, fma(dy, dy, dz2));             // |vec - vec_j|^2   HealpixRunner.py:338-341
        // ---- blended-row value at r2: :345 via ln(r_sep / a) [- ln R_com]
        const int hi = __double2hiint(r2);
        const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(r2));
        const int src = (hi >> 15) & 31;                             // top 5 mantissa bits
        const int w_hi = __shfl_sync(0xffffffffu, L2.lt_hi, src), w_lo = __shfl_sync(0xffffffffu, L2.lt_lo, src);
        const double rc = __hiloint2double(((w_lo << 4) & 0x000ffff0) | 0x3fe00000, 0);
        const unsigned ei = min(((unsigned)hi >> 20) - (unsigned)(1023 + V9_EMIN), (unsigned)V9_NE);
        const double E = lds_f64(et_s + (ei << 3));
        const double fr = fma(m, rc, -1.0);                          // |fr| <= 2^-6
        double p = fma(fr, c_l2p[1], c_l2p[2]);
        p = fma(fr, p, c_l2p[3]);
        p = fma(fr, p, c_l2p[4]);
        const double uu = fma(fma(fr, p, __hiloint2double(w_hi, w_lo)), uA, E);   // (ln r - r0) / step
        const double kk = __dadd_rd(uu, 6755399441055744.0);         // 2^52 + 2^51 + floor(uu)
        const unsigned kl = (unsigned)__double2loint(kk);
        bool ok = (it >= it_lo) & (it < n) & (__double2hiint(kk) == 0x43380000) & (kl <= nrm2);   // a disc pixel, inside [r0, r1)
        const double2 cell = lds_f64x2(rowp_s + (min(kl, nrm2) << 4));             // (v_k, v_{k+1} - v_k), times a / D
        const double tt = uu - (kk - 6755399441055744.0);
        const double val = fma(tt, cell.y, cell.x);
        // ---- offset / r_sep   :345-346
        double y0;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(r2));
        const double e = fma(-r2, y0 * y0, 1.0);
        const double sc = val * fma(y0 * e, 0.5, y0);
        // BaryonCorrection.py:410-411 zero beyond the model's cut; HealpixRunner.py:347 non-finite -> 0; exact zeros add nothing
        ok = ok & (r2 < rcut2) & ((((unsigned)__double2hiint(sc) & 0x7fffffffu) - 1u) < 0x7fefffffu);
        if (CHECK) ok = ok & (p0 >= own_lo) & (p0 < own_hi);
        const double t1 = sc * r2;
        const double eps = fma(t1, sc, t1);                          // |vec + sc d|^2 - 1
        const double dl = eps * fma(eps, fma(eps, -0.3125, 0.375), -0.5);
        const double q = fma(sc, dl, sc);
        const double ox = fma(q, dx, dl * x), oy = fma(q, dy, dl * y), oz = fma(q, dz, dl * z);   // :350-355
#ifdef BFG_SHELL_NO_RED   // diagnostic build: the arithmetic without its scatter-add (where does the time go?)
        if (ok && ox + oy + oz == 1.2345e300) red_add(p0, ox);      // never true; keeps the arithmetic alive
#else
        if (ok) {
            red_add(p0, ox);
            red_add((double *)((char *)p0 + nloc8), oy);
            red_add((double *)((char *)p0 + 2 * nloc8), oz);
        }
#endif
        p0 += GW;
        const double x2 = x * rotC - y * rotS;                       // advance the azimuth by GW pixels
        y = fma(y, rotC, x * rotS);
        x = x2;
    }
}

template <bool PAINT>
__device__ __forceinline__ FastHalo make_fast(const TableView &T, const HaloSph &s, const HaloUpd &u, const double *row,
                                              const double2 *l2tab, const double *etab) {
    // (the row pairs of the lean loop follow the plain row in dynamic shared memory, 16-byte aligned)
    FastHalo f;
    f.vx = s.vx; f.vy = s.vy;
    const double rc = s.rcut * s.a / s.D;
    f.rcut2 = rc * rc;
    f.aD = PAINT ? s.scale : (SHELL_PRESCALED ? 1.0 : s.a / s.D);   // baryonify: a / D is folded into the row at blend time
    f.t.uA = u.uA;
    f.t.uB = fma(2.0 * log2(s.D), u.uA, u.uB);      // log2 r_sep^2 = log2 |d|^2 + 2 log2 D
    f.t.uMax = u.uMax;
    f.t.nrm2 = T.n[2] - 2;
    f.t.row_s = (unsigned)__cvta_generic_to_shared(row);
    f.t.l2_s = (unsigned)__cvta_generic_to_shared(l2tab);
    f.et_s = (unsigned)__cvta_generic_to_shared(etab);
    f.rowp_s = (unsigned)__cvta_generic_to_shared(row + ((T.n[2] + 1) & ~1));
    // lean loop (span_pixels_lean): in-table radii must have exponents the 64-entry table covers, and
    // |eps| = |row| (r + |row|) <= m (rmax + m) < 2^-6 with rmax = the disc's radius, m = max |row|
    f.lean0 = (f.t.uB <= -(double)V9_EMIN * f.t.uA) ? 1 : 0;
    f.lean_thr = 0.5 * (sqrt(fma(s.radius, s.radius, 4.0 * LEAN_EPS_MAX)) - s.radius);
    return f;   // handed to the other warps through shared memory (HaloCtx), which also keeps ptxas from re-deriving it
}

// Lanes per ring in the fast path: a warp walks 32 / GW rings at once, GW consecutive pixels of each per iteration.  Disc chords are short (34 pixels on average for a mass-function-like catalogue, ~105 for the flat one), so
// narrow groups keep the lanes busy (34 pixels: 53 % of the lane slots with 32-wide groups, 85 % with 8-wide ones) and
// the per-ring set-up is paid once per FOUR rings.  The REDs of a group still cover whole 32-byte sectors.
// 8 lanes per ring for small discs, 16 for large ones (measured: flat catalogue 110.3 / 105.8 / 102.5 ms with 32 / 8 / 16
// lanes, mass-function-like catalogue 25.2 / 20.6 / 22.0 ms); chosen per halo from the disc's mean chord.
constexpr int GW_SMALL = 8, GW_LARGE = 16;
constexpr double GW_CHORD_SPLIT = 80.0;       // mean chord [pixels] above which GW_LARGE is used

// One contiguous span of a ring (no wrap): lane handles pixels p0, p0 + GW, ... < pend;  (cs, sn) = azimuth of *p0.
// p0 / pend point into component 0 of the offsets; components 1, 2 live nloc8 and 2 nloc8 bytes further.
// CHECK: the ring straddles the owned pixel range [own_lo, own_hi) (ring-range sharding) -- same arithmetic, so results
// do not depend on how the map is sharded.
template <bool CHECK, int GW>
__device__ __forceinline__ void span_pixels_fast(const FastHalo &f, const RingSeg &g, double cs, double sn,
                                                 double *__restrict__ p0, const double *__restrict__ pend, i64 nloc8,
                                                 const double *own_lo = nullptr, const double *own_hi = nullptr) {
    const double z = g.z, sth = g.sth, dz = g.dz, dz2 = g.dz2, rotC = g.rotC, rotS = g.rotS;
    for (; p0 < pend; p0 += GW) {
        const double x = sth * cs, y = sth * sn;
        const double dx = x - f.vx, dy = y - f.vy;
        const double r2 = fma(dx, dx, fma(dy, dy, dz2));             // |vec - vec_j|^2   HealpixRunner.py:338-341
        bool ok;
        const double val = row_at_r2(f.t, r2, ok);                   // :345 via ln(r_sep / a) [- ln R_com], no sqrt / division
        const double sc = (val * f.aD) * rsqrt_pos(r2);              // offset / r_sep   HealpixRunner.py:345-346
        // BaryonCorrection.py:410-411 zero beyond the model's cut; HealpixRunner.py:347 non-finite -> 0; exact zeros (and
        // denormal-sized offsets) add nothing: one integer test on the exponent field covers NaN, inf and 0
        ok = ok && (r2 < f.rcut2) && ((((unsigned)__double2hiint(sc) & 0x7fffffffu) - 1u) < 0x7fefffffu);
        if (CHECK) ok = ok && (p0 >= own_lo) && (p0 < own_hi);
        const double nx = fma(sc, dx, x), ny = fma(sc, dy, y), nz = fma(sc, dz, z);      // :350 (direction of nw_pos)
        const double ninv = rsqrt_pos(fma(nx, nx, fma(ny, ny, nz * nz)));
        if (ok) {
            red_add(p0, fma(nx, ninv, -x));                          // :351-355
            red_add((double *)((char *)p0 + nloc8), fma(ny, ninv, -y));
            red_add((double *)((char *)p0 + 2 * nloc8), fma(nz, ninv, -z));
        }
        const double c2 = cs * rotC - sn * rotS;                     // advance the azimuth by GW pixels
        sn = fma(sn, rotC, cs * rotS);
        cs = c2;
    }
}

// PaintProfilesShell counterpart of span_pixels_fast (HealpixRunner.py:464-481): map[p] += exp(table(ln(r_sep / a))) * SCALE,
// non-finite read-outs (outside the table, log of a zero or negative profile) contribute nothing.
template <bool CHECK, int GW>
__device__ __forceinline__ void span_pixels_paint(const FastHalo &f, const RingSeg &g, double cs, double sn,
                                                  double *__restrict__ p0, const double *__restrict__ pend,
                                                  const double *own_lo = nullptr, const double *own_hi = nullptr) {
    const double sth = g.sth, dz2 = g.dz2, rotC = g.rotC, rotS = g.rotS;
    for (; p0 < pend; p0 += GW) {
        const double dx = fma(sth, cs, -f.vx), dy = fma(sth, sn, -f.vy);
        const double r2 = fma(dx, dx, fma(dy, dy, dz2));             // |vec - vec_j|^2   HealpixRunner.py:466-469
        bool ok;
        double val = exp(row_at_r2(f.t, r2, ok));                    // Tabulate.py:319
        val *= f.aD;                                                 // :478 (SCALE)
        // :473 non-finite -> 0; exact zeros add nothing (one integer test on the exponent field covers NaN, inf and 0)
        ok = ok && ((((unsigned)__double2hiint(val) & 0x7fffffffu) - 1u) < 0x7fefffffu);
        if (CHECK) ok = ok && (p0 >= own_lo) && (p0 < own_hi);
        if (ok) red_add(p0, val);                                    // :481
        const double c2 = cs * rotC - sn * rotS;                     // advance the azimuth by GW pixels
        sn = fma(sn, rotC, cs * rotS);
        cs = c2;
    }
}

// Per-halo constants, computed ONCE per halo by warp 0 and read by the other warps from shared memory (they used to be
// re-derived by all four warps: ~700 instructions each -- 5 cosines, 3 square roots, 2 divisions, a log2 -- which is a
// fifth of all instructions for a catalogue of small discs).
struct HaloCtx {
    DiscRings d;
    HaloUpd u;
    FastHalo fh;
    int gw_small, touches;
};

// Fast ring walk of one staged chunk: warp w takes ring groups w, w + 4, ...; inside a group of 32 / GW rings each ring
// gets GW lanes.  Returns the number of (halo, pixel) updates owned by this lane.
template <int GW, bool PAINT>
__device__ __forceinline__ i64 walk_rings_fast(const FastHalo &fh, const RingSeg *__restrict__ segs, int nseg, bool valid,
                                               bool sharded, double eqC, double eqS, double *__restrict__ out, i64 nloc,
                                               i64 nloc8, int wi = (int)(threadIdx.x >> 5), int nw = SHELL_THREADS / 32) {
    // warp wi of the nw warps that share `segs` takes ring groups wi, wi + nw, ...  (nw = 1: the warp-per-halo kernel)
    constexpr int NG = 32 / GW;
    const int lane = threadIdx.x & 31;
    const int li = lane & (GW - 1), gi = lane / GW;
    i64 done = 0;
    for (int rb0 = wi * NG; rb0 < nseg; rb0 += nw * NG) {
        const int r = rb0 + gi;
        if (r >= nseg) continue;
        const RingSeg &g = segs[r];
        if (!g.active) continue;
        const int cnt = g.cnt;
        // updates of this ring owned by this lane (the pixel loop itself carries no counter)
        if (!(g.flags & 1)) {
            done += (cnt > li) ? ((cnt - li + GW - 1) / GW) : 0;
        } else {
            int ip = g.ip_lo + li;
            if (ip >= g.nr) ip -= g.nr;
            for (int i = li; i < cnt; i += GW) {
                done += ((unsigned long long)(g.lbase + ip) < (unsigned long long)nloc) ? 1 : 0;
                ip += GW;
                if (ip >= g.nr) ip -= g.nr;
            }
        }
        if (!valid) continue;   // halo outside the table in (z, M, extras): every read-out is NaN -> adds nothing
        double cs, sn;
        if (g.flags & 2) {      // equatorial: rotate the staged (c0, s0) by this lane's cached step
            cs = g.c0 * eqC - g.s0 * eqS;
            sn = fma(g.s0, eqC, g.c0 * eqS);
        } else {
            sincospi(fma((double)li, g.inv2nr, g.phase0), &sn, &cs);
        }
        // span A: [ip_lo, min(ip_lo + cnt, nr)); span B (disc straddles phi = 0): [0, ip_lo + cnt - nr)
        const int endA = min(g.ip_lo + cnt, g.nr);
        const int endB = g.ip_lo + cnt - g.nr;
        double *rbp = out + g.lbase;
        if (PAINT) {
            if (!sharded) {
                span_pixels_paint<false, GW>(fh, g, cs, sn, rbp + g.ip_lo + li, rbp + endA);
                if (endB > 0) {
                    sincospi(((double)li + ((g.flags & 4) ? 0.5 : 0.0)) * g.inv2nr, &sn, &cs);
                    span_pixels_paint<false, GW>(fh, g, cs, sn, rbp + li, rbp + endB);
                }
            } else {
                span_pixels_paint<true, GW>(fh, g, cs, sn, rbp + g.ip_lo + li, rbp + endA, out, out + nloc);
                if (endB > 0) {
                    sincospi(((double)li + ((g.flags & 4) ? 0.5 : 0.0)) * g.inv2nr, &sn, &cs);
                    span_pixels_paint<true, GW>(fh, g, cs, sn, rbp + li, rbp + endB, out, out + nloc);
                }
            }
        } else if (!sharded) {
            span_pixels_fast<false, GW>(fh, g, cs, sn, rbp + g.ip_lo + li, rbp + endA, nloc8);
            if (endB > 0) {
                sincospi(((double)li + ((g.flags & 4) ? 0.5 : 0.0)) * g.inv2nr, &sn, &cs);
                span_pixels_fast<false, GW>(fh, g, cs, sn, rbp + li, rbp + endB, nloc8);
            }
        } else {
            span_pixels_fast<true, GW>(fh, g, cs, sn, rbp + g.ip_lo + li, rbp + endA, nloc8, out, out + nloc);
            if (endB > 0) {
                sincospi(((double)li + ((g.flags & 4) ? 0.5 : 0.0)) * g.inv2nr, &sn, &cs);
                span_pixels_fast<true, GW>(fh, g, cs, sn, rbp + li, rbp + endB, nloc8, out, out + nloc);
            }
        }
    }
    return done;
}

// Ring walk of the lean loop: the same assignment of rings to lane groups as walk_rings_fast, but written CONVERGENT -- every
// lane of the warp reaches every span_pixels_lean call, lanes without a ring (or without a wrap-around span) with n = 0.
template <int GW>
__device__ __forceinline__ i64 walk_rings_lean(const FastHalo &fh, const Log2Lane &L2, const RingSeg *__restrict__ segs, int nseg,
                                               bool sharded, const double2 *__restrict__ eq, double *__restrict__ out, i64 nloc,
                                               i64 nloc8) {
    constexpr int NG = 32 / GW;
    const int lane = threadIdx.x & 31;
    const int li = lane & (GW - 1), gi = lane / GW;
    i64 done = 0;
    for (int rb0 = (threadIdx.x >> 5) * NG; rb0 < nseg; rb0 += (SHELL_THREADS / 32) * NG) {   // warp-uniform
        const int r = rb0 + gi;
        const RingSeg &g = segs[min(r, nseg - 1)];
        const bool have = (r < nseg) && g.active;
        const int cnt = have ? g.cnt : 0;
        // span A: [ip_lo, min(ip_lo + cnt, nr)); span B (disc straddles phi = 0): [0, ip_lo + cnt - nr).
        // Span A is walked from the 32-byte sector boundary at or before ip_lo (ring starts are multiples of 4 pixels), so that
        // every RED of a lane group covers whole sectors of the offsets array: the L2 atomic unit is what binds this kernel
        // (84 % busy), and an unaligned 16-lane group touches 5 sectors instead of 4.  Lanes before ip_lo idle for one iteration.
        const int endA = min(g.ip_lo + cnt, g.nr);
        const int endB = g.ip_lo + cnt - g.nr;
#ifdef BFG_SHELL_NO_ALIGN   // A/B build
        const int s_al = 0;
#else
        const int s_al = g.ip_lo & 3;
#endif
        const int first = g.ip_lo - s_al + li;
        const int itA = (li < s_al) ? 1 : 0;
        const int nA = have ? max(0, (endA - first + GW - 1) / GW) : 0;
        const int nB = (have && endB > li) ? (endB - li + GW - 1) / GW : 0;
        const bool straddles = have && (g.flags & 1);
        if (!straddles) {
            done += max(0, nA - itA) + nB;
        } else {   // updates of this ring owned by this lane (the pixel loop itself carries no counter)
            int ip = g.ip_lo + li;
            if (ip >= g.nr) ip -= g.nr;
            for (int i = li; i < cnt; i += GW) {
                done += ((unsigned long long)(g.lbase + ip) < (unsigned long long)nloc) ? 1 : 0;
                ip += GW;
                if (ip >= g.nr) ip -= g.nr;
            }
        }
        double cs = 1.0, sn = 0.0;
        if (have) {
            if (g.flags & 2) {      // equatorial: rotate the staged (c0, s0) (azimuth of pixel ip_lo) by li - s_al pixels
                const double2 e = eq[3 + li - s_al];
                cs = g.c0 * e.x - g.s0 * e.y;
                sn = fma(g.s0, e.x, g.c0 * e.y);
            } else {
                sincospi(fma((double)(li - s_al), g.inv2nr, g.phase0), &sn, &cs);
            }
        }
        double *rbp = out + g.lbase;
        const bool any_straddle = sharded && __any_sync(0xffffffffu, straddles);   // warp-uniform
        if (!any_straddle) span_pixels_lean<false, GW>(fh, g, L2, cs, sn, rbp + first, itA, nA, nloc8);
        else span_pixels_lean<true, GW>(fh, g, L2, cs, sn, rbp + first, itA, nA, nloc8, out, out + nloc);
        if (__any_sync(0xffffffffu, nB > 0)) {
            if (nB > 0) sincospi(((double)li + ((g.flags & 4) ? 0.5 : 0.0)) * g.inv2nr, &sn, &cs);
            if (!any_straddle) span_pixels_lean<false, GW>(fh, g, L2, cs, sn, rbp + li, 0, nB, nloc8);
            else span_pixels_lean<true, GW>(fh, g, L2, cs, sn, rbp + li, 0, nB, nloc8, out, out + nloc);
        }
    }
    return done;
}

template <int MODE, bool UNIFORM>
__global__ void __launch_bounds__(SHELL_THREADS, SHELL_MIN_CTAS)
k_shell_halos(TableView T, Hpx h, i64 n_halo, const double *__restrict__ halos, const double *__restrict__ extras,
              int n_extra, double *__restrict__ out, i64 pix_lo, i64 pix_hi, unsigned long long *nupd,
              const double2 *__restrict__ g_l2tab, AnisArgs A, unsigned long long *queue, const int *__restrict__ d_min_rings) {
    // min_rings: discs of at most this many rings belong to k_shell_halos_warp (0: this kernel takes every halo); decided on the
    // device by k_warp_vote, so the launch sequence does not depend on the catalogue
    const int min_rings = d_min_rings ? __ldg(d_min_rings) : 0;
    constexpr bool PAINT = (MODE != MODE_BARYONIFY);
    constexpr bool FAST = (MODE != MODE_ANIS) && UNIFORM;        // span_pixels_fast / _paint: 8 or 16 lanes per ring
    extern __shared__ double row[];
    __shared__ i64 s_j;
    const double *row2 = row + ((MODE == MODE_ANIS) ? T.n[2] : 0);   // anis: the tracer row follows the paint row
    __shared__ RingSeg segs[RING_CHUNK];
    __shared__ double2 l2tab[BFG_LOG2_TAB];
    __shared__ int s_cnt[SHELL_THREADS / 32];
    load_log2_table(l2tab, g_l2tab);   // visible after the first __syncthreads() below
    const Log2Lane L2 = make_log2_lane();
    const int lane = threadIdx.x & 31;
    const i64 nloc = pix_hi - pix_lo;
    i64 nloc8 = nloc * 8;
    asm volatile("" : "+l"(nloc8));   // keep the component stride in a register pair (else recomputed per pixel)
    // azimuth of `lane` pixels on an equatorial ring (every equatorial ring has 4 nside pixels): computed once; the fast
    // path needs the offsets 0 .. GW-1 of its lane group instead, tabulated once for both group widths
    double eqC, eqS;
    sincospi((double)lane * (2.0 / (double)h.nl4), &eqS, &eqC);
    // rotation by k pixels of an equatorial ring, k = -3 .. GW-1, for both group widths (entry k + 3 of each set): the lean
    // walk starts its spans on a 32-byte sector boundary, up to 3 pixels before the first pixel of the disc
    __shared__ double2 s_eq[(GW_SMALL + 3) + (GW_LARGE + 3)];
    __shared__ HaloCtx s_ctx;
    __shared__ double s_etab[V9_NE + 1];   // lean loop: uB + uA * exponent per halo; the last entry stays NaN
    if (threadIdx.x == 0) s_etab[V9_NE] = CUDART_NAN;
    if (FAST && threadIdx.x < (GW_SMALL + 3) + (GW_LARGE + 3)) {
        const int k = ((threadIdx.x < GW_SMALL + 3) ? threadIdx.x : threadIdx.x - (GW_SMALL + 3)) - 3;
        double sk, ck;
        sincospi((double)k * (2.0 / (double)h.nl4), &sk, &ck);
        s_eq[threadIdx.x] = make_double2(ck, sk);
    }
    i64 done = 0;
    const bool sharded = pix_lo > 0 || pix_hi < h.npix;

    // Persistent CTAs (SHELL_MIN_CTAS per SM) pull halos from a global queue: consecutive halos of the sky-sorted batch go to whichever
    // CTA is free, so the halos in flight stay neighbours on the sky (L2 locality) and the tail is balanced.
    for (;;) {
        __syncthreads();  // previous halo's row / segments / context no longer in use
        if (threadIdx.x < 32) {   // warp 0: next halo from the queue + its per-halo constants
            i64 jn = 0;
            if (lane == 0) jn = (i64)atomicAdd(queue, 1ULL);
            jn = ((i64)__shfl_sync(0xffffffffu, (int)(jn >> 32), 0) << 32) | (unsigned)__shfl_sync(0xffffffffu, (int)jn, 0);
            if (jn < n_halo) {
                const HaloSph s0 = load_halo(halos + jn * BFG_HALO_STRIDE);
                if (s0.skip != 0.0) {
                    jn = n_halo;   // bfg_halo_sort_owned puts the halos of other ranks last and marks them: stop here
                } else {
                    const DiscRings d0 = disc_rings(h, s0.theta, s0.phi, s0.radius);
                    // (a disc without any ring -- rb < ra -- still owes the < 4-pixel fallback: only min_rings > 0 may skip it)
                    const int tch = (min_rings == 0 || d0.rb - d0.ra + 1 > (i64)min_rings) &&
                                    (!sharded || disc_touches_range(h, d0, pix_lo, pix_hi));
                    if (lane == 0) s_ctx.touches = tch;
                    // a halo this kernel does not take (k_shell_halos_warp had it, or its disc misses the owned range) costs
                    // no more than the ring range of its disc: the constants below are ~500 instructions
                    if (tch) {
                        const HaloUpd u0 = make_upd(T, s0);
                        // lane-group width of the fast path from the disc's mean chord (pi/4 of its diameter) in pixels
                        const int gws = (0.7853981633974483 * 2.0 * s0.radius) * sqrt((double)h.npix * 0.07957747154594767)
                                        < GW_CHORD_SPLIT;
                        if (FAST) {
                            const FastHalo f0 = make_fast<PAINT>(T, s0, u0, row, l2tab, s_etab);
                            if (lane == 0) s_ctx.fh = f0;
                            if (!PAINT) {
#pragma unroll
                                for (int i = lane; i < V9_NE; i += 32) s_etab[i] = fma((double)(i + V9_EMIN), f0.t.uA, f0.t.uB);
                            }
                        }
                        if (lane == 0) { s_ctx.d = d0; s_ctx.u = u0; s_ctx.gw_small = gws; }
                    }
                }
            }
            if (lane == 0) s_j = jn;
        }
        __syncthreads();
        const i64 j = s_j;
        if (j >= n_halo) break;
        if (!s_ctx.touches) continue;   // uniform across the block (ring-range sharding: the disc misses the owned range)
        const HaloSph s = load_halo(halos + j * BFG_HALO_STRIDE);
        const DiscRings d = s_ctx.d;
        bool valid;
        // fast baryonify loops: the row holds displacement * a / D (the unit-sphere offset per unit chord is row / |d|)
        constexpr bool PRESCALED = FAST && !PAINT && SHELL_PRESCALED;
        // (value, step) pairs for the lean loop, unless the radial axis is too long for a second copy in shared memory
        // ... and unless the disc is small: a disc of ~1500 pixels is 11 updates per thread, the lean loop's extra per-halo work
        // (second row copy, lean test, convergent walk) then costs more than its loop saves (measured on the dn/dlogM ~ M^-0.9
        // catalogue: 34.2 ms with, 27.8 ms without)
        const bool gw_small = s_ctx.gw_small != 0;
        const bool pairs = PRESCALED && T.n[2] <= SHELL_MAX_PAIR_NODES && !gw_small;
        if (pairs) blend_row_pairs(T, s.lnz, s.lnM, extras ? extras + j * n_extra : nullptr, row,
                                   (double2 *)(row + ((T.n[2] + 1) & ~1)), valid, s.a / s.D);
        else blend_row(T, s.lnz, s.lnM, extras ? extras + j * n_extra : nullptr, row, valid, PRESCALED ? s.a / s.D : 1.0);
        HaloUpd u = s_ctx.u;
        if (PRESCALED) u.a = u.D;       // generic update on a prescaled row: sc = (row * D) / r_sep   (the < 4-pixel fallback)
        HaloUpd u2 = u;
        if (MODE == MODE_ANIS) {
            bool valid2;
            blend_row(A.T2, s.lnz, s.lnM, extras ? extras + j * n_extra : nullptr, row + T.n[2], valid2);
            valid = valid && valid2;   // a NaN Painting or a NaN Canvas both contribute nothing
            u2 = make_upd(A.T2, s);
        }
        FastHalo fh;
        int lean_bad = 0;   // this thread saw a finite row node too large for the lean loop's series (span_pixels_lean)
        if (FAST) {
            fh = s_ctx.fh;
            const double2 e = gw_small ? s_eq[3 + (lane & (GW_SMALL - 1))] : s_eq[(GW_SMALL + 3) + 3 + (lane & (GW_LARGE - 1))];
            eqC = e.x; eqS = e.y;
            if (PRESCALED && SHELL_LEAN && pairs) {
                if (!fh.lean0) lean_bad = 1;
                for (int k = threadIdx.x; k < T.n[2]; k += SHELL_THREADS) {   // the nodes this thread blended itself
                    const double av = fabs(row[k]);
                    if (av >= fh.lean_thr && av < CUDART_INF) lean_bad = 1;
                }
            } else {
                lean_bad = 1;
            }
        }
        // `if pixind.size < 4` (HealpixRunner.py:333) can only trigger for discs of a few pixels (<= ~12 rings)
        const bool tiny = !PAINT && (s.radius * s.radius * (double)h.npix * 0.25 < 64.0);

        if (tiny) {   // count the disc (block-wide); discs with no pixel centre at all land here too
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
                        ++done;
                        if (valid) {
                            double x, y, z;
                            pix2vec(h, p, x, y, z);
                            double *q = out + (p - pix_lo);
                            shell_update<MODE, UNIFORM>(T, row, u, x, y, z, x * u.D, y * u.D, z * u.D, q, q + nloc,
                                                        q + 2 * nloc, l2tab, A, row2, u2);
                        }
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
                g.cnt = 0; g.active = 0;
                if (iz <= d.rb) {
                    i64 start, nr, ip_lo, cnt; bool sh;
                    disc_ring_span(h, d, iz, start, nr, sh, ip_lo, cnt);
                    if (cnt > 0 && start < pix_hi && start + nr > pix_lo) {
                        g.active = 1;
                        g.lbase = start - pix_lo;
                        g.nr = (int)nr; g.ip_lo = (int)ip_lo; g.cnt = (int)cnt;
                        g.flags = ((start < pix_lo || start + nr > pix_hi) ? 1 : 0) | ((nr == h.nl4) ? 2 : 0) | (sh ? 4 : 0);
                        ring_z_sth(h, iz, g.z, g.sth);
                        g.pz = g.z * u.D; g.sD = g.sth * u.D;
                        g.dz = g.z - s.vz; g.dz2 = g.dz * g.dz;
                        g.inv2nr = 2.0 / (double)nr;
                        g.phase0 = ((double)ip_lo + (sh ? 0.5 : 0.0)) * g.inv2nr;
                        sincospi(g.phase0, &g.s0, &g.c0);
                        sincospi((FAST ? (gw_small ? (double)GW_SMALL : (double)GW_LARGE) : 32.0) * g.inv2nr, &g.rotS, &g.rotC);
                    }
                }
                segs[threadIdx.x] = g;
            }
            const bool lean = !__syncthreads_or(lean_bad);   // barrier: segments + row ready; and is the halo fit for the lean loop?
            const int nseg = (int)min((i64)RING_CHUNK, d.rb - base + 1);

            // ---- fast path: each warp walks 32 / GW rings at once, GW lanes per ring ------------------------------------
            if (FAST) {
                if (!PAINT && lean && valid) {   // block-uniform; lean implies a large disc (16 lanes per ring)
                    done += walk_rings_lean<GW_LARGE>(fh, L2, segs, nseg, sharded, s_eq + (GW_SMALL + 3), out, nloc, nloc8);
                } else {
                    if (gw_small) done += walk_rings_fast<GW_SMALL, PAINT>(fh, segs, nseg, valid, sharded, eqC, eqS, out, nloc, nloc8);
                    else done += walk_rings_fast<GW_LARGE, PAINT>(fh, segs, nseg, valid, sharded, eqC, eqS, out, nloc, nloc8);
                }
            }
            // ---- generic path: warps take rings round-robin; lanes walk consecutive pixels ------------------------
            // static round-robin: neighbouring rings have neighbouring lengths, so the 4 warps stay balanced without a
            // shared work counter (whose atomic + shuffle latency was ~12 % of the kernel's stall samples)
            for (int r = threadIdx.x >> 5; !FAST && r < nseg; r += SHELL_THREADS / 32) {
                const RingSeg &g = segs[r];
                if (!g.active) continue;
                const int cnt = g.cnt;
                // updates of this ring owned by this lane (the pixel loop itself carries no counter)
                if (!(g.flags & 1)) {
                    done += (cnt > lane) ? ((cnt - lane + 31) >> 5) : 0;
                } else {
                    int ip = g.ip_lo + lane;
                    if (ip >= g.nr) ip -= g.nr;
                    for (int i = lane; i < cnt; i += 32) {
                        done += ((unsigned long long)(g.lbase + ip) < (unsigned long long)nloc) ? 1 : 0;
                        ip += 32;
                        if (ip >= g.nr) ip -= g.nr;
                    }
                }
                if (!valid) continue;   // halo outside the table in (z, M, extras): every read-out is NaN -> adds nothing
                double cs, sn;
                if (g.flags & 2) {      // equatorial: rotate the staged (c0, s0) by this lane's cached step
                    cs = g.c0 * eqC - g.s0 * eqS;
                    sn = fma(g.s0, eqC, g.c0 * eqS);
                } else {
                    sincospi(fma((double)lane, g.inv2nr, g.phase0), &sn, &cs);
                }
                if (g.flags & 1) {
                    ring_pixels<MODE, UNIFORM, true>(T, row, u, g, cs, sn, lane, out, nloc, l2tab, A, row2, u2);
                } else {
                    ring_pixels<MODE, UNIFORM, false>(T, row, u, g, cs, sn, lane, out, nloc, l2tab, A, row2, u2);
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

// ------------------------------------------------------------------------------------------------------------------
// Small discs: ONE WARP PER HALO.  A disc of the mass-function-like catalogue has ~1500 pixels on ~40 rings: in the CTA-per-halo
// kernel above that is 11 updates per thread behind three block barriers, a per-halo set-up done by one warp while three wait,
// and a ring-staging pass that uses a third of the threads -- the fixed cost per halo, not the pixel loop, sets the pace
// (4.7e10 updates/s against 1.7e11 for large discs).  Here every warp pulls its own halos from the queue, blends its own row,
// stages its rings 32 at a time and walks them with the same lane groups and the same exact pixel loop -- no block barrier
// anywhere, four independent halos in flight per CTA.  Discs of more than WARP_MAX_RINGS rings are left to k_shell_halos
// (launched with min_rings = WARP_MAX_RINGS), which skips the ones taken here; both add into the same array.
// ------------------------------------------------------------------------------------------------------------------
constexpr int WARP_MAX_RINGS = 64;

// Which catalogues go through the warp kernel?  Measured (B200, NSIDE = 4096, 10^6 halos): it pays when nearly every disc is small
// (dn/dlogM ~ M^-0.9: 22.2 -> 20.0 ms) and costs when the catalogue is mixed (flat in log M: 96.3 -> 99.6 ms; two persistent
// kernels in a row, two tails), whatever the ring threshold.  So the decision is a vote over the batch, taken on the device:
// vote[0] = max_rings when at least 80 % of the (owned) halos have discs of at most max_rings rings, else 0.
__global__ void k_warp_vote_count(Hpx h, i64 n_halo, const double *__restrict__ halos, int max_rings, unsigned long long *cnt) {
    unsigned long long small = 0, total = 0;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < n_halo; j += (i64)gridDim.x * blockDim.x) {
        const double *H = halos + j * BFG_HALO_STRIDE;
        if (__ldg(H + BFG_HS_SKIP) != 0.0) continue;
        const DiscRings d = disc_rings(h, __ldg(H + BFG_HS_THETA), __ldg(H + BFG_HS_PHI), __ldg(H + BFG_HS_RADIUS));
        ++total;
        small += (d.rb - d.ra + 1 <= (i64)max_rings) ? 1 : 0;
    }
    small = (unsigned long long)warp_sum_i64((i64)small);
    total = (unsigned long long)warp_sum_i64((i64)total);
    if ((threadIdx.x & 31) == 0 && total) { atomicAdd(cnt, small); atomicAdd(cnt + 1, total); }
}

__global__ void k_warp_vote_finish(const unsigned long long *cnt, int max_rings, int force, int *vote) {
    vote[0] = (force > 0 || (force == 0 && cnt[0] * 10 >= cnt[1] * 8 && cnt[1] > 0)) ? max_rings : 0;
}
constexpr int WARP_KERNEL_CTAS = 6;

template <int MODE>
__global__ void __launch_bounds__(SHELL_THREADS, WARP_KERNEL_CTAS)
k_shell_halos_warp(TableView T, Hpx h, i64 n_halo, const double *__restrict__ halos, const double *__restrict__ extras,
                   int n_extra, double *__restrict__ out, i64 pix_lo, i64 pix_hi, unsigned long long *nupd,
                   const double2 *__restrict__ g_l2tab, unsigned long long *queue, int row_stride,
                   const int *__restrict__ d_max_rings) {
    constexpr bool PAINT = (MODE != MODE_BARYONIFY);
    constexpr bool UNIFORM = true;
    const int max_rings = __ldg(d_max_rings);
    if (max_rings <= 0) return;                            // k_warp_vote: this catalogue is left to k_shell_halos
    extern __shared__ double rows[];                       // [warps][row_stride]
    __shared__ RingSeg segs_all[SHELL_THREADS / 32][32];
    __shared__ double2 l2tab[BFG_LOG2_TAB];
    __shared__ double2 s_eq[GW_SMALL];
    load_log2_table(l2tab, g_l2tab);
    if (threadIdx.x < GW_SMALL) {
        double sk, ck;
        sincospi((double)threadIdx.x * (2.0 / (double)h.nl4), &sk, &ck);
        s_eq[threadIdx.x] = make_double2(ck, sk);
    }
    __syncthreads();                                       // the only block barrier: tables ready
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double *row = rows + (size_t)wid * row_stride;
    RingSeg *segs = segs_all[wid];
    const i64 nloc = pix_hi - pix_lo;
    i64 nloc8 = nloc * 8;
    asm volatile("" : "+l"(nloc8));
    const bool sharded = pix_lo > 0 || pix_hi < h.npix;
    const double2 e = s_eq[lane & (GW_SMALL - 1)];
    const double eqC = e.x, eqS = e.y;
    const AnisArgs A = {};
    i64 done = 0;
    for (;;) {
        i64 j = 0;
        if (lane == 0) j = (i64)atomicAdd(queue, 1ULL);
        j = ((i64)__shfl_sync(0xffffffffu, (int)(j >> 32), 0) << 32) | (unsigned)__shfl_sync(0xffffffffu, (int)j, 0);
        if (j >= n_halo) break;
        const HaloSph s = load_halo(halos + j * BFG_HALO_STRIDE);
        if (s.skip != 0.0) break;                          // bfg_halo_sort_owned: the halos of other ranks come last
        const DiscRings d = disc_rings(h, s.theta, s.phi, s.radius);
        if (d.rb - d.ra + 1 > (i64)max_rings) continue;                          // a large disc: k_shell_halos takes it
        if (sharded && !disc_touches_range(h, d, pix_lo, pix_hi)) continue;
        HaloUpd u = make_upd(T, s);
        const FastHalo fh = make_fast<PAINT>(T, s, u, row, l2tab, nullptr);
        // blend this warp's row (a / D folded in for the displacement table, like the CTA kernel's fast loops)
        const RowBlender B(T, s.lnz, s.lnM, extras ? extras + j * n_extra : nullptr);
        const bool valid = B.valid;
        const double post = (!PAINT && SHELL_PRESCALED) ? s.a / s.D : 1.0;
        __syncwarp();                                      // the previous halo's row is no longer read
        for (int k = lane; k < B.NR; k += 32) row[k] = B.node(T, k) * post;
        if (!PAINT && SHELL_PRESCALED) u.a = u.D;          // generic update on a prescaled row (the < 4-pixel fallback)
        __syncwarp();
        // `if pixind.size < 4` (HealpixRunner.py:333): count the disc when it could be that small
        const bool tiny = !PAINT && (s.radius * s.radius * (double)h.npix * 0.25 < 64.0);
        if (tiny) {
            i64 c = 0;
            for (i64 iz = d.ra + lane; iz <= d.rb; iz += 32) {
                i64 start, nr, ip_lo, cnt; bool sh;
                disc_ring_span(h, d, iz, start, nr, sh, ip_lo, cnt);
                c += cnt;
            }
            c = warp_sum_i64(c);
            if (c < 4) {
                if (lane < 4) {
                    i64 pix[4]; double w[4];
                    get_interpol(h, s.theta_ll, s.phi_ll, pix, w);   // HealpixRunner.py:334
                    const i64 p = pix[lane];
                    if (p >= pix_lo && p < pix_hi) {
                        ++done;
                        if (valid) {
                            double x, y, z;
                            pix2vec(h, p, x, y, z);
                            double *q = out + (p - pix_lo);
                            shell_update<MODE, UNIFORM>(T, row, u, x, y, z, x * u.D, y * u.D, z * u.D, q, q + nloc,
                                                        q + 2 * nloc, l2tab, A, row, u);
                        }
                    }
                }
                continue;
            }
        }
        for (i64 base = d.ra; base <= d.rb; base += 32) {
            {   // stage up to 32 ring segments: one ring per lane
                const i64 iz = base + lane;
                RingSeg g;
                g.cnt = 0; g.active = 0;
                if (iz <= d.rb) {
                    i64 start, nr, ip_lo, cnt; bool sh;
                    disc_ring_span(h, d, iz, start, nr, sh, ip_lo, cnt);
                    if (cnt > 0 && start < pix_hi && start + nr > pix_lo) {
                        g.active = 1;
                        g.lbase = start - pix_lo;
                        g.nr = (int)nr; g.ip_lo = (int)ip_lo; g.cnt = (int)cnt;
                        g.flags = ((start < pix_lo || start + nr > pix_hi) ? 1 : 0) | ((nr == h.nl4) ? 2 : 0) | (sh ? 4 : 0);
                        ring_z_sth(h, iz, g.z, g.sth);
                        g.pz = g.z * u.D; g.sD = g.sth * u.D;
                        g.dz = g.z - s.vz; g.dz2 = g.dz * g.dz;
                        g.inv2nr = 2.0 / (double)nr;
                        g.phase0 = ((double)ip_lo + (sh ? 0.5 : 0.0)) * g.inv2nr;
                        if (g.flags & 2) sincospi(g.phase0, &g.s0, &g.c0);       // only equatorial rings rotate a staged start
                        sincospi((double)GW_SMALL * g.inv2nr, &g.rotS, &g.rotC);
                    }
                }
                segs[lane] = g;
            }
            __syncwarp();
            const int nseg = (int)min((i64)32, d.rb - base + 1);
            done += walk_rings_fast<GW_SMALL, PAINT>(fh, segs, nseg, valid, sharded, eqC, eqS, out, nloc, nloc8, 0, 1);
            __syncwarp();
        }
    }
    if (nupd) {
        done = warp_sum_i64(done);
        if (lane == 0 && done) atomicAdd(nupd, (unsigned long long)done);
    }
}

// Re-binning: one thread per source pixel.
__global__ void __launch_bounds__(256)
k_shell_regrid(Hpx h, const RingTabEntry *__restrict__ rt, const double *__restrict__ map_in, const double *__restrict__ off,
               double *__restrict__ map_out, i64 pix_lo, i64 pix_hi) {
    const i64 nloc = pix_hi - pix_lo;
    for (i64 lp = (i64)blockIdx.x * blockDim.x + threadIdx.x; lp < nloc; lp += (i64)gridDim.x * blockDim.x) {
        double m = map_in[lp];
        if (m == 0.0) continue;                                  // HealpixRunner.py:359
        i64 pix[4]; double w[4];
        // :357-361  displaced direction -> 4 neighbours + bilinear weights (regrid_target: fast small-angle form or literal chain)
        regrid_target(h, rt, pix_lo + lp, off[lp], off[nloc + lp], off[2 * nloc + lp], pix, w);
#pragma unroll
        for (int k = 0; k < 4; ++k) red_add(map_out + pix[k], w[k] * m);   // :17-71
    }
}

// Re-binning of a SOURCE pixel range of a full-size offsets array (component stride given explicitly): the pipelined
// end-to-end path re-bins the rings whose offsets are final while the halo loop works further south.
__global__ void __launch_bounds__(256)
k_shell_regrid_range(Hpx h, const RingTabEntry *__restrict__ rt, const double *__restrict__ map_in, const double *__restrict__ off,
                     i64 comp_stride, double *__restrict__ map_out, i64 src_lo, i64 src_hi) {
    for (i64 p = src_lo + (i64)blockIdx.x * blockDim.x + threadIdx.x; p < src_hi; p += (i64)gridDim.x * blockDim.x) {
        const double m = map_in[p];
        if (m == 0.0) continue;                                  // HealpixRunner.py:359
        i64 pix[4]; double w[4];
        regrid_target(h, rt, p, off[p], off[comp_stride + p], off[2 * comp_stride + p], pix, w);   // :357-361
#pragma unroll
        for (int k = 0; k < 4; ++k) red_add(map_out + pix[k], w[k] * m);   // :17-71
    }
}

// max_p |offset_p|^2 over [lo, hi) of a component-major offsets array (bounds how far the re-binning can move mass)
__global__ void __launch_bounds__(256)
k_max_norm3(const double *__restrict__ off, i64 comp_stride, i64 lo, i64 hi, unsigned long long *out_bits) {
    double m = 0.0;
    for (i64 p = lo + (i64)blockIdx.x * blockDim.x + threadIdx.x; p < hi; p += (i64)gridDim.x * blockDim.x) {
        const double a = off[p], b = off[comp_stride + p], c = off[2 * comp_stride + p];
        const double n2 = a * a + b * b + c * c;
        m = fmax(m, (n2 == n2) ? n2 : CUDART_INF);               // a NaN offset counts as unbounded
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, (unsigned long long)__double_as_longlong(m));   // m >= 0: bit order = value order
}

// Sky-sorted halo records: first index whose colatitude band (floor(theta / band)) reaches each band edge, and the
// largest disc radius -- what the host needs to cut the halo loop into latitude chunks.
__global__ void k_band_bounds(i64 n, const double *__restrict__ halos, double band, int n_edges,
                              const i64 *__restrict__ edge_band, i64 *__restrict__ bounds, unsigned long long *rho_bits) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_edges) {
        const i64 eb = edge_band[t];
        i64 lo = 0, hi = n;                                       // first i with band(i) >= eb
        while (lo < hi) {
            const i64 mid = (lo + hi) >> 1;
            // halos marked by bfg_halo_sort_owned (other ranks') sort behind every band: they count as band 2^20, so an edge
            // of 2^20 returns the number of owned halos
            const i64 bm = (halos[mid * BFG_HALO_STRIDE + BFG_HS_SKIP] != 0.0) ? 1048576
                           : (i64)fmin(fmax(halos[mid * BFG_HALO_STRIDE + BFG_HS_THETA] / band, 0.0), 1048575.0);
            if (bm >= eb) hi = mid; else lo = mid + 1;
        }
        bounds[t] = lo;
    }
    double m = 0.0;
    for (i64 i = t; i < n; i += (i64)gridDim.x * blockDim.x)
        if (halos[i * BFG_HALO_STRIDE + BFG_HS_SKIP] == 0.0) m = fmax(m, halos[i * BFG_HALO_STRIDE + BFG_HS_RADIUS]);
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(rho_bits, (unsigned long long)__double_as_longlong(m));
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

__global__ void k_reorder_index(Hpx h, i64 n, const i64 *__restrict__ in, i64 *__restrict__ out, int to_nest) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const i64 p = in[i];
        out[i] = (p < 0 || p >= h.npix) ? -1 : (to_nest ? ring2nest(h, p) : nest2ring(h, p));
    }
}

int check_nside(int nside) {
    if (nside < 1 || nside > (1 << 24)) { set_error("nside out of range"); return BFG_ERR_INVALID; }
    return BFG_OK;
}

int grid_for(i64 n, int threads, int max_blocks = 148 * 32) {
    return (int)std::max<i64>(1, std::min<i64>((n + threads - 1) / threads, max_blocks));
}

template <int MODE>
int launch_shell(const bfg_table *t, int nside, i64 n_halo, const double *d_halos, const double *d_extras, int n_extra,
                 double *d_out, i64 pix_lo, i64 pix_hi, i64 *d_nupdates, cudaStream_t st,
                 const bfg_table *t2 = nullptr, const double *d_mtot = nullptr, const double *d_orig = nullptr,
                 double mtot_add = 0.0) {
    constexpr bool PAINT = (MODE != MODE_BARYONIFY);
    BFG_REQUIRE(t && (d_halos || n_halo == 0) && (d_out || pix_lo == pix_hi), "null argument");
    AnisArgs A;
    memset(&A, 0, sizeof(A));
    if (MODE == MODE_ANIS) {
        BFG_REQUIRE(t2 && (pix_lo == pix_hi || (d_mtot && d_orig)), "anisotropic paint needs the tracer table and both maps");
        BFG_REQUIRE(t2->view.ndim == t->view.ndim && (t2->view.flags & BFG_TABLE_LOG_VALUES),
                    "tracer table must be a log-profile table with the paint table's extra axes");
        BFG_REQUIRE(t2->device == t->device, "tables live on different devices");
        A.T2 = t2->view; A.mtot = d_mtot; A.orig = d_orig; A.mtot_add = mtot_add;
    }
    if (int rc = check_nside(nside)) return rc;
    Hpx h(nside);
    BFG_REQUIRE(pix_lo >= 0 && pix_hi <= h.npix && pix_lo <= pix_hi, "bad pixel range");
    BFG_REQUIRE(n_extra == t->view.ndim - 3, "n_extra must equal the table's extra axes");
    BFG_REQUIRE(n_extra == 0 || d_extras, "extras missing");
    BFG_REQUIRE(PAINT == ((t->view.flags & BFG_TABLE_LOG_VALUES) != 0),
                "paint needs a log-profile table, baryonify a displacement table");
    if (d_nupdates) BFG_CUDA_OK(cudaMemsetAsync(d_nupdates, 0, sizeof(i64), st));
    if (n_halo == 0 || pix_lo == pix_hi) return BFG_OK;
    size_t smem = sizeof(double) * (t->view.n[2] + (MODE == MODE_ANIS ? t2->view.n[2] : 0));
    // baryonify: the row once more as (value, step) pairs behind the plain row (fast loops, blend_row_pairs)
    if (MODE == MODE_BARYONIFY && t->view.n[2] <= SHELL_MAX_PAIR_NODES)
        smem = sizeof(double) * (((t->view.n[2] + 1) & ~1) + 2 * (size_t)t->view.n[2]);
    BFG_REQUIRE(smem <= 200 * 1024, "radial axis too long for the shared-memory row (max 25600 nodes)");
    int sms = 148;
    BFG_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, t->device));
    int blocks = (int)std::min<i64>(n_halo, (i64)sms * SHELL_MIN_CTAS);   // persistent: every CTA resident
    const double2 *g_l2tab = nullptr;
    if (int rc = get_log2_table(&g_l2tab)) return rc;
    if (int rc = retain_async_pool()) return rc;
    unsigned long long *queue = nullptr;   // halo queue head (stream-ordered scratch)
    BFG_CUDA_OK(cudaMallocAsync(&queue, sizeof(unsigned long long), st));
    BFG_CUDA_OK(cudaMemsetAsync(queue, 0, sizeof(unsigned long long), st));
    const bool uni = t->view.uniform_r && (MODE != MODE_ANIS || t2->view.uniform_r);
    // small discs first, one warp per halo (k_shell_halos_warp); the CTA-per-halo kernel then skips them.  PaintProfilesShell:
    // always (NSIDE=1024: 1.72 -> 1.59 ms).  BaryonifyShell: when the batch votes for it (k_warp_vote_*), i.e. for catalogues of
    // small discs.  BFG_SHELL_WARP_KERNEL=1 / 0 forces it on / off, BFG_SHELL_WARP_MAX_RINGS sets the ring threshold.
    const char *renv = getenv("BFG_SHELL_WARP_MAX_RINGS");
    const int warp_max_rings = renv ? std::max(1, std::min(4096, atoi(renv))) : WARP_MAX_RINGS;
    const char *wenv = getenv("BFG_SHELL_WARP_KERNEL");
    const int force = wenv ? (wenv[0] != '0' ? 1 : -1) : (MODE == MODE_PAINT ? 1 : 0);
    const size_t wsmem = sizeof(double) * (SHELL_THREADS / 32) * (size_t)((t->view.n[2] + 1) & ~1);
    StreamScratch s_vote(st), s_queue_w(st);
    const int *d_vote = nullptr;
    if (uni && MODE != MODE_ANIS && force >= 0 && wsmem <= 64 * 1024) {
        BFG_CUDA_OK(s_vote.alloc(32));
        BFG_CUDA_OK(s_queue_w.alloc(sizeof(unsigned long long)));
        BFG_CUDA_OK(cudaMemsetAsync(s_vote.p, 0, 32, st));
        BFG_CUDA_OK(cudaMemsetAsync(s_queue_w.p, 0, sizeof(unsigned long long), st));
        unsigned long long *cnt = s_vote.as<unsigned long long>();
        int *vote = (int *)(cnt + 2);
        if (force == 0)
            k_warp_vote_count<<<(int)std::min<i64>((n_halo + 255) / 256, (i64)sms * 8), 256, 0, st>>>(h, n_halo, d_halos,
                                                                                                   warp_max_rings, cnt);
        k_warp_vote_finish<<<1, 1, 0, st>>>(cnt, warp_max_rings, force, vote);
        auto kw = k_shell_halos_warp<MODE == MODE_ANIS ? MODE_PAINT : MODE>;
        BFG_CUDA_OK(cudaFuncSetAttribute(kw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
        const int wblocks = (int)std::min<i64>((n_halo + SHELL_THREADS / 32 - 1) / (SHELL_THREADS / 32), (i64)sms * WARP_KERNEL_CTAS);
        kw<<<wblocks, SHELL_THREADS, wsmem, st>>>(t->view, h, n_halo, d_halos, d_extras, n_extra, d_out, pix_lo, pix_hi,
                                                  (unsigned long long *)d_nupdates, g_l2tab, s_queue_w.as<unsigned long long>(),
                                                  (int)((t->view.n[2] + 1) & ~1), vote);
        BFG_CUDA_OK(cudaGetLastError());
        d_vote = vote;
    }
    auto go = [&](auto kern) -> int {
        BFG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, SHELL_THREADS, smem, st>>>(t->view, h, n_halo, d_halos, d_extras, n_extra, d_out, pix_lo, pix_hi,
                                                  (unsigned long long *)d_nupdates, g_l2tab, A, queue, d_vote);
        BFG_CUDA_OK(cudaGetLastError());
        BFG_CUDA_OK(cudaFreeAsync(queue, st));
        return BFG_OK;
    };
    return uni ? go(k_shell_halos<MODE, true>) : go(k_shell_halos<MODE, false>);
}

}  // namespace

extern "C" int bfg_shell_offsets(const bfg_table *t, int nside, int64_t n_halo, const double *d_halos,
                                 const double *d_extras, int n_extra, double *d_offsets, int64_t pix_lo, int64_t pix_hi,
                                 int64_t *d_nupdates, void *stream) {
    BFG_ENTRY();
    return launch_shell<MODE_BARYONIFY>(t, nside, n_halo, d_halos, d_extras, n_extra, d_offsets, pix_lo, pix_hi,
                               (i64 *)d_nupdates, (cudaStream_t)stream);
}

extern "C" int bfg_shell_paint(const bfg_table *t, int nside, int64_t n_halo, const double *d_halos,
                               const double *d_extras, int n_extra, double *d_map, int64_t pix_lo, int64_t pix_hi,
                               int64_t *d_nupdates, void *stream) {
    BFG_ENTRY();
    return launch_shell<MODE_PAINT>(t, nside, n_halo, d_halos, d_extras, n_extra, d_map, pix_lo, pix_hi, (i64 *)d_nupdates,
                              (cudaStream_t)stream);
}

extern "C" int bfg_shell_paint_anis(const bfg_table *t_paint, const bfg_table *t_tracer, int nside, int64_t n_halo,
                                    const double *d_halos, const double *d_extras, int n_extra, const double *d_mtot,
                                    double mtot_add, const double *d_orig, double *d_map, int64_t pix_lo,
                                    int64_t pix_hi, int64_t *d_nupdates, void *stream) {
    BFG_ENTRY();
    return launch_shell<MODE_ANIS>(t_paint, nside, n_halo, d_halos, d_extras, n_extra, d_map, pix_lo, pix_hi,
                                   (i64 *)d_nupdates, (cudaStream_t)stream, t_tracer, d_mtot, d_orig, mtot_add);
}

namespace {
// new_map = (new_map + coef * (bg / Mtot where Mtot > 0 else 0) * orig) * final_scale
//   HealpixRunner.py:633-636 (final_scale = 1) ; Map2DRunner.py:1004-1015 (final_scale = res^2 when include_pixel_size)
__global__ void __launch_bounds__(256)
k_anis_background(i64 n, const double *__restrict__ mtot, double mtot_add, const double *__restrict__ orig, double coef,
                  double final_scale, double *__restrict__ map) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const double m = mtot[i] + mtot_add;
        double f = (m > 0.0) ? mtot_add / m : 0.0;
        f *= orig[i];
        map[i] = (map[i] + coef * f) * final_scale;
    }
}
}  // namespace

extern "C" int bfg_anis_background(int64_t n, const double *d_mtot, double mtot_add, const double *d_orig, double coef,
                                   double final_scale, double *d_map, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(n >= 0 && (n == 0 || (d_mtot && d_orig && d_map)), "null argument");
    if (n == 0) return BFG_OK;
    k_anis_background<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, d_mtot, mtot_add, d_orig, coef, final_scale,
                                                                         d_map);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_shell_regrid(int nside, const double *d_map_in, const double *d_offsets, double *d_map_out,
                                int64_t pix_lo, int64_t pix_hi, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(d_map_in && d_offsets && d_map_out, "null argument");
    if (int rc = check_nside(nside)) return rc;
    Hpx h(nside);
    BFG_REQUIRE(pix_lo >= 0 && pix_hi <= h.npix && pix_lo <= pix_hi, "bad pixel range");
    if (pix_lo == pix_hi) return BFG_OK;
    const RingTabEntry *rt = nullptr;
    if (int rc = get_ring_table(nside, &rt, stream)) return rc;
    k_shell_regrid<<<grid_for(pix_hi - pix_lo, 256), 256, 0, (cudaStream_t)stream>>>(h, rt, d_map_in, d_offsets, d_map_out,
                                                                                    pix_lo, pix_hi);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_shell_regrid_range(int nside, const double *d_map_in, const double *d_offsets, int64_t comp_stride,
                                      double *d_map_out, int64_t src_lo, int64_t src_hi, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(d_map_in && d_offsets && d_map_out, "null argument");
    if (int rc = check_nside(nside)) return rc;
    Hpx h(nside);
    BFG_REQUIRE(src_lo >= 0 && src_hi <= h.npix && src_lo <= src_hi && comp_stride >= src_hi, "bad pixel range");
    if (src_lo == src_hi) return BFG_OK;
    const RingTabEntry *rt = nullptr;
    if (int rc = get_ring_table(nside, &rt, stream)) return rc;
    k_shell_regrid_range<<<grid_for(src_hi - src_lo, 256), 256, 0, (cudaStream_t)stream>>>(h, rt, d_map_in, d_offsets, comp_stride,
                                                                                          d_map_out, src_lo, src_hi);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_offsets_max_norm2(const double *d_offsets, int64_t comp_stride, int64_t lo, int64_t hi, double *d_out,
                                     void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(d_offsets && d_out && lo >= 0 && lo <= hi && hi <= comp_stride, "bad argument");
    BFG_CUDA_OK(cudaMemsetAsync(d_out, 0, sizeof(double), (cudaStream_t)stream));
    if (lo == hi) return BFG_OK;
    k_max_norm3<<<grid_for(hi - lo, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(d_offsets, comp_stride, lo, hi,
                                                                                  (unsigned long long *)d_out);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_halo_band_bounds(int64_t n_halo, const double *d_sorted_halos, double band, int n_edges,
                                    const int64_t *d_edge_band, int64_t *d_bounds, double *d_rho_max, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(d_sorted_halos && d_edge_band && d_bounds && d_rho_max && band > 0 && n_edges >= 1 && n_edges <= 4096,
                "bad argument");
    BFG_CUDA_OK(cudaMemsetAsync(d_rho_max, 0, sizeof(double), (cudaStream_t)stream));
    k_band_bounds<<<32, 256, 0, (cudaStream_t)stream>>>(n_halo, d_sorted_halos, band, n_edges, (const i64 *)d_edge_band,
                                                        (i64 *)d_bounds, (unsigned long long *)d_rho_max);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_healpix_disc_counts(int nside, int64_t n_halo, const double *d_halos, int64_t *d_npix, void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(d_halos && d_npix, "null argument");
    if (int rc = check_nside(nside)) return rc;
    if (n_halo == 0) return BFG_OK;
    k_disc_counts<<<grid_for(n_halo * 32, 256), 256, 0, (cudaStream_t)stream>>>(Hpx(nside), n_halo, d_halos, (i64 *)d_npix);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_healpix_query_disc(int nside, const double *d_halo, int64_t *d_pix, int64_t cap, int64_t *d_count,
                                      void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(d_halo && d_count && (d_pix || cap == 0), "null argument");
    if (int rc = check_nside(nside)) return rc;
    k_query_disc<<<1, 32, 0, (cudaStream_t)stream>>>(Hpx(nside), d_halo, (i64 *)d_pix, cap, (i64 *)d_count);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_healpix_pix2vec(int nside, int64_t pix_lo, int64_t pix_hi, double *d_xyz, void *stream) {
    BFG_ENTRY();
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
    BFG_ENTRY();
    BFG_REQUIRE(d_theta && d_phi && d_pix && d_w, "null argument");
    if (int rc = check_nside(nside)) return rc;
    if (n == 0) return BFG_OK;
    k_interp_weights<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(Hpx(nside), n, d_theta, d_phi, (i64 *)d_pix, d_w);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_healpix_ang2pix(int nside, int64_t n, const double *d_theta, const double *d_phi, int64_t *d_pix,
                                   void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(d_theta && d_phi && d_pix, "null argument");
    if (int rc = check_nside(nside)) return rc;
    if (n == 0) return BFG_OK;
    k_ang2pix<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(Hpx(nside), n, d_theta, d_phi, (i64 *)d_pix);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_healpix_reorder(int nside, int to_nest, int64_t n, const int64_t *d_pix_in, int64_t *d_pix_out,
                                   void *stream) {
    BFG_ENTRY();
    BFG_REQUIRE(d_pix_in && d_pix_out, "null argument");
    if (int rc = check_nside(nside)) return rc;
    BFG_REQUIRE((nside & (nside - 1)) == 0, "the NESTED scheme needs nside to be a power of two");
    if (n == 0) return BFG_OK;
    k_reorder_index<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(Hpx(nside), n, (const i64 *)d_pix_in, (i64 *)d_pix_out,
                                                                        to_nest);
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
    BFG_ENTRY();
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
    BFG_ENTRY();
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

// ------------------------------------------------------------------------------------------------ multi-GPU regrid
// Re-binning fused with the exchange step of ring-range sharding: every rank owns a slice of the NEW map in memory the
// other ranks of the box have mapped through CUDA IPC; a displaced pixel is deposited straight into the owner's slice
// (fp64 RED over NVLink peer memory for the few deposits that cross a range border), so no full-size partial map and
// no all-reduce are needed (SURVEY.md §8e option B; BaryonForge/utils/Parallelize.py:318 sums whole maps instead).
namespace {
struct OwnerTable {
    int world;
    i64 bounds[9];        // bounds[r] .. bounds[r+1] = pixels owned by rank r
    double *slice[8];     // slice[r][p - bounds[r]] ; slice[self] is local memory, the others are peer mappings
    int self;
};

__global__ void __launch_bounds__(256)
k_shell_regrid_p2p(Hpx h, const RingTabEntry *__restrict__ rt, const double *__restrict__ map_in, const double *__restrict__ off,
                   OwnerTable own, i64 pix_lo, i64 pix_hi, i64 src_lo, i64 src_hi, unsigned long long *remote_count) {
    const i64 nloc = pix_hi - pix_lo;
    const double inv_span = (double)own.world / (double)h.npix;
    unsigned long long nrem = 0;
    // source pixels [src_lo, src_hi) of the owned range (the pipelined end-to-end path re-bins the rings whose offsets are final)
    for (i64 lp = src_lo - pix_lo + (i64)blockIdx.x * blockDim.x + threadIdx.x; lp < src_hi - pix_lo;
         lp += (i64)gridDim.x * blockDim.x) {
        double m = map_in[lp];
        if (m == 0.0) continue;                                  // HealpixRunner.py:359
        i64 pix[4]; double w[4];
        regrid_target(h, rt, pix_lo + lp, off[lp], off[nloc + lp], off[2 * nloc + lp], pix, w);   // :357-361
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const i64 p = pix[k];
            int r = min(own.world - 1, (int)((double)p * inv_span));
            while (p < own.bounds[r]) --r;
            while (p >= own.bounds[r + 1]) ++r;
            double *dst = own.slice[r] + (p - own.bounds[r]);
            if (r == own.self) {
                red_add(dst, w[k] * m);                          // :17-71
            } else {
                atomicAdd_system(dst, w[k] * m);                 // owner's HBM through NVLink
                ++nrem;
            }
        }
    }
    if (remote_count && nrem) atomicAdd(remote_count, nrem);
}
}  // namespace

static int regrid_p2p_impl(int nside, const double *d_map_in, const double *d_offsets, int64_t pix_lo, int64_t pix_hi,
                           int64_t src_lo, int64_t src_hi, int world, int self, const int64_t *h_bounds,
                           double *const *h_slices, int64_t *d_remote_count, bool zero_count, void *stream) {
    BFG_REQUIRE(d_map_in && d_offsets && h_bounds && h_slices, "null argument");
    BFG_REQUIRE(world >= 1 && world <= 8 && self >= 0 && self < world, "world must be 1..8");
    if (int rc = check_nside(nside)) return rc;
    Hpx h(nside);
    BFG_REQUIRE(pix_lo >= 0 && pix_hi <= h.npix && pix_lo <= pix_hi, "bad pixel range");
    BFG_REQUIRE(src_lo >= pix_lo && src_hi <= pix_hi && src_lo <= src_hi, "source range outside the owned range");
    BFG_REQUIRE(h_bounds[0] == 0 && h_bounds[world] == h.npix && h_bounds[self] == pix_lo && h_bounds[self + 1] == pix_hi,
                "bounds must tile the map and agree with the owned range");
    OwnerTable own;
    own.world = world; own.self = self;
    for (int r = 0; r <= world; ++r) own.bounds[r] = h_bounds[r];
    for (int r = world + 1; r < 9; ++r) own.bounds[r] = h.npix;
    for (int r = 0; r < 8; ++r) own.slice[r] = (r < world) ? h_slices[r] : nullptr;
    if (d_remote_count && zero_count) BFG_CUDA_OK(cudaMemsetAsync(d_remote_count, 0, sizeof(i64), (cudaStream_t)stream));
    if (src_lo == src_hi) return BFG_OK;
    const RingTabEntry *rt = nullptr;
    if (int rc = get_ring_table(nside, &rt, stream)) return rc;
    k_shell_regrid_p2p<<<grid_for(src_hi - src_lo, 256), 256, 0, (cudaStream_t)stream>>>(
        h, rt, d_map_in, d_offsets, own, pix_lo, pix_hi, src_lo, src_hi, (unsigned long long *)d_remote_count);
    BFG_CUDA_OK(cudaGetLastError());
    return BFG_OK;
}

extern "C" int bfg_shell_regrid_p2p(int nside, const double *d_map_in, const double *d_offsets, int64_t pix_lo,
                                    int64_t pix_hi, int world, int self, const int64_t *h_bounds,
                                    double *const *h_slices, int64_t *d_remote_count, void *stream) {
    BFG_ENTRY();
    return regrid_p2p_impl(nside, d_map_in, d_offsets, pix_lo, pix_hi, pix_lo, pix_hi, world, self, h_bounds, h_slices,
                           d_remote_count, true, stream);
}

extern "C" int bfg_shell_regrid_p2p_range(int nside, const double *d_map_in, const double *d_offsets, int64_t pix_lo,
                                          int64_t pix_hi, int64_t src_lo, int64_t src_hi, int world, int self,
                                          const int64_t *h_bounds, double *const *h_slices, int64_t *d_remote_count,
                                          void *stream) {
    BFG_ENTRY();
    return regrid_p2p_impl(nside, d_map_in, d_offsets, pix_lo, pix_hi, src_lo, src_hi, world, self, h_bounds, h_slices,
                           d_remote_count, false, stream);
}



// ---------------------------------------------------------------------------------------------------- host test entry
// Pure host, no GPU: ONE halo's updates of a list of pixels with the generic per-pixel update of the shell kernels
// (shell_update<MODE, UNIFORM> above -- the literal arithmetic of HealpixRunner.py:336-355 / :464-481 the fast loops reorganise),
// its table read-out and its halo constants, all from the kernels' own source.  h_record = the 16-double halo record
// (bfg_shell_records), h_vec [n][3] = pixel unit vectors; mode 0: h_out [n][3] += nw_vec - vec, mode 1: h_out [n] += profile.
extern "C" int bfg_test_shell_update_host(int ndim, const int64_t *shape, const double *const *h_axes, const double *h_values,
                                          int flags, int force_search, int mode, const double *h_record, const double *h_extras,
                                          int64_t n, const double *h_vec, double *h_out) {
    BFG_REQUIRE(shape && h_axes && h_values && h_record && (n == 0 || (h_vec && h_out)), "null argument");
    BFG_REQUIRE(ndim >= 3 && ndim <= BFG_MAX_TABLE_DIM && (ndim == 3 || h_extras), "bad table / extras");
    BFG_REQUIRE(mode == 0 || mode == 1, "mode: 0 = baryonify, 1 = paint");
    TableView T;
    host_table_view(ndim, shape, h_axes, h_values, flags, T);
    if (force_search) T.uniform_r = 0;
    const HaloSph s = load_halo(h_record);
    const HaloUpd u = make_upd(T, s);
    const RowBlender B(T, s.lnz, s.lnM, h_extras);
    std::vector<double> row((size_t)B.NR);
    for (int k = 0; k < B.NR; ++k) row[(size_t)k] = B.node(T, k);
    double2 l2tab[BFG_LOG2_TAB];
    fill_log2_table(l2tab);
    const AnisArgs A = {};
    if (!B.valid) return BFG_OK;            // outside the table in (z, M, extras): every read-out is NaN -> adds nothing
    for (int64_t i = 0; i < n; ++i) {
        const double x = h_vec[3 * i], y = h_vec[3 * i + 1], z = h_vec[3 * i + 2];
        double *o = mode == 0 ? h_out + 3 * i : h_out + i;
        if (mode == 0) {
            if (T.uniform_r) shell_update<MODE_BARYONIFY, true>(T, row.data(), u, x, y, z, x * u.D, y * u.D, z * u.D, o, o + 1, o + 2, l2tab, A, row.data(), u);
            else shell_update<MODE_BARYONIFY, false>(T, row.data(), u, x, y, z, x * u.D, y * u.D, z * u.D, o, o + 1, o + 2, l2tab, A, row.data(), u);
        } else {
            if (T.uniform_r) shell_update<MODE_PAINT, true>(T, row.data(), u, x, y, z, x * u.D, y * u.D, z * u.D, o, o, o, l2tab, A, row.data(), u);
            else shell_update<MODE_PAINT, false>(T, row.data(), u, x, y, z, x * u.D, y * u.D, z * u.D, o, o, o, l2tab, A, row.data(), u);
        }
    }
    return BFG_OK;
}
