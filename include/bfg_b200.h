/*
 * bfg_b200.h -- C ABI of libbfg_b200.so: the B200 (sm_100a) back-end of BaryonForge's runner hot path.
 *
 * The reference (DhayaaAnbajagane/BaryonForge) is pure Python and has no FFI; the boundary it offers for
 * this path is the `Runners` classes' constructor + process().  Each entry point below replaces the body
 * of one stage of those process() methods (cited per function as file:line under BaryonForge/); the thin
 * Python mirror in baryonforge_b200/runners.py binds them with ctypes (see INTEGRATION.md for the stub a
 * reference maintainer would add).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no C++/torch types.  Every call returns 0 on success or a
 *    negative bfg_status; bfg_last_error() gives the thread-local message.  Nothing throws.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are asynchronous
 *    with respect to the host unless the name ends in `_host`, and never allocate unless stated.
 *  - Pointers named d_* are DEVICE pointers on the table's device; h_* are HOST pointers.
 *  - Maps are HEALPix RING order float64 (LightconeShell.map, utils/io.py:290-379) or C-order float64
 *    N^d grids (GriddedMap.map, utils/io.py:382-494).  Pixel / cell / particle ids are int64.
 *  - Halo records are rows of BFG_HALO_STRIDE float64 (128 B, one cache line) -- the per-halo scalars the
 *    reference computes at the top of each loop iteration (Runners/HealpixRunner.py:317-329,
 *    Runners/Map2DRunner.py:484-503, Runners/SnapshotRunner.py:219-228), built once per process() on the device
 *    (bfg_shell_records / bfg_box_records) or, for slab-sharded grids, vectorised on the host.
 */
#ifndef BFG_B200_H
#define BFG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BFG_ABI_VERSION 1

typedef enum {
    BFG_OK = 0,
    BFG_ERR_INVALID = -1,   /* bad argument */
    BFG_ERR_CUDA = -2,      /* CUDA runtime error, see bfg_last_error() */
    BFG_ERR_UNSUPPORTED = -3,
    BFG_ERR_NOMEM = -4
} bfg_status;

/* table flags */
#define BFG_TABLE_LOG_VALUES 1      /* values are log(profile): read-out applies exp() (utils/Tabulate.py:270-271,318-319) */
#define BFG_TABLE_RDELTA 2          /* radial axis is ln(r/R_delta) (Profiles/BaryonCorrection.py:407-408) */

#define BFG_MAX_TABLE_DIM 6
#define BFG_HALO_STRIDE 16

/* ---- spherical-shell halo record (float64 fields) ----------------------------------------------- */
enum {
    BFG_HS_VX = 0, BFG_HS_VY = 1, BFG_HS_VZ = 2, /* hp.ang2vec(ra, dec, lonlat=True)           HealpixRunner.py:327 */
    BFG_HS_THETA = 3, BFG_HS_PHI = 4,            /* pointing(vec): what query_disc works from   HealpixRunner.py:330 */
    BFG_HS_D = 5,                                /* D_A(z_j), physical Mpc                      HealpixRunner.py:321 */
    BFG_HS_A = 6,                                /* a_j = 1/(1+z_j)                             HealpixRunner.py:319 */
    BFG_HS_RADIUS = 7,                           /* R_j*epsilon_max/D_j [rad]                   HealpixRunner.py:329 */
    BFG_HS_LNZ = 8,                              /* np.log(1/a_j)                               BaryonCorrection.py:371 */
    BFG_HS_LNM = 9,                              /* np.log(M_j)                                 BaryonCorrection.py:398 */
    BFG_HS_RCUT = 10,                            /* model.epsilon_max * R_com (+inf for paint)  BaryonCorrection.py:399,410 */
    BFG_HS_LNRCOM = 11,                          /* ln R_com, used when BFG_TABLE_RDELTA        BaryonCorrection.py:408 */
    BFG_HS_SCALE = 12,                           /* paint: pixarea*D_j^2 or 1                   HealpixRunner.py:478 */
    BFG_HS_THETA_LL = 13, BFG_HS_PHI_LL = 14,    /* pi/2-radians(dec), radians(ra): fallback    HealpixRunner.py:334 */
    BFG_HS_SKIP = 15                             /* 0; set != 0 by bfg_halo_sort_owned on the halos of OTHER ranks, which it
                                                    puts last: the halo-loop kernels stop at the first such record */
};

/* ---- box (grid / snapshot) halo record ----------------------------------------------------------- */
enum {
    BFG_HB_X = 0, BFG_HB_Y = 1, BFG_HB_Z = 2,    /* float32-rounded halo position [Mpc]         utils/io.py:204-205 */
    BFG_HB_RQ = 3,                               /* clipped query radius, comoving              Map2DRunner.py:492-493, SnapshotRunner.py:227-228 */
    BFG_HB_NSIZE = 4,                            /* grid: even cutout size (as float64)         Map2DRunner.py:500-503 */
    BFG_HB_CX = 5, BFG_HB_CY = 6, BFG_HB_CZ = 7, /* grid: argmin|bins - x_j| centre cells       Map2DRunner.py:512-513,548-550 */
    BFG_HB_LNZ = 8, BFG_HB_LNM = 9, BFG_HB_RCUT = 10, BFG_HB_LNRCOM = 11,
    BFG_HB_DX = 12, BFG_HB_DY = 13, BFG_HB_DZ = 14, /* grid: bins[cen] - x_j                    Map2DRunner.py:519-520,557-559 */
    BFG_HB_PAINTCUT = 15                         /* paint: R_com*epsilon_max mask radius        Map2DRunner.py:815 */
};

typedef struct bfg_table bfg_table; /* opaque: a (ln(1+z), ln M, ln r[, extras...]) table resident in HBM */

/* ---- library ------------------------------------------------------------------------------------ */
int bfg_abi_version(void);
const char *bfg_last_error(void);
int bfg_device_count(void);
/* SM count, HBM bytes (total, free) of `device` */
int bfg_device_info(int device, int *sm_count, int64_t *mem_total, int64_t *mem_free);

/* ---- tables: RegularGridInterpolator((ln(1+z), ln M, ln r, *extras), values, bounds_error=False, fill=nan)
 *      Profiles/BaryonCorrection.py:307-323 ; utils/Tabulate.py:261-271,582-590.
 *      Axis 2 is the radial axis; axes 0,1,3.. are per-halo constants.  Copies host -> device once. */
int bfg_table_create(bfg_table **out, int ndim, const int64_t *shape, const double *const *h_axes,
                     const double *h_values, int flags, int device);
int bfg_table_destroy(bfg_table *t);
int bfg_table_info(const bfg_table *t, int *ndim, int64_t *shape, int *flags, int *device, int *uniform_r);

/* Read-out only (unit-test entry): out[i] = table(lnz, lnM, x_i, extras) with scipy's N-linear rule, NaN outside,
 * exp() when BFG_TABLE_LOG_VALUES.  Replaces BaryonificationClass._readout's `table(p_in)` (BaryonCorrection.py:404-408)
 * and TabulatedProfile._readout (Tabulate.py:318-319).  d_x, d_out: device, n values.  */
int bfg_table_readout(const bfg_table *t, double lnz, double lnM, const double *h_extras, int64_t n,
                      const double *d_x, double *d_out, void *stream);

/* ---- HEALPix RING geometry on the device (replaces healpy at HealpixRunner.py:330,334,336,357-361) -- */
/* Per-halo disc size: d_npix[j] = |query_disc(nside, vec_j, radius_j, inclusive=False)| (before the <4 fallback). */
int bfg_healpix_disc_counts(int nside, int64_t n_halo, const double *d_halos, int64_t *d_npix, void *stream);
/* Pixel list of ONE halo record (ascending), for parity tests of the index sets: writes <= cap ids, returns
 * the full count in *d_count. */
int bfg_healpix_query_disc(int nside, const double *d_halo, int64_t *d_pix, int64_t cap, int64_t *d_count, void *stream);
/* pix2vec over [pix_lo, pix_hi): d_xyz is [3][pix_hi-pix_lo]. */
int bfg_healpix_pix2vec(int nside, int64_t pix_lo, int64_t pix_hi, double *d_xyz, void *stream);
/* get_interp_weights(theta, phi): d_pix [4][n], d_w [4][n]. */
int bfg_healpix_interp_weights(int nside, int64_t n, const double *d_theta, const double *d_phi, int64_t *d_pix,
                               double *d_w, void *stream);
/* Pixel indices RING -> NEST (to_nest != 0) or NEST -> RING (nside a power of two); out-of-range ids map to -1.
 * ang2pix in the NESTED scheme = bfg_healpix_ang2pix followed by this.  The runners themselves work on RING maps, like the
 * reference (utils/io.py:302). */
int bfg_healpix_reorder(int nside, int to_nest, int64_t n, const int64_t *d_pix_in, int64_t *d_pix_out, void *stream);
/* ang2pix (RING), used for shard assignment. */
int bfg_healpix_ang2pix(int nside, int64_t n, const double *d_theta, const double *d_phi, int64_t *d_pix, void *stream);

/* ---- shells ------------------------------------------------------------------------------------- */
/* Per-halo scalars of the shell runners on the device (HealpixRunner.py:317-329, :451-462; BaryonCorrection.py:371,
 * 398-399,410): fills n_halo shell records from the raw catalogue columns.
 *   d_cols        [6][n_halo] float64: M, z, ra [deg], dec [deg], np.log(1/a), np.log(M) (the two logs come from the host
 *                 so that halos exactly on a table edge land on the reference's side)
 *   D_A spline    the reference's CubicSpline(z_t, D_A) (HealpixRunner.py:297-299) as scipy PPoly: n_DA breakpoints
 *                 d_DA_x, coefficients d_DA_c [4][n_DA-1]; evaluated in scipy's operation order (bit-identical D_j)
 *   radius splines g(u), u = ln(1+z): R_delta(M, a) = cbrt(M) * g(u) (physical Mpc) for the runner's cosmology/mass_def
 *                 (d_g_run_c) and the model's (d_g_mod_c, NULL for paint), sharing the breakpoints d_g_x [n_g]
 *   paint != 0    PaintProfilesShell records: RCUT = +inf, SCALE = pixarea * D^2 when pixarea > 0 (:478) else 1
 *   d_aux         optional [3][n_halo]: R_j (physical), D_j, R_com of the model (NaN for paint) -- for parity tests. */
int bfg_shell_records(int64_t n_halo, const double *d_cols, int paint, double eps_run, double eps_model, double pixarea,
                      int n_DA, const double *d_DA_x, const double *d_DA_c, int n_g, const double *d_g_x,
                      const double *d_g_run_c, const double *d_g_mod_c, double *d_halos, double *d_aux, void *stream);
/* Halo loop of BaryonifyShell.process (HealpixRunner.py:315-355): accumulates the unit-vector offsets of all
 * halos into d_offsets, laid out [3][pix_hi-pix_lo] (component-major, so REDs of a warp are contiguous), for the
 * owned RING range [pix_lo, pix_hi).  d_extras: [n_halo][n_extra] p_keys values or NULL.  d_offsets must be zeroed
 * by the caller.  If d_nupdates != NULL it receives sum_j |pixind_j| restricted to the range (int64, device). */
int bfg_shell_offsets(const bfg_table *t, int nside, int64_t n_halo, const double *d_halos, const double *d_extras,
                      int n_extra, double *d_offsets, int64_t pix_lo, int64_t pix_hi, int64_t *d_nupdates,
                      void *stream);
/* Halo loop of PaintProfilesShell.process (HealpixRunner.py:449-481): d_map[pix - pix_lo] += profile. */
int bfg_shell_paint(const bfg_table *t, int nside, int64_t n_halo, const double *d_halos, const double *d_extras,
                    int n_extra, double *d_map, int64_t pix_lo, int64_t pix_hi, int64_t *d_nupdates, void *stream);
/* Halo loop of PaintProfilesAnisShell.process (HealpixRunner.py:589-631): per (halo, pixel)
 *   d_map[p] += Paint(r) * SCALE * ( Tracer(r) / (d_mtot[p] + mtot_add)  if that total mass > 0 else 0 ) * d_orig[p]
 * with non-finite Paint / Tracer read-outs contributing nothing.  t_paint = model.projected table, t_tracer =
 * Tracer_model.projected table (both log-profile tables with the same extra axes); d_mtot = the halo part of Mtot_map
 * (a PaintProfilesShell pass of Mtot_model with include_pixel_size, :565-570), mtot_add = dV * drho_m (:582); d_orig =
 * LightconeShell.map.  d_mtot, d_orig, d_map cover the owned range [pix_lo, pix_hi). */
int bfg_shell_paint_anis(const bfg_table *t_paint, const bfg_table *t_tracer, int nside, int64_t n_halo,
                         const double *d_halos, const double *d_extras, int n_extra, const double *d_mtot,
                         double mtot_add, const double *d_orig, double *d_map, int64_t pix_lo, int64_t pix_hi,
                         int64_t *d_nupdates, void *stream);
/* Uniform-background term and final scaling of the anisotropic painters (HealpixRunner.py:633-636,
 * Map2DRunner.py:1004-1015):  d_map[i] = (d_map[i] + coef * (mtot_add / (d_mtot[i] + mtot_add) if > 0 else 0) * d_orig[i])
 * * final_scale, coef = background_val * global_tracer_fraction. */
int bfg_anis_background(int64_t n, const double *d_mtot, double mtot_add, const double *d_orig, double coef,
                        double final_scale, double *d_map, void *stream);
/* Re-binning of BaryonifyShell.process (HealpixRunner.py:357-365 + regrid_pixels_hpix :17-71): for source pixels
 * p in [pix_lo, pix_hi) with d_map_in[p - pix_lo] != 0, deposit onto the 4 interpolation neighbours of the displaced
 * direction.  d_map_out is a FULL map (npix) that must be zeroed by the caller; d_offsets is [3][pix_hi-pix_lo]. */
int bfg_shell_regrid(int nside, const double *d_map_in, const double *d_offsets, double *d_map_out, int64_t pix_lo,
                     int64_t pix_hi, void *stream);

/* bfg_shell_regrid for a SOURCE pixel range [src_lo, src_hi) of FULL-size arrays: d_map_in (npix), d_offsets [3][comp_stride]
 * (comp_stride >= src_hi), d_map_out (npix, zeroed once by the caller).  The pipelined BaryonifyShell.process re-bins and
 * downloads the rings whose offsets are already final while the halo loop is still working on more southern halos. */
int bfg_shell_regrid_range(int nside, const double *d_map_in, const double *d_offsets, int64_t comp_stride,
                           double *d_map_out, int64_t src_lo, int64_t src_hi, void *stream);
/* *d_out = max_p |offset_p|^2 over p in [lo, hi) (NaN counts as +inf): bounds how far the re-binning moves mass. */
int bfg_offsets_max_norm2(const double *d_offsets, int64_t comp_stride, int64_t lo, int64_t hi, double *d_out, void *stream);
/* For sky-sorted records (bfg_halo_sort / bfg_halo_sort_owned, band width `band`): d_bounds[e] = first record whose colatitude
 * band floor(theta / band) >= d_edge_band[e]; *d_rho_max = largest disc radius.  Cuts the halo loop into latitude chunks.
 * Records marked BFG_HS_SKIP (other ranks' halos, sorted last) count as band 2^20, so an edge of 2^20 yields the number of
 * owned halos; they do not enter *d_rho_max. */
int bfg_halo_band_bounds(int64_t n_halo, const double *d_sorted_halos, double band, int n_edges, const int64_t *d_edge_band,
                         int64_t *d_bounds, double *d_rho_max, void *stream);

/* Multi-GPU form of bfg_shell_regrid (one process per GPU, ring-range sharding): the re-binning fused with its exchange
 * step.  h_slices[r] points at rank r's owned slice [h_bounds[r], h_bounds[r+1]) of the NEW map -- local memory for
 * r == self, CUDA-IPC peer mappings otherwise (bfg_shared_alloc / bfg_ipc_export / bfg_ipc_import) -- and every deposit
 * goes straight to its owner (fp64 RED over NVLink for the few that cross a range border).  Replaces "each worker holds a
 * full map, the parent sums them" (utils/Parallelize.py:318).  Slices must be zeroed and all ranks synchronised before,
 * and synchronised again after, the call.  *d_remote_count (optional) receives the number of cross-GPU deposits. */
int bfg_shell_regrid_p2p(int nside, const double *d_map_in, const double *d_offsets, int64_t pix_lo, int64_t pix_hi,
                         int world, int self, const int64_t *h_bounds, double *const *h_slices,
                         int64_t *d_remote_count, void *stream);
/* The same for the source pixels [src_lo, src_hi) of the owned range only (d_remote_count is accumulated, not reset): the
 * pipelined sharded end-to-end path re-bins the rings whose offsets are final while the halo loop works further south
 * (the ring-range counterpart of bfg_shell_regrid_range; HealpixRunner.py:357-365). */
int bfg_shell_regrid_p2p_range(int nside, const double *d_map_in, const double *d_offsets, int64_t pix_lo, int64_t pix_hi,
                               int64_t src_lo, int64_t src_hi, int world, int self, const int64_t *h_bounds,
                               double *const *h_slices, int64_t *d_remote_count, void *stream);

/* ---- peer memory between the processes of one box (CUDA IPC) --------------------------------------------------------- */
int bfg_shared_alloc(void **d_ptr, int64_t bytes, int device);            /* cudaMalloc'd, exportable */
int bfg_shared_free(void *d_ptr);
int bfg_ipc_export(const void *d_ptr, unsigned char *handle64);           /* 64-byte handle to ship to the peers */
int bfg_ipc_import(const unsigned char *handle64, void **d_peer_ptr);     /* maps a peer's allocation, enables P2P */
int bfg_ipc_close(void *d_peer_ptr);

/* Page-lock / release a host range shared by the processes of the box (memfd mapping), and a stream-ordered D2H copy into
 * it: every rank copies its owned slice of a result map into ONE host map (replaces the parent-side `sum(outputs)` of
 * utils/Parallelize.py:318 and the per-rank full-map D2H). */
int bfg_host_register(void *h_ptr, int64_t bytes);
int bfg_host_unregister(void *h_ptr);
int bfg_copy_to_host_async(void *h_dst, const void *d_src, int64_t bytes, void *stream);

/* ---- periodic grids (2-D / 3-D) ------------------------------------------------------------------- */
/* Per-halo scalars of the grid (grid = 1) / snapshot (grid = 0) runners on the device (Map2DRunner.py:484-520, :727-760;
 * SnapshotRunner.py:219-228; BaryonCorrection.py:398-399,410): fills n_halo box records.
 *   d_cols   [5][n_halo] float64: M, x, y, z (the float32-rounded catalogue values, utils/io.py:204-205) and np.log(M) as the
 *            reference evaluates it, in float32
 *   a, lnz   scale factor of the catalogue's single redshift and np.log(1/a)
 *   g_run, g_mod   R_delta(M, a) = cbrt(M) * g for the runner's / the model's cosmology and mass definition (physical Mpc)
 *   rq_clip  np.max(bins)/2 (grids, Map2DRunner.py:493) or L/2 (snapshots, SnapshotRunner.py:228)
 *   d_bins   [N] cell centres (grids): centre cells = argmin|bins - x|, offsets bins[cen] - x, even cutout size clipped to
 *            [2, N/2]
 *   d_aux    optional [2][n_halo]: R_phys, R_com of the model (NaN for paint) -- for parity tests. */
int bfg_box_records(int64_t n_halo, const double *d_cols, int ndim, int grid, int paint, double a, double lnz, double g_run,
                    double g_mod, double eps_run, double eps_mod, double res, double rq_clip, int64_t N, const double *d_bins,
                    double *d_halos, double *d_aux, void *stream);
/* Halo loop of BaryonifyGrid.process (Map2DRunner.py:482-586).  d_offsets is [ndim][N^ndim] in units of cells,
 * restricted to axis-0 planes [plane_lo, plane_hi) (slab); NaNs propagate as in the reference.
 * 3-D grids whose size is a multiple of 16 (and plane_lo of 8) with a uniform ln r axis run the tile-centric gather
 * (csrc/grid_tile_kernels.cu: every cell written once); that path sizes its halo-per-tile list on the host, i.e. the call
 * synchronises `stream` once and uses stream-ordered scratch (n_halo x NR doubles for the blended rows).  BFG_GRID_TILES=0
 * in the environment forces the halo-centric scatter kernels.  Same for bfg_grid_paint.
 * use_ell (2-D only, Map2DRunner.py:281-350,531-536): d_extras rows carry, after the table's p_keys values, the 4 entries
 * (row-major) of the halo's shear matrix build_Rmat(A_ell, q_ell); n_extra counts them. */
int bfg_grid_offsets(const bfg_table *t, int ndim, int64_t N, double res, int64_t n_halo, const double *d_halos,
                     const double *d_extras, int n_extra, int use_ell, double *d_offsets, int64_t plane_lo,
                     int64_t plane_hi, int64_t *d_nupdates, void *stream);
/* Halo loop of PaintProfilesGrid.process (Map2DRunner.py:725-821); `scale` folds in the final *res^d of :825
 * (pass 1.0 when include_pixel_size is False). */
int bfg_grid_paint(const bfg_table *t, int ndim, int64_t N, double res, double scale, int64_t n_halo,
                   const double *d_halos, const double *d_extras, int n_extra, int use_ell, double *d_map,
                   int64_t plane_lo, int64_t plane_hi, int64_t *d_nupdates, void *stream);
/* Halo loop of PaintProfilesAnisGrid.process (Map2DRunner.py:895-1001, 2-D maps only :847): per (halo, cell) with a finite
 * Paint read-out and r < R_com*epsilon_max,
 *   d_map[c] += Paint(r) * ( Tracer(r) / (d_mtot[c] + mtot_add) if that total mass > 0 else 0 ) * d_orig[c].
 * d_mtot = halo part of Mtot_map (a PaintProfilesGrid pass of Mtot_model with include_pixel_size=False, :866-871),
 * mtot_add = dV * drho_m (:888).  The background term and the final *res^2 are bfg_anis_background. */
int bfg_grid_paint_anis(const bfg_table *t_paint, const bfg_table *t_tracer, int64_t N, double res, int64_t n_halo,
                        const double *d_halos, const double *d_extras, int n_extra, int use_ell, const double *d_mtot,
                        double mtot_add, const double *d_orig, double *d_map, int64_t plane_lo, int64_t plane_hi,
                        int64_t *d_nupdates, void *stream);
/* Re-binning of BaryonifyGrid.process (Map2DRunner.py:589-613 + regrid_pixels_2D/3D :13-162): non-finite offsets -> 0,
 * add cell coordinates (xy-meshgrid convention), periodic overlap deposit into the FULL grid d_map_out (zeroed by caller). */
int bfg_grid_regrid(int ndim, int64_t N, const double *d_map_in, const double *d_offsets, double *d_map_out,
                    int64_t plane_lo, int64_t plane_hi, void *stream);

/* ---- particle snapshots --------------------------------------------------------------------------- */
/* Cell list replacing scipy.spatial.KDTree (SnapshotRunner.py:95-100): counting-sorts n_part particles (0 <= x < L)
 * into ncell^ndim cells.  Outputs: d_cell_start [ncell^ndim + 1], d_order [n_part] (sorted slot -> caller's index)
 * and the cell-ordered copies d_xs/d_ys/d_zs.  Allocates stream-ordered scratch internally. */
int bfg_snap_build_cells(int ndim, int64_t n_part, const double *d_x, const double *d_y, const double *d_z, double L,
                         int ncell, int64_t *d_cell_start, int64_t *d_order, double *d_xs, double *d_ys, double *d_zs,
                         void *stream);
/* The same with an element stride on the caller's coordinates: stride 4 reads them out of the reference's own particle
 * container, ONE structured array of 32-byte records (M, x, y, z; utils/io.py:588), uploaded as raw bytes (d_x = records +
 * slot of 'x', ...).  stride 1 == bfg_snap_build_cells. */
int bfg_snap_build_cells_strided(int ndim, int64_t n_part, const double *d_x, const double *d_y, const double *d_z,
                                 int64_t stride, double L, int ncell, int64_t *d_cell_start, int64_t *d_order, double *d_xs,
                                 double *d_ys, double *d_zs, void *stream);
/* Halo loop of BaryonifySnapshot.process (SnapshotRunner.py:217-260) over the cell-ordered particles:
 * d_tot [ndim][n_part] (cell order, zeroed by the caller) += displacement * unit vector. */
int bfg_snap_offsets(const bfg_table *t, int ndim, int64_t n_part, const double *d_xs, const double *d_ys,
                     const double *d_zs, double L, int ncell, const int64_t *d_cell_start, int64_t n_halo,
                     const double *d_halos, const double *d_extras, int n_extra, double *d_tot, int64_t *d_npairs,
                     void *stream);
/* out[order[p]] = wrap_once(xs[p] + tot[p])  (SnapshotRunner.py:263-273), back in the caller's particle order. */
int bfg_snap_apply(int ndim, int64_t n_part, const double *d_xs, const double *d_ys, const double *d_zs,
                   const double *d_tot, const int64_t *d_order, double L, double *d_x_out, double *d_y_out,
                   double *d_z_out, void *stream);
/* bfg_snap_apply for particles held as 32-byte records of 4 doubles (32-byte aligned): rec_out[order[p]] = rec_in[order[p]]
 * with the doubles in slots fx, fy(, fz) replaced by the displaced, wrapped coordinates -- the reference's
 * `new_cat = cat.copy(); new_cat['x'] = ...` (SnapshotRunner.py:263-273).  d_rec_out may be d_rec_in (in place). */
int bfg_snap_apply_records(int ndim, int64_t n_part, const double *d_xs, const double *d_ys, const double *d_zs,
                           const double *d_tot, const int64_t *d_order, double L, const double *d_rec_in,
                           double *d_rec_out, int fx, int fy, int fz, void *stream);
/* NGP mass deposit = ParticleSnapshot.make_map / np.histogramdd (utils/io.py:629-677); d_grid zeroed by caller. */
int bfg_snap_deposit_ngp(int ndim, int64_t n_part, const double *d_x, const double *d_y, const double *d_z,
                         const double *d_mass, double L, int64_t n_grid, double *d_grid, void *stream);

/* bfg_snap_apply + bfg_snap_deposit_ngp in one pass over the CELL-ORDERED particles, for callers that only need the grid
 * (BaryonifySnapshot.process() followed by ParticleSnapshot.make_map, SnapshotRunner.py:263-273 + utils/io.py:629-677): the
 * displaced positions are not scattered back to the caller's order.  d_mass [n_part] is in the CALLER's order (gathered
 * through d_order) or NULL for equal-mass particles (mass_const).  d_grid zeroed by the caller. */
int bfg_snap_apply_deposit(int ndim, int64_t n_part, const double *d_xs, const double *d_ys, const double *d_zs,
                           const double *d_tot, const int64_t *d_order, const double *d_mass, double mass_const, double L,
                           int64_t n_grid, double *d_grid, void *stream);

/* ---- P(k) of a particle set: the measurement that follows BaryonifySnapshot.process() in the reference's workflow ------
 * The reference has no library function for this step; it is the cell code of examples/10_Reproduce_Schneider_deltaPk.ipynb
 * (cells 1, 12, 15), cited as nb10:cell.  Keeping it on the device means displaced particles never leave HBM. */
/* `numba_histogram3d(Part % L_fold, bins = n_grid, min_vals = 0, max_vals = L_fold)` (nb10:1, nb10:15; L_fold = L / factor):
 * d_grid [n_grid^3] float64 (zeroed by the caller) += 1 at cell int((x mod L_fold) / (L_fold / n_grid)) per axis, C order.
 * Non-finite coordinates are skipped and counted in *d_ndropped (optional); a folded coordinate that rounds up to L_fold
 * (undefined behaviour in the notebook's numba loop) goes to the last cell. */
int bfg_snap_deposit_folded(int64_t n_part, const double *d_x, const double *d_y, const double *d_z, double L_fold,
                            int64_t n_grid, double *d_grid, int64_t *d_ndropped, void *stream);
/* bfg_snap_deposit_folded over the CELL-ORDERED particles of bfg_snap_build_cells and their accumulated offsets d_tot
 * [3][n_part] (3-D): position = wrap_once(xs + tot) as bfg_snap_apply computes it (SnapshotRunner.py:263-273), then the
 * folded cell -- the displaced particles are never scattered back to the caller's order.  Same grid as
 * bfg_snap_apply + bfg_snap_deposit_folded (tests/test_gpu_spectrum.py: bit-identical grids). */
int bfg_snap_apply_deposit_folded(int64_t n_part, const double *d_xs, const double *d_ys, const double *d_zs,
                                  const double *d_tot, double L, double L_fold, int64_t n_grid, double *d_grid,
                                  int64_t *d_ndropped, void *stream);
/* Shell sums of |F|^2 over the half spectrum d_spec = rfftn(grid): complex128 [N][N][N/2+1] (interleaved re, im).
 *   |k|(a, b, c) = sqrt(klin[a]^2 + klin[c]^2 + klin[b]^2)   (the notebook's axis order, nb10:12)
 *   shell        = floor((|k| - k0) / dk), kept when 0 <= shell < Nk        (k0 = kbins[0], dk = kbins[1] - kbins[0])
 *   d_pk_sum[s] = sum |F|^2, d_k_sum[s] = sum |k|, d_count[s] = number of modes of the FULL N^3 spectrum in shell s
 * (modes whose mirror image is outside the half spectrum count twice), i.e. np.bincount(kinds[kmsk], weights = ...) of
 * nb10:12,15 before the division by k_c.  d_klin [N] = np.fft.fftfreq(N, ...) from the host.  Outputs are zeroed here.
 * d_spec may be NULL: only d_k_sum and d_count are produced (the notebook's k_c, k_cen). */
int bfg_power_bin_spectrum(int64_t N, const double *d_spec, const double *d_klin, double k0, double dk, int64_t Nk,
                           double *d_pk_sum, double *d_k_sum, int64_t *d_count, void *stream);
/* np.fft.fftn(grid) (nb10:15) as a cuFFT D2Z transform into stream-ordered scratch (N^2 (N/2+1) complex128), followed by
 * bfg_power_bin_spectrum.  d_grid [N^3] float64 is left untouched.  cuFFT is loaded at first use (dlopen; BFG_CUFFT_LIB
 * overrides the library name); BFG_ERR_UNSUPPORTED if it cannot be found.  The D2Z plan is cached per (device, N). */
int bfg_grid_power_spectrum(int64_t N, const double *d_grid, const double *d_klin, double k0, double dk, int64_t Nk,
                            double *d_pk_sum, double *d_k_sum, int64_t *d_count, void *stream);

/* ---- C_l of a shell: the measurement that follows BaryonifyShell.process() in the reference's workflow ------------------
 * `hp.anafast(map)` (examples/04_Baryonify_Density_Shell.ipynb cell 18) = healpix_cxx map2alm_iter (lmax = 3 nside - 1, three
 * Jacobi iterations, unit ring weights) + alm2cl.  First GPU run in round 2 (tests/test_gpu_harmonics.py, parity unpinned: healpy absent); the algorithm
 * is oracle/anafast_rings.py.  a_lm are complex128 in healpy's packing idx(l, m) = m (2 lmax + 1 - m) / 2 + l, m >= 0.
 *   d_ln_mm [lmax + 1]  ln of sqrt((2m+1)/(4 pi) prod_{k<=m} (2k-1)/(2k)), from the host
 *   d_work              complex128 [lmax + 1][4 nside - 1] ring coefficients, bfg_sht_workspace_elems() elements */
int64_t bfg_sht_workspace_elems(int nside, int lmax);
/* One quadrature pass ADDED to d_alm: a_lm += 4 pi / npix * sum_r lambda_lm(cos theta_r) F_m(r)  (map2alm, add = true). */
int bfg_sht_map2alm_pass(int nside, int lmax, const double *d_map, const double *d_ln_mm, double *d_work, double *d_alm,
                         void *stream);
/* d_map = sum_l a_l0 Y_l0 + 2 Re sum_{m>0} a_lm Y_lm  (alm2map; overwrites d_map). */
int bfg_sht_alm2map(int nside, int lmax, const double *d_alm, const double *d_ln_mm, double *d_work, double *d_map,
                    void *stream);
/* d_cl[l] = (|a_l0|^2 + 2 sum_{m=1..l} |a_lm|^2) / (2 l + 1), l = 0 .. lmax  (hp.alm2cl). */
int bfg_sht_alm2cl(int lmax, const double *d_alm, double *d_cl, void *stream);
/* Unit-test entry on the HOST (no GPU needed): the scaled Legendre recursion exactly as the kernels run it;
 * h_out[l - m] = lambda_lm(x), l = m .. lmax, for sin^2(theta) = sin2. */
int bfg_test_sht_lambda_host(int m, int lmax, double ln_mm, double x, double sin2, double *h_out);
/* Unit-test entry on the HOST: the per-ring sums of the two ring kernels (same functions, compiled for the host) for one ring
 * of n pixels at phi_j = (2 j + odd) pi / n:  h_F [lmax + 1][2] = sum_j h_ring[j] exp(-i m phi_j);
 * h_out [n] = sum_m c_m Re(h_b[m] exp(i m phi_j)) with c_0 = 1, c_{m>0} = 2 and h_b [lmax + 1][2]. */
int bfg_test_sht_ring_host(int64_t n, int odd, int lmax, const double *h_ring, double *h_F, const double *h_b, double *h_out);
/* Unit-test entry on the HOST: the Legendre stage of ONE m over all ring pairs with the lane functions of the two Legendre kernels
 * (north/south pairing, parity, packing).  h_F_m [4 nside - 1][2] -> h_alm_m [lmax + 1][2] (entries l >= m; the quadrature
 * weight 4 pi / npix included);  h_alm_in [lmax + 1][2] (indexed by l) -> h_B_m [4 nside - 1][2]. */
int bfg_test_sht_legendre_host(int nside, int lmax, int m, const double *h_ln_mm, const double *h_F_m, double *h_alm_m,
                               const double *h_alm_in, double *h_B_m);

/* ---- locality ordering ----------------------------------------------------------------------------- */
/* Re-orders halo records (and their extras rows) so that neighbours on the sky / in the box are adjacent: north_star (b)
 * "halo batches sorted by sky or box cell for locality".  The reference walks the catalogue in the given order
 * (HealpixRunner.py:315, Map2DRunner.py:482, SnapshotRunner.py:217); its sums do not depend on that order beyond fp64
 * round-off.  mode 0: sky, p0 = colatitude band width [rad] (serpentine in azimuth).  mode 1: box, p0 = L, p1 = number
 * of coarse cells per side, ndim = 2|3.  d_out must not alias d_in.  Uses stream-ordered scratch. */
int bfg_halo_sort(int mode, int64_t n_halo, const double *d_in, double *d_out, const double *d_extras_in,
                  double *d_extras_out, int n_extra, double p0, double p1, int ndim, void *stream);

/* Sky ordering + ownership for ring-range sharding: as mode 0 with band width `band`, and the halos whose disc cannot
 * touch the RING range [pix_lo, pix_hi) are put LAST with BFG_HS_SKIP set, so bfg_shell_offsets / _paint / _paint_anis
 * stop at the first of them -- the catalogue is compacted per rank on the device, without a host round trip. */
int bfg_halo_sort_owned(int nside, int64_t pix_lo, int64_t pix_hi, int64_t n_halo, const double *d_in, double *d_out,
                        const double *d_extras_in, double *d_extras_out, int n_extra, double band, void *stream);

/* ---- small utilities ------------------------------------------------------------------------------ */
/* *d_out (one double) = sum of n doubles (mass-conservation assert, HealpixRunner.py:368-370). */
int bfg_sum_f64(const double *d_x, int64_t n, double *d_out, void *stream);
/* Test entry, pure host (no GPU): the re-binning target the shell re-binning kernels compute per displaced pixel
 * (HealpixRunner.py:357-361: pix2vec + offset -> vec2ang -> get_interp_weights), i.e. the device source regrid_target_fast compiled
 * for the CPU.  h_off [3][n]; outputs [n][4]; h_fast[i] = 0 where the function declines and the kernel takes the literal chain. */
int bfg_test_regrid_target_host(int nside, int64_t n, const int64_t *h_pix, const double *h_off, int64_t *h_out_pix,
                                double *h_out_w, int *h_fast);
/* Test entry, pure host (no GPU): the per-halo scalar prep of bfg_shell_records (same source, same argument meaning) on HOST
 * buffers. */
int bfg_test_shell_records_host(int64_t n_halo, const double *h_cols, int paint, double eps_run, double eps_model, double pixarea,
                                int n_DA, const double *h_DA_x, const double *h_DA_c, int n_g, const double *h_g_x,
                                const double *h_g_run_c, const double *h_g_mod_c, double *h_halos, double *h_aux);
/* Test entry, pure host (no GPU): the table read-out of the halo-loop kernels (corner rows of the non-radial axes blended into one
 * radial row, then interpolation along ln r; BaryonCorrection.py:331-419, Tabulate.py:279-327 == scipy RegularGridInterpolator with
 * bounds_error=False, fill_value=nan) with the kernels' own source on HOST buffers.  force_search != 0: the non-uniform radial branch. */
int bfg_test_table_readout_host(int ndim, const int64_t *shape, const double *const *h_axes, const double *h_values, int flags,
                                int force_search, double lnz, double lnM, const double *h_extras, int64_t n, const double *h_x,
                                double *h_out);
/* Test entry, pure host (no GPU): one halo's per-pixel updates with the shell kernels' own generic update (HealpixRunner.py:336-355 /
 * :464-481), read-out and halo constants.  h_record = the 16-double halo record, h_vec [n][3] = pixel unit vectors; mode 0 (baryonify):
 * h_out [n][3] += nw_vec - vec; mode 1 (paint): h_out [n] += profile * scale.  h_out is accumulated into (zero it first). */
int bfg_test_shell_update_host(int ndim, const int64_t *shape, const double *const *h_axes, const double *h_values, int flags,
                               int force_search, int mode, const double *h_record, const double *h_extras, int64_t n,
                               const double *h_vec, double *h_out);
/* Test entry, pure host (no GPU): one axis of the grid re-binning (regrid_pixels_2D/3D, Map2DRunner.py:13-162) -- the two cells that
 * overlap [x, x + 1) after the periodic wrap and their overlap lengths -- from the kernel's own source.  h_c, h_w: [n][2]. */
int bfg_test_axis_deposit_host(int64_t n, const double *h_pos, int64_t N, int64_t *h_c, double *h_w);
/* Test entry, pure host (no GPU): the lean read-out by SQUARED radius (row_at_r2: uniform ln r axis, table-driven log2) of the default
 * grid / particle / exact shell loops, from the kernels' own source.  offset = what the kernels add to ln r (ln(1/a), -ln R_com, ...).
 * h_out[i] = table value (NaN outside), h_ok[i] = inside [r0, r1]. */
int bfg_test_row_at_r2_host(int ndim, const int64_t *shape, const double *const *h_axes, const double *h_values, int flags, double lnz,
                            double lnM, const double *h_extras, double offset, int64_t n, const double *h_r2, double *h_out,
                            int *h_ok);
/* Test entry, pure host (no GPU): index helpers of the grid and particle kernels on the CPU.  what = 0 NGP cell (np.histogramdd edges,
 * utils/io.py:629-677), 1 wrap_once (SnapshotRunner.py:272-273), 2 cell-list cell, 3 cutout coordinates + periodic indices of one axis
 * (n = Nsize, L = res, h_x[0] = centre cell; Map2DRunner.py:400-429, :500-528). */
int bfg_test_index_helpers_host(int what, int64_t n, const double *h_x, double L, int64_t N, int64_t *h_out_i, double *h_out_d);
/* Test entry, pure host (no GPU): the HEALPix RING device functions (csrc/bfg_common.cuh: query_disc rings and spans, pix2vec,
 * get_interpol, ang2pix, ring2nest / nest2ring -- healpy's C++ T_Healpix_Base algorithms) compiled for the CPU.
 * what = 0 query_disc (h_a = {theta, phi, radius}; h_out_i [cap + 1], last = count), 1 pix2vec (h_idx; h_out_d [n][3]),
 * 2 get_interpol (h_a = theta, h_b = phi; h_out_i, h_out_d [n][4]), 3 ang2pix (h_a, h_b; h_out_i), 4 ring2nest, 5 nest2ring. */
int bfg_test_healpix_host(int what, int nside, int64_t n, const int64_t *h_idx, const double *h_a, const double *h_b, int64_t cap,
                          int64_t *h_out_i, double *h_out_d);
/* Unit-test entry for the table-driven log2 used inside the pixel loops: d_out[i] = log2(d_x[i]). */
int bfg_test_fast_log2(int64_t n, const double *d_x, double *d_out, void *stream);
/* The same on the CPU (pure host, no GPU): fast_log2's own source with the table the device gets. */
int bfg_test_fast_log2_host(int64_t n, const double *h_x, double *h_out);
/* out[i][c] = in[c][i] : component-major offsets -> the reference's (n, ncomp) layout, for tests. */
int bfg_transpose_offsets(const double *d_in, double *d_out, int64_t n, int ncomp, void *stream);

/* ---- host-buffer convenience calls (the end-to-end boundary; allocate, copy in, run, copy out, free) ---- */
int bfg_shell_baryonify_host(const bfg_table *t, int nside, int64_t n_halo, const double *h_halos,
                             const double *h_extras, int n_extra, const double *h_map_in, double *h_map_out,
                             int64_t *h_nupdates, double *h_sums /* [2]: sum(new), sum(old) */);
int bfg_shell_paint_host(const bfg_table *t, int nside, int64_t n_halo, const double *h_halos, const double *h_extras,
                         int n_extra, double *h_map_out, int64_t *h_nupdates);

#ifdef __cplusplus
}
#endif
#endif /* BFG_B200_H */
