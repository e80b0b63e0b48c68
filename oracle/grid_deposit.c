/*
 * oracle/grid_deposit.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Literal C restatement of the reference's two numba re-binning loops
 *   regrid_pixels_2D  /root/reference/BaryonForge/Runners/Map2DRunner.py:13-82
 *   regrid_pixels_3D  /root/reference/BaryonForge/Runners/Map2DRunner.py:85-162
 * (window scan of int(start)-2 .. int(end)+2 cells per axis, overlap length with the
 * +N / -N periodic retries, strict `> 0` test, grid indexed [i=y][j=x][k=z]),
 * checked against the numba originals in tests (when /root/reference is present) and
 * through the committed golden fixtures.
 */
#include <math.h>
#include <stdint.h>

typedef int64_t i64;

static double pymod(double x, double n) { /* Python float % for n > 0 */
    double r = fmod(x, n);
    if (r != 0.0 && r < 0.0) r += n;
    return r;
}

static i64 wrap(i64 i, i64 N) {
    if (i < 0) i += N;
    if (i + 1 > N) i = i % N;
    return i;
}

static double overlap(i64 j, double s, double e, double N) {
    double a = ((double)(j + 1) < e ? (double)(j + 1) : e) - ((double)j > s ? (double)j : s);
    if (a < 0) a = ((double)(j + 1) < e + N ? (double)(j + 1) : e + N) - ((double)j > s + N ? (double)j : s + N);
    if (a < 0) a = ((double)(j + 1) < e - N ? (double)(j + 1) : e - N) - ((double)j > s - N ? (double)j : s - N);
    return a;
}

void grido_regrid_2d(double *grid, i64 N, i64 n, const double *pos /*[n][2]*/, const double *val) {
    for (i64 p = 0; p < n; ++p) {
        double xs = pymod(pos[2 * p + 0], (double)N), ys = pymod(pos[2 * p + 1], (double)N);
        double xe = xs + 1, ye = ys + 1;
        i64 x_min = (i64)xs - 2, x_max = (i64)xe + 2;
        i64 y_min = (i64)ys - 2, y_max = (i64)ye + 2;
        for (i64 i0 = y_min; i0 < y_max; ++i0) {
            i64 i = wrap(i0, N);
            double dy = overlap(i, ys, ye, (double)N);
            for (i64 j0 = x_min; j0 < x_max; ++j0) {
                i64 j = wrap(j0, N);
                double dx = overlap(j, xs, xe, (double)N);
                if ((dx > 0) && (dy > 0)) grid[i * N + j] += (dx * dy) * val[p];
            }
        }
    }
}

void grido_regrid_3d(double *grid, i64 N, i64 n, const double *pos /*[n][3]*/, const double *val) {
    for (i64 p = 0; p < n; ++p) {
        double xs = pymod(pos[3 * p + 0], (double)N), ys = pymod(pos[3 * p + 1], (double)N),
               zs = pymod(pos[3 * p + 2], (double)N);
        double xe = xs + 1, ye = ys + 1, ze = zs + 1;
        i64 x_min = (i64)xs - 2, x_max = (i64)xe + 2;
        i64 y_min = (i64)ys - 2, y_max = (i64)ye + 2;
        i64 z_min = (i64)zs - 2, z_max = (i64)ze + 2;
        for (i64 i0 = y_min; i0 < y_max; ++i0) {
            i64 i = wrap(i0, N);
            double dy = overlap(i, ys, ye, (double)N);
            for (i64 j0 = x_min; j0 < x_max; ++j0) {
                i64 j = wrap(j0, N);
                double dx = overlap(j, xs, xe, (double)N);
                for (i64 k0 = z_min; k0 < z_max; ++k0) {
                    i64 k = wrap(k0, N);
                    double dz = overlap(k, zs, ze, (double)N);
                    if ((dx > 0) && (dy > 0) && (dz > 0)) grid[(i * N + j) * N + k] += ((dx * dy) * dz) * val[p];
                }
            }
        }
    }
}
