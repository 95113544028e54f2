/*
 * lbm_oracle.c -- CPU restatement of the jviquerat/lbm D2Q9 time step.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke()
 * check in __graft_entry__.py and the cpu_baseline / --impl reference legs of
 * bench.py may load it.  The product path (lbm_b200/) never links, imports or
 * calls anything in oracle/ and fails loudly when its CUDA library is missing.
 *
 * Every function restates one phase of the reference, phase by phase and array
 * by array (no fusion), so that it can be compared against the reference's own
 * Numba/NumPy execution.  Citations are into /root/reference (read-only):
 *
 *   orc_macro                       lbm/src/core/lattice.py:178-189
 *   orc_equilibrium                 lbm/src/core/nb.py:7-17
 *   orc_col_str                     lbm/src/core/nb.py:21-45
 *   orc_drag_lift                   lbm/src/core/nb.py:49-73
 *   orc_bounce_back_obstacle        lbm/src/core/nb.py:77-117
 *   orc_zou_he_left_wall_velocity   lbm/src/core/nb.py:121-143
 *   orc_zou_he_right_wall_velocity  lbm/src/core/nb.py:147-169
 *   orc_zou_he_right_wall_pressure  lbm/src/core/nb.py:173-195
 *   orc_zou_he_top_wall_velocity    lbm/src/core/nb.py:199-221
 *   orc_zou_he_bottom_wall_velocity lbm/src/core/nb.py:225-247
 *   orc_zou_he_*_corner             lbm/src/core/nb.py:251-344
 *   D2Q9 tables c, w, ns            lbm/src/core/lattice.py:135-152
 *
 * Parity pin: tests/test_oracle_vs_reference.py runs the reference itself (when
 * /root/reference is present) beside this file on identical inputs, and
 * tests/golden/ holds vectors generated from the reference by
 * tests/golden/make_golden.py; tests/test_oracle_golden.py checks this file
 * against them everywhere (no reference needed).
 *
 * Layout: every field is C-ordered [q][i][j] with i = x (0..nx-1) and j = y
 * (0..ny-1), j contiguous, float64 -- exactly lattice.py:155-174.
 * Compiled with -ffp-contract=off: strict IEEE evaluation of the expressions as
 * written (the reference's fastmath JIT agrees with strict evaluation to 6e-16,
 * SURVEY.md section 9.7).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define Q 9
static const int    CX[Q] = {0, 1, -1, 0, 0, 1, -1, -1, 1};
static const int    CY[Q] = {0, 0, 0, 1, -1, 1, -1, 1, -1};
static const int    NS[Q] = {0, 2, 1, 4, 3, 6, 5, 8, 7};
static const double W[Q]  = {4.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0,
                             1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0};

#define IDX(q, i, j) (((size_t)(q) * nx + (size_t)(i)) * ny + (size_t)(j))
#define IDU(d, i, j) (((size_t)(d) * nx + (size_t)(i)) * ny + (size_t)(j))
#define IDR(i, j) ((size_t)(i) * ny + (size_t)(j))

int orc_abi_version(void) { return 1; }

void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_get_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* lattice.py:178-189 -- rho = sum_q g_q (np.sum over axis 0 adds the planes in
 * index order); u = (c . g) / rho with the zero entries of c dropped. */
void orc_macro(int64_t nx, int64_t ny, const double *g, double *rho, double *u)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nx; i++) {
        for (int64_t j = 0; j < ny; j++) {
            double r = g[IDX(0, i, j)];
            for (int q = 1; q < Q; q++) r += g[IDX(q, i, j)];
            double mx = g[IDX(1, i, j)] - g[IDX(2, i, j)] + g[IDX(5, i, j)] - g[IDX(6, i, j)] -
                        g[IDX(7, i, j)] + g[IDX(8, i, j)];
            double my = g[IDX(3, i, j)] - g[IDX(4, i, j)] + g[IDX(5, i, j)] - g[IDX(6, i, j)] +
                        g[IDX(7, i, j)] - g[IDX(8, i, j)];
            rho[IDR(i, j)]  = r;
            u[IDU(0, i, j)] = mx / r;
            u[IDU(1, i, j)] = my / r;
        }
    }
}

/* nb.py:7-17 */
void orc_equilibrium(int64_t nx, int64_t ny, const double *u, const double *rho, double *g_eq)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nx; i++) {
        for (int64_t j = 0; j < ny; j++) {
            double ux = u[IDU(0, i, j)], uy = u[IDU(1, i, j)];
            double v  = 1.5 * (ux * ux + uy * uy);
            double r  = rho[IDR(i, j)];
            for (int q = 0; q < Q; q++) {
                double t = 3.0 * (ux * (double)CX[q] + uy * (double)CY[q]);
                double e = (1.0 + t + 0.5 * (t * t) - v);
                g_eq[IDX(q, i, j)] = e * (r * W[q]);
            }
        }
    }
}

/* nb.py:21-45 -- TRT collision of every node (no mask), then push-streaming into
 * g.  Entries of g with no in-domain source keep their previous value. */
void orc_col_str(int64_t nx, int64_t ny, double *g, const double *g_eq, double *g_up,
                 double om_p, double om_m)
{
    const double a_self = 1.0 - 0.5 * (om_p + om_m);
    const double a_opp  = 0.5 * (om_p - om_m);
    const double a_eq   = 0.5 * (om_p + om_m);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nx; i++) {
        for (int64_t j = 0; j < ny; j++) {
            g_up[IDX(0, i, j)] = (1.0 - om_p) * g[IDX(0, i, j)] + om_p * g_eq[IDX(0, i, j)];
            for (int q = 1; q < Q; q++) {
                int qb = NS[q];
                g_up[IDX(q, i, j)] = (a_self * g[IDX(q, i, j)] - a_opp * g[IDX(qb, i, j)] +
                                      a_eq * g_eq[IDX(q, i, j)] + a_opp * g_eq[IDX(qb, i, j)]);
            }
        }
    }
    /* stream: g_q[x + c_q] = g_up_q[x] wherever the destination is inside */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nx; i++) {
        for (int q = 0; q < Q; q++) {
            int64_t is = i - CX[q];
            if (is < 0 || is >= nx) continue;
            int64_t j0 = CY[q] > 0 ? 1 : 0;
            int64_t j1 = CY[q] < 0 ? ny - 1 : ny;
            for (int64_t j = j0; j < j1; j++) g[IDX(q, i, j)] = g_up[IDX(q, is, j - CY[q])];
        }
    }
}

/* nb.py:49-73 -- momentum exchange over the link list of one obstacle.  The
 * reference's prange reduction order is unspecified; this sums in list order. */
void orc_drag_lift(int64_t nx, int64_t ny, int64_t K, const int64_t *boundary, const double *g_up,
                   const double *g, double R_ref, double U_ref, double L_ref, double *CxCy)
{
    double fx = 0.0, fy = 0.0;
    for (int64_t k = 0; k < K; k++) {
        int64_t i = boundary[3 * k + 0], j = boundary[3 * k + 1];
        int     q = (int)boundary[3 * k + 2];
        if (i < 0) i += nx; /* NumPy/Numba negative-index wrap, SURVEY.md section 10.3 */
        if (j < 0) j += ny;
        int    qb = NS[q];
        double g0 = g_up[IDX(q, i, j)] + g[IDX(qb, i, j)];
        fx += g0 * (double)CX[q];
        fy += g0 * (double)CY[q];
    }
    CxCy[0] = -2.0 * fx / (R_ref * L_ref * (U_ref * U_ref));
    CxCy[1] = -2.0 * fy / (R_ref * L_ref * (U_ref * U_ref));
}

static inline int64_t wrap(int64_t a, int64_t n) { return a < 0 ? a + n : a; }

/* nb.py:77-117 -- (interpolated) bounce-back on the link list of one obstacle.
 * Link row = (i, j, q): fluid node (i,j), q points from it into the solid. */
void orc_bounce_back_obstacle(int64_t nx, int64_t ny, int IBB, int64_t K, const int64_t *boundary,
                              const double *obs_ibb, const double *g_up, double *g)
{
    for (int64_t k = 0; k < K; k++) {
        int64_t i = wrap(boundary[3 * k + 0], nx), j = wrap(boundary[3 * k + 1], ny);
        int     q  = (int)boundary[3 * k + 2];
        int     qb = NS[q];
        if (!IBB) {
            g[IDX(qb, i, j)] = g_up[IDX(q, i, j)];
            continue;
        }
        int64_t im  = wrap(i + CX[qb], nx), jm = wrap(j + CY[qb], ny);
        int64_t imm = wrap(i + 2 * CX[qb], nx), jmm = wrap(j + 2 * CY[qb], ny);
        double  p   = obs_ibb[k];
        double  pp  = 2.0 * p;
        if (p < 0.5) {
            g[IDX(qb, i, j)] = (p * (pp + 1.0) * g_up[IDX(q, i, j)] +
                                (1.0 + pp) * (1.0 - pp) * g_up[IDX(q, im, jm)] -
                                p * (1.0 - pp) * g_up[IDX(q, imm, jmm)]);
        } else {
            g[IDX(qb, i, j)] = ((1.0 / (p * (pp + 1.0))) * g_up[IDX(q, i, j)] +
                                ((pp - 1.0) / p) * g_up[IDX(qb, i, j)] +
                                ((1.0 - pp) / (1.0 + pp)) * g_up[IDX(qb, im, jm)]);
        }
    }
}

static const double cst1 = 2.0 / 3.0, cst2 = 1.0 / 6.0, cst3 = 1.0 / 2.0;

/* nb.py:121-143 */
void orc_zou_he_left_wall_velocity(int64_t nx, int64_t ny, double *u, const double *u_left,
                                   double *rho, double *g)
{
    for (int64_t j = 0; j < ny; j++) {
        double ux = u_left[j], uy = u_left[ny + j];
        u[IDU(0, 0, j)] = ux;
        u[IDU(1, 0, j)] = uy;
        double r = (g[IDX(0, 0, j)] + g[IDX(3, 0, j)] + g[IDX(4, 0, j)] + 2.0 * g[IDX(2, 0, j)] +
                    2.0 * g[IDX(6, 0, j)] + 2.0 * g[IDX(7, 0, j)]) / (1.0 - ux);
        rho[IDR(0, j)] = r;
        double d = g[IDX(3, 0, j)] - g[IDX(4, 0, j)];
        g[IDX(1, 0, j)] = (g[IDX(2, 0, j)] + cst1 * r * ux);
        g[IDX(5, 0, j)] = (g[IDX(6, 0, j)] - cst3 * d + cst2 * r * ux + cst3 * r * uy);
        g[IDX(8, 0, j)] = (g[IDX(7, 0, j)] + cst3 * d + cst2 * r * ux - cst3 * r * uy);
    }
}

/* nb.py:147-169 */
void orc_zou_he_right_wall_velocity(int64_t nx, int64_t ny, double *u, const double *u_right,
                                    double *rho, double *g)
{
    const int64_t lx = nx - 1;
    for (int64_t j = 0; j < ny; j++) {
        double ux = u_right[j], uy = u_right[ny + j];
        u[IDU(0, lx, j)] = ux;
        u[IDU(1, lx, j)] = uy;
        double r = (g[IDX(0, lx, j)] + g[IDX(3, lx, j)] + g[IDX(4, lx, j)] +
                    2.0 * g[IDX(1, lx, j)] + 2.0 * g[IDX(5, lx, j)] + 2.0 * g[IDX(8, lx, j)]) /
                   (1.0 + ux);
        rho[IDR(lx, j)] = r;
        double d = g[IDX(3, lx, j)] - g[IDX(4, lx, j)];
        g[IDX(2, lx, j)] = (g[IDX(1, lx, j)] - cst1 * r * ux);
        g[IDX(6, lx, j)] = (g[IDX(5, lx, j)] + cst3 * d - cst2 * r * ux - cst3 * r * uy);
        g[IDX(7, lx, j)] = (g[IDX(8, lx, j)] - cst3 * d - cst2 * r * ux + cst3 * r * uy);
    }
}

/* nb.py:173-195 */
void orc_zou_he_right_wall_pressure(int64_t nx, int64_t ny, double *u, const double *rho_right,
                                    const double *u_right, double *rho, double *g)
{
    const int64_t lx = nx - 1;
    for (int64_t j = 0; j < ny; j++) {
        double r  = rho_right[j];
        double uy = u_right[ny + j];
        rho[IDR(lx, j)]  = r;
        u[IDU(1, lx, j)] = uy;
        double ux = (g[IDX(0, lx, j)] + g[IDX(3, lx, j)] + g[IDX(4, lx, j)] +
                     2.0 * g[IDX(1, lx, j)] + 2.0 * g[IDX(5, lx, j)] + 2.0 * g[IDX(8, lx, j)]) / r -
                    1.0;
        u[IDU(0, lx, j)] = ux;
        double d = g[IDX(3, lx, j)] - g[IDX(4, lx, j)];
        g[IDX(2, lx, j)] = (g[IDX(1, lx, j)] - cst1 * r * ux);
        g[IDX(6, lx, j)] = (g[IDX(5, lx, j)] + cst3 * d - cst2 * r * ux - cst3 * r * uy);
        g[IDX(7, lx, j)] = (g[IDX(8, lx, j)] - cst3 * d - cst2 * r * ux + cst3 * r * uy);
    }
}

/* nb.py:199-221 */
void orc_zou_he_top_wall_velocity(int64_t nx, int64_t ny, double *u, const double *u_top,
                                  double *rho, double *g)
{
    const int64_t ly = ny - 1;
    for (int64_t i = 0; i < nx; i++) {
        double ux = u_top[i], uy = u_top[nx + i];
        u[IDU(0, i, ly)] = ux;
        u[IDU(1, i, ly)] = uy;
        double r = (g[IDX(0, i, ly)] + g[IDX(1, i, ly)] + g[IDX(2, i, ly)] +
                    2.0 * g[IDX(3, i, ly)] + 2.0 * g[IDX(5, i, ly)] + 2.0 * g[IDX(7, i, ly)]) /
                   (1.0 + uy);
        rho[IDR(i, ly)] = r;
        double d = g[IDX(1, i, ly)] - g[IDX(2, i, ly)];
        g[IDX(4, i, ly)] = (g[IDX(3, i, ly)] - cst1 * r * uy);
        g[IDX(8, i, ly)] = (g[IDX(7, i, ly)] - cst3 * d + cst3 * r * ux - cst2 * r * uy);
        g[IDX(6, i, ly)] = (g[IDX(5, i, ly)] + cst3 * d - cst3 * r * ux - cst2 * r * uy);
    }
}

/* nb.py:225-247 */
void orc_zou_he_bottom_wall_velocity(int64_t nx, int64_t ny, double *u, const double *u_bot,
                                     double *rho, double *g)
{
    for (int64_t i = 0; i < nx; i++) {
        double ux = u_bot[i], uy = u_bot[nx + i];
        u[IDU(0, i, 0)] = ux;
        u[IDU(1, i, 0)] = uy;
        double r = (g[IDX(0, i, 0)] + g[IDX(1, i, 0)] + g[IDX(2, i, 0)] + 2.0 * g[IDX(4, i, 0)] +
                    2.0 * g[IDX(6, i, 0)] + 2.0 * g[IDX(8, i, 0)]) / (1.0 - uy);
        rho[IDR(i, 0)] = r;
        double d = g[IDX(1, i, 0)] - g[IDX(2, i, 0)];
        g[IDX(3, i, 0)] = (g[IDX(4, i, 0)] + cst1 * r * uy);
        g[IDX(5, i, 0)] = (g[IDX(6, i, 0)] - cst3 * d + cst3 * r * ux + cst2 * r * uy);
        g[IDX(7, i, 0)] = (g[IDX(8, i, 0)] + cst3 * d - cst3 * r * ux + cst2 * r * uy);
    }
}

/* Corners, nb.py:251-344.  (ci,cj) is the corner, (ni,cj) its x-neighbour on
 * the same horizontal wall whose u and rho are copied; sx, sy = +1 when the
 * unknown populations point towards +x / +y. */
static void corner(int64_t nx, int64_t ny, double *u, double *rho, double *g, int64_t ci,
                   int64_t cj, int64_t ni)
{
    double ux = u[IDU(0, ni, cj)], uy = u[IDU(1, ni, cj)];
    double r  = rho[IDR(ni, cj)];
    u[IDU(0, ci, cj)] = ux;
    u[IDU(1, ci, cj)] = uy;
    rho[IDR(ci, cj)]  = r;
    const int left = (ci == 0), bottom = (cj == 0);
    if (left) g[IDX(1, ci, cj)] = (g[IDX(2, ci, cj)] + (2.0 / 3.0) * r * ux);
    else      g[IDX(2, ci, cj)] = (g[IDX(1, ci, cj)] - (2.0 / 3.0) * r * ux);
    if (bottom) g[IDX(3, ci, cj)] = (g[IDX(4, ci, cj)] + (2.0 / 3.0) * r * uy);
    else        g[IDX(4, ci, cj)] = (g[IDX(3, ci, cj)] - (2.0 / 3.0) * r * uy);
    if (left && bottom) { /* nb.py:251-271 */
        g[IDX(5, ci, cj)] = (g[IDX(6, ci, cj)] + (1.0 / 6.0) * r * ux + (1.0 / 6.0) * r * uy);
        g[IDX(7, ci, cj)] = 0.0;
        g[IDX(8, ci, cj)] = 0.0;
    } else if (left && !bottom) { /* nb.py:275-296 */
        g[IDX(8, ci, cj)] = (g[IDX(7, ci, cj)] + (1.0 / 6.0) * r * ux - (1.0 / 6.0) * r * uy);
        g[IDX(5, ci, cj)] = 0.0;
        g[IDX(6, ci, cj)] = 0.0;
    } else if (!left && !bottom) { /* nb.py:300-320 */
        g[IDX(6, ci, cj)] = (g[IDX(5, ci, cj)] - (1.0 / 6.0) * r * ux - (1.0 / 6.0) * r * uy);
        g[IDX(7, ci, cj)] = 0.0;
        g[IDX(8, ci, cj)] = 0.0;
    } else { /* nb.py:324-344 */
        g[IDX(7, ci, cj)] = (g[IDX(8, ci, cj)] - (1.0 / 6.0) * r * ux + (1.0 / 6.0) * r * uy);
        g[IDX(5, ci, cj)] = 0.0;
        g[IDX(6, ci, cj)] = 0.0;
    }
    g[IDX(0, ci, cj)] = (r - g[IDX(1, ci, cj)] - g[IDX(2, ci, cj)] - g[IDX(3, ci, cj)] -
                         g[IDX(4, ci, cj)] - g[IDX(5, ci, cj)] - g[IDX(6, ci, cj)] -
                         g[IDX(7, ci, cj)] - g[IDX(8, ci, cj)]);
}

void orc_zou_he_bottom_left_corner(int64_t nx, int64_t ny, double *u, double *rho, double *g)
{
    corner(nx, ny, u, rho, g, 0, 0, 1);
}
void orc_zou_he_top_left_corner(int64_t nx, int64_t ny, double *u, double *rho, double *g)
{
    corner(nx, ny, u, rho, g, 0, ny - 1, 1);
}
void orc_zou_he_top_right_corner(int64_t nx, int64_t ny, double *u, double *rho, double *g)
{
    corner(nx, ny, u, rho, g, nx - 1, ny - 1, nx - 2);
}
void orc_zou_he_bottom_right_corner(int64_t nx, int64_t ny, double *u, double *rho, double *g)
{
    corner(nx, ny, u, rho, g, nx - 1, 0, nx - 2);
}
