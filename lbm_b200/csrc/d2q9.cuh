// d2q9.cuh -- per-cell arithmetic of the D2Q9 TRT time step (device side).
//
// Restates, for one lattice cell held in registers, what the reference does with whole-array
// passes: lattice.macro (lbm/src/core/lattice.py:178-189), nb_equilibrium (nb.py:7-17), the
// collision half of nb_col_str (nb.py:25-35), the five Zou-He wall kernels (nb.py:121-247) and
// the four corner kernels (nb.py:251-344).  Expression order follows the reference text so that
// the STRICT arithmetic (no FMA contraction) is bit-identical to oracle/lbm_oracle.c.
#pragma once
#include <cstdint>

namespace lbm {

// D2Q9 tables, lattice.py:135-152
__device__ __constant__ const int kCx[9] = {0, 1, -1, 0, 0, 1, -1, -1, 1};
__device__ __constant__ const int kCy[9] = {0, 0, 0, 1, -1, 1, -1, 1, -1};
__host__ __device__ constexpr int opp(int q) { return q == 0 ? 0 : (q & 1 ? q + 1 : q - 1); }
__host__ __device__ constexpr int cx_of(int q) { return q == 1 || q == 5 || q == 8 ? 1 : (q == 2 || q == 6 || q == 7 ? -1 : 0); }
__host__ __device__ constexpr int cy_of(int q) { return q == 3 || q == 5 || q == 7 ? 1 : (q == 4 || q == 6 || q == 8 ? -1 : 0); }

// Arithmetic policy.  STRICT: every operation separately rounded (the *_rn intrinsics are never
// contracted by nvcc).  FUSED: plain operators, nvcc contracts mul+add into FMA (-fmad=true).
template <typename T, bool STRICT> struct Ar;
__device__ __forceinline__ double rcp_fast(double b);
template <> struct Ar<double, true> {
    static constexpr bool strict = true;
    static __device__ __forceinline__ double mulr(double a, double b) { return __dmul_rn(a, b); }   // (collide_fused is never strict)
    static __device__ __forceinline__ double fmad(double a, double b, double c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double rcp(double b) { return __ddiv_rn(1.0, b); }   // (collide_fused is never strict)
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ void div2(double a0, double a1, double b, double &q0, double &q1)
    {
        q0 = __ddiv_rn(a0, b);
        q1 = __ddiv_rn(a1, b);
    }
};
template <> struct Ar<float, true> {
    static constexpr bool strict = true;
    static __device__ __forceinline__ float mulr(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float fmad(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float rcp(float b) { return __fdiv_rn(1.0f, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ void div2(float a0, float a1, float b, float &q0, float &q1)
    {
        q0 = __fdiv_rn(a0, b);
        q1 = __fdiv_rn(a1, b);
    }
};
// FUSED division: the IEEE double division of nvcc is a ~60-instruction subroutine call, and the
// multi-update kernels are bound by FP64 issue and dependent-issue latency once the HBM traffic is
// cut (profiles/README.md).  The divisor is always a density or 1 +- u (normal range, far from
// 0/inf), so a reciprocal seed (MUFU.RCP64H, relative error < 1e-6 measured on B200 with
// tools/probe/rcp_seed.cu) refined by ONE cubic step  y (1 + e + e^2),  e = 1 - b y  is used: three
// dependent DFMAs, no special-case handling, |y b - 1| <= 2.3e-16 (measured over [1e-3, 1e3]).
// Single quotients (wall cells) add a residual correction.  (The reference's Numba kernels run with
// fastmath, which licenses the same reciprocal rewrite.)
__device__ __forceinline__ double rcp_fast(double b)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
    const double e = fma(-b, y, 1.0);
    const double t = fma(e, e, e);
    return fma(y, t, y);
}
__device__ __forceinline__ double mul_rcp(double a, double b, double y /* ~1/b */)
{
    const double q = a * y;
    return fma(fma(-b, q, a), y, q);
}
template <> struct Ar<double, false> {
    static constexpr bool strict = false;
    // explicit forms for collide_fused: which product of  a*b + c*d  gets fused must not be left to the compiler
    // (it may choose differently in different kernels, and the kernels have to agree bit for bit)
    static __device__ __forceinline__ double mulr(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double fmad(double a, double b, double c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ double mul(double a, double b) { return a * b; }
    static __device__ __forceinline__ double add(double a, double b) { return a + b; }
    static __device__ __forceinline__ double sub(double a, double b) { return a - b; }
    static __device__ __forceinline__ double rcp(double b) { return rcp_fast(b); }
    static __device__ __forceinline__ double div(double a, double b) { return mul_rcp(a, b, rcp_fast(b)); }
    static __device__ __forceinline__ void div2(double a0, double a1, double b, double &q0, double &q1)
    {
        const double y = rcp_fast(b);   // |y b - 1| <= 2 ulp: the two quotients are good to 2.5 ulp
        q0 = a0 * y;
        q1 = a1 * y;
    }
};
template <> struct Ar<float, false> {
    static constexpr bool strict = false;
    static __device__ __forceinline__ float mulr(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float fmad(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
    static __device__ __forceinline__ float rcp(float b) { return __frcp_rn(b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdividef(a, b); }
    static __device__ __forceinline__ void div2(float a0, float a1, float b, float &q0, float &q1)
    {
        const float y = __frcp_rn(b);
        q0 = a0 * y;
        q1 = a1 * y;
    }
};

// Storage of a population.  f64: the population itself (the reference's arrays).  f32: its
// DEVIATION from the rest-state weight, h_q = f_q - w_q: a float keeps ~7 digits, and f_q ~ 0.03..0.44
// while the physics lives in deviations of 1e-2 and below, so storing f_q would waste two digits and
// let rounding drift into u (measured: 1.1e-5 of max|u| after 120 iterations).  Every operation of
// the update is affine in the populations with weights that sum to one (TRT, IBB, Zou-He, see the
// notes at each function), so the update is restated exactly on h; host arrays stay populations.
template <typename T> struct Stored { static constexpr bool dev = false; };
template <> struct Stored<float> { static constexpr bool dev = true; };
__host__ __device__ constexpr double weight_of(int q) { return q == 0 ? 4.0 / 9.0 : (q < 5 ? 1.0 / 9.0 : 1.0 / 36.0); }

template <typename T> struct Coef {
    T one_m_omp, om_p;        // q = 0:      (1-om_p) g0 + om_p geq0                 nb.py:26
    T a_self, a_opp, a_eq;    // q >= 1:     1-(om_p+om_m)/2, (om_p-om_m)/2, (om_p+om_m)/2   nb.py:31-35
    // FUSED arithmetic (collide_fused): w_q om_p for the three weight classes, 4.5 w_q om_p,
    // 3 w_q om_m, and the pair coefficients (1 - om_p)/2, (1 - om_m)/2
    T wp0, wp1, wp5, wq1, wq5, wm1, wm5, cs, cd;
};

// rho = sum_q g_q in index order; u = (c . g) / rho     (lattice.py:181-189, oracle orc_macro)
// Deviation storage: dr = sum_q h_q = rho - 1 (sum_q w_q = 1), momentum unchanged (sum_q c_q w_q = 0).
template <typename A, typename T>
__device__ __forceinline__ void macro(const T (&G)[9], T &r, T &ux, T &uy, T &dr)
{
    if (A::strict) {
        r = A::add(G[0], G[1]);
#pragma unroll
        for (int q = 2; q < 9; q++) r = A::add(r, G[q]);
    } else {   // same sum as a tree: 4 dependent additions instead of 8 (the kernels are latency bound)
        r = (((G[0] + G[1]) + (G[2] + G[3])) + ((G[4] + G[5]) + (G[6] + G[7]))) + G[8];
    }
    dr = r;
    if (Stored<T>::dev) r = A::add(T(1.0), dr);
    T mx, my;
    if (A::strict) {
        mx = A::add(A::sub(A::sub(A::add(A::sub(G[1], G[2]), G[5]), G[6]), G[7]), G[8]);
        my = A::sub(A::add(A::sub(A::add(A::sub(G[3], G[4]), G[5]), G[6]), G[7]), G[8]);
    } else {   // same sums with the two diagonal differences shared
        const T d56 = G[5] - G[6], d78 = G[7] - G[8];
        mx = ((G[1] - G[2]) + d56) - d78;
        my = ((G[3] - G[4]) + d56) + d78;
    }
    A::div2(mx, my, r, ux, uy);
}

// g_eq (nb.py:10-17) followed by the TRT collision (nb.py:25-35); G -> F in place.
// Deviation storage: g_eq_q - w_q = w_q (dr + rho (t + t^2/2 - v)); the TRT combination has
// coefficients a_self - a_opp + a_eq + a_opp = 1 and w_q = w_qbar, so it maps h to h unchanged.
//
// FUSED arithmetic (collide_fused): lattice.macro + nb_equilibrium + the TRT collision in one
// expression tree, written in the MOMENTS  rho, m = rho u  and, per opposite pair (q, qbar), in the
// pair's SUM and DIFFERENCE.  With  eq_s = (geq_q + geq_qbar)/2 = w rho (1 + 4.5 s^2 - v),  eq_a = (geq_q -
// geq_qbar)/2 = 3 w rho s  (s = c_q.u, v = 1.5 u.u)  and  a_eq + a_opp = om_p,  a_eq - a_opp = om_m,
// a_self - a_opp = 1 - om_p,  a_self + a_opp = 1 - om_m  the TRT update (nb.py:31-35) of a pair is
//     (F_q + F_qbar)/2 = (1 - om_p)/2 (g_q + g_qbar) + om_p eq_s        =: Fs
//     (F_q - F_qbar)/2 = (1 - om_m)/2 (g_q - g_qbar) + om_m eq_a        =: Fd
//     F_q = Fs + Fd,   F_qbar = Fs - Fd
// and, with  m_s = c_q.m = rho s,  y = 1/rho  and  E = rho - 1.5 (m.m) y,
//     om_p eq_s = om_p w E + (4.5 om_p w y) m_s^2,      om_m eq_a = 3 om_m w m_s .
// The pair sums and differences are the ones the moments are built from anyway (rho = g_0 + sum of the
// four pair sums, m from the four pair differences), and E is shared by the three weight classes, so the
// whole cell costs 59 FP64 instructions (the plain pair form  a_self g_q - a_opp g_qbar + ...  of round 1
// took 76; keeping everything but one multiply-add independent of the reciprocal costs 63 and measured
// 1.8 % slower: the kernels are bound by energy and shared-memory traffic rather than by this chain).  Algebraically identical to
// lattice.py:181-189 + nb.py:10-17 + 25-35; rounding differs at the 1e-16 level like any FMA
// contraction does.  Deviation storage (f32): om_p (eq_s - w) = om_p w dr + y (...), same form.
template <typename A, typename T>
__device__ __forceinline__ void collide_fused(T (&G)[9], const Coef<T> &c, bool want_u, T &r, T &ux, T &uy)
{
    constexpr bool dev = Stored<T>::dev;
    const T S[4] = {G[1] + G[2], G[3] + G[4], G[5] + G[6], G[7] + G[8]};
    const T D[4] = {G[1] - G[2], G[3] - G[4], G[5] - G[6], G[7] - G[8]};
    const T sum = G[0] + ((S[0] + S[1]) + (S[2] + S[3]));
    r = dev ? T(1.0) + sum : sum;
    const T mx = (D[0] + D[2]) - D[3];
    const T my = (D[1] + D[2]) + D[3];
    const T y = A::rcp(r);
    const T ms[4] = {mx, my, mx + my, my - mx};
    // every multiply-add below is written out (A::fmad = one fused operation, A::mulr = a rounded product): the
    // compiler must not be the one to decide which product of a sum of two products is fused
    const T m2 = A::fmad(my, my, A::mulr(mx, mx));
    // E = rho - 1.5 (m.m) y is shared by all weight classes; b = 4.5 om_p w y
    const T E = A::fmad(-A::mulr(T(1.5), m2), y, sum);
    const T e0 = A::mulr(E, c.wp0), e1 = A::mulr(E, c.wp1), e5 = A::mulr(E, c.wp5);   // om_p w (rho - 1.5 m.m / rho)   (f32: rho -> dr)
    const T b1 = A::mulr(c.wq1, y), b5 = A::mulr(c.wq5, y);
    G[0] = A::fmad(c.one_m_omp, G[0], e0);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int q = 2 * k + 1, qb = q + 1;
        const T Fs = A::fmad(k < 2 ? b1 : b5, A::mulr(ms[k], ms[k]), A::fmad(c.cs, S[k], k < 2 ? e1 : e5));
        const T Fd = A::fmad(c.cd, D[k], A::mulr(k < 2 ? c.wm1 : c.wm5, ms[k]));
        G[q] = Fs + Fd;
        G[qb] = Fs - Fd;
    }
    if (want_u) {   // lattice.macro's u, only where it is stored
        ux = A::mulr(mx, y);
        uy = A::mulr(my, y);
    }
}

template <typename A, typename T>
__device__ __forceinline__ void collide(T (&G)[9], T r, T dr, T ux, T uy, const Coef<T> &c)
{
    constexpr bool dev = Stored<T>::dev;
    const T w0 = T(4.0 / 9.0), w1 = T(1.0 / 9.0), w5 = T(1.0 / 36.0);
    const T v = A::mul(T(1.5), A::add(A::mul(ux, ux), A::mul(uy, uy)));
    const T rw0 = A::mul(r, w0), rw1 = A::mul(r, w1), rw5 = A::mul(r, w5);
    // q = 0: t = 0
    {
        T geq;
        if (dev) geq = A::mul(w0, A::sub(dr, A::mul(r, v)));
        else     geq = A::mul(A::sub(A::add(A::add(T(1.0), T(0.0)), T(0.0)), v), rw0);
        G[0] = A::add(A::mul(c.one_m_omp, G[0]), A::mul(c.om_p, geq));
    }
    // opposite pairs (q, q+1): t_q = 3 c_q.u, t_qbar = -t_q
    const T s[4] = {ux, uy, A::add(ux, uy), A::sub(uy, ux)};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int q = 2 * k + 1, qb = q + 1;
        const T rw = k < 2 ? rw1 : rw5;
        const T t = A::mul(T(3.0), s[k]);
        const T h = A::mul(T(0.5), A::mul(t, t));
        T eq, eb;
        if (dev) {
            const T w = k < 2 ? w1 : w5;
            const T hv = A::sub(h, v);
            eq = A::mul(w, A::add(dr, A::mul(r, A::add(hv, t))));
            eb = A::mul(w, A::add(dr, A::mul(r, A::sub(hv, t))));
        } else {
            eq = A::mul(A::sub(A::add(A::add(T(1.0), t), h), v), rw);
            eb = A::mul(A::sub(A::add(A::sub(T(1.0), t), h), v), rw);
        }
        const T gq = G[q], gb = G[qb];
        G[q]  = A::add(A::add(A::sub(A::mul(c.a_self, gq), A::mul(c.a_opp, gb)), A::mul(c.a_eq, eq)),
                       A::mul(c.a_opp, eb));
        G[qb] = A::add(A::add(A::sub(A::mul(c.a_self, gb), A::mul(c.a_opp, gq)), A::mul(c.a_eq, eb)),
                       A::mul(c.a_opp, eq));
    }
}

// macro + equilibrium + collision of one cell, G -> F in place; (r, ux, uy) are lattice.macro's
// fields (ux, uy only if want_u).  STRICT: the reference's expression order; FUSED: collide_fused.
template <typename A, typename T>
__device__ __forceinline__ void collide_cell(T (&G)[9], const Coef<T> &c, bool want_u, T &r, T &ux, T &uy)
{
    if (A::strict) {
        T dr;
        macro<A, T>(G, r, ux, uy, dr);
        collide<A, T>(G, r, dr, ux, uy, c);
    } else {
        collide_fused<A, T>(G, c, want_u, r, ux, uy);
    }
}

// equilibrium only (nb.py:10-17)
template <typename A, typename T>
__device__ __forceinline__ void equilibrium(T (&E)[9], T r, T ux, T uy)
{
    const T w[3] = {T(4.0 / 9.0), T(1.0 / 9.0), T(1.0 / 36.0)};
    const T v = A::mul(T(1.5), A::add(A::mul(ux, ux), A::mul(uy, uy)));
#pragma unroll
    for (int q = 0; q < 9; q++) {
        const T cu = A::add(A::mul(ux, T(cx_of(q))), A::mul(uy, T(cy_of(q))));
        const T t = A::mul(T(3.0), cu);
        const T e = A::sub(A::add(A::add(T(1.0), t), A::mul(T(0.5), A::mul(t, t))), v);
        E[q] = A::mul(e, A::mul(r, w[q == 0 ? 0 : (q < 5 ? 1 : 2)]));
    }
}

// ---- Zou-He walls -----------------------------------------------------------------------
// Each fills the three unknown populations of an off-corner wall cell and returns rho/ux/uy as
// the reference stores them in its rho/u arrays.
template <typename A, typename T> struct ZouHe {
    static __device__ __forceinline__ T sum6(T a, T b, T c, T d, T e, T f)
    {   // a + b + c + 2d + 2e + 2f, left to right
        return A::add(A::add(A::add(A::add(A::add(a, b), c), A::mul(T(2.0), d)), A::mul(T(2.0), e)),
                      A::mul(T(2.0), f));
    }
    static constexpr double c1 = 2.0 / 3.0, c2 = 1.0 / 6.0, c3 = 0.5;
    // Deviation storage: in each wall's density sum the weights add up to exactly one
    // (w0 + 2 w_card + 2 (w_card + 2 w_diag) = 1), so  sum(f) = 1 + sum(h);  the three unknown
    // populations are differences of populations with equal weights plus terms in rho u, unchanged.
    static __device__ __forceinline__ T full(T s) { return Stored<T>::dev ? A::add(T(1.0), s) : s; }

    // nb.py:121-143
    static __device__ __forceinline__ void left(T (&G)[9], T ux, T uy, T &r)
    {
        r = A::div(full(sum6(G[0], G[3], G[4], G[2], G[6], G[7])), A::sub(T(1.0), ux));
        const T d = A::sub(G[3], G[4]);
        const T k1 = A::mul(T(c1), r), k2 = A::mul(T(c2), r), k3 = A::mul(T(c3), r);
        G[1] = A::add(G[2], A::mul(k1, ux));
        G[5] = A::add(A::add(A::sub(G[6], A::mul(T(c3), d)), A::mul(k2, ux)), A::mul(k3, uy));
        G[8] = A::sub(A::add(A::add(G[7], A::mul(T(c3), d)), A::mul(k2, ux)), A::mul(k3, uy));
    }
    // nb.py:147-169 (velocity) and nb.py:173-195 (pressure: rho given, ux solved)
    static __device__ __forceinline__ void right(T (&G)[9], T &ux, T uy, T &r, bool pressure)
    {
        const T s = full(sum6(G[0], G[3], G[4], G[1], G[5], G[8]));
        if (pressure) ux = A::sub(A::div(s, r), T(1.0));
        else          r = A::div(s, A::add(T(1.0), ux));
        const T d = A::sub(G[3], G[4]);
        const T k1 = A::mul(T(c1), r), k2 = A::mul(T(c2), r), k3 = A::mul(T(c3), r);
        G[2] = A::sub(G[1], A::mul(k1, ux));
        G[6] = A::sub(A::sub(A::add(G[5], A::mul(T(c3), d)), A::mul(k2, ux)), A::mul(k3, uy));
        G[7] = A::add(A::sub(A::sub(G[8], A::mul(T(c3), d)), A::mul(k2, ux)), A::mul(k3, uy));
    }
    // nb.py:199-221
    // dr = rho - 1 without cancellation (deviation storage; used by the corner cells)
    static __device__ __forceinline__ T top_rho(T g0, T g1, T g2, T g3, T g5, T g7, T uy, T &dr)
    {
        const T s = sum6(g0, g1, g2, g3, g5, g7);
        if (Stored<T>::dev) dr = A::div(A::sub(s, uy), A::add(T(1.0), uy));
        else dr = T(0.0);
        return A::div(full(s), A::add(T(1.0), uy));
    }
    static __device__ __forceinline__ void top(T (&G)[9], T ux, T uy, T &r)
    {
        T dr;
        r = top_rho(G[0], G[1], G[2], G[3], G[5], G[7], uy, dr);
        const T d = A::sub(G[1], G[2]);
        const T k1 = A::mul(T(c1), r), k2 = A::mul(T(c2), r), k3 = A::mul(T(c3), r);
        G[4] = A::sub(G[3], A::mul(k1, uy));
        G[8] = A::sub(A::add(A::sub(G[7], A::mul(T(c3), d)), A::mul(k3, ux)), A::mul(k2, uy));
        G[6] = A::sub(A::sub(A::add(G[5], A::mul(T(c3), d)), A::mul(k3, ux)), A::mul(k2, uy));
    }
    // nb.py:225-247
    static __device__ __forceinline__ T bottom_rho(T g0, T g1, T g2, T g4, T g6, T g8, T uy, T &dr)
    {
        const T s = sum6(g0, g1, g2, g4, g6, g8);
        if (Stored<T>::dev) dr = A::div(A::add(s, uy), A::sub(T(1.0), uy));
        else dr = T(0.0);
        return A::div(full(s), A::sub(T(1.0), uy));
    }
    static __device__ __forceinline__ void bottom(T (&G)[9], T ux, T uy, T &r)
    {
        T dr;
        r = bottom_rho(G[0], G[1], G[2], G[4], G[6], G[8], uy, dr);
        const T d = A::sub(G[1], G[2]);
        const T k1 = A::mul(T(c1), r), k2 = A::mul(T(c2), r), k3 = A::mul(T(c3), r);
        G[3] = A::add(G[4], A::mul(k1, uy));
        G[5] = A::add(A::add(A::sub(G[6], A::mul(T(c3), d)), A::mul(k3, ux)), A::mul(k2, uy));
        G[7] = A::add(A::sub(A::add(G[8], A::mul(T(c3), d)), A::mul(k3, ux)), A::mul(k2, uy));
    }
    // corners nb.py:251-344; (r, ux, uy) copied from the x-neighbour on the horizontal wall.
    // Deviation storage: a population set to 0.0 is h = -w; g0 = rho - sum(others) becomes
    // h0 = (rho - 1) - sum(h others).
    static __device__ __forceinline__ void corner(T (&G)[9], bool is_left, bool is_bottom, T r, T dr, T ux, T uy)
    {
        const T zero = Stored<T>::dev ? T(-1.0 / 36.0) : T(0.0);
        const T k23 = A::mul(T(2.0 / 3.0), r), k16 = A::mul(T(1.0 / 6.0), r);
        if (is_left) G[1] = A::add(G[2], A::mul(k23, ux));
        else         G[2] = A::sub(G[1], A::mul(k23, ux));
        if (is_bottom) G[3] = A::add(G[4], A::mul(k23, uy));
        else           G[4] = A::sub(G[3], A::mul(k23, uy));
        if (is_left && is_bottom) {
            G[5] = A::add(A::add(G[6], A::mul(k16, ux)), A::mul(k16, uy));
            G[7] = zero; G[8] = zero;
        } else if (is_left) {
            G[8] = A::sub(A::add(G[7], A::mul(k16, ux)), A::mul(k16, uy));
            G[5] = zero; G[6] = zero;
        } else if (!is_bottom) {
            G[6] = A::sub(A::sub(G[5], A::mul(k16, ux)), A::mul(k16, uy));
            G[7] = zero; G[8] = zero;
        } else {
            G[7] = A::add(A::sub(G[8], A::mul(k16, ux)), A::mul(k16, uy));
            G[5] = zero; G[6] = zero;
        }
        T g0 = A::sub(Stored<T>::dev ? dr : r, G[1]);
#pragma unroll
        for (int q = 2; q < 9; q++) g0 = A::sub(g0, G[q]);
        G[0] = g0;
    }
};

}  // namespace lbm
