// kernels.cuh -- sm_100a kernels of the D2Q9 time step.
//
//   step_kernel   one lattice update per launch: pull-stream from the post-collision array F,
//                 obstacle (I)BB (link blocks), Zou-He walls/corners, macro, equilibrium, TRT, store.
//   step2_kernel  TWO lattice updates per launch (temporal blocking): a thread block computes the
//                 first update on its tile plus a one-cell rim into shared memory and the second
//                 update from shared memory, so the populations cross HBM once per two updates.
//                 Same per-cell device functions, hence bit-identical to two step_kernel launches.
//   stepw_kernel  up to FOUR updates per launch (wavefront temporal blocking): a block sweeps a strip
//                 of rows along x, the updates form a pipeline of thread groups through rings of
//                 columns in shared memory, the source streams in through TMA tensor copies issued
//                 by a producer warp.  Bit-identical again; the default on obstacle-free lattices.
//
// Indexing: every population plane is addressed as  base_q[idx]  with a 32-bit cell index
// idx = x*pitch + y and per-plane base pointers that live in the kernel parameters (constant
// bank); the pull shift -c_q is folded into the base pointer on the host.  One IMAD.WIDE per
// access instead of 64-bit index arithmetic per population.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "d2q9.cuh"

namespace lbm {

enum Mode { kFused = 0, kCollideOnly = 1, kStreamOnly = 2 };
constexpr int kBlock = 256;
constexpr int kHalo = 4;       // halo columns on each side of a slab (= deepest temporal blocking)

template <typename T> struct StepParams {
    const T *pull[9];       // pull[q][x*pitch + y] == F_q(x - cx_q, y - cy_q)   (local x, may be -1 .. nxl)
    const T *ctr[9];        // ctr[q][x*pitch + y]  == F_q(x, y)
    T *dst[9];              // dst[q][x*pitch + y]  == F'_q(x, y)
    int pitch;              // elements between x columns
    int nxl, ny;            // local slab width, height
    int xa, xb;             // local columns processed [xa, xb)
    int x_wl, x_wr;         // local column of the global left / right wall (out of range if not in this slab)
    int x_lo, x_hi;         // local columns that exist in the global lattice: [x_lo, x_hi) within [-2, nxl+2)
    int gx0, gnx;           // global column of local x = 0, global width
    Coef<T> coef;
    const T *walls;         // wall row of this update: u_left[2][ny] u_right[2][ny] u_top[2][gnx] u_bot[2][gnx] rho_right[ny]
    const T *walls2;        // step2_kernel: wall row of the second update
    T *rho_out, *u_out, *uy_out;  // optional macro output (pitched [nxl][pitch]; u_out = x component, uy_out = y)
    const unsigned char *mask;  // optional: nonzero = cell is handled by the link blocks
    int right_pressure;
    int write_macro;
    int pf_ahead;           // step2_kernel: L2 prefetch distance in blocks (0 = off)
    const T *wrow[4];       // stepw_kernel: wall rows of the (up to four) updates of one launch
    int chunk;              // stepw_kernel: columns swept by one block
    // Ramp: the velocity entries of a wall row are multiplied by *scale (device scalar, 1 when no ramp
    // table is set): the apps' inlet ramp  u_wall(it) = ret(it) * base profile  (cavity.py:70-73,
    // turek.py:99-104) costs 8 bytes of host->device traffic per update instead of a whole row.
    const T *scale;         // of `walls`
    const T *scale2;        // of `walls2`
    const T *wscale[4];     // of `wrow[k]`
    // stepw_kernel, slab runs: the last stage stores its first / last kHalo columns also into the halo
    // columns of the left / right neighbour's destination buffer (peer memory over NVLink).
    // peer_X + q*peer_plane_X + xc*pitch + y  is the neighbour's copy of my cell (q, xc, y).
    T *peer_l, *peer_r;
    long long peer_plane_l, peer_plane_r;
    int peer_nxl_l;         // width of the left neighbour's slab (its halo column of my column xc is peer_nxl_l + xc)
    // non-uniform chunks: blockIdx.y < n_main sweeps `chunk` columns, later blocks `chunk_tail` columns
    // (short blocks at the end of the launch shorten its tail)
    int n_main, chunk_tail;
    int wave_l2;            // stepw_kernel: columns of L2 prefetch ahead of the TMA ring (0 = off)
};

struct LinkParams {
    int n_cells;            // distinct boundary cells in this slab
    int n_links;
    int n_links_total;      // links of the caller's whole list (slots), this slab's or not
    int n_obs;
    const int *cell_x, *cell_y, *cell_off;   // [n_cells], [n_cells], [n_cells+1]
    const int *grp_cell;                     // [n_groups+1] groups of <= kBlock cells with <= kBlock links
    const int *link_idx;                     // [n_links] x*pitch + y of the link's cell
    int n_groups;
    const int *link_q;                       // [n_links] direction fluid -> solid
    const int *link_kind;                    // 0 plain BB, 1 IBB p<1/2, 2 IBB p>=1/2
    const int *link_slot;                    // position in the caller's concatenated list
    const void *link_c;                      // [n_links][3] coefficients (T)
    const int *obs_off;                      // [n_obs+1] ranges of the caller's list
    double *link_f;                          // [n_links_total][2] per-link momentum exchange
    double *forces;                          // [n_obs][2] output slot
    unsigned int *done;                      // block completion counter
    int n_link_blocks;
    int defer;                               // 1: leave the per-link terms in link_f (a slot of link_fs), no reduction here
};

// ---- population sources ------------------------------------------------------------------
// A source hands out the nine populations ARRIVING at cell (x, y): G_q = F_q((x,y) - c_q).
// Entries with no in-domain source are garbage and are overwritten by the wall code
// (SURVEY.md section 9.3).
// CG: loads that bypass L1 (ld.global.cg) -- for the resident kernel, where the source was written by other blocks of
// the SAME launch (an L1 line or a non-coherent load could be stale); everywhere else the read-only path.
template <bool CG, typename T> __device__ __forceinline__ T load_pop(const T *a) { return CG ? __ldcg(a) : __ldg(a); }
// shift: elements from the buffer p.pull describes to the buffer actually read (the resident kernel alternates between
// the two buffers, which have the same layout; 0 -- and folded away -- everywhere else).
template <typename T, bool CG = false> struct GlobalSource {
    const StepParams<T> &p;
    long long shift = 0;
    __device__ __forceinline__ void operator()(int x, int y, T (&G)[9]) const
    {
        const int idx = x * p.pitch + y;
#pragma unroll
        for (int q = 0; q < 9; q++) G[q] = load_pop<CG>(p.pull[q] + shift + idx);
    }
};

template <typename T, int SP, int NS> struct SharedSource {
    const T *f;             // [9][NS], cell (xr, yr) at xr*SP + yr
    int x0, y0;             // lattice coordinates of (xr, yr) = (0, 0)
    __device__ __forceinline__ void operator()(int x, int y, T (&G)[9]) const
    {
        const int i = (x - x0) * SP + (y - y0);
#pragma unroll
        for (int q = 0; q < 9; q++) G[q] = f[q * NS + i - cx_of(q) * SP - cy_of(q)];
    }
};

// Wall / corner treatment of the streamed populations of cell (x, y).  Returns true when the
// cell is on a wall; then (r, ux, uy) are what the reference writes into rho/u there.
// Corner cells copy rho and u from the x-neighbour on the same horizontal wall
// (nb.py:254-257 and siblings); its Zou-He density is recomputed here from the source.
template <typename A, typename T, typename Src>
__device__ __forceinline__ bool apply_walls(const StepParams<T> &p, const T *walls, const T *scale, const Src &src,
                                            int x, int y, T (&G)[9], T &r, T &ux, T &uy)
{
    const bool L = x == p.x_wl, R = x == p.x_wr, B = y == 0, Tp = y == p.ny - 1;
    if (!(L | R | B | Tp)) return false;
    const T sc = *scale;            // ramp factor of this update (1 without a ramp table: x * 1 is exact)
    const int gx = p.gx0 + x;
    const T *ul = walls, *ur = ul + 2 * p.ny, *ut = ur + 2 * p.ny, *ub = ut + 2 * p.gnx,
            *rr = ub + 2 * p.gnx;
    if ((L | R) & (B | Tp)) {
        const int xn = L ? x + 1 : x - 1;
        const int gxn = L ? gx + 1 : gx - 1;
        const T *uw = B ? ub : ut;
        ux = A::mul(uw[gxn], sc);
        uy = A::mul(uw[p.gnx + gxn], sc);
        T N[9], dr;
        src(xn, B ? 0 : p.ny - 1, N);
        r = B ? ZouHe<A, T>::bottom_rho(N[0], N[1], N[2], N[4], N[6], N[8], uy, dr)
              : ZouHe<A, T>::top_rho(N[0], N[1], N[2], N[3], N[5], N[7], uy, dr);
        ZouHe<A, T>::corner(G, L, B, r, dr, ux, uy);
    } else if (B) {
        ux = A::mul(ub[gx], sc); uy = A::mul(ub[p.gnx + gx], sc);
        ZouHe<A, T>::bottom(G, ux, uy, r);
    } else if (Tp) {
        ux = A::mul(ut[gx], sc); uy = A::mul(ut[p.gnx + gx], sc);
        ZouHe<A, T>::top(G, ux, uy, r);
    } else if (L) {
        ux = A::mul(ul[y], sc); uy = A::mul(ul[p.ny + y], sc);
        ZouHe<A, T>::left(G, ux, uy, r);
    } else {
        ux = A::mul(ur[y], sc); uy = A::mul(ur[p.ny + y], sc);
        r = rr[y];  // only used by the pressure variant; the outlet density is not ramped (turek.py:109)
        ZouHe<A, T>::right(G, ux, uy, r, p.right_pressure != 0);
    }
    return true;
}

// Everything after the populations of a cell are in registers: walls, macro, collision, store.
// (walls / scale: the wall row and ramp factor of this update -- p.walls / p.scale except in the resident kernel;
// sshift / dshift: see GlobalSource, for the source and the destination buffer)
template <typename T, bool STRICT, int MODE, bool CG = false>
__device__ __forceinline__ void finish_cell_w(const StepParams<T> &p, const T *walls, const T *scale, int x, int y, T (&G)[9],
                                              long long sshift = 0, long long dshift = 0)
{
    using A = Ar<T, STRICT>;
    const int idx = x * p.pitch + y;
    if (MODE != kCollideOnly) {
        T r, ux, uy;
        const bool on_wall = apply_walls<A, T>(p, walls, scale, GlobalSource<T, CG>{p, sshift}, x, y, G, r, ux, uy);
        if (MODE == kStreamOnly) {
            if (on_wall && p.rho_out) {
                p.rho_out[idx] = r;
                p.u_out[idx] = ux;
                p.uy_out[idx] = uy;
            }
        }
    }
    if (MODE != kStreamOnly) {
        T r, ux, uy;
        collide_cell<A, T>(G, p.coef, p.write_macro != 0, r, ux, uy);
        if (p.write_macro) {
            p.rho_out[idx] = r;
            p.u_out[idx] = ux;
            p.uy_out[idx] = uy;
        }
    }
#pragma unroll
    for (int q = 0; q < 9; q++) (p.dst[q] + dshift)[idx] = G[q];
}
template <typename T, bool STRICT, int MODE>
__device__ __forceinline__ void finish_cell(const StepParams<T> &p, int x, int y, T (&G)[9])
{
    finish_cell_w<T, STRICT, MODE, false>(p, p.walls, p.scale, x, y, G);
}

// ---------------------------------------------------------------------------------------
// Obstacle links: one thread per distinct boundary cell (the extra blocks of the step kernel)
// ---------------------------------------------------------------------------------------
// Fixed-order sum of per-link terms f[2*k], f[2*k+1] over each obstacle's range [obs_off[o], obs_off[o+1]): warp w sums
// the links a + w*32 + lane + 256 k (strided partial sums, then a shuffle tree), one thread adds the eight warp sums in
// warp order; one barrier per pass of (up to) kObsPass obstacles.  The same code serves the three places where forces
// are summed, so they agree bit for bit.  CG: the terms were written by other blocks of this launch (bypass L1).
template <bool CG>
__device__ __forceinline__ void reduce_obstacles(const double *f, const int *obs_off, int n_obs, double *out, int nthreads)
{
    constexpr int kObsPass = 16;
    __shared__ double red[kObsPass][kBlock / 32][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = nthreads >> 5;
    for (int o0 = 0; o0 < n_obs; o0 += kObsPass) {
        const int no = min(kObsPass, n_obs - o0);
        for (int o = 0; o < no; o++) {
            double fx = 0.0, fy = 0.0;
            for (int k = obs_off[o0 + o] + threadIdx.x; k < obs_off[o0 + o + 1]; k += nthreads) {
                fx += CG ? __ldcg(f + 2 * k) : f[2 * k];
                fy += CG ? __ldcg(f + 2 * k + 1) : f[2 * k + 1];
            }
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) {
                fx += __shfl_xor_sync(0xffffffffu, fx, m);
                fy += __shfl_xor_sync(0xffffffffu, fy, m);
            }
            if (lane == 0) { red[o][warp][0] = fx; red[o][warp][1] = fy; }
        }
        __syncthreads();
        if (threadIdx.x < 2 * no) {
            const int o = threadIdx.x >> 1, c = threadIdx.x & 1;
            double v = red[o][0][c];
            for (int w = 1; w < nwarps; w++) v += red[o][w][c];
            out[2 * (o0 + o) + c] = v;
        }
        __syncthreads();
    }
}

constexpr int kLinkLocal = 1024;     // per-link terms of up to this many links stay in the link block's shared memory

template <typename T, bool STRICT, int MODE, bool FORCE_ONLY>
__device__ void link_block(const StepParams<T> &p, const LinkParams &lp, int block, int nthreads)
{
    using A = Ar<T, STRICT>;
    // Up to kLinkLocal links (every BASELINE config: 234 / 468 / 928) can be handled by ONE block that keeps the
    // per-link momentum-exchange terms in shared memory and reduces them after a block barrier: no fence,
    // atomic and second trip through L2 (the immediate sums of lbm_forces_now / lbm_apply_bc).
    __shared__ double sf[2 * kLinkLocal];
    __shared__ T sval[kBlock];           // bounced-back population of link l0 + i of the current group
    const bool local = !lp.defer && lp.n_link_blocks == 1 && lp.n_links_total <= kLinkLocal;
    if (local) {
        for (int k = threadIdx.x; k < 2 * lp.n_links_total; k += nthreads) sf[k] = 0.0;   // (links of other slabs stay 0)
        __syncthreads();
    }
    // The boundary cells come in GROUPS of at most kBlock cells and kBlock links (lp.grp_cell, built by the host).  Two
    // phases per group, so that the dependent loads of a cell with eight links do not queue up behind each other
    // (these lattices are latency bound): one thread per LINK evaluates the (interpolated) bounce-back value and the
    // momentum-exchange term; after a barrier one thread per CELL puts the values into its pulled populations and
    // finishes the cell (walls, collision, store).
    for (int g = block; g < lp.n_groups; g += lp.n_link_blocks) {
        const int c0 = lp.grp_cell[g], c1 = lp.grp_cell[g + 1];
        const int l0 = lp.cell_off[c0], l1 = lp.cell_off[c1];
        const int c = c0 + threadIdx.x, l = l0 + threadIdx.x;
        int x = 0, y = 0;
        T G[9];
        if (!FORCE_ONLY && c < c1) {
            x = lp.cell_x[c]; y = lp.cell_y[c];
            GlobalSource<T>{p}(x, y, G);
        }
        if (l < l1) {
            const int q = lp.link_q[l], qb = opp(q), kind = lp.link_kind[l], idx = lp.link_idx[l];
            const T *coef = static_cast<const T *>(lp.link_c) + 3 * l;
            const int o1 = kCx[qb] * p.pitch + kCy[qb];            // (im, jm) = (i, j) + c_qbar
            const T *Fq = p.ctr[q] + idx, *Fb = p.ctr[qb] + idx;
            const T a = __ldg(Fq);
            const T c0f = coef[0], c1f = coef[1], c2f = coef[2];
            T val;
            if (kind == 1)        // nb.py:98-100
                val = A::strict ? A::sub(A::add(A::mul(c0f, a), A::mul(c1f, __ldg(Fq + o1))), A::mul(c2f, __ldg(Fq + 2 * o1)))
                                : A::fmad(-c2f, __ldg(Fq + 2 * o1), A::fmad(c1f, __ldg(Fq + o1), A::mulr(c0f, a)));   // (explicit: see collide_fused)
            else if (kind == 2)   // nb.py:102-104
                val = A::strict ? A::add(A::add(A::mul(c0f, a), A::mul(c1f, __ldg(Fb))), A::mul(c2f, __ldg(Fb + o1)))
                                : A::fmad(c2f, __ldg(Fb + o1), A::fmad(c1f, __ldg(Fb), A::mulr(c0f, a)));
            else                  // nb.py:117
                val = a;
            sval[threadIdx.x] = val;
            // nb.py:64-67: (g_up_q + g_qbar) c_q   (deviation storage: both carry -w_q; the constant
            // sum_links 2 w_q c_q is added once, in double, on the host)
            const T g0 = A::add(a, val);
            const int s = lp.link_slot[l];
            double *f = local ? sf : lp.link_f;
            f[2 * s] = (double)A::mul(g0, T(kCx[q]));
            f[2 * s + 1] = (double)A::mul(g0, T(kCy[q]));
        }
        if (!FORCE_ONLY) {
            __syncthreads();
            if (c < c1) {
                for (int k = lp.cell_off[c]; k < lp.cell_off[c + 1]; k++) {     // later links of a cell overwrite earlier ones
                    const int qb = opp(lp.link_q[k]);
                    const T v = sval[k - l0];
#pragma unroll
                    for (int m = 1; m < 9; m++)
                        if (m == qb) G[m] = v;
                }
                finish_cell<T, STRICT, MODE>(p, x, y, G);
            }
            __syncthreads();
        }
    }
    if (lp.defer) return;       // the terms of this update stay in their slot of link_fs: force_reduce_kernel sums them later
    // Reduction of the per-link terms, per obstacle -- by this block if it is the only one, else by the last block to
    // finish (no dependence on block scheduling).
    __shared__ bool last;
    if (!local) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) last = atomicAdd(lp.done, 1u) == (unsigned)lp.n_link_blocks - 1;
        __syncthreads();
        if (!last) return;
        __threadfence();
        reduce_obstacles<true>(lp.link_f, lp.obs_off, lp.n_obs, lp.forces, nthreads);
        if (threadIdx.x == 0) *lp.done = 0;
    } else {
        __syncthreads();
        reduce_obstacles<false>(sf, lp.obs_off, lp.n_obs, lp.forces, nthreads);
    }
}

// The per-link terms of update slots slot0 + blockIdx.x (written by the link blocks of lbm_step's launches) summed per
// obstacle, one block per slot: the reduction is off the critical path of the updates, which matters on the small,
// latency-bound lattices of the reference's own cases (a Turek update took 8.9 us with the sums inside, 2.6 us are
// the bulk alone).
__global__ void __launch_bounds__(kBlock)
force_reduce_kernel(const double *link_fs, int n_links_total, const int *obs_off, int n_obs, double *forces, int slot0)
{
    const size_t slot = (size_t)slot0 + blockIdx.x;
    reduce_obstacles<false>(link_fs + slot * n_links_total * 2, obs_off, n_obs, forces + slot * n_obs * 2, kBlock);
}

// ---------------------------------------------------------------------------------------
// One update per launch.  grid.x = y tiles, grid.y = local columns xa..xb-1 (+ link blocks
// appended along grid.y when obstacles are present); one thread per cell, threadIdx.x along y
// (the contiguous axis) so every population plane is read and written as full 128 B lines.
// ---------------------------------------------------------------------------------------
template <typename T, bool STRICT, int MODE>
__global__ void __launch_bounds__(kBlock)
step_kernel(const __grid_constant__ StepParams<T> p, const __grid_constant__ LinkParams lp)
{
    // Programmatic dependent launch (small, launch-bound lattices: lbm_b200.cu launches with the stream-serialisation
    // attribute): the NEXT update's grid may be scheduled while this one runs, and this grid does not touch memory
    // before the previous one has completed and flushed.  Both instructions are no-ops in a plain launch.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int ncols = p.xb - p.xa;
    if ((int)blockIdx.y >= ncols) {
        if (MODE != kCollideOnly) {
            const int b = ((int)blockIdx.y - ncols) * gridDim.x + blockIdx.x;
            if (b < lp.n_link_blocks) link_block<T, STRICT, MODE, false>(p, lp, b, kBlock);
        }
        return;
    }
    const int y = blockIdx.x * kBlock + threadIdx.x;
    const int x = p.xa + blockIdx.y;
    if (y >= p.ny) return;
    const int idx = x * p.pitch + y;
    if (MODE != kCollideOnly && p.mask && p.mask[idx]) return;
    T G[9];
    if (MODE == kCollideOnly) {
#pragma unroll
        for (int q = 0; q < 9; q++) G[q] = __ldg(p.ctr[q] + idx);
    } else {
        GlobalSource<T>{p}(x, y, G);
    }
    finish_cell<T, STRICT, MODE>(p, x, y, G);
}

// ---------------------------------------------------------------------------------------
// Two updates per launch (temporal blocking).  Tile = TX columns x TY rows; phase 1 computes the
// first update on the tile plus a one-cell rim ((TX+2) x (TY+2) cells, pulled from global memory,
// i.e. from the tile plus a two-cell rim) into shared memory; phase 2 computes the second update
// of the tile from shared memory and stores it.  Rim cells outside the global lattice are skipped:
// nothing reads them that Zou-He does not overwrite.  No obstacles, no macro output on this path.
// ---------------------------------------------------------------------------------------
template <typename T, bool STRICT, int TX, int TY, int NT, int CPT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
step2_kernel(const __grid_constant__ StepParams<T> p)
{
    using A = Ar<T, STRICT>;
    constexpr int SX = TX + 2, SP = TY + 2, NS = SX * SP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *f = reinterpret_cast<T *>(smem_raw);
    const int tx0 = p.xa + blockIdx.y * TX, ty0 = blockIdx.x * TY;
    const int txe = min(tx0 + TX, p.xb);                 // tile columns [tx0, txe)
    const GlobalSource<T> gsrc{p};

    // ---- phase 1: first update on the rimmed tile -> shared memory --------------------------
    // CPT independent cells per thread and iteration: their loads are issued together and the
    // straight-line macro/collide sequences interleave (ILP), which hides the FP64 and memory
    // latency that 2 resident blocks per SM cannot hide by themselves.
    const int r_lo = max(tx0 - 1, p.x_lo), r_hi = min(txe + 1, p.x_hi);
    for (int c0 = threadIdx.x; c0 < NS; c0 += NT * CPT) {
        T G[CPT][9];
        int x[CPT], y[CPT];
        bool ok[CPT];
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int c = c0 + k * NT;
            const int xr = c / SP, yr = c - xr * SP;
            x[k] = tx0 - 1 + xr;
            y[k] = ty0 - 1 + yr;
            ok[k] = c < NS && x[k] >= r_lo && x[k] < r_hi && y[k] >= 0 && y[k] < p.ny;
            if (!ok[k]) { x[k] = r_lo; y[k] = min(max(y[k], 0), p.ny - 1); }   // compute something valid, store nothing
            gsrc(x[k], y[k], G[k]);
        }
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            T r, ux, uy;
            apply_walls<A, T>(p, p.walls, p.scale, gsrc, x[k], y[k], G[k], r, ux, uy);
        }
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            T r, ux, uy;
            collide_cell<A, T>(G[k], p.coef, false, r, ux, uy);
        }
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            if (ok[k]) {
                const int c = c0 + k * NT;
#pragma unroll
                for (int q = 0; q < 9; q++) f[q * NS + c] = G[k][q];
            }
        }
    }
    __syncthreads();

    // ---- L2 prefetch of the source region of the tile that the grid reaches `pf_ahead` blocks
    // later: phase 1 of that block then finds its populations in L2 (~300 cycles) instead of HBM
    // (~800), which is what the 4 resident blocks per SM cannot hide otherwise.
    if (p.pf_ahead > 0) {
        const long long lin = (long long)blockIdx.y * gridDim.x + blockIdx.x + p.pf_ahead;
        const int by = (int)(lin / gridDim.x), bx = (int)(lin - (long long)by * gridDim.x);
        constexpr int PADE = 16 / (int)sizeof(T);                       // keeps the row start 16 B aligned
        constexpr unsigned ROWB = (TY + 2 * PADE) * (unsigned)sizeof(T); // bytes of one source row
        if (by < (int)gridDim.y && threadIdx.x < 9 * (TX + 4)) {        // one bulk prefetch per plane and row
            const int q = threadIdx.x / (TX + 4), xr = threadIdx.x - q * (TX + 4);
            const int x = p.xa + by * TX - 2 + xr, y0 = bx * TY - PADE;
            if (x >= p.x_lo - 1 && x < p.x_hi + 1 && y0 >= 0 && y0 + TY + 2 * PADE <= p.pitch) {
                const T *a = p.ctr[q] + (long long)x * p.pitch + y0;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(ROWB) : "memory");
            }
        }
    }

    // ---- phase 2: second update of the tile from shared memory -> global ----------------------
    const SharedSource<T, SP, NS> ssrc{f, tx0 - 1, ty0 - 1};
    for (int c0 = threadIdx.x; c0 < TX * TY; c0 += NT * CPT) {
        T G[CPT][9];
        int x[CPT], y[CPT];
        bool ok[CPT];
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int c = c0 + k * NT;
            const int xt = c / TY, yt = c - xt * TY;
            x[k] = tx0 + xt;
            y[k] = ty0 + yt;
            ok[k] = c < TX * TY && x[k] < txe && y[k] < p.ny;
            if (!ok[k]) { x[k] = tx0; y[k] = min(y[k], p.ny - 1); }
            ssrc(x[k], y[k], G[k]);
        }
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            T r, ux, uy;
            apply_walls<A, T>(p, p.walls2, p.scale2, ssrc, x[k], y[k], G[k], r, ux, uy);
        }
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            T r, ux, uy;
            collide_cell<A, T>(G[k], p.coef, false, r, ux, uy);
        }
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            if (ok[k]) {
                const int idx = x[k] * p.pitch + y[k];
#pragma unroll
                for (int q = 0; q < 9; q++) p.dst[q][idx] = G[k][q];
            }
        }
    }
}


// ---------------------------------------------------------------------------------------
// D updates per launch: wavefront temporal blocking (stepw_kernel).
//
// A block owns a strip of TYB rows (y, the contiguous axis) and sweeps it along x over a chunk of
// columns.  The D updates form a software pipeline of D stages, one group of TYB threads each:
// at sweep step s stage k computes update k+1 of column  xs0 + s - 2k  from the three columns
// x-1, x, x+1 of the previous level and writes the result into its own ring of 4 columns in shared
// memory (the last stage stores to global memory).  Level 0 -- the source populations -- is
// streamed into an 8-column shared-memory ring by TMA: one 3-D tensor copy (box = rows x 1 column x
// 9 planes, cp.async.bulk.tensor + mbarrier) per column, issued by a producer warp five columns
// ahead of its use, so the compute threads never touch global memory on the load side and never
// compute a global load address; rows and columns outside the allocation are zero-filled by the
// TMA unit.
//
//   * HBM traffic: one read + one write of the populations per D updates (144/D bytes per lattice
//     update in f64), plus the margin rows of each strip (served by L2: neighbouring strips run
//     side by side).
//   * Redundant work: only along y.  A strip loses one row per side and update, so TYB rows yield
//     TO = TYB - 2(D-1) output rows: 64 -> 58 for D = 4 in f64, 10 % (the 8 x 64 tiles of
//     step2_kernel recompute 14.5 % for D = 2).  Along x a chunk recomputes 2(D-1) columns per 512.
//   * One block-wide barrier per sweep step orders the ring traffic: stage k reads columns
//     x-1..x+1 of level k-1 while stage k-1 writes column x+2 = slot (x-2) & 3.
//   * Same per-cell device functions as step_kernel, hence bit-identical results.
//
// Corner cells need the pulled populations of their x-neighbour on the horizontal wall
// (nb.py:254-257).  For the right wall those are columns x-2..x of the previous level, all still in
// the rings (the stage before does not run past the wall).  The LEFT corners would need column x+2,
// which the stage before is writing in the same step, so they are deferred by one step: the two
// corner threads process cell (x_wl, y) together with column x_wl + 1.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *b, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity)
{
    while (!mbar_try_wait(b, parity)) {}
}
// TMA tile copy global -> shared of the box at (c0, c1, c2) of a 3-D tensor map; completion is
// counted in bytes on an mbarrier (out-of-range elements are zero-filled and counted).
__device__ __forceinline__ void tma_load_3d(void *dst, const void *tmap, int c0, int c1, int c2, unsigned long long *b)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(b)) : "memory");
}

// TMA tile copy shared -> global of a dense box; part of the issuing thread's current bulk async-group.
__device__ __forceinline__ void tma_store_3d(const void *tmap, const void *src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(tmap), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// L2 prefetch of a box (no shared memory, no completion tracking): brings a column in from HBM many steps before
// the ring has a free slot for it, so that the later tensor load hits in L2
__device__ __forceinline__ void tma_prefetch_3d(const void *tmap, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

struct alignas(64) TensorMap { unsigned long long opaque[16]; };   // CUtensorMap (cuda.h), built by the host

// Geometry of a strip.  The TMA unit wants the box to start on a 16-byte boundary along the
// contiguous axis (measured: odd start rows trap in f64), i.e. the first loaded row ys - M0 =
// strip * TO - PAD - M0 has to be a multiple of AL = 16 / sizeof(T) rows, and a box row has to fill
// whole 16-byte units.  PAD rows are lost on each side of a strip (PAD >= D - 1).
template <typename T, int TYB, int D> struct Wave {
    static constexpr int AL = 16 / (int)sizeof(T);                       // rows per 16 bytes: 2 (f64), 4 (f32)
    static constexpr int PAD = (D - 1 + AL / 2 - 1) / (AL / 2) * (AL / 2);   // multiple of AL/2 so that TO is a multiple of AL
    static constexpr int M0 = AL - PAD % AL == AL && AL == 2 ? 2 : AL - PAD % AL;   // margin rows: PAD + M0 multiple of AL, M0 >= 1
    static constexpr int TO = TYB - 2 * PAD;            // output rows of a strip
    static constexpr int ROWS = TYB + 2 * M0;           // rows of one plane-column in a ring: t = -M0 .. TYB+M0-1
    static constexpr int COL = 9 * ROWS;                // elements of one ring slot (a column of the strip)
    static constexpr int COL0 = (COL * (int)sizeof(T) + 127) / 128 * 128 / (int)sizeof(T);   // level 0: TMA wants 128-byte aligned slots
    static constexpr int R = 4;                         // slots of the rings between stages
    static constexpr int LAG = 2;                       // columns between consecutive stages
    static constexpr int HDR = 128;                     // mbarriers
    // staging ring of the last stage: NSTG dense slots [9][TO] that TMA tensor stores copy to global memory
    static constexpr int NSTG = 3;
    static constexpr int STG = (9 * TO * (int)sizeof(T) + 127) / 128 * 128 / (int)sizeof(T);
    static __host__ __device__ constexpr size_t stg_off(int R0) { return (HDR + ((size_t)R0 * COL0 + (size_t)(D - 1) * R * COL) * sizeof(T) + 127) / 128 * 128; }
    static __host__ __device__ constexpr size_t smem(int R0) { return stg_off(R0) + (size_t)NSTG * STG * sizeof(T); }
    static_assert(TO > 0 && TO % AL == 0 && (PAD + M0) % AL == 0 && M0 >= 1 && (2 * M0) % AL == 0, "strip geometry");
};

template <typename T, int ROWS> struct RingSource {
    const T *base;          // (slot 0, q 0, row of t = 0)
    int mask;               // slots - 1
    int stride;             // elements per slot
    int ys;                 // lattice y of t = 0
    __device__ __forceinline__ void operator()(int x, int y, T (&G)[9]) const
    {
        const int t = y - ys;
        const T *c0 = base + (x & mask) * stride + t;
        const T *cm = base + ((x - 1) & mask) * stride + t;    // column x-1 feeds the populations with c_x = +1
        const T *cp = base + ((x + 1) & mask) * stride + t;
#pragma unroll
        for (int q = 0; q < 9; q++) {
            const T *c = cx_of(q) > 0 ? cm : (cx_of(q) < 0 ? cp : c0);
            G[q] = c[q * ROWS - cy_of(q)];
        }
    }
};

// A cell on a wall (or one that shares its step with a deferred corner): the general per-cell
// sequence, kept out of line so that the bulk path of stepw_kernel stays a straight run of
// loads, arithmetic and stores (the two would otherwise be merged back into one by the compiler).
template <typename A, typename T, typename Src>
__device__ __noinline__ void wall_cell(const StepParams<T> &p, const T *walls, const T *scale, const Src &src, int xc, int y, T *out)
{
    T G[9], r, ux, uy;
    src(xc, y, G);
    apply_walls<A, T>(p, walls, scale, src, xc, y, G, r, ux, uy);
    collide_cell<A, T>(G, p.coef, false, r, ux, uy);
#pragma unroll
    for (int q = 0; q < 9; q++) out[q] = G[q];
}

template <typename T, bool STRICT, int D, int TYB, int R0, int MINB>
__global__ void __launch_bounds__(D * TYB + 32, MINB)
stepw_kernel(const __grid_constant__ StepParams<T> p, const __grid_constant__ TensorMap tmap,
             const __grid_constant__ TensorMap tmap_st, const __grid_constant__ TensorMap tmap_pl,
             const __grid_constant__ TensorMap tmap_pr)
{
    using A = Ar<T, STRICT>;
    using W = Wave<T, TYB, D>;
    constexpr int M0 = W::M0, ROWS = W::ROWS, COL = W::COL, COL0 = W::COL0, R = W::R, LAG = W::LAG;
    constexpr int PAD = W::PAD, TO = W::TO;
    constexpr int NC = D * TYB, NT = NC + 32;            // compute threads + one producer warp
    static_assert(TO > 0 && (R0 & (R0 - 1)) == 0 && R0 >= 8, "bad wavefront geometry");
    extern __shared__ __align__(1024) unsigned char wave_smem[];
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(wave_smem);
    T *lvl0 = reinterpret_cast<T *>(wave_smem + W::HDR);     // [R0][COL0]         level 0 (TMA ring)
    T *lvl = lvl0 + R0 * COL0;                               // [D-1][R][9][ROWS]  levels 1 .. D-1
    T *stg = reinterpret_cast<T *>(wave_smem + W::stg_off(R0));   // [NSTG][9][TO]     output columns of the last stage
    constexpr int NSTG = W::NSTG, STG = W::STG;

    const int stage = threadIdx.x / TYB, t = threadIdx.x - stage * TYB;
    const int ys = (int)blockIdx.x * TO - PAD, y = ys + t;
    const int by = (int)blockIdx.y;                          // chunk: `chunk` columns, the last blocks of the launch `chunk_tail`
    const int ca = by < p.n_main ? p.xa + by * p.chunk : p.xa + p.n_main * p.chunk + (by - p.n_main) * p.chunk_tail;
    const int cb = min(ca + (by < p.n_main ? p.chunk : p.chunk_tail), p.xb);         // output columns [ca, cb)
    const int xs0 = ca - (D - 1);                            // first column of stage 0
    const int c_first = xs0 - 2, c_last = cb + D - 1;        // level-0 columns that are loaded
    const int nsteps = (cb - ca) + 2 * (D - 1) + LAG * (D - 1);

    if (threadIdx.x == NC) {
        for (int i = 0; i < R0; i++) mbar_init(mbar + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // slab runs: this chunk's share of the kHalo edge columns also goes to the neighbour's halo (peer memory)
    const bool push_l = p.peer_l != nullptr && ca < kHalo, push_r = p.peer_r != nullptr && cb > p.nxl - kHalo;

    // ---- producer warp: one lane streams the level-0 columns in, five columns ahead of stage 0, and sends the
    // columns the last stage has finished to global memory: ONE tensor copy per column in either direction ----
    if (threadIdx.x >= NC) {
        auto load_column = [&](int c) {                      // tensor coordinates: (row, column + halo, plane)
            unsigned long long *b = mbar + (c & (R0 - 1));
            mbar_expect_tx(b, 9u * ROWS * (unsigned)sizeof(T));
            tma_load_3d(lvl0 + (c & (R0 - 1)) * COL0, &tmap, ys - M0, c + kHalo, 0, b);
        };
        // column xc of the last stage sits in staging slot `slot` (dense [9][TO], rows ys+PAD ..): one tensor store into
        // the destination buffer -- rows beyond ny are clipped by the map -- and, for the kHalo edge columns of a
        // slab, one more into the neighbour's halo (NVLink peer store)
        auto store_column = [&](int xc, int slot) {
            const T *src = stg + slot * STG;
            tma_store_3d(&tmap_st, src, ys + PAD, xc + kHalo, 0);
            if (push_l && xc < kHalo) tma_store_3d(&tmap_pl, src, ys + PAD, p.peer_nxl_l + xc + kHalo, 0);
            if (push_r && xc >= p.nxl - kHalo) tma_store_3d(&tmap_pr, src, ys + PAD, xc - p.nxl + kHalo, 0);
            tma_commit();
        };
        const int l2 = p.wave_l2;                            // columns of L2 look-ahead beyond the ring (0 = off)
        if (threadIdx.x == NC) {
            for (int c = c_first; c < c_first + R0 && c <= c_last; c++) load_column(c);
            for (int c = c_first + R0; c < c_first + R0 + l2 && c <= c_last; c++) tma_prefetch_3d(&tmap, ys - M0, c + kHalo, 0);
            for (int i = 0; i < 4 && c_first + i <= c_last; i++) mbar_wait(mbar + ((c_first + i) & (R0 - 1)), 0);   // step 0 reads columns xs0-1 .. xs0+1
        }
        __syncwarp();
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");          // level 0 is ready for step 0
        // The last stage wrote column xl = xl0 + s - 1 into slot (s-1) % NSTG during step s-1.  A LEFT-wall column
        // is completed one step late (its two corner cells are computed together with column x_wl + 1), so it is
        // sent one iteration later, before the column that follows it.
        const int xl0 = xs0 - LAG * (D - 1);
        int slot = NSTG - 1, prev = NSTG - 2;                // slots of steps s-1 and s-2 (s = 0: none)
        for (int s = 0; s < nsteps; s++) {
            if (threadIdx.x == NC) {
                // the ring holds columns x-2 .. x+R0-3 of stage 0's column x = xs0 + s; x-3 was last read in step s-1
                if (s >= 1 && xs0 + s + R0 - 3 <= c_last) load_column(xs0 + s + R0 - 3);
                if (l2 > 0 && s >= 1 && xs0 + s + R0 - 3 + l2 <= c_last) tma_prefetch_3d(&tmap, ys - M0, xs0 + s + R0 - 3 + l2 + kHalo, 0);
                const int xl = xl0 + s - 1;
                if (s >= 2 && xl - 1 == p.x_wl && xl - 1 >= ca && xl - 1 < cb) store_column(xl - 1, prev);
                if (s >= 1 && xl != p.x_wl && xl >= ca && xl < cb) store_column(xl, slot);
                tma_wait_read<1>();                          // every group but the newest has been read out of its slot
                // the column stage 0 reads first in step s+1 (x+2): the compute threads never touch the mbarriers --
                // a try_wait round trip at the head of every step made stage 0 the slowest stage (ncu: it waited
                // least at the block barrier) -- the block barrier that ends step s publishes the data to them
                const int c = xs0 + s + 2;
                if (c <= c_last) mbar_wait(mbar + (c & (R0 - 1)), ((c - c_first) / R0) & 1);
            }
            prev = slot;
            slot = slot == NSTG - 1 ? 0 : slot + 1;
            __syncwarp();                                    // (the barrier instruction wants the whole warp)
            asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        }
        // (the last stage finishes D-1 steps before the sweep ends: nothing is left in the staging ring)
        if (threadIdx.x == NC) tma_wait_read<0>();
        return;
    }

    asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");              // (the producer's: level 0 is ready for step 0)
    // stage k computes columns [ca - (D-1-k), cb + (D-1-k)) that exist in the global lattice
    const int lo = max(ca - (D - 1 - stage), p.x_lo), hi = min(cb + (D - 1 - stage), p.x_hi);
    const bool row_ok = y >= 0 && y < p.ny;
    const bool edge_row = y == 0 || y == p.ny - 1;
    const bool store_row = t >= PAD && t < PAD + TO;
    const T *const in_base = (stage == 0 ? lvl0 : lvl + (stage - 1) * R * COL) + M0;
    const RingSource<T, ROWS> src{in_base, stage == 0 ? R0 - 1 : R - 1, stage == 0 ? COL0 : COL, ys};
    const T *const in_t = in_base + t;
    T *const out = lvl + stage * R * COL + M0 + t;           // own ring (stages 0 .. D-2)
    const T *const walls = p.wrow[stage];
    const T *const wscale = p.wscale[stage];
    const Coef<T> cf = p.coef;
    // columns whose cells see no wall and no deferred corner, for a row that is not a wall row: the bulk path
    const bool inner_row = y > 0 && y < p.ny - 1;
    const int f_lo = inner_row ? max(lo, p.x_wl + 2) : 0x7fffffff;
    const int f_n = inner_row ? max(min(hi, p.x_wr >= 0 ? p.x_wr : 0x7fffffff) - f_lo, 0) : 0;

    // The sweep, specialised by the role of the stage so that the ring geometry is a compile-time
    // constant and the first stage's mbarrier waits / the last stage's global stores are not in the
    // instruction stream of the others: FIRST reads the TMA ring, LAST stores to global memory.
    auto sweep = [&](auto first_c, auto last_c) {
        constexpr bool FIRST = decltype(first_c)::value, LAST = decltype(last_c)::value;
        constexpr int MASK = FIRST ? R0 - 1 : R - 1, STRIDE = FIRST ? COL0 : COL;
        // LAST: the column goes to the staging slot of its step (sl = slot of the current step; a deferred left
        // corner belongs to the column of the step before); the producer warp sends the slot to global memory
        auto store = [&](int xc, int x, int sl, const T (&G)[9]) {
            if (LAST) {
                if (store_row) {
                    const int k = xc == x ? sl : (sl == 0 ? NSTG - 1 : sl - 1);
                    T *o = stg + k * STG + (t - PAD);
#pragma unroll
                    for (int q = 0; q < 9; q++) o[q * TO] = G[q];
                }
            } else {
                T *o = out + (xc & (R - 1)) * COL;
#pragma unroll
                for (int q = 0; q < 9; q++) o[q * ROWS] = G[q];
            }
        };
        // one sweep step on column x: cm / c0 / cp = this thread's row in the ring slots of columns x-1 / x / x+1,
        // o = its row in the output slot of column x (unused by the last stage)
        auto step = [&](int s, int x, const T *cm, const T *c0, const T *cp, T *o, int sl) {
            if ((unsigned)(x - f_lo) < (unsigned)f_n) {
                // bulk cell: pull, collide, store -- nothing else
                T G[9], r, ux, uy;
#pragma unroll
                for (int q = 0; q < 9; q++) {
                    const T *c = cx_of(q) > 0 ? cm : (cx_of(q) < 0 ? cp : c0);
                    G[q] = c[q * ROWS - cy_of(q)];
                }
                collide_cell<A, T>(G, cf, false, r, ux, uy);
                if (LAST) {
                    store(x, x, sl, G);
                } else {
#pragma unroll
                    for (int q = 0; q < 9; q++) o[q * ROWS] = G[q];
                }
            } else if (row_ok && x >= lo && x < hi) {
                // wall cells.  Left corners wait for the next column (see above); then two cells in one step
                int n = 1, xc = x;
                if (edge_row && x == p.x_wl) n = 0;
                if (edge_row && x == p.x_wl + 1 && x - 1 >= lo) n = 2;
                for (; n > 0; n--, xc--) {
                    T G[9];
                    wall_cell<A, T>(p, walls, wscale, src, xc, y, G);
                    store(xc, x, sl, G);
                }
            }
            // the staging slot is read by the async proxy (TMA store) after the barrier
            if (LAST) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        };
        int x = xs0 - LAG * stage, s = 0, sl = 0;            // sl = s % NSTG (staging slot of the step, last stage)
        auto next_slot = [&]() { if (LAST) sl = sl == NSTG - 1 ? 0 : sl + 1; };
        if (!FIRST) {
            // the 4-slot rings repeat every four columns: four steps per loop iteration with the slot
            // addresses kept in registers (no address arithmetic in the bulk path)
            const T *ib[R];
            T *ob[R];
#pragma unroll
            for (int j = 0; j < R; j++) {
                ib[j] = in_t + ((x + j) & (R - 1)) * COL;
                ob[j] = out + ((x + j) & (R - 1)) * COL;
            }
            for (; s + R <= nsteps; s += R, x += R) {
#pragma unroll
                for (int j = 0; j < R; j++) {
                    step(s + j, x + j, ib[(j + R - 1) & (R - 1)], ib[j], ib[(j + 1) & (R - 1)], ob[j], sl);
                    next_slot();
                }
            }
        }
        for (; s < nsteps; s++, x++) {
            step(s, x, in_t + ((x - 1) & MASK) * STRIDE, in_t + (x & MASK) * STRIDE, in_t + ((x + 1) & MASK) * STRIDE,
                 out + (x & (R - 1)) * COL, sl);
            next_slot();
        }
    };
    using Yes = std::integral_constant<bool, true>;
    using No = std::integral_constant<bool, false>;
    if (stage == 0) sweep(Yes{}, No{});                      // D >= 2: the first stage is never the last
    else if (stage == D - 1) sweep(No{}, Yes{});
    else sweep(No{}, No{});
}

template <typename T, bool STRICT>
__global__ void __launch_bounds__(kBlock)
force_kernel(const __grid_constant__ StepParams<T> p, const __grid_constant__ LinkParams lp)
{
    link_block<T, STRICT, kFused, true>(p, lp, blockIdx.x, kBlock);
}

// Macroscopic fields of the streamed + boundary-treated populations along one lattice line
// (what lattice.macro() of the next iteration computes there): axis 0 -> column x = index
// (ny cells), axis 1 -> row y = index (nxl cells).  out = [rho | ux | uy], each n long.
template <typename T, bool STRICT>
__global__ void __launch_bounds__(kBlock)
probe_kernel(const __grid_constant__ StepParams<T> p, int axis, int index, int n, T *out)
{
    using A = Ar<T, STRICT>;
    const int k = blockIdx.x * kBlock + threadIdx.x;
    if (k >= n) return;
    const int x = axis == 0 ? index : k, y = axis == 0 ? k : index;
    T G[9], r, ux, uy;
    const GlobalSource<T> gsrc{p};
    gsrc(x, y, G);
    apply_walls<A, T>(p, p.walls, p.scale, gsrc, x, y, G, r, ux, uy);
    T dr;
    macro<A, T>(G, r, ux, uy, dr);
    out[k] = r;
    out[n + k] = ux;
    out[2 * n + k] = uy;
}

// |u| of the stored macro fields, -1 on masked (solid) cells: what plot_norm (lbm/src/plot/plot.py:9-16)
// computes on the host from lattice.u and lattice.lattice; separately rounded operations, so the
// result equals np.sqrt(u[0]**2 + u[1]**2) bit for bit.
template <typename T>
__global__ void __launch_bounds__(kBlock)
speed_kernel(const T *ux, const T *uy, const unsigned char *solid, T *out, int pitch, int ny)
{
    const int y = blockIdx.x * kBlock + threadIdx.x;
    if (y >= ny) return;
    const long long c = (long long)blockIdx.y * pitch + y;
    using A = Ar<T, true>;
    const T v2 = A::add(A::mul(ux[c], ux[c]), A::mul(uy[c], uy[c]));
    T v = sizeof(T) == 8 ? (T)__dsqrt_rn((double)v2) : (T)__fsqrt_rn((float)v2);
    if (solid && solid[c]) v = T(-1.0);
    out[c] = v;
}

// Wrap-around 64-bit sum of the bit patterns of the owned cells of one population buffer (all nine
// planes): an order-independent fingerprint that is equal for equal arrays -- parity of runs too large
// to compare on the host (32768^2) and of slab runs against single-GPU runs.
template <typename T>
__global__ void __launch_bounds__(kBlock)
checksum_kernel(const T *f, long long plane, int pitch, int nxl, int ny, unsigned long long *out)
{
    unsigned long long acc = 0;
    const long long cells = (long long)nxl * ny;
    for (long long c = (long long)blockIdx.x * kBlock + threadIdx.x; c < cells; c += (long long)gridDim.x * kBlock) {
        const int x = (int)(c / ny), y = (int)(c - (long long)x * ny);
        const long long idx = (long long)x * pitch + y;
#pragma unroll
        for (int q = 0; q < 9; q++) {
            const T v = f[q * plane + idx];
            if (sizeof(T) == 8) acc += (unsigned long long)__double_as_longlong((double)v);
            else acc += (unsigned long long)(unsigned int)__float_as_int((float)v);
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// uniform equilibrium fill (initial state of every reference app: g = w_q rho at u = 0)
template <typename T, bool STRICT>
__global__ void __launch_bounds__(kBlock)
init_kernel(T *dst, long long plane, int pitch, int nxl, int ny, T r, T ux, T uy)
{
    const int y = blockIdx.x * kBlock + threadIdx.x;
    const int x = blockIdx.y;
    if (y >= ny) return;
    T E[9];
    equilibrium<Ar<T, STRICT>, T>(E, r, ux, uy);
#pragma unroll
    for (int q = 0; q < 9; q++)   // device state: deviation from the weight in f32 (d2q9.cuh: Stored)
        dst[q * plane + (long long)x * pitch + y] = Stored<T>::dev ? E[q] - T(weight_of(q)) : E[q];
}

// nb_equilibrium on pitched device fields
template <typename T, bool STRICT>
__global__ void __launch_bounds__(kBlock)
equilibrium_kernel(T *dst, long long plane, int pitch, int nxl, int ny, const T *rho, const T *u)
{
    const int y = blockIdx.x * kBlock + threadIdx.x;
    const int x = blockIdx.y;
    if (y >= ny) return;
    const long long cell = (long long)x * pitch + y;
    T E[9];
    equilibrium<Ar<T, STRICT>, T>(E, rho[cell], u[cell], u[(long long)nxl * pitch + cell]);
#pragma unroll
    for (int q = 0; q < 9; q++) dst[q * plane + cell] = E[q];
}

// ---------------------------------------------------------------------------------------
// Slab runs, one process per GPU: halo exchange through peer memory (NVLink), no host in the loop.
//
//   peer_push_kernel    copies my first / last kHalo columns (all nine planes) into the neighbours' halo
//                       columns -- used after launches that do not store there themselves
//                       (step_kernel, step2_kernel; stepw_kernel's last stage does).
//   peer_signal_kernel  after the launches of one update group: publishes the group's sequence number
//                       in the neighbours' flag words (release at system scope).
//   peer_wait_kernel    before the next launch: spins until both neighbours have published that
//                       number (their stores into my halos are then visible, and they are done reading
//                       the buffer my next launch will store into).  A time-out records an error word
//                       instead of hanging the device.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kBlock)
peer_push_kernel(const T *src, T *peer_l, T *peer_r, long long plane, long long peer_plane_l, long long peer_plane_r,
                 int pitch, int nxl, int ny)
{
    const int y = blockIdx.x * kBlock + threadIdx.x;
    if (y >= ny) return;
    const int q = blockIdx.y / (2 * kHalo), k = blockIdx.y % (2 * kHalo);
    const bool left = k < kHalo;
    const int xc = left ? k : nxl - 2 * kHalo + k;           // my column: 0..kHalo-1 or nxl-kHalo..nxl-1
    T *peer = left ? peer_l : peer_r;
    if (!peer) return;
    const long long idx = (long long)xc * pitch + y;
    peer[q * (left ? peer_plane_l : peer_plane_r) + idx] = src[q * plane + idx];
}

__global__ void peer_signal_kernel(unsigned int *flag_l, unsigned int *flag_r, unsigned int seq)
{
    __threadfence_system();
    if (flag_l) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag_l), "r"(seq) : "memory");
    if (flag_r) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag_r), "r"(seq) : "memory");
}

__global__ void peer_wait_kernel(const unsigned int *flags, int need_l, int need_r, unsigned int seq,
                                 unsigned int *err, unsigned long long timeout_ns)
{
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int side = 0; side < 2; side++) {
        if (!(side == 0 ? need_l : need_r)) continue;
        for (;;) {
            unsigned int v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + side) : "memory");
            if ((int)(v - seq) >= 0) break;
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > timeout_ns) { *err = seq; return; }
            __nanosleep(200);
        }
    }
}

}  // namespace lbm
