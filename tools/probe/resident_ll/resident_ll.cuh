// resident_ll.cuh -- the flag-free "LL" form of stepr_kernel, AS MEASURED AND DROPPED (not part of the library; kept as
// the record behind DESIGN.md section 7 and profiles/r2_resident_jitter.txt).  It was wired into lbm_b200.cu in place of
// lbm_b200/csrc/resident.cuh (host side: column ranges, per-block boundary-cell ranges, "extra" entries for link
// operands and corner inputs in other blocks' columns, cell tags, the LL areas), passed tests/test_gpu_resident.py
// bit for bit, and ran at 4-12 us per update where CUDA-graph replay takes 2.6-7 and the progress-word form 3.0-5.1.
//
// resident.cuh -- stepr_kernel: a whole BATCH of lattice updates in one launch on a small lattice.
//
// The reference's own cases (cavity 200^2, Turek 2D-1/2D-2, the array: 2.9-15.5 MB of populations, BASELINE configs
// 1-4) are launch bound: an update is a few hundred cycles of work, and a kernel boundary per update -- even as a
// CUDA-graph node with programmatic dependent launch -- costs 2.6-7 us.  Here the grid stays resident for the whole
// batch (cooperative launch: every block is on an SM) and neighbouring blocks hand their edge data over through L2.
// Measured on B200 (tools/probe/l2_handshake.cu): an L2 load takes 488 cycles, a flag store seen by a polling load
// 930, with release / acquire fences around it 2000-2700 -- a "data, fence, flag" protocol costs more than the update
// (the first form of this kernel: 3.1 us per update on 200^2).  So there are no flags and no fences:
//
//   * block b owns the columns [col_a[b], col_a[b+1]) in every update -- bulk cells and obstacle boundary cells alike
//     -- and keeps them in the two global population buffers (these lattices live in L2); nobody else reads or
//     writes those columns during the launch;
//   * what a neighbour needs travels in "LL" entries, value and sequence number in ONE store (the NCCL low-latency
//     protocol): an f64 goes as two 64-bit words {low half | seq}, {high half | seq} (each word is single-copy
//     atomic), an f32 as one.  The thread that finishes a cell of the block's first / last column stores the three
//     populations that cross the interface (q = 2,6,7 / 1,5,8) into the neighbour's halo area straight from its
//     registers; the neighbour's thread that pulls them polls the entry until the sequence number is the one of the
//     update it needs -- arrival of the data IS the synchronisation;
//   * operands of interpolated bounce-back links that lie in another block's columns (nb.py:98-104 reads up to two
//     cells away) travel the same way through per-operand "extra" entries (tables built by the host);
//   * entries are rings of four versions: a block can be at most two updates ahead of a block that still reads
//     (data two columns away: lead 2; the end-of-update barrier of the reader bounds the rest), so version k+4 never
//     overwrites version k before it was read; sequence numbers grow monotonically over launches, a prologue
//     exports the launch's initial edge data (so that update 0 reads halos like every other update and never looks
//     into a neighbour's columns);
//   * a poll that does not come true within `timeout_clk` cycles (a bug, or a grid that is not co-resident) sets
//     the abort word and the host's error word; every block leaves instead of hanging the device (lbm_sync reports it).
//
// Same per-cell functions as step_kernel (apply_walls, collide_cell; ibb_value restates link_block's expressions
// operation by operation), hence bit-identical to single updates (tests/test_gpu_resident.py).  The per-link
// momentum-exchange terms (nb.py:64-67) go to the update's slot of link_fs and are summed per obstacle by
// force_reduce_kernel when the forces are fetched, as with lbm_step's other launches.
#pragma once
#include "kernels.cuh"

namespace lbm {

constexpr int kLLVersions = 4;
constexpr unsigned int kTagExtra = 1u << 30;        // cell_tag: the cell's populations are exported as link operands
constexpr unsigned int kTagCell = (1u << 24) - 1;   // cell_tag: boundary-cell number + 1 (0 = bulk cell)

// One LL entry: value + sequence number, written and read with single stores / loads.
template <typename T> struct LL;
template <> struct LL<double> {
    typedef ulonglong2 E;
    static __device__ __forceinline__ void put(E *e, double v, unsigned int seq)
    {
        const unsigned long long b = (unsigned long long)__double_as_longlong(v), s = (unsigned long long)seq << 32;
        asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(e), "l"((b & 0xffffffffull) | s), "l"((b >> 32) | s) : "memory");
    }
    static __device__ __forceinline__ void load(const E *e, E &r)
    {
        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(e) : "memory");
    }
    static __device__ __forceinline__ bool ok(const E &r, unsigned int seq) { return (unsigned int)(r.x >> 32) == seq && (unsigned int)(r.y >> 32) == seq; }
    static __device__ __forceinline__ double value(const E &r) { return __longlong_as_double((long long)((r.x & 0xffffffffull) | (r.y << 32))); }
};
template <> struct LL<float> {
    typedef unsigned long long E;
    static __device__ __forceinline__ void put(E *e, float v, unsigned int seq)
    {
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(e), "l"((unsigned long long)__float_as_uint(v) | ((unsigned long long)seq << 32)) : "memory");
    }
    static __device__ __forceinline__ void load(const E *e, E &r) { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(r) : "l"(e) : "memory"); }
    static __device__ __forceinline__ bool ok(const E &r, unsigned int seq) { return (unsigned int)(r >> 32) == seq; }
    static __device__ __forceinline__ float value(const E &r) { return __uint_as_float((unsigned int)r); }
};

template <typename T> struct ResidentParams {
    int n_updates;
    const int *col_a;                   // [n_blocks + 1] column ranges
    const int *blk_cell;                // [n_blocks + 1] boundary cells of block b: lp.cell_x[blk_cell[b] .. blk_cell[b+1])
    const int *blk_ex;                  // [n_blocks + 1] extras PRODUCED by block b (sorted by block)
    const int *ex_idx, *ex_q;           // [n_extras] producer cell (x * pitch + y) and population
    const int *link_ex1, *link_ex2;     // [n_links] extra that carries the link's second / third operand, -1: own columns
    const int *corner_ex;               // [4 corners: left-bottom, left-top, right-bottom, right-top][9] extra that carries population q
                                        // arriving at the corner's x-neighbour on the horizontal wall, -1: own columns (or not needed)
    const unsigned int *cell_tag;       // [nxl * pitch]
    void *halo;                         // LL entries [n_blocks][2 sides][kLLVersions][3][ny]: side 0 = from the left neighbour (q = 1,5,8)
    void *extra;                        // LL entries [n_extras][kLLVersions]
    unsigned int seq_base;              // sequence numbers of this launch: data of update k (k = -1: prologue) carries seq_base + k + 2
    unsigned int *abort_w;              // device word, zeroed before the launch
    unsigned int *err;                  // host-visible time-out word
    long long timeout_clk;
    // wall row and ramp factor of update k: row = first_row + k * row_stride; with a ramp table the profile row is
    // row % wall_rows and the factor ramp[row] (ramp = table - it0), without one the profile row is `row`, the factor *one
    const T *walls;
    long long row_len, wall_rows, first_row, row_stride;
    const T *ramp, *one;
    double *link_fs;                    // per-link terms of update k at link_fs + (slot0 + k) * fs_stride
    long long fs_stride, slot0;
    long long buf_delta;                // elements from the source buffer of update 0 to its destination buffer
};

// nb.py:98-100 (kind 1), 102-104 (kind 2), 117 (kind 0) with the precomputed coefficients of lbm_set_links
template <typename A, typename T>
__device__ __forceinline__ T ibb_value(int kind, T c0f, T c1f, T c2f, T a, T n1, T n2)
{
    if (kind == 1)
        return A::strict ? A::sub(A::add(A::mul(c0f, a), A::mul(c1f, n1)), A::mul(c2f, n2))
                         : A::fmad(-c2f, n2, A::fmad(c1f, n1, A::mulr(c0f, a)));
    if (kind == 2)
        return A::strict ? A::add(A::add(A::mul(c0f, a), A::mul(c1f, n1)), A::mul(c2f, n2))
                         : A::fmad(c2f, n2, A::fmad(c1f, n1, A::mulr(c0f, a)));
    return a;
}

// What a block knows during one update: its columns, the version of the halo / extra entries it reads (in_*) and
// writes (out_*).
template <typename T> struct ResidentCtx {
    typedef typename LL<T>::E E;
    const StepParams<T> &p;
    const ResidentParams<T> &rp;
    int xa, xb;                         // own columns
    long long sshift, dshift;           // this update's source / destination buffer relative to p.pull / p.dst
    const E *in_l, *in_r;               // [3][ny] entries from the left (q = 1,5,8) / right (q = 2,6,7) neighbour; null at a lattice wall
    E *out_l, *out_r;                   // the left / right neighbour's entries for my first / last column
    unsigned int seq_in, seq_out;
    int ver_in, ver_out;                // ring positions of the extra entries
    int *fail;                          // shared word: a poll timed out

    // Values of N entries (null: none) once they all carry seq_in.  All loads of a round are issued before the first
    // result is looked at -- a poll costs an L2 round trip, so polls must not queue up behind each other.
    // WARP (all 32 lanes call together): when a first round finds entries missing, only ONE lane of the warp keeps
    // polling (its own entries), the others look again once it is through -- a producer's warp writes its entries
    // together, and 1600 warps that all spin on six entries per lane saturate L2 (measured: 12 us per update on 200^2).
    template <int N, bool WARP>
    __device__ __forceinline__ void gather(const E *const (&e)[N], T (&v)[N]) const
    {
        E r[N];
        bool pend = false;
#pragma unroll
        for (int i = 0; i < N; i++)
            if (e[i]) LL<T>::load(e[i], r[i]);
#pragma unroll
        for (int i = 0; i < N; i++)
            if (e[i] && !LL<T>::ok(r[i], seq_in)) pend = true;
        unsigned int any = WARP ? __ballot_sync(0xffffffffu, pend) : (pend ? 1u : 0u);
        if (any) {
            const long long t0 = clock64();
            const int lane = threadIdx.x & 31;
            bool dead = false;
            while (any && !dead) {
                const bool poller = !WARP || lane == __ffs(any) - 1;
                if (poller) {
                    for (unsigned int spins = 1; pend; spins++) {
#pragma unroll
                        for (int i = 0; i < N; i++)
                            if (e[i] && !LL<T>::ok(r[i], seq_in)) LL<T>::load(e[i], r[i]);
                        pend = false;
#pragma unroll
                        for (int i = 0; i < N; i++)
                            if (e[i] && !LL<T>::ok(r[i], seq_in)) pend = true;
                        if (pend && (spins & 255u) == 0 && (clock64() - t0 > rp.timeout_clk || *(volatile unsigned int *)rp.abort_w)) {
                            *fail = 1;
                            dead = true;
                            break;
                        }
                    }
                }
                if (!WARP) break;
                __syncwarp();
                if (pend && !dead) {            // the others: one more look
#pragma unroll
                    for (int i = 0; i < N; i++)
                        if (e[i] && !LL<T>::ok(r[i], seq_in)) LL<T>::load(e[i], r[i]);
                    pend = false;
#pragma unroll
                    for (int i = 0; i < N; i++)
                        if (e[i] && !LL<T>::ok(r[i], seq_in)) pend = true;
                }
                dead = __any_sync(0xffffffffu, dead);
                any = __ballot_sync(0xffffffffu, pend);
            }
        }
#pragma unroll
        for (int i = 0; i < N; i++) v[i] = e[i] ? LL<T>::value(r[i]) : T(0);
    }
    __device__ __forceinline__ const E *extra_entry(int e) const
    {
        return e < 0 ? nullptr : static_cast<const E *>(rp.extra) + (size_t)e * kLLVersions + ver_in;
    }

    // The nine populations arriving at cell (x, y) of my columns: own columns from the source buffer, the neighbours'
    // edge columns from the halo entries.  (Sources outside the lattice: any value, the wall code overwrites it.)
    // Called by all lanes of a warp together (active = false: a lane without a cell).
    __device__ __forceinline__ void operator()(int x, int y, bool active, T (&G)[9]) const
    {
        const int idx = x * p.pitch + y;
        const bool own_l = x > xa, own_r = x < xb - 1;
#pragma unroll
        for (int q = 0; q < 9; q++) {       // own loads first: their latency overlaps the polls
            const bool own = cx_of(q) == 0 || (cx_of(q) > 0 ? own_l : own_r);
            G[q] = T(0);
            if (own && active) G[q] = __ldcg(p.pull[q] + sshift + idx);
        }
        constexpr int qs[6] = {1, 5, 8, 2, 6, 7};
        const E *e[6];
        T v[6];
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const int q = qs[i], ys = y - cy_of(q);
            const E *in = i < 3 ? (own_l ? nullptr : in_l) : (own_r ? nullptr : in_r);
            e[i] = active && in && ys >= 0 && ys < p.ny ? in + (i % 3) * p.ny + ys : nullptr;
        }
        gather<6, true>(e, v);
#pragma unroll
        for (int i = 0; i < 6; i++)
            if (!(i < 3 ? own_l : own_r)) G[qs[i]] = v[i];
    }

    // Hands the populations of cell (x, y) over: F(q) = the cell's post-collision population q.
    template <typename Get>
    __device__ __forceinline__ void export_cell(int x, int y, unsigned int tag, int ex0, int ex1, const Get &F) const
    {
        if (x == xa && out_l) {
            LL<T>::put(out_l + y, F(2), seq_out);
            LL<T>::put(out_l + p.ny + y, F(6), seq_out);
            LL<T>::put(out_l + 2 * p.ny + y, F(7), seq_out);
        }
        if (x == xb - 1 && out_r) {
            LL<T>::put(out_r + y, F(1), seq_out);
            LL<T>::put(out_r + p.ny + y, F(5), seq_out);
            LL<T>::put(out_r + 2 * p.ny + y, F(8), seq_out);
        }
        if (tag & kTagExtra) {
            const int idx = x * p.pitch + y;
            for (int e = ex0; e < ex1; e++)
                if (rp.ex_idx[e] == idx) LL<T>::put(static_cast<E *>(rp.extra) + (size_t)e * kLLVersions + ver_out, F(rp.ex_q[e]), seq_out);
        }
    }
};

// apply_walls' source of the populations arriving at a corner cell's x-neighbour on the horizontal wall (nb.py:254-257: the
// corner copies that cell's rho and u): own columns from the source buffer, the rest through the corner's extras.  Only
// the six populations the neighbour's Zou-He density is built from are fetched.
template <typename T> struct CornerSource {
    typedef typename LL<T>::E E;
    const ResidentCtx<T> &cx;
    __device__ __forceinline__ void operator()(int xn, int yn, T (&N)[9]) const
    {
        const bool bottom = yn == 0, left = cx.xa == 0;
        const int *tab = cx.rp.corner_ex + ((left ? 0 : 2) + (bottom ? 0 : 1)) * 9;
        const int idx = xn * cx.p.pitch + yn;
        constexpr int qb[6] = {0, 1, 2, 4, 6, 8};       // bottom wall: 0,1,2,4,6,8; top wall: 0,1,2,3,5,7
        const E *e[6];
        T v[6];
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const int q = qb[i] - (i >= 3 && !bottom ? 1 : 0), ex = tab[q];
            e[i] = cx.extra_entry(ex);
            v[i] = T(0);
            if (ex < 0) v[i] = __ldcg(cx.p.pull[q] + cx.sshift + idx);
        }
        T w[6];
        cx.template gather<6, false>(e, w);       // (a corner thread is alone in its warp here)
#pragma unroll
        for (int q = 0; q < 9; q++) N[q] = T(0);
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const T val = e[i] ? w[i] : v[i];
            if (i < 3) N[qb[i]] = val;
            else if (bottom) N[qb[i]] = val;
            else N[qb[i] - 1] = val;
        }
    }
};

template <typename T> __device__ __forceinline__ T pick(const T (&G)[9], int q)
{
    T v = G[0];
#pragma unroll
    for (int m = 1; m < 9; m++)
        if (m == q) v = G[m];
    return v;
}

// MINB: resident blocks per SM the register allocation is bounded for (lbm_set_tuning "resident_blocks").
template <typename T, bool STRICT, int MINB>
__global__ void __launch_bounds__(kBlock, MINB)
stepr_kernel(const __grid_constant__ StepParams<T> pa /* source / destination buffer of updates 0, 2, .. of the launch */,
             const __grid_constant__ LinkParams lp, const __grid_constant__ ResidentParams<T> rp)
{
    using A = Ar<T, STRICT>;
    typedef typename LL<T>::E E;
    __shared__ T sval[kBlock];
    __shared__ int s_fail;
    const int b = blockIdx.x, nb = gridDim.x, ny = pa.ny;
    const int xa = rp.col_a[b], xb = rp.col_a[b + 1];
    const int ncells = (xb - xa) * ny;
    // boundary cells and links of this block (contiguous in the lists: they are sorted by column)
    const int c0 = rp.blk_cell[b], c1 = rp.blk_cell[b + 1];
    const int l0 = c0 < c1 ? lp.cell_off[c0] : 0, l1 = c0 < c1 ? lp.cell_off[c1] : 0;
    const int ex0 = rp.blk_ex[b], ex1 = rp.blk_ex[b + 1];
    // this thread's link (registers, loaded once)
    int lq = 0, lkind = 0, lidx = 0, lslot = 0, lo1 = 0, lex1 = -1, lex2 = -1;
    T c0f = T(0), c1f = T(0), c2f = T(0);
    if (l0 + (int)threadIdx.x < l1) {
        const int l = l0 + (int)threadIdx.x;
        lq = lp.link_q[l]; lkind = lp.link_kind[l]; lidx = lp.link_idx[l]; lslot = lp.link_slot[l];
        const int qb = opp(lq);
        lo1 = kCx[qb] * pa.pitch + kCy[qb];                 // (im, jm) = (i, j) + c_qbar
        const T *coef = static_cast<const T *>(lp.link_c) + 3 * l;
        c0f = coef[0]; c1f = coef[1]; c2f = coef[2];
        lex1 = rp.link_ex1[l]; lex2 = rp.link_ex2[l];
    }
    if (threadIdx.x == 0) s_fail = 0;
    const size_t side = (size_t)kLLVersions * 3 * ny;       // entries of one side of one block's halo area
    E *halo = static_cast<E *>(rp.halo);
    __syncthreads();

    for (int k = -1; k < rp.n_updates; k++) {
        ResidentCtx<T> cx{pa, rp, xa, xb, 0, 0, nullptr, nullptr, nullptr, nullptr, 0u, 0u, 0, 0, &s_fail};
        // updates 1, 3, .. run the other way round: same layout, so every address moves by the distance of the buffers
        // (k = -1, the prologue, reads what update 0 reads)
        cx.sshift = (k > 0 && (k & 1)) ? rp.buf_delta : 0;
        cx.dshift = (k > 0 && (k & 1)) ? -rp.buf_delta : 0;
        cx.ver_in = k & 3; cx.ver_out = (k + 1) & 3;
        cx.seq_in = rp.seq_base + (unsigned int)(k + 1); cx.seq_out = rp.seq_base + (unsigned int)(k + 2);
        cx.in_l = b > 0 ? halo + ((size_t)b * 2 + 0) * side + (size_t)cx.ver_in * 3 * ny : nullptr;
        cx.in_r = b < nb - 1 ? halo + ((size_t)b * 2 + 1) * side + (size_t)cx.ver_in * 3 * ny : nullptr;
        cx.out_l = b > 0 ? halo + ((size_t)(b - 1) * 2 + 1) * side + (size_t)cx.ver_out * 3 * ny : nullptr;
        cx.out_r = b < nb - 1 ? halo + ((size_t)(b + 1) * 2 + 0) * side + (size_t)cx.ver_out * 3 * ny : nullptr;

        if (k < 0) {
            // ---- prologue: the launch's initial edge columns and link operands, from the source buffer ----------
            for (int c = threadIdx.x; c < ncells; c += kBlock) {
                const int xr = c / ny, y = c - xr * ny, x = xa + xr, idx = x * pa.pitch + y;
                const unsigned int tag = rp.cell_tag[idx];
                if (x == xa || x == xb - 1 || (tag & kTagExtra))
                    cx.export_cell(x, y, tag, ex0, ex1, [&](int q) { return __ldcg(pa.ctr[q] + idx); });
            }
            continue;
        }
        const long long row = rp.first_row + (long long)k * rp.row_stride;
        const long long prow = rp.ramp ? (rp.wall_rows == 1 ? 0 : row % rp.wall_rows) : row;
        const T *walls = rp.walls + prow * rp.row_len;
        const T *scale = rp.ramp ? rp.ramp + row : rp.one;

        // ---- the block's links: bounced-back value and momentum-exchange term, one thread per link ------------------
        if (l0 < l1) {
            // operands in other blocks' columns (all threads look together, most of them at nothing)
            T x1 = T(0), x2 = T(0);
            {
                const E *e[2] = {cx.extra_entry(lex1), cx.extra_entry(lex2)};
                T v[2];
                cx.template gather<2, true>(e, v);
                x1 = v[0]; x2 = v[1];
            }
            if (lq) {
                const int qb = opp(lq);
                const T *Fq = pa.ctr[lq] + cx.sshift + lidx, *Fb = pa.ctr[qb] + cx.sshift + lidx;
                const T a = __ldcg(Fq);
                T n1 = T(0), n2 = T(0);
                if (lkind == 1) {
                    if (lex1 < 0) n1 = __ldcg(Fq + lo1);
                    if (lex2 < 0) n2 = __ldcg(Fq + 2 * lo1);
                } else if (lkind == 2) {
                    n1 = __ldcg(Fb);
                    if (lex2 < 0) n2 = __ldcg(Fb + lo1);
                }
                if (lex1 >= 0) n1 = x1;
                if (lex2 >= 0) n2 = x2;
                const T val = ibb_value<A, T>(lkind, c0f, c1f, c2f, a, n1, n2);
                sval[threadIdx.x] = val;
                const T g0 = A::add(a, val);                // nb.py:64-67 (deviation storage: see link_block)
                double *f = rp.link_fs + (rp.slot0 + k) * rp.fs_stride;
                f[2 * lslot] = (double)A::mul(g0, T(kCx[lq]));
                f[2 * lslot + 1] = (double)A::mul(g0, T(kCy[lq]));
            }
            __syncthreads();
        }
        // ---- the block's cells, kBlock per pass ---------------------------------------------------------------------
        // (edge columns first: their exports are what the neighbours wait for; the inner columns' work then overlaps the
        // hand-over)
        for (int cw = threadIdx.x & ~31; cw < ncells; cw += kBlock) {       // (warp-uniform trip count: the polls are warp-wide)
            const int c = cw + (threadIdx.x & 31);
            const bool active = c < ncells;
            const int cc = active ? c : 0;
            const int j = cc / ny, y = cc - j * ny, x = j == 0 ? xa : (j == 1 ? xb - 1 : xa + j - 1), idx = x * pa.pitch + y;
            T G[9];
            cx(x, y, active, G);
            if (!active) continue;
            const unsigned int tag = rp.cell_tag[idx];
            if (tag & kTagCell) {                           // a boundary cell: later links overwrite earlier ones
                const int cell = (int)(tag & kTagCell) - 1;
                for (int l = lp.cell_off[cell]; l < lp.cell_off[cell + 1]; l++) {
                    const int qb = opp(lp.link_q[l]);
                    const T v = sval[l - l0];
#pragma unroll
                    for (int m = 1; m < 9; m++)
                        if (m == qb) G[m] = v;
                }
            }
            T r, ux, uy;
            apply_walls<A, T>(pa, walls, scale, CornerSource<T>{cx}, x, y, G, r, ux, uy);
            collide_cell<A, T>(G, pa.coef, false, r, ux, uy);
#pragma unroll
            for (int q = 0; q < 9; q++) (pa.dst[q] + cx.dshift)[idx] = G[q];
            cx.export_cell(x, y, tag, ex0, ex1, [&](int q) { return pick<T>(G, q); });
        }
        // own columns written by other threads of the block are read in the next update; sval is free again
        __syncthreads();
        if (*(volatile int *)&s_fail) {
            if (threadIdx.x == 0) {
                atomicExch(rp.abort_w, 1u);
                *(volatile unsigned int *)rp.err = 0x80000000u | (unsigned int)k;      // (mapped host memory; lbm_sync reports it)
                __threadfence_system();
            }
            return;
        }
    }
}

}  // namespace lbm
