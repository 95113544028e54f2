// resident.cuh -- stepr_kernel: a whole BATCH of lattice updates in one launch on a small lattice.
//
// The reference's own cases (cavity 200^2, Turek 2D-1/2D-2, the array: 2.9-15.5 MB of populations, BASELINE configs
// 1-4) live in L2, and a kernel boundary per update -- even as a CUDA-graph node with programmatic dependent launch
// -- is a large part of the 2.6-7 us an update takes.  Here the grid stays resident for the whole batch (cooperative
// launch: every block is on an SM) and the kernel boundary becomes a NEIGHBOUR hand-shake:
//
//   * block b < n_col_blocks owns the columns [col_a[b], col_a[b+1]) in every update; each remaining block owns one
//     group of obstacle boundary cells (the link groups of lbm_set_links; column blocks skip the masked cells);
//   * the populations stay in the two global buffers and are read with ld.global.cg (strong, served by L2): L2 is
//     the point of coherence, no L1 line or non-coherent load can be stale inside the launch;
//   * after update k a block publishes k+1 in its progress word (bar.sync, then st.release.gpu by one thread); before
//     update k it waits until every block it DEPENDS on has published k (ld.relaxed.gpu polls + fence.acq_rel.gpu, one
//     thread per dependency, then bar.sync).  i depends on j when i reads columns that j writes or the other way round
//     (host side: column blocks read one column beyond their own -- two at the lattice's left / right wall, where a
//     corner cell looks at its neighbour's pulled populations --, link groups two: the interpolated bounce-back
//     stencil, nb.py:98-104).  The relation is symmetric, and with the two buffers alternating that one condition
//     orders both the reads of update k after the neighbours' writes of update k-1 and the writes of update k after
//     the neighbours' reads of update k-1.  There is no grid-wide barrier: a block only waits for its neighbours,
//     distant blocks may be several updates apart;
//   * a wait that does not come true within `timeout_clk` cycles (a bug, or a grid that is not co-resident) sets the
//     abort word and the host's error word, and every block leaves at its next wait instead of hanging the device
//     (lbm_sync reports it).
//
// What it buys, measured (tools/resident_bench.py, tools/probe/l2_handshake.cu on B200): an L2 load takes 488 cycles,
// a relaxed flag store seen by a polling load 930, with the release / acquire fences that make the data visible
// 2000-2700 -- the hand-shake costs as much as the update.  Lattices with obstacle links gain 17-26 % over graph
// replay (their link blocks' chain of dependent loads no longer sits behind a kernel boundary): Turek 2D-1 4.72 -> 3.93
// us per update, 2D-2 6.97 -> 5.13, array 6.17 -> 5.08 (median of 100 launches of 1024 updates, p90 within 0.5 % of
// it: tools/resident_jitter.py); the obstacle-free cavity loses (2.6 -> 3.0), so `resident = auto` uses this kernel
// only where there are links.  A pause between two polls (__nanosleep) changes nothing.  A second form without flags and fences -- value and sequence number
// in ONE store ("LL" entries) for the three populations that cross a block interface, link operands and corner inputs
// through per-operand entries, rings of four versions -- was built, was bit-identical too and was SLOWER (4-12 us): the
// per-entry polls and the extra L2 traffic of 16-byte entries cost more than the fences they replace.  ncu on config 3:
// L2 throughput 39 % of peak, 9.1 warps per issue at a barrier -- latency chains, not bandwidth; the link blocks (two
// dependent phases, 4500-5300 cycles) set the pace.  What is left is fewer trips through L2 per update: populations that
// stay in shared memory.
//
// Same per-cell functions as step_kernel (finish_cell_w; ibb_value restates link_block's expressions operation by
// operation), hence bit-identical to single updates (tests/test_gpu_resident.py).  The per-link momentum-exchange
// terms (nb.py:64-67) go to the update's slot of link_fs and are summed per obstacle by force_reduce_kernel when the
// forces are fetched, as with lbm_step's other launches.
#pragma once
#include "kernels.cuh"

namespace lbm {

constexpr int kProgStride = 8;          // one 32-byte sector per progress word

template <typename T> struct ResidentParams {
    int n_updates;
    int n_col_blocks;                   // blocks [0, n_col_blocks) own columns, block n_col_blocks + g owns link group g
    const int *col_a;                   // [n_col_blocks + 1]
    const int *dep_off, *dep;           // CSR: the blocks that block b depends on
    unsigned int *prog;                 // [n_blocks][kProgStride] progress words, then the abort word (all zeroed before the launch)
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

// One update of a column block's cells: flattened over its columns, kBlock cells per pass; (x0, y0) = this thread's
// first cell.
template <typename T, bool STRICT>
__device__ __forceinline__ void resident_columns(const StepParams<T> &p, const T *walls, const T *scale, int x0, int y0, int x_end,
                                                 long long sshift, long long dshift)
{
    int x = x0, y = y0;
    while (x < x_end) {
        const int idx = x * p.pitch + y;
        // (the mask byte and the populations are loaded TOGETHER: the acquire fence of the wait has emptied L1, a mask load
        // that decides whether to load at all would put a second L2 round trip in front of every pass)
        const unsigned char m = p.mask ? p.mask[idx] : (unsigned char)0;
        T G[9];
        GlobalSource<T, true>{p, sshift}(x, y, G);
        if (!m) finish_cell_w<T, STRICT, kFused, true>(p, walls, scale, x, y, G, sshift, dshift);
        y += kBlock;
        while (y >= p.ny) { y -= p.ny; x++; }
    }
}

// A thread's share of a link group, loaded once per launch: thread i is link l0 + i and boundary cell c0 + i.
template <typename T> struct ResidentLink {
    int q, kind, idx, slot, o1;         // q == 0: no link
    T c0f, c1f, c2f;
    int x, y, k0, k1;                   // [k0, k1) = the cell's links relative to l0; k0 > k1: no cell
};

// One update of a link group: one thread per LINK evaluates the bounced-back value and the momentum-exchange term,
// then one thread per CELL puts the values into its pulled populations and finishes the cell (as link_block does).
template <typename T, bool STRICT>
__device__ __forceinline__ void resident_links(const StepParams<T> &p, const T *walls, const T *scale, double *f,
                                               const ResidentLink<T> &rl, T *sval, const unsigned char *sqb,
                                               long long sshift, long long dshift)
{
    using A = Ar<T, STRICT>;
    T G[9];
    const bool cell = rl.k0 <= rl.k1;
    if (cell) GlobalSource<T, true>{p, sshift}(rl.x, rl.y, G);
    if (rl.q) {
        const int q = rl.q, qb = opp(q);
        const T *Fq = p.ctr[q] + sshift + rl.idx, *Fb = p.ctr[qb] + sshift + rl.idx;
        const T a = __ldcg(Fq);
        T n1 = T(0), n2 = T(0);
        if (rl.kind == 1) { n1 = __ldcg(Fq + rl.o1); n2 = __ldcg(Fq + 2 * rl.o1); }
        else if (rl.kind == 2) { n1 = __ldcg(Fb); n2 = __ldcg(Fb + rl.o1); }
        const T val = ibb_value<A, T>(rl.kind, rl.c0f, rl.c1f, rl.c2f, a, n1, n2);
        sval[threadIdx.x] = val;
        const T g0 = A::add(a, val);            // nb.py:64-67 (deviation storage: see link_block)
        f[2 * rl.slot] = (double)A::mul(g0, T(kCx[q]));
        f[2 * rl.slot + 1] = (double)A::mul(g0, T(kCy[q]));
    }
    __syncthreads();
    if (cell) {
        for (int k = rl.k0; k < rl.k1; k++) {   // later links of a cell overwrite earlier ones
            const int qb = sqb[k];
            const T v = sval[k];
#pragma unroll
            for (int m = 1; m < 9; m++)
                if (m == qb) G[m] = v;
        }
        finish_cell_w<T, STRICT, kFused, true>(p, walls, scale, rl.x, rl.y, G, sshift, dshift);
    }
}

// Wait until every block this one depends on has published `need` completed updates.  Block-uniform result; false = timed
// out or aborted (the caller leaves).
template <typename T>
__device__ __forceinline__ bool resident_wait(const ResidentParams<T> &rp, const unsigned int *my_dep, int d0, int d1,
                                              unsigned int need, unsigned int *abort_w)
{
    bool ok = true;
    for (int d = d0 + (int)threadIdx.x; d < d1 && ok; d += kBlock) {
        const unsigned int *w = d < d0 + kBlock ? my_dep : rp.prog + (size_t)rp.dep[d] * kProgStride;
        const long long t0 = clock64();
        for (unsigned int spins = 1;; spins++) {        // relaxed polls (each one a trip to L2), ONE fence on success
            unsigned int v;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(w) : "memory");
            if (v >= need) break;
            if (((spins & 63u) == 0 || rp.timeout_clk == 0) && (clock64() - t0 > rp.timeout_clk || *(volatile unsigned int *)abort_w)) {
                ok = false;
                break;
            }
        }
    }
    if (my_dep) asm volatile("fence.acq_rel.gpu;" ::: "memory");  // polls + fence = acquire of the neighbours' releases
    if (__syncthreads_and(ok)) return true;
    if (threadIdx.x == 0) {
        atomicExch(abort_w, 1u);
        *(volatile unsigned int *)rp.err = 0x80000000u | need;     // (mapped host memory; lbm_sync reports it)
        __threadfence_system();
    }
    return false;
}

// Every thread's stores of this update are issued (bar.sync) and published by one thread: the release is cumulative over
// what the barrier ordered.
__device__ __forceinline__ void resident_publish(unsigned int *my_prog, unsigned int done)
{
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(my_prog), "r"(done) : "memory");
}

// Wall row, ramp factor and force slot of update k of the launch.
template <typename T>
__device__ __forceinline__ void resident_inputs(const ResidentParams<T> &rp, int k, const T *&walls, const T *&scale, double *&f)
{
    const long long row = rp.first_row + (long long)k * rp.row_stride;
    const long long prow = rp.ramp ? (rp.wall_rows == 1 ? 0 : row % rp.wall_rows) : row;
    walls = rp.walls + prow * rp.row_len;
    scale = rp.ramp ? rp.ramp + row : rp.one;
    f = rp.link_fs + (rp.slot0 + k) * rp.fs_stride;
}

// MINB: resident blocks per SM the register allocation is bounded for (2: 128 registers, 3: 80; lbm_set_tuning
// "resident_blocks").
template <typename T, bool STRICT, int MINB>
__global__ void __launch_bounds__(kBlock, MINB)
stepr_kernel(const __grid_constant__ StepParams<T> pa /* source / destination buffer of updates 0, 2, .. of the launch */,
             const __grid_constant__ LinkParams lp, const __grid_constant__ ResidentParams<T> rp)
{
    __shared__ T sval[kBlock];
    __shared__ unsigned char sqb[kBlock];
    const int b = blockIdx.x;
    unsigned int *abort_w = rp.prog + (size_t)gridDim.x * kProgStride;
    // the progress words this block waits for: thread t polls dependency d0 + t (+ kBlock, ..)
    const int d0 = rp.dep_off[b], d1 = rp.dep_off[b + 1];
    const unsigned int *my_dep = d0 + (int)threadIdx.x < d1 ? rp.prog + (size_t)rp.dep[d0 + threadIdx.x] * kProgStride : nullptr;
    unsigned int *my_prog = rp.prog + (size_t)b * kProgStride;
    // updates 1, 3, .. run the other way round: same layout, so every address moves by the distance of the buffers

    if (b < rp.n_col_blocks) {
        // ---- a column block ----------------------------------------------------------------------
        const int x_end = rp.col_a[b + 1];
        const int x0 = rp.col_a[b] + (int)threadIdx.x / pa.ny, y0 = (int)threadIdx.x % pa.ny;
        for (int k = 0; k < rp.n_updates; k++) {
            const T *walls, *scale;
            double *f;
            resident_inputs<T>(rp, k, walls, scale, f);     // (before the wait: off the critical path)
            if (k > 0 && !resident_wait<T>(rp, my_dep, d0, d1, (unsigned int)k, abort_w)) return;   // (update 0 follows the previous launch in stream order)
            const long long sshift = (k & 1) ? rp.buf_delta : 0, dshift = (k & 1) ? -rp.buf_delta : 0;
            resident_columns<T, STRICT>(pa, walls, scale, x0, y0, x_end, sshift, dshift);
            resident_publish(my_prog, (unsigned int)(k + 1));
        }
    } else {
        // ---- a link group ------------------------------------------------------------------------
        ResidentLink<T> rl;
        rl.q = 0; rl.kind = 0; rl.idx = 0; rl.slot = 0; rl.o1 = 0;
        rl.c0f = rl.c1f = rl.c2f = T(0);
        rl.x = rl.y = 0; rl.k0 = 1; rl.k1 = 0;
        const int g = b - rp.n_col_blocks;
        const int c0 = lp.grp_cell[g], c1 = lp.grp_cell[g + 1];
        const int l0 = lp.cell_off[c0], l1 = lp.cell_off[c1];
        const int c = c0 + (int)threadIdx.x, l = l0 + (int)threadIdx.x;
        if (l < l1) {
            rl.q = lp.link_q[l]; rl.kind = lp.link_kind[l]; rl.idx = lp.link_idx[l]; rl.slot = lp.link_slot[l];
            const int qb = opp(rl.q);
            rl.o1 = kCx[qb] * pa.pitch + kCy[qb];               // (im, jm) = (i, j) + c_qbar
            const T *coef = static_cast<const T *>(lp.link_c) + 3 * l;
            rl.c0f = coef[0]; rl.c1f = coef[1]; rl.c2f = coef[2];
            sqb[threadIdx.x] = (unsigned char)qb;
        }
        if (c < c1) {
            rl.x = lp.cell_x[c]; rl.y = lp.cell_y[c];
            rl.k0 = lp.cell_off[c] - l0; rl.k1 = lp.cell_off[c + 1] - l0;
        }
        __syncthreads();
        for (int k = 0; k < rp.n_updates; k++) {
            const T *walls, *scale;
            double *f;
            resident_inputs<T>(rp, k, walls, scale, f);
            if (k > 0 && !resident_wait<T>(rp, my_dep, d0, d1, (unsigned int)k, abort_w)) return;
            const long long sshift = (k & 1) ? rp.buf_delta : 0, dshift = (k & 1) ? -rp.buf_delta : 0;
            resident_links<T, STRICT>(pa, walls, scale, f, rl, sval, sqb, sshift, dshift);
            resident_publish(my_prog, (unsigned int)(k + 1));       // (its barrier also frees sval for the next update)
        }
    }
}

}  // namespace lbm
