// lbm_b200.cu -- sm_100a kernels and the C ABI (include/lbm_b200.h) of the D2Q9 time step.
//
// One fused kernel per lattice update -- or per two / three / four updates on obstacle-free
// lattices (temporal blocking, kernels.cuh): pull-stream from the post-collision array F of the
// previous update, obstacle (interpolated) bounce-back, Zou-He walls/corners, macro,
// equilibrium, TRT collision, store.  See DESIGN.md for layout and roofline.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <unistd.h>

#include "../../include/lbm_b200.h"

#include "kernels.cuh"
#include "resident.cuh"

// =========================================================================================
// Host side: handle + C ABI
// =========================================================================================
using namespace lbm;

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return fail(e_ == cudaErrorMemoryAllocation ? LBM_E_NOMEM : LBM_E_CUDA, "%s: %s",   \
                        #expr, cudaGetErrorString(e_));                                         \
    } while (0)

enum StateKind { kNone = 0, kHaveG = 1, kHaveF = 2 };
struct lbm_handle;
static void invalidate_graphs(lbm_handle *h);   // captured launches hold pointers and launch shapes

struct lbm_handle {
    lbm_cfg cfg;
    lbm_layout lay;
    size_t esz;
    void *buf[2] = {nullptr, nullptr};
    bool own_buf = false;
    int cur = 0;
    StateKind kind = kNone;
    bool other_has_g = false;     // other buffer holds stream+BC of current F (lbm_apply_bc)
    bool temporal = true;         // several updates per launch (step2_kernel / stepw_kernel) where possible
    int depth = 4;                // most updates per launch (lbm_set_temporal_depth)
    int wave_chunk = 512;         // columns swept by one block of stepw_kernel
    bool wave_auto = true;        // balance the chunk count against the number of resident blocks
    int n_sm = 148;               // SMs of the handle's device
    int wave_rows = 64;           // rows of a strip of stepw_kernel: 64 with two blocks per SM (measured faster), or 128 with one
    bool wave_attr_set[5] = {false, false, false, false, false};
    int wave_occ[5] = {1, 1, 1, 1, 1};   // resident blocks per SM of stepw_kernel<.., D, ..>
    TensorMap tmap[2];            // one 3-D tensor map per population buffer (stepw_kernel's TMA loads)
    TensorMap tmap_st[2];         // ... and one for its TMA stores (box = output rows of a strip x 1 column x 9 planes)
    TensorMap tmap_peer[2][2];    // [side][buffer]: store maps into the neighbours' buffers (peer halo exchange)
    int tmap_peer_rows = 0;
    int tmap_rows = 0;            // strip geometry (rows, depth) the maps were built for (0 = not built)
    int pf_ahead = 2 * 148;       // L2 prefetch distance of step2_kernel in blocks: two blocks per SM (+3.5 % measured at
                                  // 16384^2 f64); lbm_create sets it from the device's SM count
    bool tb_force = false;        // pair updates even on small lattices (tests)
    bool smem_attr_set = false;
    cudaStream_t stream = nullptr;
    // walls
    void *walls = nullptr;        // device table, element type T
    int64_t wall_rows = 0, wall_cap = 0, row_len = 0;
    // ramp: per-update scale of the velocity entries of a wall row (lbm_set_ramp); d_one = device scalar 1
    void *d_ramp = nullptr, *d_one = nullptr;
    int64_t ramp_n = 0, ramp_cap = 0, ramp_it0 = 0;
    bool pdl = true;              // step_kernel: programmatic dependent launch on small lattices
    int wave_l2 = 6;              // stepw_kernel: columns of L2 prefetch ahead of the TMA ring
    int wave_tail = -1;           // stepw_kernel: width of the short chunks at the end of a launch (-1 = auto, 0 = uniform chunks)
    // peer halo exchange (slab runs, one process per GPU; lbm_peer_*)
    struct Peer {
        bool attached = false, ipc = false;
        void *buf[2] = {nullptr, nullptr};   // the neighbour's population buffers, mapped here
        unsigned int *flags = nullptr;       // the neighbour's flag words
        int64_t nxl = 0, plane = 0, origin = 0;
    } peer[2];                    // 0 = left neighbour (x0 - 1), 1 = right neighbour
    unsigned int *d_flags = nullptr;   // [0] written by the left neighbour, [1] by the right one
    unsigned int *h_err = nullptr, *d_err = nullptr;   // time-out word of the flag waits: pinned host memory, mapped
    unsigned int peer_seq = 0, peer_waited = 0;
    int64_t peer_timeout_ms = 20000;
    // macro
    void *rho = nullptr, *u = nullptr;
    bool macro_valid = false;
    // links
    int n_obs = 0, n_cells = 0, n_links = 0, n_links_total = 0, n_link_blocks = 0;
    int *d_cell_x = nullptr, *d_cell_y = nullptr, *d_cell_off = nullptr, *d_link_q = nullptr,
        *d_link_kind = nullptr, *d_link_slot = nullptr, *d_obs_off = nullptr, *d_grp_cell = nullptr, *d_link_idx = nullptr;
    int n_groups = 0;
    void *d_link_c = nullptr;
    double *d_link_f = nullptr;
    // deferred force sums (lbm_step batches): per-link terms of every update slot [force_cap][n_links_total][2]; the slots
    // in [force_dirty_lo, force_dirty_hi) have not been summed into d_forces yet (lbm_get_forces does it)
    double *d_link_fs = nullptr;
    int64_t link_fs_cap = 0, force_dirty_lo = 0, force_dirty_hi = 0;
    bool defer_now = false;       // the launches being enqueued leave their sums to force_reduce_kernel
    unsigned int *d_done = nullptr;
    unsigned char *d_mask = nullptr;
    // Obstacle band (multi-update launches on lattices with bodies): a child handle that owns local columns
    // [band_ca, band_cb) of this slab with its own pair of buffers; band_a .. band_b are the columns it delivers
    lbm_handle *band = nullptr;
    bool is_band = false;
    int band_a = 0, band_b = 0, band_ca = 0, band_cb = 0;
    // forces
    double *d_forces = nullptr;
    int64_t force_cap = 0, force_n = 0;
    bool force_skip0 = false;     // slot 0 of the last lbm_step was a collide-only update (no link blocks ran)
    double *d_force_now = nullptr;     // [n_obs][2] scratch of lbm_forces_now
    std::vector<double> force_const;   // f32 deviation storage: sum over an obstacle's owned links of 2 w_q c_q
    void *d_probe = nullptr;
    int64_t probe_cap = 0;
    // CUDA graphs of whole lbm_step batches on small (launch-bound) lattices
    struct StepGraph {
        int64_t n = 0, first_row = 0, stride = 0, ramp_it0 = 0;
        uint32_t flags = 0;
        int cur0 = 0, cur1 = 0;
        bool ramp = false;            // captured with a ramp table (&ramp[row - it0] baked in) or without (&one)
        int64_t launches = 0;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
    };
    std::vector<StepGraph> graphs;
    bool use_graph = true;
    // resident batches on small lattices (stepr_kernel, resident.cuh): block layout and dependency lists, built on first use
    struct Resident {
        bool built = false, usable = false;
        int n_blocks = 0, n_col_blocks = 0;
        int *d_col_a = nullptr, *d_dep_off = nullptr, *d_dep = nullptr;
        unsigned int *d_prog = nullptr;
    } res;
    int resident = -1;                 // lbm_set_tuning("resident", ..): -1 = where it is the faster form (lattices with obstacle
                                       // links, measured: profiles/README.md), 1 = every small lattice, 0 = never
    int resident_blocks = 0;           // blocks per SM: 0 = default (3, the 80-register build), 1 / 2 = the 128-register build
    int64_t resident_timeout_ms = 4000;
    int clock_khz = 0;                 // (cudaDevAttrClockRate is a slow query -- up to 100 ms, measured --: asked once)
    std::vector<int> grp_x0, grp_x1;   // first / last column of each link group
    // accounting
    int64_t launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_valid = false;
};

static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

static void free_resident(lbm_handle *h)
{
    void *ptrs[] = {h->res.d_col_a, h->res.d_dep_off, h->res.d_dep, h->res.d_prog};
    for (void *p : ptrs) if (p) cudaFree(p);
    h->res = lbm_handle::Resident();
}

static void invalidate_graphs(lbm_handle *h)
{
    for (auto &g : h->graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        if (g.graph) cudaGraphDestroy(g.graph);
    }
    h->graphs.clear();
}

static void compute_layout(const lbm_cfg &c, lbm_layout &l)
{
    l.elem_size = c.dtype == LBM_F64 ? 8 : 4;
    l.pitch = round_up(c.ny, 128 / l.elem_size);
    l.halo = kHalo;
    l.plane = (c.nxl + 2 * kHalo) * l.pitch;
    // a guard before and after everything so that y-1 / y+1 pulls at the first and last halo
    // column, and the margin rows of the bulk copies of stepw_kernel, stay inside the allocation
    const int64_t guard = std::max<int64_t>(l.pitch, 256);
    l.origin = guard + kHalo * l.pitch /*halo columns x = -kHalo .. -1*/;
    l.elems = 9 * l.plane + 2 * guard;
}

static int check_cfg(const lbm_cfg *c)
{
    if (!c) return fail(LBM_E_INVALID, "cfg is NULL");
    if (c->nx < 3 || c->ny < 3) return fail(LBM_E_INVALID, "lattice must be at least 3x3 (got %lld x %lld)", (long long)c->nx, (long long)c->ny);
    if (c->ny > (1 << 30) || c->nx > (1LL << 31) - 2) return fail(LBM_E_INVALID, "lattice too large");
    if ((c->nxl + 2 * kHalo + 2) * round_up(c->ny, 32) >= (1LL << 31))
        return fail(LBM_E_UNSUPPORTED, "slab of %lld x %lld cells exceeds the 32-bit cell index of one population plane; use more slabs", (long long)c->nxl, (long long)c->ny);
    if (c->x0 < 0 || c->nxl < 1 || c->x0 + c->nxl > c->nx) return fail(LBM_E_INVALID, "slab [%lld, %lld) outside [0, %lld)", (long long)c->x0, (long long)(c->x0 + c->nxl), (long long)c->nx);
    if (c->nxl < 2 && c->nxl != c->nx) return fail(LBM_E_INVALID, "slab must be at least 2 columns wide");
    if (c->dtype != LBM_F64 && c->dtype != LBM_F32) return fail(LBM_E_INVALID, "dtype must be LBM_F64 or LBM_F32");
    if (c->arith != LBM_ARITH_FUSED && c->arith != LBM_ARITH_STRICT) return fail(LBM_E_INVALID, "bad arith");
    if (c->right_wall != LBM_RIGHT_VELOCITY && c->right_wall != LBM_RIGHT_PRESSURE) return fail(LBM_E_INVALID, "bad right_wall");
    if (!(c->om_p > 0.0) || !(c->om_m > 0.0)) return fail(LBM_E_INVALID, "relaxation rates must be positive");
    return LBM_OK;
}

#define CHECK_H(h) do { if (!(h)) return fail(LBM_E_INVALID, "handle is NULL"); CUDA_TRY(cudaSetDevice((h)->cfg.device)); } while (0)

static void *elem_ptr(const lbm_handle *h, int which, int64_t off)
{
    return static_cast<char *>(h->buf[which]) + (h->lay.origin + off) * h->esz;
}

static int ensure_state(lbm_handle *h)
{
    if (h->buf[0]) return LBM_OK;
    const size_t bytes = (size_t)h->lay.elems * h->esz;
    CUDA_TRY(cudaMalloc(&h->buf[0], bytes));
    CUDA_TRY(cudaMalloc(&h->buf[1], bytes));
    h->own_buf = true;
    CUDA_TRY(cudaMemsetAsync(h->buf[0], 0, bytes, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->buf[1], 0, bytes, h->stream));
    return LBM_OK;
}

static int ensure_macro(lbm_handle *h)
{
    if (h->rho) return LBM_OK;
    invalidate_graphs(h);
    const size_t n = (size_t)h->cfg.nxl * h->lay.pitch * h->esz;
    CUDA_TRY(cudaMalloc(&h->rho, n));
    CUDA_TRY(cudaMalloc(&h->u, 2 * n));
    CUDA_TRY(cudaMemsetAsync(h->rho, 0, n, h->stream));
    CUDA_TRY(cudaMemsetAsync(h->u, 0, 2 * n, h->stream));
    return LBM_OK;
}

static int ensure_forces(lbm_handle *h, int64_t n)
{
    const int nobs = std::max(h->n_obs, 1);
    if (h->d_forces && h->force_cap >= n) return LBM_OK;
    invalidate_graphs(h);
    if (h->d_forces) { CUDA_TRY(cudaStreamSynchronize(h->stream)); CUDA_TRY(cudaFree(h->d_forces)); h->d_forces = nullptr; }
    const int64_t cap = std::max<int64_t>(n, 64);
    CUDA_TRY(cudaMalloc(&h->d_forces, (size_t)cap * nobs * 2 * sizeof(double)));
    CUDA_TRY(cudaMemsetAsync(h->d_forces, 0, (size_t)cap * nobs * 2 * sizeof(double), h->stream));
    h->force_cap = cap;
    if (h->d_link_fs) { CUDA_TRY(cudaFree(h->d_link_fs)); h->d_link_fs = nullptr; h->link_fs_cap = 0; }
    h->force_dirty_lo = h->force_dirty_hi = 0;
    const size_t fs_bytes = (size_t)cap * (size_t)h->n_links_total * 2 * sizeof(double);
    if (h->n_links_total > 0 && !h->is_band && fs_bytes <= ((size_t)128 << 20)) {      // (zeros: links of other slabs)
        CUDA_TRY(cudaMalloc(&h->d_link_fs, fs_bytes));
        CUDA_TRY(cudaMemsetAsync(h->d_link_fs, 0, fs_bytes, h->stream));
        h->link_fs_cap = cap;
    }
    return LBM_OK;
}

static void mark_forces_dirty(lbm_handle *h, int64_t lo, int64_t hi)
{
    if (!h->d_link_fs || hi <= lo) return;
    if (h->force_dirty_hi <= h->force_dirty_lo) { h->force_dirty_lo = lo; h->force_dirty_hi = hi; }
    else { h->force_dirty_lo = std::min(h->force_dirty_lo, lo); h->force_dirty_hi = std::max(h->force_dirty_hi, hi); }
}

static int reduce_dirty_forces(lbm_handle *h)
{
    if (!h->d_link_fs || h->force_dirty_hi <= h->force_dirty_lo) return LBM_OK;
    const int64_t lo = h->force_dirty_lo, hi = std::min(h->force_dirty_hi, h->link_fs_cap);
    if (hi > lo) {
        force_reduce_kernel<<<(unsigned)(hi - lo), kBlock, 0, h->stream>>>(h->d_link_fs, h->n_links_total, h->d_obs_off, h->n_obs,
                                                                            h->d_forces, (int)lo);
        h->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    h->force_dirty_lo = h->force_dirty_hi = 0;
    return LBM_OK;
}

// Time-out word of the device-side waits (peer flags, resident batches): pinned host memory, mapped
static int ensure_err(lbm_handle *h)
{
    if (h->h_err) return LBM_OK;
    CUDA_TRY(cudaHostAlloc((void **)&h->h_err, sizeof(unsigned int), cudaHostAllocMapped));
    *h->h_err = 0;
    CUDA_TRY(cudaHostGetDevicePointer((void **)&h->d_err, h->h_err, 0));
    return LBM_OK;
}

// After a stream synchronisation: did a device-side wait give up?  Every call that hands results to the host checks this --
// data that passed a timed-out wait is garbage and must not be returned as a result.
static int wait_error(const lbm_handle *h)
{
    if (!h->h_err || !*(volatile unsigned int *)h->h_err) return LBM_OK;
    const unsigned int e = *(volatile unsigned int *)h->h_err;
    if (e & 0x80000000u)
        return fail(LBM_E_STATE, "resident batch (stepr_kernel) timed out waiting for a neighbour block at update %u of the launch", e & 0x7fffffffu);
    return fail(LBM_E_STATE, "peer halo exchange timed out waiting for update group %u of a neighbour", e);
}

// Slab runs with peer halos: before a launch reads the halo columns (or stores into a neighbour's), both
// neighbours must have published the sequence number of the last update group (lbm_peer_signal).
static int peer_wait(lbm_handle *h)
{
    if (h->peer_waited == h->peer_seq || !(h->peer[0].attached || h->peer[1].attached)) return LBM_OK;
    peer_wait_kernel<<<1, 1, 0, h->stream>>>(h->d_flags, h->peer[0].attached ? 1 : 0, h->peer[1].attached ? 1 : 0,
                                             h->peer_seq, h->d_err, (unsigned long long)h->peer_timeout_ms * 1000000ull);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    h->peer_waited = h->peer_seq;
    return LBM_OK;
}

static void peer_detach(lbm_handle *h)
{
    for (auto &pe : h->peer) {
        if (pe.attached && pe.ipc) {
            for (void *b : pe.buf) if (b) cudaIpcCloseMemHandle(b);
            if (pe.flags) cudaIpcCloseMemHandle(pe.flags);
        }
        pe = lbm_handle::Peer();
    }
}

// Wall row index r of the API -> profile row of the table and device address of its ramp factor.  Without a
// ramp table: row r, factor 1.  With one (lbm_set_ramp): factor ramp[r - it0], profile row r % wall_rows (one
// base row serves every update).
template <typename T> static const T *wall_row_ptr(const lbm_handle *h, int64_t row)
{
    if (!h->walls) return nullptr;
    const int64_t pr = h->ramp_n > 0 ? row % h->wall_rows : row;
    return static_cast<const T *>(h->walls) + pr * h->row_len;
}
template <typename T> static const T *wall_scale_ptr(const lbm_handle *h, int64_t row)
{
    if (h->ramp_n > 0) return static_cast<const T *>(h->d_ramp) + (row - h->ramp_it0);
    return static_cast<const T *>(h->d_one);
}

template <typename T> static void fill_params(const lbm_handle *h, StepParams<T> &p, LinkParams &lp,
                                               int src, int dst, int xa, int xb, int64_t row, int64_t slot)
{
    const T *s0 = static_cast<const T *>(elem_ptr(h, src, 0));
    T *d0 = static_cast<T *>(elem_ptr(h, dst, 0));
    for (int q = 0; q < 9; q++) {
        p.ctr[q] = s0 + q * h->lay.plane;
        p.pull[q] = p.ctr[q] - cx_of(q) * h->lay.pitch - cy_of(q);
        p.dst[q] = d0 + q * h->lay.plane;
    }
    p.pitch = (int)h->lay.pitch;
    p.nxl = (int)h->cfg.nxl;
    p.ny = (int)h->cfg.ny;
    p.xa = xa;
    p.xb = xb;
    const bool has_l = h->cfg.x0 == 0, has_r = h->cfg.x0 + h->cfg.nxl == h->cfg.nx;
    p.x_wl = has_l ? 0 : -(1 << 30);
    p.x_wr = has_r ? (int)h->cfg.nxl - 1 : -(1 << 30);
    p.x_lo = has_l ? 0 : -1;
    p.x_hi = has_r ? (int)h->cfg.nxl : (int)h->cfg.nxl + 1;
    p.gx0 = (int)h->cfg.x0;
    p.gnx = (int)h->cfg.nx;
    const double op = h->cfg.om_p, om = h->cfg.om_m;
    p.coef.one_m_omp = T(1.0 - op);
    p.coef.om_p = T(op);
    p.coef.a_self = T(1.0 - 0.5 * (op + om));
    p.coef.a_opp = T(0.5 * (op - om));
    p.coef.a_eq = T(0.5 * (op + om));
    p.coef.wp0 = T(op * 4.0 / 9.0); p.coef.wp1 = T(op / 9.0); p.coef.wp5 = T(op / 36.0);
    p.coef.wq1 = T(4.5 * op / 9.0); p.coef.wq5 = T(4.5 * op / 36.0);
    p.coef.cs = T(0.5 * (1.0 - op)); p.coef.cd = T(0.5 * (1.0 - om));
    p.coef.wm1 = T(3.0 * om / 9.0); p.coef.wm5 = T(3.0 * om / 36.0);
    p.walls = wall_row_ptr<T>(h, row);
    p.scale = wall_scale_ptr<T>(h, row);
    p.walls2 = nullptr;
    p.scale2 = p.scale;
    p.rho_out = static_cast<T *>(h->rho);
    p.u_out = static_cast<T *>(h->u);
    p.uy_out = h->u ? static_cast<T *>(h->u) + (size_t)h->cfg.nxl * h->lay.pitch : nullptr;
    p.mask = h->d_mask;
    p.right_pressure = h->cfg.right_wall == LBM_RIGHT_PRESSURE;
    p.write_macro = 0;
    p.pf_ahead = h->pf_ahead;
    for (int k = 0; k < 4; k++) { p.wrow[k] = nullptr; p.wscale[k] = p.scale; }
    p.chunk = h->wave_chunk;
    p.wave_l2 = h->wave_l2;
    p.n_main = 1 << 30;
    p.chunk_tail = h->wave_chunk;
    p.peer_l = p.peer_r = nullptr;
    p.peer_plane_l = p.peer_plane_r = 0;
    p.peer_nxl_l = 0;
    lp.n_cells = h->n_cells;
    lp.n_links = h->n_links;
    lp.n_links_total = h->n_links_total;
    lp.n_obs = h->n_obs;
    lp.cell_x = h->d_cell_x; lp.cell_y = h->d_cell_y; lp.cell_off = h->d_cell_off;
    lp.link_q = h->d_link_q; lp.link_kind = h->d_link_kind; lp.link_slot = h->d_link_slot;
    lp.grp_cell = h->d_grp_cell; lp.link_idx = h->d_link_idx; lp.n_groups = h->n_groups;
    lp.link_c = h->d_link_c;
    lp.obs_off = h->d_obs_off;
    lp.defer = h->defer_now && h->d_link_fs && slot < h->link_fs_cap ? 1 : 0;
    lp.link_f = lp.defer ? h->d_link_fs + (size_t)slot * h->n_links_total * 2 : h->d_link_f;
    lp.forces = h->d_forces ? h->d_forces + slot * std::max(h->n_obs, 1) * 2 : nullptr;
    lp.done = h->d_done;
    lp.n_link_blocks = h->n_link_blocks;
}

// ---- temporal blocking: two updates per launch ------------------------------------------
template <typename T, bool STRICT, int TX, int TY, int NT, int CPT, int MINB>
static int launch_step2_v(lbm_handle *h, int src, int dst, int xa, int xb, int64_t row1, int64_t row2)
{
    StepParams<T> p;
    LinkParams lp;
    fill_params<T>(h, p, lp, src, dst, xa, xb, row1, 0);
    p.walls2 = wall_row_ptr<T>(h, row2);
    p.scale2 = wall_scale_ptr<T>(h, row2);
    // a corner cell reads its x-neighbour's pulled populations from shared memory: the right wall
    // column must not be the first column of its tile -> give the last two columns their own launch
    if (p.x_wr >= xa + 1 && p.x_wr < xb && (p.x_wr - xa) % TX == 0) {
        int rc = launch_step2_v<T, STRICT, TX, TY, NT, CPT, MINB>(h, src, dst, xa, p.x_wr - 1, row1, row2);
        if (rc) return rc;
        return launch_step2_v<T, STRICT, TX, TY, NT, CPT, MINB>(h, src, dst, p.x_wr - 1, xb, row1, row2);
    }
    constexpr size_t smem = 9 * (size_t)(TX + 2) * (TY + 2) * sizeof(T);
    if (!h->smem_attr_set) {   // per handle: the attribute belongs to the handle's device
        CUDA_TRY(cudaFuncSetAttribute(step2_kernel<T, STRICT, TX, TY, NT, CPT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->smem_attr_set = true;
    }
    dim3 grid((unsigned)((h->cfg.ny + TY - 1) / TY), (unsigned)((xb - xa + TX - 1) / TX)), block(NT);
    if (grid.y > 65535) return fail(LBM_E_UNSUPPORTED, "slab too wide for one temporal-blocking launch");
    { int rcw = peer_wait(h); if (rcw) return rcw; }
    step2_kernel<T, STRICT, TX, TY, NT, CPT, MINB><<<grid, block, smem, h->stream>>>(p);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return LBM_OK;
}

template <typename T, bool STRICT>
static int launch_step2_t(lbm_handle *h, int src, int dst, int xa, int xb, int64_t row1, int64_t row2)
{
    // Tile / block shape chosen by measurement on B200 at 16384^2 (profiles/README.md): 8 x 64 cells,
    // 256 threads, 4 blocks per SM (<= 64 registers) beat 16x64, 32x32, 16x32, 12x32, 8x32 tiles,
    // 384/512-thread blocks, two cells per thread, and a persistent cp.async-staged variant.
    return launch_step2_v<T, STRICT, 8, 64, 256, 1, 4>(h, src, dst, xa, xb, row1, row2);
}

static int launch_step2(lbm_handle *h, int src, int dst, int xa, int xb, int64_t row1, int64_t row2)
{
    const bool strict = h->cfg.arith == LBM_ARITH_STRICT;
    if (h->cfg.dtype == LBM_F64)
        return strict ? launch_step2_t<double, true>(h, src, dst, xa, xb, row1, row2)
                      : launch_step2_t<double, false>(h, src, dst, xa, xb, row1, row2);
    return strict ? launch_step2_t<float, true>(h, src, dst, xa, xb, row1, row2)
                  : launch_step2_t<float, false>(h, src, dst, xa, xb, row1, row2);
}



// ---- TMA tensor maps of the population buffers (stepw_kernel) -----------------------------
// 3-D tensor (row y, column x + halo, plane q) over one buffer; box = ROWS rows x 1 column x 9
// planes = one ring slot.  cuTensorMapEncodeTiled comes from the driver through the runtime's
// entry-point query, so the library does not link libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_map(const lbm_handle *h, TensorMap *out, void *base, int64_t rows, int64_t cols, int64_t plane, int box_rows)
{
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) return fail(LBM_E_UNSUPPORTED, "driver has no cuTensorMapEncodeTiled");
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    static_assert(sizeof(TensorMap) == sizeof(CUtensorMap), "TensorMap must mirror CUtensorMap");
    const cuuint64_t dims[3] = {(cuuint64_t)rows, (cuuint64_t)cols, 9};
    const cuuint64_t strides[2] = {(cuuint64_t)h->lay.pitch * h->esz, (cuuint64_t)plane * h->esz};
    const cuuint32_t box[3] = {(cuuint32_t)box_rows, 1, 9}, estr[3] = {1, 1, 1};
    CUresult r = encode(reinterpret_cast<CUtensorMap *>(out),
                        h->cfg.dtype == LBM_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                        3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LBM_E_CUDA, "cuTensorMapEncodeTiled failed with code %d", (int)r);
    return LBM_OK;
}

// Load maps: box = all rows of a ring slot (margins included; rows outside the allocation's pitch are zero-filled).
// Store maps: box = the output rows of a strip, tensor height = ny so that rows beyond the lattice are clipped.
// Tensor column 0 = local column -halo.  The neighbours' maps describe THEIR buffers (peer memory).
static int ensure_tensor_maps(lbm_handle *h, int rows_per_slot, int out_rows, int key /* strip rows and depth */)
{
    if (h->tmap_rows != key) {
        for (int b = 0; b < 2; b++) {
            void *base = elem_ptr(h, b, -(int64_t)kHalo * h->lay.pitch);     // (q = 0, x = -halo, y = 0)
            int rc = encode_map(h, &h->tmap[b], base, h->lay.pitch, h->cfg.nxl + 2 * kHalo, h->lay.plane, rows_per_slot);
            if (!rc) rc = encode_map(h, &h->tmap_st[b], base, h->cfg.ny, h->cfg.nxl + 2 * kHalo, h->lay.plane, out_rows);
            if (rc) return rc;
        }
        h->tmap_rows = key;
        h->tmap_peer_rows = 0;
    }
    if (h->tmap_peer_rows != key) {
        for (int side = 0; side < 2; side++) {
            const lbm_handle::Peer &pe = h->peer[side];
            for (int b = 0; b < 2; b++) {
                if (!pe.attached) { h->tmap_peer[side][b] = h->tmap_st[b]; continue; }   // (never used: a valid placeholder)
                void *base = static_cast<char *>(pe.buf[b]) + (pe.origin - (int64_t)kHalo * h->lay.pitch) * (int64_t)h->esz;
                int rc = encode_map(h, &h->tmap_peer[side][b], base, h->cfg.ny, pe.nxl + 2 * kHalo, pe.plane, out_rows);
                if (rc) return rc;
            }
        }
        h->tmap_peer_rows = key;
    }
    return LBM_OK;
}

// ---- wavefront temporal blocking: D updates per launch ----------------------------------
constexpr int kWaveR0 = 8;

template <typename T, bool STRICT, int D, int kWaveRows, int MINB>
static int launch_stepw_v(lbm_handle *h, int src, int dst, int xa, int xb, const int64_t *rows)
{
    using W = Wave<T, kWaveRows, D>;
    StepParams<T> p;
    LinkParams lp;
    fill_params<T>(h, p, lp, src, dst, xa, xb, rows[0], 0);
    for (int k = 0; k < D; k++) { p.wrow[k] = wall_row_ptr<T>(h, rows[k]); p.wscale[k] = wall_scale_ptr<T>(h, rows[k]); }
    // columns that exist in the global lattice; without a wall the slab continues into the halo
    p.x_lo = h->cfg.x0 == 0 ? 0 : -(1 << 20);
    p.x_hi = h->cfg.x0 + h->cfg.nxl == h->cfg.nx ? (int)h->cfg.nxl : (int)h->cfg.nxl + (1 << 20);
    // Chunks of columns.  wave_chunk (512) is the target: longer chunks amortise the 3(D-1) fill/drain
    // steps of the stage pipeline, shorter ones give more blocks.  Blocks run ~0.3 ms each, so a launch
    // of few waves (slabs at 4-8 GPUs) loses up to one wave at its tail: pick the chunk count near
    // the target that fills the last wave best.  A right-wall corner reads its x-neighbour's pulled
    // populations from the rings, so the chunk that holds the right wall must be at least two wide.
    const int n = xb - xa;
    constexpr int TOc = W::TO;
    constexpr size_t smem = W::smem(kWaveR0);
    auto kern = stepw_kernel<T, STRICT, D, kWaveRows, kWaveR0, MINB>;
    if (!h->wave_attr_set[D]) {   // per handle: the attribute belongs to the handle's device
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int occ = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, D * kWaveRows + 32, smem));
        h->wave_occ[D] = std::max(occ, 1);
        h->wave_attr_set[D] = true;
    }
    const long long strips = (h->cfg.ny + TOc - 1) / TOc, slots = (long long)h->wave_occ[D] * h->n_sm;
    int chunk = std::max(16, h->wave_chunk);
    if (h->wave_auto) {
        // measured on B200 (f64, D = 4, GLUPS at chunk 128 / 256 / 512): 4096^2 93 / 83 / 66, 8192^2 - / 112 / 104,
        // 32768^2 - / 127 / 128: short launches want short blocks
        const long long cells = (long long)n * h->cfg.ny;
        const int target = cells >= (1LL << 28) ? 512 : (cells >= (1LL << 26) ? 256 : 128);
        chunk = target;
        double best = 1e30;
        for (int nc = (n + target - 1) / target; nc <= std::max(1, 2 * n / target); nc++) {   // chunks of target/2 .. target columns
            const int c = (n + nc - 1) / nc;
            if (c < 64) break;
            const long long blocks = strips * ((n + c - 1) / c);
            const double waves = (double)blocks / slots;
            const double cost = std::ceil(waves) / waves * (1.0 + 3.0 * (D - 1) / c);
            if (cost < best - 1e-9) { best = cost; chunk = c; }
        }
        chunk = std::min(chunk, std::max(n, 16));
    }
    while (n > chunk && n % chunk == 1) chunk++;
    // The launch ends when its last block does, and the hardware hands blocks out in grid order (strip
    // fastest, chunk slowest): the last chunks of the slab are made short -- at least one full set of
    // resident blocks of a quarter of the width -- so that the SMs run dry within a short block's time
    // (a 4096-column slab at 8 GPUs is only ~15 sets of resident blocks of 0.15-0.3 ms each).
    int n_main = (n + chunk - 1) / chunk, chunk_tail = chunk, n_chunks = n_main;
    const int tail_w = h->wave_tail < 0 ? (h->wave_auto ? std::max(32, chunk / 4) : 0) : h->wave_tail;
    if (tail_w >= 16 && tail_w < chunk && n >= 4 * chunk) {
        const int ntc = (int)std::max<long long>(1, (slots + strips - 1) / strips);
        const int tail_cols = std::min(ntc * tail_w, n / 4);
        const int nm = std::max(1, (n - tail_cols + chunk / 2) / chunk);
        const int cm = (n - tail_cols + nm - 1) / nm;          // main chunks, all full
        const int rest = n - nm * cm;
        if (rest >= 2) {
            int ct = std::min(tail_w, rest);
            while (rest > ct && rest % ct == 1) ct++;
            chunk = cm; n_main = nm; chunk_tail = ct;
            n_chunks = nm + (rest + ct - 1) / ct;
        }
    }
    p.chunk = chunk;
    p.n_main = n_main;
    p.chunk_tail = chunk_tail;
    // slab runs: the last stage also stores the kHalo edge columns into the neighbours' halos
    for (int side = 0; side < 2; side++) {
        const lbm_handle::Peer &pe = h->peer[side];
        if (!pe.attached) continue;
        T *base = static_cast<T *>(pe.buf[dst]) + pe.origin;
        if (side == 0) { p.peer_l = base + pe.nxl * h->lay.pitch; p.peer_plane_l = pe.plane; p.peer_nxl_l = (int)pe.nxl; }
        else           { p.peer_r = base - h->cfg.nxl * h->lay.pitch; p.peer_plane_r = pe.plane; }
    }
    int rc = peer_wait(h);
    if (rc) return rc;
    rc = ensure_tensor_maps(h, W::ROWS, W::TO, kWaveRows * 8 + D);
    if (rc) return rc;
    constexpr int TO = W::TO;
    dim3 grid((unsigned)((h->cfg.ny + TO - 1) / TO), (unsigned)n_chunks), block(D * kWaveRows + 32);
    if (grid.y > 65535) return fail(LBM_E_UNSUPPORTED, "slab too wide for one wavefront launch");
    kern<<<grid, block, smem, h->stream>>>(p, h->tmap[src], h->tmap_st[dst], h->tmap_peer[0][dst], h->tmap_peer[1][dst]);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return LBM_OK;
}

template <typename T, bool STRICT, int D>
static int launch_stepw_t(lbm_handle *h, int src, int dst, int xa, int xb, const int64_t *rows)
{
    if (h->wave_rows == 64) return launch_stepw_v<T, STRICT, D, 64, 2>(h, src, dst, xa, xb, rows);
    return launch_stepw_v<T, STRICT, D, 128, 1>(h, src, dst, xa, xb, rows);
}

template <typename T, bool STRICT>
static int launch_stepw_d(lbm_handle *h, int src, int dst, int xa, int xb, int depth, const int64_t *rows)
{
    switch (depth) {
    case 2: return launch_stepw_t<T, STRICT, 2>(h, src, dst, xa, xb, rows);
    case 3: return launch_stepw_t<T, STRICT, 3>(h, src, dst, xa, xb, rows);
    default: return launch_stepw_t<T, STRICT, 4>(h, src, dst, xa, xb, rows);
    }
}

static int launch_stepw(lbm_handle *h, int src, int dst, int xa, int xb, int depth, const int64_t *rows)
{
    const bool strict = h->cfg.arith == LBM_ARITH_STRICT;
    if (h->cfg.dtype == LBM_F64)
        return strict ? launch_stepw_d<double, true>(h, src, dst, xa, xb, depth, rows)
                      : launch_stepw_d<double, false>(h, src, dst, xa, xb, depth, rows);
    return strict ? launch_stepw_d<float, true>(h, src, dst, xa, xb, depth, rows)
                  : launch_stepw_d<float, false>(h, src, dst, xa, xb, depth, rows);
}

template <typename T, bool STRICT>
static int launch_step_t(lbm_handle *h, int mode, int src, int dst, int xa, int xb, int64_t row,
                         int64_t slot, bool write_macro, bool with_links = true)
{
    StepParams<T> p;
    LinkParams lp;
    fill_params<T>(h, p, lp, src, dst, xa, xb, row, slot);
    p.write_macro = write_macro ? 1 : 0;
    const int ytiles = (int)((h->cfg.ny + kBlock - 1) / kBlock);
    int extra = 0;
    // grid.y is limited to 65535: wide slabs are split into launches of 32768 columns; the link blocks (they own the
    // boundary cells wherever those lie: the bulk threads skip masked cells) ride with the last of them
    if (xa == 0 && xb == (int)h->cfg.nxl && xb - xa > 32768) {
        for (int a = xa; a < xb; a += 32768) {
            const int b = std::min(xb, a + 32768);
            int rc = launch_step_t<T, STRICT>(h, mode, src, dst, a, b, row, slot, write_macro, b == xb);
            if (rc) return rc;
        }
        return LBM_OK;
    }
    // link blocks ride along only when the whole slab is processed by this call
    const bool whole = (xa == 0 && xb == (int)h->cfg.nxl) || (h->cfg.nxl > 32768 && xb == (int)h->cfg.nxl && (xa % 32768) == 0);
    const bool links_here = mode != kCollideOnly && h->n_link_blocks > 0 && whole && with_links;
    if (lp.defer) lp.n_link_blocks = std::max(1, h->n_groups);   // no reduction in the launch: one block per group of cells
    if (links_here) extra = (lp.n_link_blocks + ytiles - 1) / ytiles;
    else lp.n_link_blocks = 0;
    const bool part_of_split = h->cfg.nxl > 32768 && (xa % 32768) == 0 && (xb == (int)h->cfg.nxl || xb - xa == 32768);
    if (mode != kCollideOnly && h->n_link_blocks > 0 && !links_here && !part_of_split)
        return fail(LBM_E_UNSUPPORTED, "column-restricted updates with obstacles are not supported (the link blocks ride with whole-slab updates)");
    dim3 grid(ytiles, (xb - xa) + extra), block(kBlock);
    if (grid.y > 65535) return fail(LBM_E_UNSUPPORTED, "more than 65535 columns (+ link blocks) in one launch");
    { int rcw = peer_wait(h); if (rcw) return rcw; }
    // small lattices are launch bound: programmatic dependent launch lets the next update's grid be scheduled while this
    // one runs (the kernel waits for its predecessor before it reads anything)
    cudaLaunchConfig_t lc = {};
    lc.gridDim = grid; lc.blockDim = block; lc.dynamicSmemBytes = 0; lc.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    // (measured, us per update with / without: 536 x 100 2.65 / 2.86, 1073 x 200 4.38 / 4.92, 900 x 200 4.02 / 4.53, but
    // 200 x 200 -- 200 blocks, less than two per SM -- 2.97 / 2.61: only where a grid is more than one set of resident blocks)
    const bool pdl = h->pdl && h->cfg.nxl * h->cfg.ny <= (1LL << 21) && (long long)grid.x * grid.y >= 2LL * h->n_sm;
    lc.attrs = at; lc.numAttrs = pdl ? 1 : 0;
    cudaError_t le;
    switch (mode) {
    case kFused: le = cudaLaunchKernelEx(&lc, step_kernel<T, STRICT, kFused>, p, lp); break;
    case kCollideOnly: le = cudaLaunchKernelEx(&lc, step_kernel<T, STRICT, kCollideOnly>, p, lp); break;
    default: le = cudaLaunchKernelEx(&lc, step_kernel<T, STRICT, kStreamOnly>, p, lp); break;
    }
    h->launches++;
    CUDA_TRY(le);
    CUDA_TRY(cudaGetLastError());
    return LBM_OK;
}

static int launch_step(lbm_handle *h, int mode, int src, int dst, int xa, int xb, int64_t row,
                       int64_t slot, bool write_macro)
{
    const bool strict = h->cfg.arith == LBM_ARITH_STRICT;
    if (h->cfg.dtype == LBM_F64)
        return strict ? launch_step_t<double, true>(h, mode, src, dst, xa, xb, row, slot, write_macro)
                      : launch_step_t<double, false>(h, mode, src, dst, xa, xb, row, slot, write_macro);
    return strict ? launch_step_t<float, true>(h, mode, src, dst, xa, xb, row, slot, write_macro)
                  : launch_step_t<float, false>(h, mode, src, dst, xa, xb, row, slot, write_macro);
}

// ---------------------------------------------------------------------------------------
// Obstacles and multi-update launches.  The wavefront kernel keeps three columns of the previous level in its
// rings; an interpolated bounce-back link reads two columns away (nb.py:98-104), and the link cells are few.
// So the bodies get a BAND of columns of their own: all link cells lie in [x_min, x_max]; the band delivers
// columns [a, b) = [x_min - 4, x_max + 5) and is computed by a child handle that owns [a - 4, b + 4) (clipped at
// the lattice walls) with its own pair of buffers and the same link lists: per group of D <= 4 updates its
// columns are copied out of the source buffer, updated D times by step_kernel (link blocks ride along, drag/lift
// sums go to the parent's force slots) -- no halo refresh, so the valid range shrinks by one column per update
// and side and ends as [a, b) -- and copied into the destination buffer, while ONE wavefront launch per side
// covers the obstacle-free columns [0, a) and [b, nxl).  Those launches recompute up to three columns inside
// the band's margin at intermediate levels; the margin of four columns beyond the outermost link cell is what
// keeps every cell they touch out of reach of a link for that many updates.  Same per-cell functions
// everywhere, hence bit-identical to single updates (tests/test_gpu_temporal.py).
static int build_band(lbm_handle *h, int32_t n_obstacles, const int64_t *offsets, const int64_t *ijq, const double *ibb,
                      int32_t use_ibb, int xmin, int xmax)
{
    const int nxl = (int)h->cfg.nxl;
    const bool wall_l = h->cfg.x0 == 0, wall_r = h->cfg.x0 + h->cfg.nxl == h->cfg.nx;
    int a = xmin - kHalo, b = xmax + kHalo + 1;
    if (a < 2) a = 0;                          // (a wavefront launch is at least two columns wide)
    if (b > nxl - 2) b = nxl;
    if ((a == 0 && !wall_l) || (b == nxl && !wall_r)) return LBM_OK;     // band on a slab interface: single updates only
    // peer halos: the kHalo edge columns must come out of the wavefront launches (their last stage stores them into the
    // neighbour's halo), so the band has to stay clear of them
    if ((h->peer[0].attached && a < kHalo) || (h->peer[1].attached && b > nxl - kHalo)) return LBM_OK;
    const int ca = std::max(a - kHalo, 0), cb = std::min(b + kHalo, nxl);
    if (cb - ca < 2 || (a == 0 && b == nxl)) return LBM_OK;
    lbm_cfg c = h->cfg;
    c.x0 = h->cfg.x0 + ca;
    c.nxl = cb - ca;
    lbm_handle *child = nullptr;
    int rc = lbm_create(&c, &child);
    if (rc) return rc;
    child->is_band = true;
    child->stream = h->stream;
    child->temporal = false;
    child->use_graph = false;
    rc = lbm_set_links(child, n_obstacles, offsets, ijq, ibb, use_ibb);
    if (!rc) rc = ensure_state(child);
    if (rc) { lbm_destroy(child); return rc; }
    child->kind = kHaveF;
    h->band = child;
    h->band_a = a; h->band_b = b; h->band_ca = ca; h->band_cb = cb;
    return LBM_OK;
}

static bool band_usable(const lbm_handle *h)
{
    if (!h->band) return false;
    return !(h->peer[0].attached && h->band_a < kHalo) && !(h->peer[1].attached && h->band_b > (int)h->cfg.nxl - kHalo);
}
// obstacle links are set, but none of them belongs to this slab: the slab is updated like an obstacle-free one
static bool links_elsewhere(const lbm_handle *h) { return h->n_obs > 0 && h->n_cells == 0; }

// One group of d = 2..4 updates of the whole slab (src -> dst buffer), bodies in the band; force slots slot0 ..
static int launch_band_group(lbm_handle *h, int src, int dst, int d, const int64_t *rows, int64_t slot0)
{
    lbm_handle *b = h->band;
    const int64_t pitch = h->lay.pitch, esz = (int64_t)h->esz;
    // the child shares the parent's wall table, ramp and force slots
    b->walls = h->walls; b->wall_rows = h->wall_rows; b->wall_cap = 0;
    b->d_ramp = h->d_ramp; b->ramp_n = h->ramp_n; b->ramp_it0 = h->ramp_it0; b->ramp_cap = 0;
    b->d_forces = h->d_forces; b->force_cap = h->force_cap;
    b->d_link_fs = h->d_link_fs; b->link_fs_cap = h->link_fs_cap; b->defer_now = h->defer_now;
    b->cfg.right_wall = h->cfg.right_wall;
    b->stream = h->stream;
    // 1. the band's columns out of the source buffer
    for (int q = 0; q < 9; q++)     // (a column range of one plane is contiguous; planes may be > 2 GB apart: no 2-D copy)
        CUDA_TRY(cudaMemcpyAsync(elem_ptr(b, b->cur, q * b->lay.plane), elem_ptr(h, src, q * h->lay.plane + (int64_t)h->band_ca * pitch),
                                 (size_t)((h->band_cb - h->band_ca) * pitch * esz), cudaMemcpyDeviceToDevice, h->stream));
    // 2. d single updates with the links
    for (int k = 0; k < d; k++) {
        int rc = launch_step(b, kFused, b->cur, b->cur ^ 1, 0, (int)b->cfg.nxl, rows[k], slot0 + k, false);
        if (rc) return rc;
        b->cur ^= 1;
    }
    h->launches += d;
    // 3. the obstacle-free columns on either side: one wavefront launch each
    if (h->band_a > 0) { int rc = launch_stepw(h, src, dst, 0, h->band_a, d, rows); if (rc) return rc; }
    if (h->band_b < (int)h->cfg.nxl) { int rc = launch_stepw(h, src, dst, h->band_b, (int)h->cfg.nxl, d, rows); if (rc) return rc; }
    // 4. the band's share of the result
    for (int q = 0; q < 9; q++)
        CUDA_TRY(cudaMemcpyAsync(elem_ptr(h, dst, q * h->lay.plane + (int64_t)h->band_a * pitch),
                                 elem_ptr(b, b->cur, q * b->lay.plane + (int64_t)(h->band_a - h->band_ca) * pitch),
                                 (size_t)((h->band_b - h->band_a) * pitch * esz), cudaMemcpyDeviceToDevice, h->stream));
    return LBM_OK;
}

// ---------------------------------------------------------------------------------------
// Resident batches (stepr_kernel, resident.cuh): which lattices they are for, the block plan, the launch.
static bool resident_candidate(const lbm_handle *h)
{
    return (h->resident > 0 || (h->resident < 0 && h->n_cells > 0)) && h->temporal && !h->tb_force && !h->is_band && h->stream != nullptr && h->cfg.x0 == 0 && h->cfg.nxl == h->cfg.nx &&
           h->cfg.nxl * h->cfg.ny <= (1LL << 19) && !h->peer[0].attached && !h->peer[1].attached &&
           (h->n_obs == 0 || (h->n_cells > 0 && h->n_groups <= 64));
}

// The plan itself (no device needed; lbm_resident_plan exports it for tests/test_resident_plan_cpu.py, which replays the
// hand-shake with random block timing): column block i owns columns [col_a[i], col_a[i+1]) and reads one column beyond
// them -- two at the lattice's left / right wall: a corner cell takes rho and u from its x-neighbour on the horizontal
// wall, whose pulled populations come from one column further; link group g owns the boundary cells in columns
// [grp_x0[g], grp_x1[g]] and reads two columns beyond (interpolated bounce-back, nb.py:98-104).
struct ResidentPlan {
    int n_col_blocks = 0;
    std::vector<int> col_a, dep_off, dep;
};
static ResidentPlan plan_resident(int nx, int max_blocks, int ng, const int *grp_x0, const int *grp_x1)
{
    ResidentPlan pl;
    const int ncb = std::min(nx, max_blocks - ng), nb = ncb + ng;
    pl.n_col_blocks = ncb;
    pl.col_a.resize(ncb + 1);
    for (int i = 0; i <= ncb; i++) pl.col_a[i] = (int)((int64_t)i * nx / ncb);
    struct Iv { int w0, w1, r0, r1; };
    std::vector<Iv> iv(nb);
    const std::vector<int> &col_a = pl.col_a;
    for (int i = 0; i < ncb; i++)
        iv[i] = {col_a[i], col_a[i + 1] - 1, col_a[i] - (col_a[i + 1] == nx ? 2 : 1), col_a[i + 1] + (col_a[i] == 0 ? 1 : 0)};
    for (int g = 0; g < ng; g++) iv[ncb + g] = {grp_x0[g], grp_x1[g], grp_x0[g] - 2, grp_x1[g] + 2};
    auto meets = [](int a0, int a1, int b0, int b1) { return a0 <= b1 && b0 <= a1; };
    pl.dep_off.assign(nb + 1, 0);
    for (int i = 0; i < nb; i++) {
        for (int j = 0; j < nb; j++)
            if (j != i && (meets(iv[i].r0, iv[i].r1, iv[j].w0, iv[j].w1) || meets(iv[j].r0, iv[j].r1, iv[i].w0, iv[i].w1)))
                pl.dep.push_back(j);
        pl.dep_off[i + 1] = (int)pl.dep.size();
    }
    return pl;
}

template <typename T, bool STRICT>
static int build_resident(lbm_handle *h)
{
    auto &r = h->res;
    free_resident(h);
    r.built = true;
    int coop = 0, occ = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->cfg.device));
    const bool three = h->resident_blocks == 0 || h->resident_blocks >= 3;
    if (three) CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stepr_kernel<T, STRICT, 3>, kBlock, 0));
    else CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stepr_kernel<T, STRICT, 2>, kBlock, 0));
    if (h->resident_blocks > 0) occ = std::min(occ, h->resident_blocks);
    const int ng = h->n_obs > 0 ? h->n_groups : 0;
    const int max_blocks = occ * h->n_sm;
    if (!coop || max_blocks - ng < 1) return LBM_OK;                // not usable: the per-update launches stay
    ResidentPlan pl = plan_resident((int)h->cfg.nxl, max_blocks, ng, h->grp_x0.data(), h->grp_x1.data());
    const int ncb = pl.n_col_blocks, nb = ncb + ng;
    std::vector<int> &col_a = pl.col_a, &dep_off = pl.dep_off, &dep = pl.dep;
    if (dep.empty()) dep.push_back(0);
    CUDA_TRY(cudaMalloc(&r.d_col_a, col_a.size() * sizeof(int)));
    CUDA_TRY(cudaMalloc(&r.d_dep_off, dep_off.size() * sizeof(int)));
    CUDA_TRY(cudaMalloc(&r.d_dep, dep.size() * sizeof(int)));
    CUDA_TRY(cudaMalloc(&r.d_prog, ((size_t)nb * kProgStride + kProgStride) * sizeof(unsigned int)));
    CUDA_TRY(cudaMemcpy(r.d_col_a, col_a.data(), col_a.size() * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(r.d_dep_off, dep_off.data(), dep_off.size() * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(r.d_dep, dep.data(), dep.size() * sizeof(int), cudaMemcpyHostToDevice));
    r.n_blocks = nb;
    r.n_col_blocks = ncb;
    int rc = ensure_err(h);
    if (rc) return rc;
    r.usable = true;
    return LBM_OK;
}

// n consecutive fused updates in ONE launch: rows first_row, first_row + row_stride, .., force slots slot0 ..
template <typename T, bool STRICT>
static int launch_resident_t(lbm_handle *h, int64_t n, int64_t first_row, int64_t row_stride, int64_t slot0)
{
    auto &r = h->res;
    StepParams<T> pa;
    LinkParams lp;
    fill_params<T>(h, pa, lp, h->cur, h->cur ^ 1, 0, (int)h->cfg.nxl, first_row, slot0);
    ResidentParams<T> rp;
    const ptrdiff_t delta_b = static_cast<const char *>(h->buf[h->cur ^ 1]) - static_cast<const char *>(h->buf[h->cur]);
    if (delta_b % (ptrdiff_t)sizeof(T)) return fail(LBM_E_INVALID, "population buffers are not element-aligned to each other");
    rp.buf_delta = delta_b / (ptrdiff_t)sizeof(T);
    rp.n_updates = (int)n;
    rp.n_col_blocks = r.n_col_blocks;
    rp.col_a = r.d_col_a; rp.dep_off = r.d_dep_off; rp.dep = r.d_dep;
    rp.prog = r.d_prog;
    rp.err = h->d_err;
    if (!h->clock_khz) cudaDeviceGetAttribute(&h->clock_khz, cudaDevAttrClockRate, h->cfg.device);
    rp.timeout_clk = (long long)h->resident_timeout_ms * std::max(h->clock_khz, 1000000);
    rp.walls = static_cast<const T *>(h->walls);
    rp.row_len = h->row_len; rp.wall_rows = std::max<int64_t>(h->wall_rows, 1);
    rp.first_row = first_row; rp.row_stride = row_stride;
    rp.ramp = h->ramp_n > 0 ? static_cast<const T *>(h->d_ramp) - h->ramp_it0 : nullptr;
    rp.one = static_cast<const T *>(h->d_one);
    rp.link_fs = h->d_link_fs;
    rp.fs_stride = (long long)h->n_links_total * 2;
    rp.slot0 = slot0;
    CUDA_TRY(cudaMemsetAsync(r.d_prog, 0, ((size_t)r.n_blocks * kProgStride + kProgStride) * sizeof(unsigned int), h->stream));
    void *args[] = {&pa, &lp, &rp};
    const bool three = h->resident_blocks == 0 || h->resident_blocks >= 3;
    const void *fn = three ? (const void *)stepr_kernel<T, STRICT, 3> : (const void *)stepr_kernel<T, STRICT, 2>;
    CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(r.n_blocks), dim3(kBlock), args, 0, h->stream));
    h->launches++;
    return LBM_OK;
}

// Whether a run of `n` fused updates starting at force slot `slot0` can go through stepr_kernel (builds the layout on
// first use).
static bool resident_usable(lbm_handle *h, int64_t n, int64_t slot0, int *rc)
{
    *rc = LBM_OK;
    if (n < 4 || n > (1 << 30) || !resident_candidate(h)) return false;
    if (h->n_obs > 0 && (!h->d_link_fs || slot0 + n > h->link_fs_cap)) return false;    // (per-link terms need a slot each)
    if (!h->res.built) {
        const bool strict = h->cfg.arith == LBM_ARITH_STRICT;
        if (h->cfg.dtype == LBM_F64) *rc = strict ? build_resident<double, true>(h) : build_resident<double, false>(h);
        else *rc = strict ? build_resident<float, true>(h) : build_resident<float, false>(h);
        if (*rc) return false;
    }
    return h->res.usable;
}

static int launch_resident(lbm_handle *h, int64_t n, int64_t first_row, int64_t row_stride, int64_t slot0)
{
    const bool strict = h->cfg.arith == LBM_ARITH_STRICT;
    if (h->cfg.dtype == LBM_F64)
        return strict ? launch_resident_t<double, true>(h, n, first_row, row_stride, slot0)
                      : launch_resident_t<double, false>(h, n, first_row, row_stride, slot0);
    return strict ? launch_resident_t<float, true>(h, n, first_row, row_stride, slot0)
                  : launch_resident_t<float, false>(h, n, first_row, row_stride, slot0);
}

// ---------------------------------------------------------------------------------------
extern "C" {

int lbm_abi_version(void) { return LBM_ABI_VERSION; }
const char *lbm_last_error(void) { return g_err.c_str(); }

int lbm_create(const lbm_cfg *cfg, lbm_t **out)
{
    if (!out) return fail(LBM_E_INVALID, "out is NULL");
    *out = nullptr;
    int rc = check_cfg(cfg);
    if (rc) return rc;
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(LBM_E_INVALID, "device %d not present (%d visible)", cfg->device, ndev);
    CUDA_TRY(cudaSetDevice(cfg->device));
    lbm_handle *h = new (std::nothrow) lbm_handle();
    if (!h) return fail(LBM_E_NOMEM, "host allocation failed");
    h->cfg = *cfg;
    compute_layout(*cfg, h->lay);
    h->esz = (size_t)h->lay.elem_size;
    h->row_len = 5 * cfg->ny + 4 * cfg->nx;
    cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, cfg->device);
    if (h->n_sm < 1) h->n_sm = 148;
    h->pf_ahead = 2 * h->n_sm;
    cudaError_t e = cudaEventCreate(&h->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_one, 8);
    if (e == cudaSuccess) {
        const double one_d = 1.0;
        const float one_f = 1.0f;
        e = cudaMemcpy(h->d_one, cfg->dtype == LBM_F64 ? (const void *)&one_d : (const void *)&one_f, h->esz, cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) { delete h; return fail(LBM_E_CUDA, "lbm_create: %s", cudaGetErrorString(e)); }
    *out = h;
    return LBM_OK;
}

static void free_links(lbm_handle *h)
{
    free_resident(h);
    h->grp_x0.clear(); h->grp_x1.clear();
    void *ptrs[] = {h->d_cell_x, h->d_cell_y, h->d_cell_off, h->d_link_q, h->d_link_kind, h->d_link_slot,
                    h->d_obs_off, h->d_link_c, h->d_link_f, h->d_done, h->d_mask, h->d_force_now, h->is_band ? nullptr : h->d_link_fs,
                    h->d_grp_cell, h->d_link_idx};
    h->d_grp_cell = h->d_link_idx = nullptr; h->n_groups = 0;
    h->d_link_fs = nullptr; h->link_fs_cap = 0; h->force_dirty_lo = h->force_dirty_hi = 0;
    h->d_force_now = nullptr;
    for (void *p : ptrs) if (p) cudaFree(p);
    h->d_cell_x = h->d_cell_y = h->d_cell_off = h->d_link_q = h->d_link_kind = h->d_link_slot = h->d_obs_off = nullptr;
    h->d_link_c = nullptr; h->d_link_f = nullptr; h->d_done = nullptr; h->d_mask = nullptr;
    h->n_obs = h->n_cells = h->n_links = h->n_links_total = h->n_link_blocks = 0;
}

static void destroy_band(lbm_handle *h)
{
    if (!h->band) return;
    lbm_handle *b = h->band;
    h->band = nullptr;
    b->walls = nullptr; b->d_ramp = nullptr; b->d_forces = nullptr; b->d_link_fs = nullptr;   // aliases of the parent's tables
    lbm_destroy(b);
}

int lbm_destroy(lbm_t *h)
{
    if (!h) return LBM_OK;
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->stream);
    destroy_band(h);
    invalidate_graphs(h);
    if (h->own_buf) { cudaFree(h->buf[0]); cudaFree(h->buf[1]); }
    if (h->walls) cudaFree(h->walls);
    if (h->rho) cudaFree(h->rho);
    if (h->u) cudaFree(h->u);
    if (h->d_forces) cudaFree(h->d_forces);
    if (h->d_probe) cudaFree(h->d_probe);
    if (h->d_ramp) cudaFree(h->d_ramp);
    if (h->d_one) cudaFree(h->d_one);
    peer_detach(h);
    if (h->d_flags) cudaFree(h->d_flags);
    if (h->h_err) cudaFreeHost(h->h_err);
    free_links(h);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    delete h;
    return LBM_OK;
}

int lbm_get_layout(const lbm_t *h, lbm_layout *out)
{
    if (!h || !out) return fail(LBM_E_INVALID, "NULL argument");
    *out = h->lay;
    return LBM_OK;
}

int lbm_bind_state(lbm_t *h, void *dev_a, void *dev_b, size_t bytes_each)
{
    CHECK_H(h);
    if (h->buf[0]) return fail(LBM_E_STATE, "state buffers already present");
    const size_t need = (size_t)h->lay.elems * h->esz;
    if (!dev_a || !dev_b || dev_a == dev_b) return fail(LBM_E_INVALID, "need two distinct device buffers");
    if (bytes_each < need) return fail(LBM_E_INVALID, "buffers too small: %zu < %zu bytes", bytes_each, need);
    if (((uintptr_t)dev_a | (uintptr_t)dev_b) & 127) return fail(LBM_E_INVALID, "buffers must be 128-byte aligned");
    h->buf[0] = dev_a;
    h->buf[1] = dev_b;
    h->own_buf = false;
    CUDA_TRY(cudaMemsetAsync(dev_a, 0, need, h->stream));
    CUDA_TRY(cudaMemsetAsync(dev_b, 0, need, h->stream));
    return LBM_OK;
}

int lbm_state_ptrs(const lbm_t *h, void **current, void **other)
{
    if (!h) return fail(LBM_E_INVALID, "handle is NULL");
    if (current) *current = h->buf[h->cur];
    if (other) *other = h->buf[h->cur ^ 1];
    return LBM_OK;
}

int lbm_set_stream(lbm_t *h, void *cuda_stream)
{
    CHECK_H(h);
    h->stream = static_cast<cudaStream_t>(cuda_stream);
    invalidate_graphs(h);
    return LBM_OK;
}

int lbm_set_right_wall(lbm_t *h, int32_t right_wall)
{
    CHECK_H(h);
    if (right_wall != LBM_RIGHT_VELOCITY && right_wall != LBM_RIGHT_PRESSURE) return fail(LBM_E_INVALID, "bad right_wall");
    h->cfg.right_wall = right_wall;
    invalidate_graphs(h);
    return LBM_OK;
}

int lbm_sync(lbm_t *h)
{
    CHECK_H(h);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return wait_error(h);
}

static int copy_field_h2d(lbm_handle *h, void *dev_cell0, int64_t dev_plane, const void *host, int nplanes)
{
    const size_t w = (size_t)h->cfg.ny * h->esz;
    for (int q = 0; q < nplanes; q++)
        CUDA_TRY(cudaMemcpy2DAsync(static_cast<char *>(dev_cell0) + q * dev_plane * h->esz, h->lay.pitch * h->esz,
                                   static_cast<const char *>(host) + (size_t)q * h->cfg.nxl * w, w, w,
                                   (size_t)h->cfg.nxl, cudaMemcpyHostToDevice, h->stream));
    return LBM_OK;
}

static int copy_field_d2h(lbm_handle *h, const void *dev_cell0, int64_t dev_plane, void *host, int nplanes)
{
    const size_t w = (size_t)h->cfg.ny * h->esz;
    for (int q = 0; q < nplanes; q++)
        CUDA_TRY(cudaMemcpy2DAsync(static_cast<char *>(host) + (size_t)q * h->cfg.nxl * w, w,
                                   static_cast<const char *>(dev_cell0) + q * dev_plane * h->esz, h->lay.pitch * h->esz,
                                   w, (size_t)h->cfg.nxl, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return wait_error(h);
}

static int upload_populations(lbm_t *h, const void *host, StateKind kind)
{
    CHECK_H(h);
    if (!host) return fail(LBM_E_INVALID, "host array is NULL");
    int rc = ensure_state(h);
    if (rc) return rc;
    std::vector<float> shifted;
    if (h->cfg.dtype == LBM_F32) {               // f32 state is stored as deviation from the weights
        const size_t n = (size_t)h->cfg.nxl * h->cfg.ny;
        const float *g = static_cast<const float *>(host);
        shifted.resize(9 * n);
        for (int q = 0; q < 9; q++)
            for (size_t k = 0; k < n; k++) shifted[q * n + k] = (float)((double)g[q * n + k] - weight_of(q));
        host = shifted.data();
    }
    rc = copy_field_h2d(h, elem_ptr(h, h->cur, 0), h->lay.plane, host, 9);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->stream));  // the caller may free the host array on return
    h->kind = kind;
    h->other_has_g = false;
    h->macro_valid = false;
    return LBM_OK;
}

int lbm_set_populations(lbm_t *h, const void *g_host) { return upload_populations(h, g_host, kHaveG); }

int lbm_set_post_collision(lbm_t *h, const void *f_host) { return upload_populations(h, f_host, kHaveF); }

int lbm_init_equilibrium(lbm_t *h, double rho, double ux, double uy)
{
    CHECK_H(h);
    int rc = ensure_state(h);
    if (rc) return rc;
    dim3 grid((unsigned)((h->cfg.ny + kBlock - 1) / kBlock), 1), block(kBlock);
    const bool strict = h->cfg.arith == LBM_ARITH_STRICT;
    for (int64_t a = 0; a < h->cfg.nxl; a += 32768) {
        grid.y = (unsigned)std::min<int64_t>(32768, h->cfg.nxl - a);
        const int64_t off = a * h->lay.pitch;
        if (h->cfg.dtype == LBM_F64) {
            double *d = static_cast<double *>(elem_ptr(h, h->cur, off));
            if (strict) init_kernel<double, true><<<grid, block, 0, h->stream>>>(d, h->lay.plane, (int)h->lay.pitch, (int)h->cfg.nxl, (int)h->cfg.ny, rho, ux, uy);
            else init_kernel<double, false><<<grid, block, 0, h->stream>>>(d, h->lay.plane, (int)h->lay.pitch, (int)h->cfg.nxl, (int)h->cfg.ny, rho, ux, uy);
        } else {
            float *d = static_cast<float *>(elem_ptr(h, h->cur, off));
            if (strict) init_kernel<float, true><<<grid, block, 0, h->stream>>>(d, h->lay.plane, (int)h->lay.pitch, (int)h->cfg.nxl, (int)h->cfg.ny, (float)rho, (float)ux, (float)uy);
            else init_kernel<float, false><<<grid, block, 0, h->stream>>>(d, h->lay.plane, (int)h->lay.pitch, (int)h->cfg.nxl, (int)h->cfg.ny, (float)rho, (float)ux, (float)uy);
        }
        h->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    h->kind = kHaveG;
    h->other_has_g = false;
    h->macro_valid = false;
    return LBM_OK;
}

int lbm_equilibrium(lbm_t *h, const void *rho_host, const void *u_host, void *g_eq_host)
{
    CHECK_H(h);
    if (!rho_host || !u_host || !g_eq_host) return fail(LBM_E_INVALID, "NULL argument");
    // scratch: rho/u device fields + a 9-plane pitched buffer
    const size_t cells = (size_t)h->cfg.nxl * h->lay.pitch;
    void *d_rho = nullptr, *d_u = nullptr, *d_g = nullptr;
    CUDA_TRY(cudaMalloc(&d_rho, cells * h->esz));
    CUDA_TRY(cudaMalloc(&d_u, 2 * cells * h->esz));
    CUDA_TRY(cudaMalloc(&d_g, 9 * cells * h->esz));
    int rc = copy_field_h2d(h, d_rho, (int64_t)cells, rho_host, 1);
    if (!rc) rc = copy_field_h2d(h, d_u, (int64_t)cells, u_host, 2);
    if (!rc) {
        const bool strict = h->cfg.arith == LBM_ARITH_STRICT;
        const int pitch = (int)h->lay.pitch, nxl = (int)h->cfg.nxl, ny = (int)h->cfg.ny;
        for (int a = 0; a < nxl; a += 32768) {      // (grid.y <= 65535)
            dim3 grid((unsigned)((h->cfg.ny + kBlock - 1) / kBlock), (unsigned)std::min(32768, nxl - a)), block(kBlock);
            const size_t off = (size_t)a * pitch;       // the kernel indexes u's second plane with nxl * pitch from its base
            if (h->cfg.dtype == LBM_F64) {
                if (strict) equilibrium_kernel<double, true><<<grid, block, 0, h->stream>>>((double *)d_g + off, (long long)cells, pitch, nxl, ny, (const double *)d_rho + off, (const double *)d_u + off);
                else equilibrium_kernel<double, false><<<grid, block, 0, h->stream>>>((double *)d_g + off, (long long)cells, pitch, nxl, ny, (const double *)d_rho + off, (const double *)d_u + off);
            } else {
                if (strict) equilibrium_kernel<float, true><<<grid, block, 0, h->stream>>>((float *)d_g + off, (long long)cells, pitch, nxl, ny, (const float *)d_rho + off, (const float *)d_u + off);
                else equilibrium_kernel<float, false><<<grid, block, 0, h->stream>>>((float *)d_g + off, (long long)cells, pitch, nxl, ny, (const float *)d_rho + off, (const float *)d_u + off);
            }
            h->launches++;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail(LBM_E_CUDA, "equilibrium_kernel: %s", cudaGetErrorString(e));
    }
    if (!rc) rc = copy_field_d2h(h, d_g, (int64_t)cells, g_eq_host, 9);
    cudaStreamSynchronize(h->stream);
    cudaFree(d_rho); cudaFree(d_u); cudaFree(d_g);
    return rc;
}

int lbm_set_links(lbm_t *h, int32_t n_obstacles, const int64_t *offsets, const int64_t *ijq,
                  const double *ibb, int32_t use_ibb)
{
    CHECK_H(h);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    free_links(h);
    destroy_band(h);
    invalidate_graphs(h);
    h->force_const.clear();
    if (h->d_forces) { cudaFree(h->d_forces); h->d_forces = nullptr; h->force_cap = 0; }
    if (n_obstacles <= 0) return LBM_OK;
    if (!offsets || !ijq) return fail(LBM_E_INVALID, "NULL link arrays");
    if (use_ibb && !ibb) return fail(LBM_E_INVALID, "use_ibb set but ibb is NULL");
    const int64_t K = offsets[n_obstacles];
    if (offsets[0] != 0 || K < 0 || K > (1 << 28)) return fail(LBM_E_INVALID, "bad offsets");
    const int64_t nx = h->cfg.nx, ny = h->cfg.ny, x0 = h->cfg.x0, nxl = h->cfg.nxl;
    struct L { int x, y, q, kind, slot; double c[3]; };
    std::vector<L> links;
    links.reserve((size_t)K);
    for (int64_t k = 0; k < K; k++) {
        const int64_t i = ijq[3 * k], j = ijq[3 * k + 1], q = ijq[3 * k + 2];
        if (q < 1 || q > 8) return fail(LBM_E_INVALID, "link %lld: direction %lld not in 1..8", (long long)k, (long long)q);
        if (i < 0 || i >= nx || j < 0 || j >= ny)
            return fail(LBM_E_INVALID, "link %lld: node (%lld, %lld) outside the lattice (the reference wraps negative indices; this library rejects them)", (long long)k, (long long)i, (long long)j);
        L l;
        l.q = (int)q; l.slot = (int)k; l.kind = 0; l.c[0] = l.c[1] = l.c[2] = 0.0;
        if (use_ibb) {
            const int qb = opp((int)q);
            const int64_t reach = 2;  // (i, j) + 2 c_qbar is read when p < 1/2, + c_qbar otherwise
            const double p = ibb[k], pp = 2.0 * p;
            const int64_t r = p < 0.5 ? reach : 1;
            const int64_t ie = i + r * cx_of(qb), je = j + r * cy_of(qb);
            if (ie < 0 || ie >= nx || je < 0 || je >= ny)
                return fail(LBM_E_INVALID, "link %lld: IBB stencil leaves the lattice", (long long)k);
            if (p < 0.5) {  // nb.py:98-100
                l.kind = 1;
                l.c[0] = p * (pp + 1.0);
                l.c[1] = (1.0 + pp) * (1.0 - pp);
                l.c[2] = p * (1.0 - pp);
            } else {        // nb.py:102-104
                l.kind = 2;
                l.c[0] = 1.0 / (p * (pp + 1.0));
                l.c[1] = (pp - 1.0) / p;
                l.c[2] = (1.0 - pp) / (1.0 + pp);
            }
            if (i >= x0 && i < x0 + nxl) {
                const int64_t il = ie - x0;
                if (il < -2 || il > nxl + 1)   // (i, j) + 2 c_qbar at most: two of the kHalo halo columns
                    return fail(LBM_E_UNSUPPORTED, "link %lld: IBB stencil reaches beyond the slab halo", (long long)k);
            }
        }
        if (i < x0 || i >= x0 + nxl) continue;  // owned by another slab
        if (h->cfg.dtype == LBM_F32) {
            int o = 0;
            while (o + 1 < n_obstacles && offsets[o + 1] <= k) o++;
            if ((int)h->force_const.size() < 2 * n_obstacles) h->force_const.assign(2 * n_obstacles, 0.0);
            h->force_const[2 * o] += 2.0 * weight_of((int)q) * cx_of((int)q);
            h->force_const[2 * o + 1] += 2.0 * weight_of((int)q) * cy_of((int)q);
        }
        l.x = (int)(i - x0);
        l.y = (int)j;
        links.push_back(l);
    }
    h->n_obs = n_obstacles;
    h->n_links_total = (int)K;
    // group by cell, keeping the caller's order inside a cell (later links overwrite earlier ones)
    std::vector<int> order(links.size());
    for (size_t k = 0; k < order.size(); k++) order[k] = (int)k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return links[a].x != links[b].x ? links[a].x < links[b].x : links[a].y < links[b].y;
    });
    std::vector<int> cx, cy, coff, lq, lkind, lslot;
    std::vector<double> lc;
    for (size_t n = 0; n < order.size(); n++) {
        const L &l = links[order[n]];
        if (cx.empty() || cx.back() != l.x || cy.back() != l.y) {
            cx.push_back(l.x); cy.push_back(l.y); coff.push_back((int)n);
        }
        lq.push_back(l.q); lkind.push_back(l.kind); lslot.push_back(l.slot);
        lc.push_back(l.c[0]); lc.push_back(l.c[1]); lc.push_back(l.c[2]);
    }
    coff.push_back((int)order.size());
    h->n_cells = (int)cx.size();
    h->n_links = (int)order.size();
    // groups of cells for the link blocks: at most kBlock cells and kBlock links each (a cell has at most 8 links)
    std::vector<int> grp(1, 0), lidx;
    for (size_t c = 0, cells = 0, lk = 0; c < cx.size(); c++) {
        const size_t nl = (size_t)(coff[c + 1] - coff[c]);
        if (cells + 1 > (size_t)kBlock || lk + nl > (size_t)kBlock) { grp.push_back((int)c); cells = 0; lk = 0; }
        cells++; lk += nl;
        for (size_t k = 0; k < nl; k++) lidx.push_back(cx[c] * (int)h->lay.pitch + cy[c]);
    }
    grp.push_back((int)cx.size());
    h->n_groups = (int)grp.size() - 1;
    for (int g = 0; g < h->n_groups && !cx.empty(); g++) {     // (cells are sorted by column; a slab that owns none of the
        h->grp_x0.push_back(cx[grp[g]]);                         // links still has one -- empty -- group)
        h->grp_x1.push_back(cx[grp[g + 1] - 1]);
    }
    // immediate sums: one block (looping over the groups) while the per-link terms fit its shared memory, else one block
    // per group; >= 1 so that forces are always written
    h->n_link_blocks = K <= kLinkLocal ? 1 : std::max(1, h->n_groups);
    std::vector<int> obs_off(n_obstacles + 1);
    for (int o = 0; o <= n_obstacles; o++) obs_off[o] = (int)offsets[o];

    auto up = [&](auto **dptr, const void *src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc((void **)dptr, std::max<size_t>(bytes, 16));
        if (e != cudaSuccess) return e;
        return bytes ? cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice) : cudaSuccess;
    };
    CUDA_TRY(up(&h->d_cell_x, cx.data(), cx.size() * sizeof(int)));
    CUDA_TRY(up(&h->d_cell_y, cy.data(), cy.size() * sizeof(int)));
    CUDA_TRY(up(&h->d_cell_off, coff.data(), coff.size() * sizeof(int)));
    CUDA_TRY(up(&h->d_link_q, lq.data(), lq.size() * sizeof(int)));
    CUDA_TRY(up(&h->d_link_kind, lkind.data(), lkind.size() * sizeof(int)));
    CUDA_TRY(up(&h->d_link_slot, lslot.data(), lslot.size() * sizeof(int)));
    CUDA_TRY(up(&h->d_obs_off, obs_off.data(), obs_off.size() * sizeof(int)));
    CUDA_TRY(up(&h->d_grp_cell, grp.data(), grp.size() * sizeof(int)));
    CUDA_TRY(up(&h->d_link_idx, lidx.data(), lidx.size() * sizeof(int)));
    if (h->cfg.dtype == LBM_F64) {
        CUDA_TRY(up(&h->d_link_c, lc.data(), lc.size() * sizeof(double)));
    } else {
        std::vector<float> lcf(lc.begin(), lc.end());
        CUDA_TRY(up(&h->d_link_c, lcf.data(), lcf.size() * sizeof(float)));
    }
    CUDA_TRY(cudaMalloc(&h->d_link_f, std::max<size_t>((size_t)K, 1) * 2 * sizeof(double)));
    CUDA_TRY(cudaMemset(h->d_link_f, 0, std::max<size_t>((size_t)K, 1) * 2 * sizeof(double)));
    CUDA_TRY(cudaMalloc(&h->d_force_now, (size_t)n_obstacles * 2 * sizeof(double)));   // lbm_forces_now's own slot
    CUDA_TRY(cudaMalloc(&h->d_done, sizeof(unsigned int)));
    CUDA_TRY(cudaMemset(h->d_done, 0, sizeof(unsigned int)));
    // per-cell mask: cells owned by the link blocks are skipped by the bulk threads
    const size_t mbytes = (size_t)nxl * h->lay.pitch;
    std::vector<unsigned char> mask(mbytes, 0);
    for (size_t c = 0; c < cx.size(); c++) mask[(size_t)cx[c] * h->lay.pitch + cy[c]] = 1;
    CUDA_TRY(up(&h->d_mask, mask.data(), mbytes));
    if (!h->is_band && !cx.empty()) {
        int rc = build_band(h, n_obstacles, offsets, ijq, ibb, use_ibb, *std::min_element(cx.begin(), cx.end()),
                            *std::max_element(cx.begin(), cx.end()));
        if (rc) return rc;
    }
    return LBM_OK;
}

int64_t lbm_wall_row_len(const lbm_t *h) { return h ? h->row_len : 0; }

int lbm_set_walls(lbm_t *h, int64_t n_rows, const double *rows_host)
{
    CHECK_H(h);
    if (n_rows < 1 || !rows_host) return fail(LBM_E_INVALID, "need at least one wall row");
    if (n_rows > h->wall_cap) {
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        invalidate_graphs(h);
        if (h->walls) CUDA_TRY(cudaFree(h->walls));
        h->walls = nullptr;
        h->wall_cap = 0;
        CUDA_TRY(cudaMalloc(&h->walls, (size_t)n_rows * h->row_len * h->esz));
        h->wall_cap = n_rows;
    }
    const size_t n = (size_t)n_rows * h->row_len;
    if (h->cfg.dtype == LBM_F64) {
        CUDA_TRY(cudaMemcpyAsync(h->walls, rows_host, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    } else {
        std::vector<float> tmp(n);
        for (size_t k = 0; k < n; k++) tmp[k] = (float)rows_host[k];
        CUDA_TRY(cudaMemcpyAsync(h->walls, tmp.data(), n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    h->wall_rows = n_rows;
    return LBM_OK;
}

static int check_row(lbm_handle *h, int64_t row)
{
    if (h->ramp_n > 0) {          // ramp table: row = iteration index into it, profile row = row % wall_rows
        if (!h->walls || h->wall_rows < 1) return fail(LBM_E_INVALID, "no wall profiles set");
        if (row < h->ramp_it0 || row >= h->ramp_it0 + h->ramp_n)
            return fail(LBM_E_INVALID, "row %lld outside the ramp table [%lld, %lld)", (long long)row, (long long)h->ramp_it0, (long long)(h->ramp_it0 + h->ramp_n));
        return LBM_OK;
    }
    if (!h->walls || row < 0 || row >= h->wall_rows)
        return fail(LBM_E_INVALID, "wall row %lld not in the table (%lld rows set)", (long long)row, (long long)h->wall_rows);
    return LBM_OK;
}

int lbm_set_wall_profiles(lbm_t *h, const double *u_left, const double *u_right, const double *u_top,
                          const double *u_bot, const double *rho_right)
{
    CHECK_H(h);
    const int64_t nx = h->cfg.nx, ny = h->cfg.ny;
    std::vector<double> row((size_t)h->row_len, 0.0);
    if (u_left) memcpy(row.data(), u_left, 2 * ny * sizeof(double));
    if (u_right) memcpy(row.data() + 2 * ny, u_right, 2 * ny * sizeof(double));
    if (u_top) memcpy(row.data() + 4 * ny, u_top, 2 * nx * sizeof(double));
    if (u_bot) memcpy(row.data() + 4 * ny + 2 * nx, u_bot, 2 * nx * sizeof(double));
    if (rho_right) memcpy(row.data() + 4 * ny + 4 * nx, rho_right, ny * sizeof(double));
    int rc = lbm_set_walls(h, 1, row.data());
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->stream));    // `row` is a local
    return LBM_OK;
}

int lbm_set_ramp(lbm_t *h, const double *ret_host, int64_t it0, int64_t n)
{
    CHECK_H(h);
    if (n < 0 || (n > 0 && !ret_host)) return fail(LBM_E_INVALID, "bad ramp table");
    if (n == 0) {                 // (captured batches carry their ramp mode in their key: nothing to invalidate)
        h->ramp_n = 0;
        return LBM_OK;
    }
    if (n > h->ramp_cap) {
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        invalidate_graphs(h);
        if (h->d_ramp) CUDA_TRY(cudaFree(h->d_ramp));
        h->d_ramp = nullptr;
        h->ramp_cap = 0;
        const int64_t cap = std::max<int64_t>(n, 64);
        CUDA_TRY(cudaMalloc(&h->d_ramp, (size_t)cap * h->esz));
        h->ramp_cap = cap;
    }
    if (h->cfg.dtype == LBM_F64) {
        CUDA_TRY(cudaMemcpyAsync(h->d_ramp, ret_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    } else {
        std::vector<float> tmp((size_t)n);
        for (int64_t k = 0; k < n; k++) tmp[k] = (float)ret_host[k];
        CUDA_TRY(cudaMemcpyAsync(h->d_ramp, tmp.data(), (size_t)n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    h->ramp_n = n;
    h->ramp_it0 = it0;
    return LBM_OK;
}

}  // extern "C"

// The launches of n_updates consecutive updates on the handle's stream (or into a stream capture).
static int enqueue_updates(lbm_handle *h, int64_t n_updates, int64_t first_row, int64_t row_stride, uint32_t flags)
{
    int rc = LBM_OK;
    for (int64_t s = 0; s < n_updates; s++) {
        const bool last = s == n_updates - 1;
        const bool wm = last && (flags & LBM_STEP_MACRO_LAST);
        const int mode = h->kind == kHaveG ? kCollideOnly : kFused;
        // several updates in one launch when none of them needs obstacle links or macro output
        // (the update that writes rho,u is the last one of the call and runs alone) ...
        const int64_t plain = n_updates - s - ((flags & LBM_STEP_MACRO_LAST) ? 1 : 0);
        // ... and the lattice is large enough to profit: below two full waves of 8 x 64 tiles at 4
        // blocks per SM the single-update kernel is faster (small lattices are latency bound)
        const bool big = ((h->cfg.nxl + 7) / 8) * ((h->cfg.ny + 63) / 64) >= 2 * 4 * (int64_t)h->n_sm || h->tb_force;
        // small lattices: the whole run of plain updates in ONE resident launch (stepr_kernel)
        if (mode == kFused && resident_usable(h, plain, s, &rc)) {
            rc = launch_resident(h, plain, first_row + s * row_stride, row_stride, s);
            if (rc) return rc;
            h->cur ^= (int)(plain & 1);
            s += plain - 1;
            continue;
        }
        if (rc) return rc;
        if (h->temporal && big && mode == kFused && h->n_obs > 0 && !links_elsewhere(h) && band_usable(h) && plain >= 2 && h->depth >= 2 &&
            (h->tb_force || h->cfg.nxl * h->cfg.ny >= (1LL << 24))) {
            const int d = (int)std::min<int64_t>(plain, h->depth);
            int64_t rows[4];
            for (int k = 0; k < d; k++) rows[k] = first_row + (s + k) * row_stride;
            rc = launch_band_group(h, h->cur, h->cur ^ 1, d, rows, s);
            if (rc) return rc;
            h->cur ^= 1;
            s += d - 1;
            continue;
        }
        if (h->temporal && big && mode == kFused && (h->n_obs == 0 || links_elsewhere(h)) && plain >= 2 && h->cfg.nxl >= 4) {
            int d = (int)std::min<int64_t>(plain, h->depth);
            // wavefront launches pay off from ~4096^2 cells per slab (measured: 4096^2 93 vs 81 GLUPS for
            // step2_kernel, 4096 x 2048 the other way round); below that pairs of updates
            const bool wave_ok = h->tb_force || h->cfg.nxl * h->cfg.ny >= (1LL << 24);
            if (!wave_ok) d = std::min(d, 2);
            if ((d >= 3 || (d == 2 && h->depth > 2)) && wave_ok) {
                int64_t rows[4];
                for (int k = 0; k < d; k++) rows[k] = first_row + (s + k) * row_stride;
                rc = launch_stepw(h, h->cur, h->cur ^ 1, 0, (int)h->cfg.nxl, d, rows);
            } else {
                rc = launch_step2(h, h->cur, h->cur ^ 1, 0, (int)h->cfg.nxl, first_row + s * row_stride,
                                  first_row + (s + 1) * row_stride);
            }
            if (rc) return rc;
            h->cur ^= 1;
            s += d - 1;
            continue;
        }
        rc = launch_step(h, mode, h->cur, h->cur ^ 1, 0, (int)h->cfg.nxl, first_row + s * row_stride, s, wm);
        if (rc) return rc;
        h->cur ^= 1;
        h->kind = kHaveF;
    }
    return LBM_OK;
}

extern "C" {

int lbm_step(lbm_t *h, int64_t n_updates, int64_t first_row, int64_t row_stride, uint32_t flags)
{
    CHECK_H(h);
    if (n_updates < 0) return fail(LBM_E_INVALID, "n_updates < 0");
    if (h->kind == kNone) return fail(LBM_E_STATE, "no populations set");
    if (n_updates == 0) return LBM_OK;
    int rc = ensure_forces(h, n_updates);
    if (rc) return rc;
    if (flags & LBM_STEP_MACRO_LAST) { rc = ensure_macro(h); if (rc) return rc; }
    // validate the rows before enqueuing anything
    for (int64_t s = 0; s < n_updates; s++) {
        if (s == 0 && h->kind == kHaveG) continue;  // collide-only update uses no walls
        rc = check_row(h, first_row + s * row_stride);
        if (rc) return rc;
        if (row_stride == 0) break;
    }
    h->force_skip0 = h->kind == kHaveG;         // a collide-only first update runs no link blocks: slot 0 stays as it is
    h->force_dirty_lo = h->force_dirty_hi = 0;  // (sums of an earlier batch that nobody fetched are dropped)
    h->defer_now = true;                        // the link blocks leave their sums to force_reduce_kernel (lbm_get_forces)
    struct DeferOff { lbm_handle *h; ~DeferOff() { h->defer_now = false; } } defer_off{h};
    CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
    // Small lattices are launch bound (4 us per update at 200 x 200, profiles/README.md): a batch of
    // updates is captured once into a CUDA graph and replayed (the wall table, force slots and
    // population buffers keep their addresses; their contents are read when the graph runs).
    // (<= 2^19 cells: below the size at which multi-update kernels take over, so that a captured batch
    // consists of step_kernel launches only)
    // (a batch that goes through stepr_kernel is one launch already -- and a cooperative launch cannot be captured)
    const bool resident_takes_it = h->kind == kHaveF && resident_usable(h, n_updates - ((flags & LBM_STEP_MACRO_LAST) ? 1 : 0), 0, &rc);
    if (rc) return rc;
    const bool graph_ok = h->use_graph && h->stream != nullptr && h->kind == kHaveF && n_updates >= 16 &&
                          h->cfg.nxl * h->cfg.ny <= (1LL << 19) && !h->peer[0].attached && !h->peer[1].attached &&
                          !resident_takes_it;
    if (graph_ok) {
        lbm_handle::StepGraph *g = nullptr;
        for (auto &e : h->graphs)
            if (e.n == n_updates && e.first_row == first_row && e.stride == row_stride && e.flags == flags && e.cur0 == h->cur &&
                e.ramp == (h->ramp_n > 0) && (!e.ramp || e.ramp_it0 == h->ramp_it0)) g = &e;
        if (!g) {
            if (h->graphs.size() >= 4) invalidate_graphs(h);
            lbm_handle::StepGraph e;
            e.n = n_updates; e.first_row = first_row; e.stride = row_stride; e.flags = flags; e.cur0 = h->cur;
            e.ramp = h->ramp_n > 0; e.ramp_it0 = h->ramp_it0;
            const int64_t l0 = h->launches;
            CUDA_TRY(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
            rc = enqueue_updates(h, n_updates, first_row, row_stride, flags);
            cudaError_t ce = cudaStreamEndCapture(h->stream, &e.graph);
            if (rc) { if (e.graph) cudaGraphDestroy(e.graph); h->cur = e.cur0; return rc; }
            if (ce != cudaSuccess) { h->cur = e.cur0; return fail(LBM_E_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(ce)); }
            ce = cudaGraphInstantiate(&e.exec, e.graph, 0);
            if (ce != cudaSuccess) { cudaGraphDestroy(e.graph); h->cur = e.cur0; return fail(LBM_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ce)); }
            e.cur1 = h->cur;
            e.launches = h->launches - l0;
            h->graphs.push_back(e);
            g = &h->graphs.back();
        } else {
            h->cur = g->cur1;
            h->launches += g->launches;
        }
        CUDA_TRY(cudaGraphLaunch(g->exec, h->stream));
    } else {
        rc = enqueue_updates(h, n_updates, first_row, row_stride, flags);
        if (rc) return rc;
    }
    CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
    h->ev_valid = true;
    if (h->n_obs > 0) mark_forces_dirty(h, h->force_skip0 ? 1 : 0, n_updates);
    h->force_n = n_updates;
    h->other_has_g = false;
    if (flags & LBM_STEP_MACRO_LAST) h->macro_valid = true;
    return LBM_OK;
}

int lbm_step_columns(lbm_t *h, int64_t xa, int64_t xb, int64_t row, int64_t slot, uint32_t flags)
{
    CHECK_H(h);
    if (h->kind == kNone) return fail(LBM_E_STATE, "no populations set");
    if (xa < 0 || xb > h->cfg.nxl || xa >= xb) return fail(LBM_E_INVALID, "bad column range");
    int rc = ensure_forces(h, slot + 1);
    if (rc) return rc;
    if (flags & LBM_STEP_MACRO_LAST) { rc = ensure_macro(h); if (rc) return rc; }
    const int mode = h->kind == kHaveG ? kCollideOnly : kFused;
    if (mode == kFused) { rc = check_row(h, row); if (rc) return rc; }
    h->defer_now = true;
    rc = launch_step(h, mode, h->cur, h->cur ^ 1, (int)xa, (int)xb, row, slot, (flags & LBM_STEP_MACRO_LAST) != 0);
    h->defer_now = false;
    if (!rc && mode == kFused && xa == 0 && xb == h->cfg.nxl) {       // link blocks rode along
        h->force_n = std::max<int64_t>(h->force_n, slot + 1);
        if (h->n_obs > 0 && slot < h->link_fs_cap) mark_forces_dirty(h, slot, slot + 1);
    }
    return rc;
}

int lbm_step2_columns(lbm_t *h, int64_t xa, int64_t xb, int64_t row1, int64_t row2)
{
    CHECK_H(h);
    if (h->kind != kHaveF) return fail(LBM_E_STATE, "lbm_step2_columns needs post-collision populations");
    if (h->n_obs > 0) return fail(LBM_E_UNSUPPORTED, "two-update launches do not handle obstacle links");
    if (xa < 0 || xb > h->cfg.nxl || xb - xa < 2) return fail(LBM_E_INVALID, "bad column range (need at least 2 columns)");
    int rc = check_row(h, row1);
    if (!rc) rc = check_row(h, row2);
    if (rc) return rc;
    return launch_step2(h, h->cur, h->cur ^ 1, (int)xa, (int)xb, row1, row2);
}

int lbm_stepn_columns(lbm_t *h, int64_t xa, int64_t xb, int32_t depth, const int64_t *rows)
{
    CHECK_H(h);
    if (h->kind != kHaveF) return fail(LBM_E_STATE, "lbm_stepn_columns needs post-collision populations");
    if (depth < 2 || depth > 4 || !rows) return fail(LBM_E_INVALID, "depth must be 2, 3 or 4 (with one wall row each)");
    if (xa < 0 || xb > h->cfg.nxl || xb - xa < 2) return fail(LBM_E_INVALID, "bad column range (need at least 2 columns)");
    for (int k = 0; k < depth; k++) {
        int rc = check_row(h, rows[k]);
        if (rc) return rc;
    }
    if (links_elsewhere(h)) {       // the bodies lie in other slabs: plain launch, zero sums in this slab's force slots
        if (xa == 0 && xb == h->cfg.nxl) {
            int rc = ensure_forces(h, depth);
            if (rc) return rc;
            h->force_dirty_lo = h->force_dirty_hi = 0;
            h->force_n = depth; h->force_skip0 = false;
            mark_forces_dirty(h, 0, depth);
        }
        return launch_stepw(h, h->cur, h->cur ^ 1, (int)xa, (int)xb, depth, rows);
    }
    if (h->n_obs > 0) {
        // bodies: the whole slab in one group -- wavefront launches beside the obstacle band, single updates
        // with the links inside it; drag/lift sums of the updates go to force slots 0 .. depth-1
        if (xa != 0 || xb != h->cfg.nxl) return fail(LBM_E_UNSUPPORTED, "multi-update launches with obstacle links cover the whole slab");
        if (!band_usable(h)) return fail(LBM_E_UNSUPPORTED, "no obstacle band on this slab (links next to a slab interface, or peer halos attached)");
        int rc = ensure_forces(h, depth);
        if (rc) return rc;
        h->force_dirty_lo = h->force_dirty_hi = 0;
        h->defer_now = true;
        rc = launch_band_group(h, h->cur, h->cur ^ 1, depth, rows, 0);
        h->defer_now = false;
        if (!rc) { h->force_n = depth; h->force_skip0 = false; mark_forces_dirty(h, 0, depth); }
        return rc;
    }
    return launch_stepw(h, h->cur, h->cur ^ 1, (int)xa, (int)xb, depth, rows);
}

int lbm_can_stepn(const lbm_t *h)
{
    if (!h) return 0;
    return h->n_obs == 0 || links_elsewhere(h) || band_usable(h) ? 1 : 0;
}

int lbm_set_temporal_depth(lbm_t *h, int32_t depth)
{
    if (!h) return fail(LBM_E_INVALID, "handle is NULL");
    if (depth < 1 || depth > 4) return fail(LBM_E_INVALID, "depth must be 1..4");
    h->depth = depth;
    invalidate_graphs(h);
    return LBM_OK;
}

int lbm_set_tuning(lbm_t *h, const char *key, int64_t value)
{
    if (!h || !key) return fail(LBM_E_INVALID, "NULL argument");
    invalidate_graphs(h);
    if (!strcmp(key, "graph")) { h->use_graph = value != 0; return LBM_OK; }
    if (!strcmp(key, "wave_chunk")) {
        if (value < 16 || value > (1 << 20)) return fail(LBM_E_INVALID, "wave_chunk must be in [16, 2^20]");
        h->wave_chunk = (int)value;
        h->wave_auto = false;                   // an explicit chunk is taken literally
    } else if (!strcmp(key, "wave_rows")) {
        if (value != 64 && value != 128) return fail(LBM_E_INVALID, "wave_rows must be 64 or 128");
        h->wave_rows = (int)value;
        h->tmap_rows = 0;
        for (bool &b : h->wave_attr_set) b = false;
    } else if (!strcmp(key, "resident")) {
        if (value < -1 || value > 1) return fail(LBM_E_INVALID, "resident must be -1 (auto), 0 or 1");
        h->resident = (int)value;
    } else if (!strcmp(key, "resident_blocks")) {
        if (value < 0 || value > 32) return fail(LBM_E_INVALID, "resident_blocks must be in [0, 32] (blocks per SM, 0 = what fits)");
        h->resident_blocks = (int)value;
        free_resident(h);
    } else if (!strcmp(key, "resident_timeout_ms")) {
        if (value < 0) return fail(LBM_E_INVALID, "resident_timeout_ms must not be negative (0: a block gives up at its first unsuccessful poll -- tests of the error path)");
        h->resident_timeout_ms = value;
    } else if (!strcmp(key, "pdl")) {
        h->pdl = value != 0;
    } else if (!strcmp(key, "wave_l2")) {
        if (value < 0 || value > 4096) return fail(LBM_E_INVALID, "wave_l2 must be in [0, 4096]");
        h->wave_l2 = (int)value;
    } else if (!strcmp(key, "wave_tail")) {
        if (value != -1 && value != 0 && (value < 16 || value > (1 << 20))) return fail(LBM_E_INVALID, "wave_tail must be -1 (auto), 0 (off) or >= 16 columns");
        h->wave_tail = (int)value;
    } else if (!strcmp(key, "peer_timeout_ms")) {
        if (value < 1) return fail(LBM_E_INVALID, "peer_timeout_ms must be positive");
        h->peer_timeout_ms = value;
    } else if (!strcmp(key, "pf_ahead")) {
        if (value < 0 || value > (1 << 20)) return fail(LBM_E_INVALID, "pf_ahead must be in [0, 2^20]");
        h->pf_ahead = (int)value;
    } else return fail(LBM_E_INVALID, "unknown tuning key '%s'", key);
    return LBM_OK;
}

int lbm_set_temporal_blocking(lbm_t *h, int32_t enable)
{
    if (!h) return fail(LBM_E_INVALID, "handle is NULL");
    h->temporal = enable != 0;
    invalidate_graphs(h);
    h->tb_force = enable < 0;                   // negative: also on lattices too small to profit (tests)
    if (enable < 0) enable = -enable;
    return LBM_OK;
}

int lbm_flip(lbm_t *h)
{
    CHECK_H(h);
    if (h->kind == kNone) return fail(LBM_E_STATE, "no populations set");
    h->cur ^= 1;
    h->kind = kHaveF;
    h->other_has_g = false;
    return LBM_OK;
}

int lbm_apply_bc(lbm_t *h, int64_t row)
{
    CHECK_H(h);
    if (h->kind != kHaveF) return fail(LBM_E_STATE, "lbm_apply_bc needs post-collision populations");
    int rc = check_row(h, row);
    if (rc) return rc;
    rc = ensure_forces(h, 1);
    if (rc) return rc;
    rc = launch_step(h, kStreamOnly, h->cur, h->cur ^ 1, 0, (int)h->cfg.nxl, row, 0, false);
    if (rc) return rc;
    h->other_has_g = true;
    h->force_n = 1;
    h->force_skip0 = false;
    h->force_dirty_lo = h->force_dirty_hi = 0;     // slot 0 now holds this call's sums
    return LBM_OK;
}

int lbm_get_forces(lbm_t *h, int64_t first, int64_t n, double *out)
{
    CHECK_H(h);
    if (!out || first < 0 || n < 0 || first + n > h->force_n) return fail(LBM_E_INVALID, "force slots [%lld, %lld) not available (%lld written)", (long long)first, (long long)(first + n), (long long)h->force_n);
    const int nobs = std::max(h->n_obs, 1);
    if (n == 0) return LBM_OK;
    { int rc = reduce_dirty_forces(h); if (rc) return rc; }
    CUDA_TRY(cudaMemcpyAsync(out, h->d_forces + first * nobs * 2, (size_t)n * nobs * 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    { int rc = wait_error(h); if (rc) return rc; }
    if (!h->force_const.empty())
        for (int64_t s = 0; s < n; s++) {
            if (first + s == 0 && h->force_skip0) continue;      // never produced by link blocks
            for (int k = 0; k < 2 * nobs; k++) out[s * 2 * nobs + k] += h->force_const[k];
        }
    return LBM_OK;
}

int lbm_get_forces_async(lbm_t *h, int64_t first, int64_t n, double *out_pinned)
{
    CHECK_H(h);
    if (!out_pinned || first < 0 || n < 0 || first + n > h->force_n) return fail(LBM_E_INVALID, "force slots [%lld, %lld) not available (%lld written)", (long long)first, (long long)(first + n), (long long)h->force_n);
    if (!h->force_const.empty()) return fail(LBM_E_UNSUPPORTED, "lbm_get_forces_async: f32 storage adds a host-side constant to the sums; use lbm_get_forces");
    const int nobs = std::max(h->n_obs, 1);
    if (n == 0) return LBM_OK;
    { int rc = reduce_dirty_forces(h); if (rc) return rc; }
    CUDA_TRY(cudaMemcpyAsync(out_pinned, h->d_forces + first * nobs * 2, (size_t)n * nobs * 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    return LBM_OK;
}

int lbm_forces_now(lbm_t *h, double *out)
{
    CHECK_H(h);
    if (!out) return fail(LBM_E_INVALID, "out is NULL");
    if (h->kind != kHaveF) return fail(LBM_E_STATE, "lbm_forces_now needs post-collision populations");
    const int nobs = std::max(h->n_obs, 1);
    if (h->n_obs == 0) { out[0] = out[1] = 0.0; return LBM_OK; }
    double *d_out = h->d_force_now;     // the handle's own slot (allocated with the links), not a step slot
    int rc = LBM_OK;
    {
        const bool strict = h->cfg.arith == LBM_ARITH_STRICT;
        dim3 grid(h->n_link_blocks), block(kBlock);
        if (h->cfg.dtype == LBM_F64) {
            StepParams<double> p; LinkParams lp;
            fill_params<double>(h, p, lp, h->cur, h->cur ^ 1, 0, (int)h->cfg.nxl, 0, 0);
            p.walls = nullptr; lp.forces = d_out;
            if (strict) force_kernel<double, true><<<grid, block, 0, h->stream>>>(p, lp);
            else force_kernel<double, false><<<grid, block, 0, h->stream>>>(p, lp);
        } else {
            StepParams<float> p; LinkParams lp;
            fill_params<float>(h, p, lp, h->cur, h->cur ^ 1, 0, (int)h->cfg.nxl, 0, 0);
            p.walls = nullptr; lp.forces = d_out;
            if (strict) force_kernel<float, true><<<grid, block, 0, h->stream>>>(p, lp);
            else force_kernel<float, false><<<grid, block, 0, h->stream>>>(p, lp);
        }
        h->launches++;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)nobs * 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail(LBM_E_CUDA, "lbm_forces_now: %s", cudaGetErrorString(e));
        if (!rc) rc = wait_error(h);
        if (!rc && !h->force_const.empty())
            for (int k = 0; k < 2 * nobs; k++) out[k] += h->force_const[k];
    }
    return rc;
}

int lbm_get_populations(lbm_t *h, int32_t which, void *host)
{
    CHECK_H(h);
    if (!host) return fail(LBM_E_INVALID, "host is NULL");
    int src;
    if (which == LBM_POP_POST_COLLISION) {
        if (h->kind != kHaveF) return fail(LBM_E_STATE, "no post-collision populations yet");
        src = h->cur;
    } else if (which == LBM_POP_STREAMED) {
        if (h->kind == kHaveG) src = h->cur;
        else if (h->kind == kHaveF && h->other_has_g) src = h->cur ^ 1;
        else return fail(LBM_E_STATE, "streamed populations not materialised: call lbm_apply_bc first");
    } else return fail(LBM_E_INVALID, "bad `which`");
    int rc = copy_field_d2h(h, elem_ptr(h, src, 0), h->lay.plane, host, 9);
    if (!rc && h->cfg.dtype == LBM_F32) {        // deviation storage -> populations
        const size_t n = (size_t)h->cfg.nxl * h->cfg.ny;
        float *g = static_cast<float *>(host);
        for (int q = 0; q < 9; q++)
            for (size_t k = 0; k < n; k++) g[q * n + k] = (float)((double)g[q * n + k] + weight_of(q));
    }
    return rc;
}

int lbm_get_macro(lbm_t *h, void *rho_host, void *u_host)
{
    CHECK_H(h);
    if (!h->macro_valid || !h->rho) return fail(LBM_E_STATE, "no macroscopic fields stored: run lbm_step with LBM_STEP_MACRO_LAST");
    const int64_t cells = h->cfg.nxl * h->lay.pitch;
    int rc = LBM_OK;
    if (rho_host) rc = copy_field_d2h(h, h->rho, cells, rho_host, 1);
    if (!rc && u_host) rc = copy_field_d2h(h, h->u, cells, u_host, 2);
    return rc;
}

int lbm_get_speed(lbm_t *h, const unsigned char *solid_host, void *speed_host)
{
    CHECK_H(h);
    if (!speed_host) return fail(LBM_E_INVALID, "speed_host is NULL");
    if (!h->macro_valid || !h->u) return fail(LBM_E_STATE, "no macroscopic fields stored: run lbm_step with LBM_STEP_MACRO_LAST");
    const size_t cells = (size_t)h->cfg.nxl * h->lay.pitch;
    void *d_out = nullptr;
    unsigned char *d_solid = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, cells * h->esz));
    int rc = LBM_OK;
    if (solid_host) {
        cudaError_t e = cudaMalloc(&d_solid, cells);
        if (e == cudaSuccess) e = cudaMemcpy2DAsync(d_solid, h->lay.pitch, solid_host, (size_t)h->cfg.ny, (size_t)h->cfg.ny,
                                                     (size_t)h->cfg.nxl, cudaMemcpyHostToDevice, h->stream);
        if (e != cudaSuccess) rc = fail(LBM_E_CUDA, "lbm_get_speed: %s", cudaGetErrorString(e));
    }
    if (!rc) {
        const int pitch = (int)h->lay.pitch, ny = (int)h->cfg.ny, nxl = (int)h->cfg.nxl;
        for (int a = 0; a < nxl; a += 32768) {      // (grid.y <= 65535)
            dim3 grid((unsigned)((h->cfg.ny + kBlock - 1) / kBlock), (unsigned)std::min(32768, nxl - a)), block(kBlock);
            const size_t off = (size_t)a * pitch;
            if (h->cfg.dtype == LBM_F64)
                speed_kernel<double><<<grid, block, 0, h->stream>>>((const double *)h->u + off, (const double *)h->u + cells + off, d_solid ? d_solid + off : nullptr, (double *)d_out + off, pitch, ny);
            else
                speed_kernel<float><<<grid, block, 0, h->stream>>>((const float *)h->u + off, (const float *)h->u + cells + off, d_solid ? d_solid + off : nullptr, (float *)d_out + off, pitch, ny);
            h->launches++;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail(LBM_E_CUDA, "speed_kernel: %s", cudaGetErrorString(e));
    }
    if (!rc) rc = copy_field_d2h(h, d_out, (int64_t)cells, speed_host, 1);
    cudaStreamSynchronize(h->stream);
    cudaFree(d_out);
    if (d_solid) cudaFree(d_solid);
    return rc;
}

int lbm_probe_line(lbm_t *h, int32_t axis, int64_t index, int64_t row, void *out_host)
{
    CHECK_H(h);
    if (!out_host) return fail(LBM_E_INVALID, "out_host is NULL");
    if (h->kind != kHaveF) return fail(LBM_E_STATE, "lbm_probe_line needs post-collision populations");
    if (h->n_obs > 0) return fail(LBM_E_UNSUPPORTED, "lbm_probe_line ignores obstacle links; not available with obstacles");
    if (axis != 0 && axis != 1) return fail(LBM_E_INVALID, "axis must be 0 (column) or 1 (row)");
    const int64_t n = axis == 0 ? h->cfg.ny : h->cfg.nxl;
    const int64_t lim = axis == 0 ? h->cfg.nxl : h->cfg.ny;
    if (index < 0 || index >= lim) return fail(LBM_E_INVALID, "line index out of range");
    int rc = check_row(h, row);
    if (rc) return rc;
    if (!h->d_probe || h->probe_cap < 3 * n) {
        if (h->d_probe) { CUDA_TRY(cudaStreamSynchronize(h->stream)); CUDA_TRY(cudaFree(h->d_probe)); h->d_probe = nullptr; }
        CUDA_TRY(cudaMalloc(&h->d_probe, (size_t)3 * n * h->esz));
        h->probe_cap = 3 * n;
    }
    dim3 grid((unsigned)((n + kBlock - 1) / kBlock)), block(kBlock);
    const bool strict = h->cfg.arith == LBM_ARITH_STRICT;
    rc = peer_wait(h);               // slab runs: the edge cells pull from halo columns the neighbours fill
    if (rc) return rc;
    if (h->cfg.dtype == LBM_F64) {
        StepParams<double> p; LinkParams lp;
        fill_params<double>(h, p, lp, h->cur, h->cur ^ 1, 0, (int)h->cfg.nxl, row, 0);
        if (strict) probe_kernel<double, true><<<grid, block, 0, h->stream>>>(p, axis, (int)index, (int)n, (double *)h->d_probe);
        else probe_kernel<double, false><<<grid, block, 0, h->stream>>>(p, axis, (int)index, (int)n, (double *)h->d_probe);
    } else {
        StepParams<float> p; LinkParams lp;
        fill_params<float>(h, p, lp, h->cur, h->cur ^ 1, 0, (int)h->cfg.nxl, row, 0);
        if (strict) probe_kernel<float, true><<<grid, block, 0, h->stream>>>(p, axis, (int)index, (int)n, (float *)h->d_probe);
        else probe_kernel<float, false><<<grid, block, 0, h->stream>>>(p, axis, (int)index, (int)n, (float *)h->d_probe);
    }
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_host, h->d_probe, (size_t)3 * n * h->esz, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return LBM_OK;
}

// ---- peer halo exchange (slab runs, one process per GPU) --------------------------------------
int lbm_peer_export(lbm_t *h, lbm_peer_info *out)
{
    CHECK_H(h);
    if (!out) return fail(LBM_E_INVALID, "out is NULL");
    if (h->buf[0] && !h->own_buf)
        return fail(LBM_E_STATE, "peer halos need library-owned population buffers (do not call lbm_bind_state)");
    int rc = ensure_state(h);
    if (rc) return rc;
    if (!h->d_flags) {
        CUDA_TRY(cudaMalloc(&h->d_flags, 4 * sizeof(unsigned int)));
        CUDA_TRY(cudaMemset(h->d_flags, 0, 4 * sizeof(unsigned int)));
    }
    { int rce = ensure_err(h); if (rce) return rce; }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    memset(out, 0, sizeof *out);
    static_assert(sizeof(cudaIpcMemHandle_t) <= LBM_IPC_HANDLE_BYTES, "IPC handle size");
    for (int b = 0; b < 2; b++) {
        cudaIpcMemHandle_t mh;
        CUDA_TRY(cudaIpcGetMemHandle(&mh, h->buf[b]));
        memcpy(out->mem[b], &mh, sizeof mh);
        out->addr[b] = (uint64_t)(uintptr_t)h->buf[b];
    }
    cudaIpcMemHandle_t fh;
    CUDA_TRY(cudaIpcGetMemHandle(&fh, h->d_flags));
    memcpy(out->flags, &fh, sizeof fh);
    out->flags_addr = (uint64_t)(uintptr_t)h->d_flags;
    out->pid = (int64_t)getpid();
    out->device = h->cfg.device;
    out->x0 = h->cfg.x0; out->nxl = h->cfg.nxl;
    out->origin = h->lay.origin; out->plane = h->lay.plane; out->pitch = h->lay.pitch; out->elem_size = h->lay.elem_size;
    return LBM_OK;
}

int lbm_peer_attach(lbm_t *h, int32_t side, const lbm_peer_info *nb)
{
    CHECK_H(h);
    if (side != 0 && side != 1) return fail(LBM_E_INVALID, "side must be 0 (left) or 1 (right)");
    if (!nb) return fail(LBM_E_INVALID, "neighbour info is NULL");
    if (!h->d_flags || !h->own_buf) return fail(LBM_E_STATE, "call lbm_peer_export on this handle first");
    lbm_handle::Peer &pe = h->peer[side];
    if (pe.attached) return fail(LBM_E_STATE, "side %d already attached", side);
    if (nb->pitch != h->lay.pitch || nb->elem_size != h->lay.elem_size)
        return fail(LBM_E_INVALID, "neighbour has a different row pitch or element type");
    if (side == 0 ? nb->x0 + nb->nxl != h->cfg.x0 : nb->x0 != h->cfg.x0 + h->cfg.nxl)
        return fail(LBM_E_INVALID, "neighbour slab [%lld, %lld) does not touch side %d of [%lld, %lld)", (long long)nb->x0,
                    (long long)(nb->x0 + nb->nxl), side, (long long)h->cfg.x0, (long long)(h->cfg.x0 + h->cfg.nxl));
    if (nb->nxl < kHalo || h->cfg.nxl < kHalo) return fail(LBM_E_INVALID, "slabs must be at least %d columns wide", kHalo);
    if (nb->pid == (int64_t)getpid()) {     // same process (two handles on two devices): plain peer access
        if (nb->device != h->cfg.device) {
            cudaError_t e = cudaDeviceEnablePeerAccess((int)nb->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return fail(LBM_E_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", (int)nb->device, cudaGetErrorString(e));
        }
        pe.buf[0] = (void *)(uintptr_t)nb->addr[0];
        pe.buf[1] = (void *)(uintptr_t)nb->addr[1];
        pe.flags = (unsigned int *)(uintptr_t)nb->flags_addr;
        pe.ipc = false;
    } else {
        cudaIpcMemHandle_t mh;
        for (int b = 0; b < 2; b++) {
            memcpy(&mh, nb->mem[b], sizeof mh);
            CUDA_TRY(cudaIpcOpenMemHandle(&pe.buf[b], mh, cudaIpcMemLazyEnablePeerAccess));
        }
        memcpy(&mh, nb->flags, sizeof mh);
        void *fp = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&fp, mh, cudaIpcMemLazyEnablePeerAccess));
        pe.flags = static_cast<unsigned int *>(fp);
        pe.ipc = true;
    }
    pe.nxl = nb->nxl; pe.plane = nb->plane; pe.origin = nb->origin;
    pe.attached = true;
    h->tmap_peer_rows = 0;
    invalidate_graphs(h);
    return LBM_OK;
}

int lbm_peer_detach(lbm_t *h)
{
    CHECK_H(h);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    peer_detach(h);
    h->tmap_peer_rows = 0;
    return LBM_OK;
}

int lbm_peer_push(lbm_t *h, int32_t which)
{
    CHECK_H(h);
    if (which != 0 && which != 1) return fail(LBM_E_INVALID, "which must be 0 (current) or 1 (other buffer)");
    if (!(h->peer[0].attached || h->peer[1].attached)) return LBM_OK;
    int rc = peer_wait(h);
    if (rc) return rc;
    const int b = h->cur ^ which;
    const int64_t pitch = h->lay.pitch, nxl = h->cfg.nxl;
    dim3 grid((unsigned)((h->cfg.ny + kBlock - 1) / kBlock), 9 * 2 * kHalo), block(kBlock);
    auto go = [&](auto zero) {
        using T = decltype(zero);
        T *pl = nullptr, *pr = nullptr;
        if (h->peer[0].attached) pl = static_cast<T *>(h->peer[0].buf[b]) + h->peer[0].origin + h->peer[0].nxl * pitch;
        if (h->peer[1].attached) pr = static_cast<T *>(h->peer[1].buf[b]) + h->peer[1].origin - nxl * pitch;
        peer_push_kernel<T><<<grid, block, 0, h->stream>>>(static_cast<const T *>(elem_ptr(h, b, 0)), pl, pr, h->lay.plane,
                                                           h->peer[0].plane, h->peer[1].plane, (int)pitch, (int)nxl, (int)h->cfg.ny);
    };
    if (h->cfg.dtype == LBM_F64) go(double(0)); else go(float(0));
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return LBM_OK;
}

int lbm_peer_signal(lbm_t *h)
{
    CHECK_H(h);
    if (!(h->peer[0].attached || h->peer[1].attached)) return LBM_OK;
    h->peer_seq++;
    // I am the RIGHT neighbour of my left neighbour: its flag word [1]; and the left neighbour of my right one: [0]
    peer_signal_kernel<<<1, 1, 0, h->stream>>>(h->peer[0].attached ? h->peer[0].flags + 1 : nullptr,
                                               h->peer[1].attached ? h->peer[1].flags + 0 : nullptr, h->peer_seq);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return LBM_OK;
}

int lbm_state_checksum(lbm_t *h, uint64_t *out)
{
    CHECK_H(h);
    if (!out) return fail(LBM_E_INVALID, "out is NULL");
    if (h->kind == kNone) return fail(LBM_E_STATE, "no populations set");
    if (!h->d_probe || h->probe_cap < 4) {
        if (h->d_probe) { CUDA_TRY(cudaStreamSynchronize(h->stream)); CUDA_TRY(cudaFree(h->d_probe)); h->d_probe = nullptr; }
        CUDA_TRY(cudaMalloc(&h->d_probe, 64));
        h->probe_cap = 64 / (int64_t)h->esz;
    }
    unsigned long long *d = static_cast<unsigned long long *>(h->d_probe);
    CUDA_TRY(cudaMemsetAsync(d, 0, sizeof *d, h->stream));
    const int blocks = 8 * h->n_sm;
    if (h->cfg.dtype == LBM_F64)
        checksum_kernel<double><<<blocks, kBlock, 0, h->stream>>>(static_cast<const double *>(elem_ptr(h, h->cur, 0)), h->lay.plane, (int)h->lay.pitch, (int)h->cfg.nxl, (int)h->cfg.ny, d);
    else
        checksum_kernel<float><<<blocks, kBlock, 0, h->stream>>>(static_cast<const float *>(elem_ptr(h, h->cur, 0)), h->lay.plane, (int)h->lay.pitch, (int)h->cfg.nxl, (int)h->cfg.ny, d);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    unsigned long long v = 0;
    CUDA_TRY(cudaMemcpyAsync(&v, d, sizeof v, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    *out = (uint64_t)v;
    return LBM_OK;
}

int64_t lbm_launch_count(const lbm_t *h) { return h ? h->launches : 0; }

int lbm_resident_plan(int32_t nx, int32_t max_blocks, int32_t n_groups, const int32_t *grp_x0, const int32_t *grp_x1,
                      int32_t *n_col_blocks, int32_t *col_a, int32_t *dep_off, int32_t *dep, int64_t dep_cap)
{
    if (nx < 1 || n_groups < 0 || max_blocks - n_groups < 1 || (n_groups > 0 && (!grp_x0 || !grp_x1)) || !n_col_blocks || !col_a || !dep_off)
        return fail(LBM_E_INVALID, "lbm_resident_plan: bad arguments");
    const ResidentPlan pl = plan_resident(nx, max_blocks, n_groups, grp_x0, grp_x1);
    if ((int64_t)pl.dep.size() > dep_cap) return fail(LBM_E_INVALID, "lbm_resident_plan: %lld dependencies, room for %lld", (long long)pl.dep.size(), (long long)dep_cap);
    *n_col_blocks = pl.n_col_blocks;
    std::copy(pl.col_a.begin(), pl.col_a.end(), col_a);
    std::copy(pl.dep_off.begin(), pl.dep_off.end(), dep_off);
    if (dep) std::copy(pl.dep.begin(), pl.dep.end(), dep);
    return LBM_OK;
}

int lbm_last_step_ms(lbm_t *h, float *ms)
{
    CHECK_H(h);
    if (!ms) return fail(LBM_E_INVALID, "ms is NULL");
    if (!h->ev_valid) return fail(LBM_E_STATE, "no lbm_step call yet");
    CUDA_TRY(cudaEventSynchronize(h->ev1));
    CUDA_TRY(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return LBM_OK;
}

}  // extern "C"
