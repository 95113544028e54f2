/*
 * lbm_b200.h -- C ABI of the B200-native D2Q9 time step (liblbm_b200.so).
 *
 * This is the drop-in boundary for the hot path of jviquerat/lbm.  The reference
 * has no FFI of its own: its seam is the `lattice` class, whose compute methods
 * each forward to one Numba kernel (lbm/src/core/lattice.py:178-286 ->
 * lbm/src/core/nb.py).  Every entry point below names the reference interface it
 * replaces.  lbm_b200/lattice.py binds these with ctypes and presents the
 * reference's `lattice` surface; INTEGRATION.md shows the binding a maintainer of
 * the reference would add.
 *
 * Conventions
 *   - plain C types only; no C++/torch types cross the boundary;
 *   - every call returns LBM_OK (0) or a negative error code and never throws;
 *     lbm_last_error() returns a message for the last failure on this thread;
 *   - the library never takes ownership of caller memory; device buffers bound
 *     with lbm_bind_state stay owned by the caller (a torch tensor);
 *   - all device work is enqueued on the stream given to lbm_set_stream (default:
 *     the CUDA legacy stream); only the lbm_get_* calls and lbm_sync wait for it;
 *   - one host thread per handle.
 *
 * Field layout on the HOST side is the reference's: C-ordered [q][i][j] with
 * i = x (0..nx-1), j = y (0..ny-1), j contiguous (lattice.py:155-174); element
 * type = the handle's dtype.  D2Q9 numbering, weights and opposite table are
 * lattice.py:135-152.
 *
 * Time-step formulation.  The state carried between steps is
 * the post-collision population array F (= the reference's g_up).  One lbm_step
 * "update" executes, fused in one kernel,
 *     stream (pull)  ->  obstacle (I)BB  ->  Zou-He walls + corners   [of reference iteration it-1]
 *     -> macro -> equilibrium -> TRT collision                        [of reference iteration it]
 * which is exactly run.py:33-45 regrouped.  The very first update after
 * lbm_set_populations / lbm_init_equilibrium is collide-only (iteration 0).
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBM_ABI_VERSION 1

/* error codes */
#define LBM_OK            0
#define LBM_E_INVALID    -1  /* bad argument */
#define LBM_E_CUDA       -2  /* CUDA runtime error (message has the detail) */
#define LBM_E_STATE      -3  /* call not valid in the handle's current state */
#define LBM_E_NOMEM      -4
#define LBM_E_UNSUPPORTED -5

/* lbm_cfg.dtype */
#define LBM_F64 0
#define LBM_F32 1
/* lbm_cfg.arith: LBM_ARITH_FUSED lets the compiler contract a*b+c into FMAs (production);
 * LBM_ARITH_STRICT evaluates every expression with separately rounded IEEE operations in the
 * order the reference writes them -- bit-identical to oracle/lbm_oracle.c (parity tests). */
#define LBM_ARITH_FUSED  0
#define LBM_ARITH_STRICT 1
/* lbm_cfg.right_wall: which Zou-He variant closes the right side
 * (nb_zou_he_right_wall_velocity nb.py:147-169 / nb_zou_he_right_wall_pressure nb.py:173-195). */
#define LBM_RIGHT_VELOCITY 0
#define LBM_RIGHT_PRESSURE 1

/* lbm_step flags */
#define LBM_STEP_MACRO_LAST 1u /* the last update of the call also stores rho,u (lattice.macro) */

/* lbm_get_populations `which` */
#define LBM_POP_POST_COLLISION 0 /* F  == reference g_up */
#define LBM_POP_STREAMED       1 /* g  == reference g (valid after set_populations or lbm_apply_bc) */

typedef struct lbm_handle lbm_t;

typedef struct lbm_cfg {
    int64_t nx, ny;    /* global lattice size (lattice.nx, lattice.ny) */
    int64_t x0, nxl;   /* x-slab owned by this handle: global columns [x0, x0+nxl); one GPU: 0, nx */
    double  om_p, om_m;/* TRT rates lattice.om_p_lbm / om_m_lbm (lattice.py:127-131) */
    int32_t dtype;     /* LBM_F64 | LBM_F32 */
    int32_t arith;     /* LBM_ARITH_FUSED | LBM_ARITH_STRICT */
    int32_t right_wall;/* LBM_RIGHT_VELOCITY | LBM_RIGHT_PRESSURE */
    int32_t device;    /* CUDA device ordinal */
} lbm_cfg;

/* Memory layout of one population buffer, for callers that allocate it themselves (torch) and
 * for the slab halo exchange: element (q, x, y), x in [-halo, nxl+halo) (x < 0 and x >= nxl are the
 * halo columns), lives at element offset  origin + q*plane + x*pitch + y. */
typedef struct lbm_layout {
    int64_t elems;  /* elements in one buffer */
    int64_t origin; /* offset of (q=0, x=0, y=0) */
    int64_t plane;  /* elements between consecutive q */
    int64_t pitch;  /* elements between consecutive x */
    int64_t elem_size;
    int64_t halo;   /* halo columns on each side (4) */
} lbm_layout;

int lbm_abi_version(void);
const char *lbm_last_error(void);

/* Replaces lattice.__init__/set_default_lbm array allocation (lattice.py:117-174). */
int lbm_create(const lbm_cfg *cfg, lbm_t **out);
int lbm_destroy(lbm_t *h);
int lbm_get_layout(const lbm_t *h, lbm_layout *out);
/* Bind two caller-owned device buffers of layout.elems elements each (they are zero-filled).
 * Optional: without it the library allocates its own on first use. */
int lbm_bind_state(lbm_t *h, void *dev_a, void *dev_b, size_t bytes_each);
/* Device address of the buffer that holds the current / the other population array. */
int lbm_state_ptrs(const lbm_t *h, void **current, void **other);
int lbm_set_stream(lbm_t *h, void *cuda_stream);
int lbm_sync(lbm_t *h);
/* Switch the right-side Zou-He variant for the following updates (the apps choose it by which
 * lattice.zou_he_right_wall_* method their set_bc calls: cavity.py:84 vs turek.py:121). */
int lbm_set_right_wall(lbm_t *h, int32_t right_wall);

/* lattice.g = ... (cavity.py:62, turek.py:91): upload pre-collision populations g[9][nxl][ny]. */
int lbm_set_populations(lbm_t *h, const void *g_host);
/* Restart: install post-collision populations F[9][nxl][ny] (what lbm_get_populations(...,
 * LBM_POP_POST_COLLISION, ...) returned) as the current state; the next update is a fused one.
 * The reference has no checkpoint/restart; its whole state is this array. */
int lbm_set_post_collision(lbm_t *h, const void *f_host);
/* Same state without a host array: g_q = equilibrium(rho, ux, uy) everywhere, computed on the
 * device (what the apps' initialize() produces with u = 0: cavity.py:54-62). */
int lbm_init_equilibrium(lbm_t *h, double rho, double ux, double uy);
/* nb_equilibrium (nb.py:7-17) on host fields rho[nxl][ny], u[2][nxl][ny] -> g_eq[9][nxl][ny]. */
int lbm_equilibrium(lbm_t *h, const void *rho_host, const void *u_host, void *g_eq_host);

/* Obstacle link lists, replaces the per-obstacle arguments of nb_bounce_back_obstacle
 * (nb.py:77-117) and nb_drag_lift (nb.py:49-73).  Obstacle o owns links
 * offsets[o] .. offsets[o+1]-1; ijq holds rows (i, j, q) with GLOBAL i, q pointing from the fluid
 * node into the solid (lattice.py:336-341); ibb[k] is the wall distance (ignored if !use_ibb).
 * Links outside this handle's slab are dropped.  Negative or out-of-range indices (the
 * reference wraps them silently) are rejected with LBM_E_INVALID. */
int lbm_set_links(lbm_t *h, int32_t n_obstacles, const int64_t *offsets, const int64_t *ijq,
                  const double *ibb, int32_t use_ibb);

/* Wall profiles, replaces the u_left/u_right/u_top/u_bot/rho_right arguments of the nb_zou_he_*
 * kernels (lattice.py:160-164).  One row = [u_left(2*ny) | u_right(2*ny) | u_top(2*nx) |
 * u_bot(2*nx) | rho_right(ny)] doubles, component-major as in the reference, GLOBAL sizes.
 * Copies n_rows rows into the device table (replacing it).  If rows_host is pinned memory the
 * copy is asynchronous: keep it unchanged until the next lbm_sync / lbm_get_*. */
int64_t lbm_wall_row_len(const lbm_t *h);
int lbm_set_walls(lbm_t *h, int64_t n_rows, const double *rows_host);
/* The same table with ONE row given as the reference's five arrays (lattice.py:160-164: u_left[2][ny],
 * u_right[2][ny], u_top[2][nx], u_bot[2][nx], rho_right[ny]; NULL = zeros): the BASE profiles of a
 * ramped run, uploaded once. */
int lbm_set_wall_profiles(lbm_t *h, const double *u_left, const double *u_right, const double *u_top,
                          const double *u_bot, const double *rho_right);
/* Inlet ramp.  Every reference app scales its wall velocity profile by one scalar per iteration,
 * ret(it) = 1 - exp(-it^2 / (2 sigma^2))  (cavity.py:70-73, turek.py:99-104, poiseuille.py, array.py):
 * with a ramp table set, the `row` arguments of lbm_step / lbm_step*_columns / lbm_apply_bc /
 * lbm_probe_line are ITERATION indices it in [it0, it0+n): the update multiplies the velocity entries
 * of profile row (it % n_rows) by ret_host[it - it0] (rho_right is not scaled, turek.py:109), so a
 * step's host->device traffic is 8 bytes instead of a wall row.  The product is one rounded
 * multiplication, as in the apps (u_lbm*ret; (ret*u_lbm)*profile).  n = 0 removes the table.  If
 * ret_host is pinned the copy is asynchronous: keep it unchanged until the next lbm_sync / lbm_get_*. */
int lbm_set_ramp(lbm_t *h, const double *ret_host, int64_t it0, int64_t n);

/* n_updates fused updates; update s uses wall row first_row + s*row_stride and stores the
 * momentum-exchange sums of every obstacle in force slot s.  Replaces one pass of
 * run.py:33-45 (macro, equilibrium, collision_stream, set_bc) per update. */
int lbm_step(lbm_t *h, int64_t n_updates, int64_t first_row, int64_t row_stride, uint32_t flags);
/* Lower level, for slab overlap: one update restricted to local columns [xa, xb) WITHOUT
 * flipping the buffers; lbm_flip makes the written buffer current.  slot = force slot. */
int lbm_step_columns(lbm_t *h, int64_t xa, int64_t xb, int64_t row, int64_t slot, uint32_t flags);
int lbm_flip(lbm_t *h);
/* Two consecutive updates (wall rows row1, row2) of local columns [xa, xb) in ONE launch
 * (temporal blocking through shared memory), without flipping; reads columns xa-2 .. xb+1.
 * Not available with obstacle links (lbm_stepn_columns is).  lbm_step pairs updates this way by itself when the lattice
 * has no obstacles and is large enough to profit (>= 1184 tiles of 8 x 64 cells), unless
 * lbm_set_temporal_blocking(h, 0) was called; enable < 0 forces pairing on any size (tests). */
int lbm_step2_columns(lbm_t *h, int64_t xa, int64_t xb, int64_t row1, int64_t row2);
int lbm_set_temporal_blocking(lbm_t *h, int32_t enable);
/* depth = 2, 3 or 4 consecutive updates (wall rows rows[0..depth-1]) of local columns [xa, xb) in
 * ONE launch (wavefront temporal blocking: a block sweeps a strip of rows along x, the updates
 * form a pipeline through shared memory, the source is streamed in by TMA bulk copies), without
 * flipping; reads columns xa-depth .. xb+depth-1.  With obstacle links (nb_bounce_back_obstacle,
 * nb.py:77-117; nb_drag_lift, nb.py:49-73) the call must cover the whole slab: the bodies get a band of
 * columns of their own -- four columns beyond the outermost link cell on either side -- that is updated
 * `depth` times by the single-update kernel with the links (momentum-exchange sums in force slots
 * 0 .. depth-1) while one wavefront launch per side covers the obstacle-free columns; lbm_step groups
 * updates the same way on lattices of >= 2^24 cells.  Bit-identical to single updates.
 * lbm_set_temporal_depth bounds the updates per launch that lbm_step chooses by itself
 * (1 = never more than one, 2 = step2_kernel pairs, 3/4 = wavefront launches; default 4). */
int lbm_stepn_columns(lbm_t *h, int64_t xa, int64_t xb, int32_t depth, const int64_t *rows);
/* 1 if lbm_stepn_columns is available in the handle's present state: no obstacle links, links that all belong to
 * other slabs, or an obstacle band that stays clear of the slab interfaces (by kHalo columns where peer halos are
 * attached).  A slab run uses multi-update launches only if every rank says 1 (all ranks issue the same groups). */
int lbm_can_stepn(const lbm_t *h);
int lbm_set_temporal_depth(lbm_t *h, int32_t depth);
/* Launch-shape knobs (measurement, tests): "wave_chunk" = columns swept by one block of a
 * wavefront launch (default: by slab size), "wave_tail" = width of the short chunks that end a
 * wavefront launch (-1 auto, 0 uniform chunks), "wave_rows" = strip height (64 | 128), "pf_ahead" = L2
 * prefetch distance of step2_kernel in blocks, "graph" = 0 disables the CUDA-graph replay of
 * lbm_step batches on small lattices (>= 16 updates, <= 2^19 cells, non-default stream), "resident" =
 * resident batches on those lattices (one cooperative launch per run of >= 4 updates, stepr_kernel: the
 * reference's run.py:24-54 loop body executed n times without a kernel boundary): -1 where it is the faster
 * form (lattices with obstacle links; default), 0 never, 1 always; "resident_blocks" = its blocks per SM
 * (0 = default), "resident_timeout_ms" = after how long a block-to-block wait gives up (lbm_sync then fails). */
int lbm_set_tuning(lbm_t *h, const char *key, int64_t value);

/* Stream + (I)BB + Zou-He of the current F with wall row `row`, no collision: materialises the
 * reference's g (nb_col_str stream part + set_bc) in the other buffer, overwrites rho,u on the
 * walls as the nb_zou_he_* kernels do, and stores the forces in slot 0. */
int lbm_apply_bc(lbm_t *h, int64_t row);

/* nb_drag_lift numerator: per obstacle (fx, fy) = sum_k (F_q + g_qbar) c_q.  lbm_get_forces
 * reads slots first..first+n-1 written by the last lbm_step; lbm_forces_now evaluates the
 * current F directly (what lattice.drag_lift sees right after set_bc).  out: [n][n_obs][2]. */
int lbm_get_forces(lbm_t *h, int64_t first, int64_t n, double *out);
/* lbm_get_forces without the wait: the sums of the last lbm_step are reduced and copied to `out_pinned` (page-locked
 * host memory) on the handle's stream; they are there once the stream has passed this point (an event recorded
 * after the call).  Lets the caller enqueue the next batch of updates before the host looks at this one's drag/lift
 * (lbm_b200/run.py: the apps' per-iteration callbacks, turek.py:147-162, overlap the device's next batch).  f64 only. */
int lbm_get_forces_async(lbm_t *h, int64_t first, int64_t n, double *out_pinned);
int lbm_forces_now(lbm_t *h, double *out);

int lbm_get_populations(lbm_t *h, int32_t which, void *host);
/* rho[nxl][ny], u[2][nxl][ny] as stored by the last LBM_STEP_MACRO_LAST update (+ wall
 * overwrites of a later lbm_apply_bc): the composite the reference leaves in lattice.rho/u. */
int lbm_get_macro(lbm_t *h, void *rho_host, void *u_host);

/* |u| of the stored macro fields, speed[nxl][ny] in the handle dtype, with -1 on the cells flagged in
 * solid[nxl][ny] (NULL: no mask): the field plot_norm (lbm/src/plot/plot.py:9-16) draws, computed on
 * the device so that an output step transfers one plane instead of three; equals
 * sqrt(u[0]**2 + u[1]**2) of lbm_get_macro bit for bit. */
int lbm_get_speed(lbm_t *h, const unsigned char *solid_host, void *speed_host);

/* rho, ux, uy along one lattice line of the streamed + boundary-treated current populations,
 * i.e. what lattice.macro() (lattice.py:178-189) of the next iteration yields there; serves the
 * centre-line readers cavity.line_fields (cavity.py:110-133) and poiseuille.compute_error
 * (poiseuille.py:128-152) without a full-field transfer.  axis 0: column x = index (ny values),
 * axis 1: row y = index (nxl values).  out_host: [rho | ux | uy], 3*n elements of the handle dtype. */
int lbm_probe_line(lbm_t *h, int32_t axis, int64_t index, int64_t row, void *out_host);

/* ---- x-slab runs, one process per GPU: halo exchange through peer memory (NVLink) --------------------
 * The reference is one address space; its analogue over several GPUs is a split
 * along x with kHalo = 4 halo columns per side.  Instead of sending halos after a launch, the multi-update
 * kernel's last stage stores its edge columns straight into the neighbour's halo columns:
 *   lbm_peer_export   describes this handle's population buffers and flag words (CUDA IPC handles; the
 *                     buffers must be library-owned, i.e. no lbm_bind_state);
 *   lbm_peer_attach   maps a neighbour's buffers (side 0 = left, x0 - 1; side 1 = right).  From then on
 *                     every lbm_stepn_columns launch also fills that neighbour's halo, and every launch
 *                     first waits (on the device) until both neighbours have signalled the previous group;
 *   lbm_peer_push     copies the kHalo edge columns of the current (which = 0) or other (1) buffer into the
 *                     neighbours' halos of the same buffer -- for launches that do not do it themselves
 *                     (lbm_step_columns, lbm_step2_columns, initial state);
 *   lbm_peer_signal   call once after the launches (and pushes) of one update group, before lbm_flip: tells
 *                     both neighbours (flag words in their memory, release/acquire at system scope) that
 *                     their halos are complete and that this rank is done reading its source buffer.
 * All ranks must issue the same sequence of update groups.  lbm_sync reports a wait that timed out
 * (tuning key "peer_timeout_ms", default 20 s) instead of hanging the device. */
#define LBM_IPC_HANDLE_BYTES 64
typedef struct lbm_peer_info {
    unsigned char mem[2][LBM_IPC_HANDLE_BYTES];  /* cudaIpcMemHandle_t of the two population buffers */
    unsigned char flags[LBM_IPC_HANDLE_BYTES];   /* cudaIpcMemHandle_t of the flag words */
    uint64_t addr[2], flags_addr;                /* the same as addresses, for peers inside one process */
    int64_t pid, device;
    int64_t x0, nxl, origin, plane, pitch, elem_size;
} lbm_peer_info;
int lbm_peer_export(lbm_t *h, lbm_peer_info *out);
int lbm_peer_attach(lbm_t *h, int32_t side, const lbm_peer_info *neighbour);
int lbm_peer_detach(lbm_t *h);
int lbm_peer_push(lbm_t *h, int32_t which);
int lbm_peer_signal(lbm_t *h);

/* Wrap-around 64-bit sum of the bit patterns of the current population array (owned cells, nine planes;
 * f32: the 32-bit patterns, zero-extended): equal arrays give equal sums whatever the slab split, so the
 * per-rank values of a slab run add up (mod 2^64) to the single-GPU value.  Parity aid for lattices that
 * do not fit a host comparison (bench.py parity.state_bits_sum, tests at 32768^2). */
int lbm_state_checksum(lbm_t *h, uint64_t *out);

/* Launch counter: kernels launched by this handle since creation (bench.py gpu_launches). */
int64_t lbm_launch_count(const lbm_t *h);
/* The block plan of a resident batch (stepr_kernel), computed on the host without a device: nx columns, at most
 * max_blocks resident blocks, n_groups link groups with boundary cells in columns [grp_x0[g], grp_x1[g]].  Column block
 * i < *n_col_blocks owns columns [col_a[i], col_a[i+1]), block *n_col_blocks + g is link group g; block b waits for the
 * blocks dep[dep_off[b] .. dep_off[b+1]) before every update.  col_a: room for max_blocks + 1, dep_off: max_blocks + 1,
 * dep: dep_cap entries.  For tests/test_resident_plan_cpu.py, which replays the hand-shake with random block timing
 * and checks that no block reads a column before it was written or overwrites one that is still being read. */
int lbm_resident_plan(int32_t nx, int32_t max_blocks, int32_t n_groups, const int32_t *grp_x0, const int32_t *grp_x1,
                      int32_t *n_col_blocks, int32_t *col_a, int32_t *dep_off, int32_t *dep, int64_t dep_cap);
/* Device time of the updates of the last lbm_step call, from CUDA events recorded on the
 * handle's stream around them (milliseconds; waits for the stream). */
int lbm_last_step_ms(lbm_t *h, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */
