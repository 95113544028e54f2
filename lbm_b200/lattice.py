"""Drop-in replacement of the reference's ``lattice`` class on the B200 path.

The reference's seam between host logic and kernels is the ``lattice`` object
(/root/reference/lbm/src/core/lattice.py:15-286): ``run.py`` and the apps call
its methods and read/write its array attributes, and every compute method
forwards to one Numba kernel of ``nb.py``.  This class keeps that surface --
same method names, same attributes, same argument meaning -- and forwards to
the C ABI of ``include/lbm_b200.h`` through ctypes.  PyTorch is used for the two
population buffers and the stream handle only.

How the per-phase calls of the reference driver loop (run.py:24-54) map onto the
fused update of the library:

  macro()               launches ONE fused update: stream + (I)BB + Zou-He of the
                        previous iteration (as recorded by the set_bc calls) followed
                        by macro / equilibrium / TRT collision of this iteration.  The
                        very first macro() after ``lattice.g = ...`` is collide-only.
  equilibrium()         no-op after macro() (already done inside the update); before
                        any populations exist (apps' initialize()) it evaluates
                        nb_equilibrium on the host-side rho/u through lbm_equilibrium.
  collision_stream()    no-op marker: opens the boundary-condition recording window.
  zou_he_*(), bounce_back_obstacle()
                        record which boundary treatment runs and SNAPSHOT the wall
                        profile arrays at call time (the next set_inlets overwrites
                        them before the next macro(), run.py:30-33).
  drag_lift()           momentum-exchange sums of the current post-collision array
                        (nb_drag_lift only needs g_up and IBB(g_up), nb.py:64 / 98-104).
  .u / .rho / .g / .g_up / .g_eq
                        host mirrors with the reference's semantics, materialised on
                        demand.

There is no CPU fallback; without the CUDA library or a GPU the constructor raises.
"""
import ctypes
import math
import os
from datetime import datetime

import numpy as np

from . import _capi as C

_REQUIRED_BCS = ("bottom", "left", "top", "right", "bl", "tl", "tr", "br")


class lattice:
    def __init__(self, app, dtype=None, arith=None, device=None, make_dirs=None):
        # --- parameters, same defaults and hasattr ladder as lattice.py:18-78 -------------
        defaults = dict(name="lattice", x_min=0.0, x_max=1.0, y_min=0.0, y_max=1.0, nx=100, ny=100,
                        tau_lbm=1.0, dx=1.0, dt=1.0, dpi=100, u_lbm=0.03, L_lbm=100, nu_lbm=0.01,
                        Re_lbm=100.0, rho_lbm=1.0, IBB=False, stop="it", t_max=1.0, it_max=1000,
                        obs_cv_ct=1.0e-1, obs_cv_nb=500)
        for k, v in defaults.items():
            setattr(self, k, getattr(app, k, v))
        self.Cx = getattr(app, "Cx", self.dx)
        self.Ct = getattr(app, "Ct", self.dt)
        self.Cr = getattr(app, "Ct", 1.0) if hasattr(app, "Cr") else 1.0   # sic, lattice.py:63
        self.Cn = getattr(app, "Cn", self.Cx ** 2 / self.Ct)
        self.Cu = getattr(app, "Cu", self.Cx / self.Ct)
        self.Cf = getattr(app, "Cf", self.Cr * self.Cx ** 2 / self.Ct)
        # --- B200 options: keyword, else attribute on the app, else default ----------------
        self.dtype = dtype or getattr(app, "lbm_dtype", "f64")
        self.arith = arith or getattr(app, "lbm_arith", "fused")
        self.device = device if device is not None else getattr(app, "lbm_device", 0)
        if self.dtype not in ("f64", "f32"):
            raise ValueError("dtype must be 'f64' or 'f32'")
        if self.arith not in ("fused", "strict"):
            raise ValueError("arith must be 'fused' or 'strict'")
        self._np = np.float64 if self.dtype == "f64" else np.float32
        # --- output directories (lattice.py:80-91) -----------------------------------------
        if make_dirs is None:
            make_dirs = getattr(app, "lbm_make_dirs", True)
        stamp = datetime.now().strftime("%Y-%m-%d_%H_%M_%S")
        self.results_dir = "./results/"
        self.output_dir = self.results_dir + stamp + "/"
        self.png_dir = self.output_dir + "./png/"
        if make_dirs:
            os.makedirs(self.png_dir, exist_ok=True)
        self._set_default_lbm()
        self._open_device()

    # ------------------------------------------------------------------------------------
    def _set_default_lbm(self):
        """Constants and host arrays of lattice.py:117-174."""
        self.output_it = 0
        self.lx, self.ly, self.q = self.nx - 1, self.ny - 1, 9
        self.Cs = 1.0 / math.sqrt(3.0)
        self.tau_p_lbm = self.tau_lbm
        self.lambda_trt = 1.0 / 4.0
        self.tau_m_lbm = self.lambda_trt / (self.tau_p_lbm - 0.5) + 0.5
        self.om_p_lbm = 1.0 / self.tau_p_lbm
        self.om_m_lbm = 1.0 / self.tau_m_lbm
        self.om_lbm = 1.0 / self.tau_lbm
        self.c = np.array([[0, 0], [1, 0], [-1, 0], [0, 1], [0, -1], [1, 1], [-1, -1], [-1, 1], [1, -1]])
        self.w = np.array([4. / 9.] + [1. / 9.] * 4 + [1. / 36.] * 4)
        self.ns = np.array([0, 2, 1, 4, 3, 6, 5, 8, 7])
        nx, ny = self.nx, self.ny
        self.u_left = np.zeros((2, ny))
        self.u_right = np.zeros((2, ny))
        self.u_top = np.zeros((2, nx))
        self.u_bot = np.zeros((2, nx))
        self.rho_right = np.zeros(ny)
        self.lattice = np.zeros((nx, ny))
        self._rho_host = np.ones((nx, ny))
        self._u_host = np.zeros((2, nx, ny))
        self._g_eq_host = None

    def _open_device(self):
        import torch
        if not torch.cuda.is_available():
            raise C.LbmError(-2, "no CUDA device: lbm_b200 has no CPU fallback")
        L = C.lib()
        self._L = L
        self._torch = torch
        cfg = C.LbmCfg(nx=self.nx, ny=self.ny, x0=0, nxl=self.nx, om_p=self.om_p_lbm,
                       om_m=self.om_m_lbm, dtype=C.LBM_F64 if self.dtype == "f64" else C.LBM_F32,
                       arith=C.LBM_ARITH_STRICT if self.arith == "strict" else C.LBM_ARITH_FUSED,
                       right_wall=C.LBM_RIGHT_VELOCITY, device=self.device)
        self._cfg = cfg
        self._h = None
        self._right_wall = None         # decided by the first recorded right-wall call
        self._state = "fresh"           # fresh | g | macro_done | streamed
        self._bcs = set()
        self._obstacles = []            # obstacles recorded by bounce_back_obstacle, in call order
        self._links_key = None
        self._row = np.zeros(5 * self.ny + 4 * self.nx)
        self._row_dev = None            # copy of the row that is on the device
        self._cache = {}
        self._replay = None             # per-obstacle (fx, fy) of the iteration being replayed
        self._link_obstacles = []
        self.updates = 0                # fused updates executed
        self._base_dev = None           # base row of the ramp form that is on the device (batch follows batch)
        self._pipe_slot = 1             # pinned ramp / force buffers of the two batches in flight
        self._pipe_pool = [None, None]
        self._saved = {}                # slot -> [device copy of the populations, update count]

    def _create(self, right_wall):
        torch = self._torch
        self._cfg.right_wall = right_wall
        h = C.c_vp()
        C.check(self._L.lbm_create(ctypes.byref(self._cfg), ctypes.byref(h)))
        self._h = h
        self._right_wall = right_wall
        lay = C.LbmLayout()
        C.check(self._L.lbm_get_layout(h, ctypes.byref(lay)))
        self.layout = lay
        tdt = torch.float64 if self.dtype == "f64" else torch.float32
        dev = torch.device("cuda", self.device)
        self._buf = [torch.empty(lay.elems, dtype=tdt, device=dev) for _ in range(2)]
        # a stream of its own (not the legacy default stream): batches of updates on these small
        # lattices are replayed as CUDA graphs, and stream capture needs a real stream
        self._stream = torch.cuda.current_stream(dev)
        if self._stream.cuda_stream == 0:
            self._stream = torch.cuda.Stream(device=dev)
        C.check(self._L.lbm_set_stream(h, C.c_vp(self._stream.cuda_stream)))
        C.check(self._L.lbm_bind_state(h, C.c_vp(self._buf[0].data_ptr()), C.c_vp(self._buf[1].data_ptr()),
                                       lay.elems * lay.elem_size))

    def _handle(self):
        if self._h is None:
            self._create(C.LBM_RIGHT_VELOCITY)
        return self._h

    def close(self):
        if getattr(self, "_h", None) is not None:
            self._L.lbm_destroy(self._h)
            self._h = None
            self._buf = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------
    # array attributes
    # ------------------------------------------------------------------------------------
    def _ptr(self, a):
        return C.c_vp(a.ctypes.data)

    def _apply_bc(self):
        """Materialise stream + recorded BCs of the current post-collision array (other buffer)."""
        if "bc" in self._cache:
            return
        self._macro_arrays(walls=False)      # keep the pure macro() fields before walls overwrite them
        self._need_all_bcs()
        self._push_walls()
        self._push_links()
        C.check(self._L.lbm_apply_bc(self._handle(), 0))
        self._cache["bc"] = True

    @property
    def g(self):
        if self._state == "fresh":
            raise AttributeError("lattice.g has not been set yet")
        if "g" not in self._cache:
            h = self._handle()
            if self._state == "streamed":
                self._apply_bc()
            elif self._state == "macro_done":
                raise C.LbmError(-3, "lattice.g between macro() and collision_stream() is not kept "
                                     "on the fused path (the update has already collided it)")
            out = np.empty((9, self.nx, self.ny), dtype=self._np)
            C.check(self._L.lbm_get_populations(h, C.LBM_POP_STREAMED, self._ptr(out)))
            self._cache["g"] = out
        return self._cache["g"]

    @g.setter
    def g(self, value):
        arr = np.ascontiguousarray(value, dtype=self._np)
        if arr.shape != (9, self.nx, self.ny):
            raise ValueError("g must have shape (9, nx, ny)")
        C.check(self._L.lbm_set_populations(self._handle(), self._ptr(arr)))
        self._state = "g"
        self._cache = {}
        self._bcs = set()

    @property
    def g_up(self):
        if self._state in ("fresh", "g"):
            return np.zeros((9, self.nx, self.ny), dtype=self._np)
        if "g_up" not in self._cache:
            out = np.empty((9, self.nx, self.ny), dtype=self._np)
            C.check(self._L.lbm_get_populations(self._handle(), C.LBM_POP_POST_COLLISION, self._ptr(out)))
            self._cache["g_up"] = out
        return self._cache["g_up"]

    @property
    def g_eq(self):
        if self._state in ("fresh", "g"):
            if self._g_eq_host is None:
                return np.zeros((9, self.nx, self.ny), dtype=self._np)
            return self._g_eq_host
        if "g_eq" not in self._cache:
            rho, u = self._macro_arrays(walls=False)
            self._cache["g_eq"] = self._equilibrium_of(rho, u)
        return self._cache["g_eq"]

    def _macro_arrays(self, walls):
        """(rho, u) host copies; walls=True adds the Zou-He overwrites of the recorded BCs."""
        key = "macro_w" if walls else "macro"
        if key not in self._cache:
            if walls:
                self._apply_bc()
            rho = np.empty((self.nx, self.ny), dtype=self._np)
            u = np.empty((2, self.nx, self.ny), dtype=self._np)
            C.check(self._L.lbm_get_macro(self._handle(), self._ptr(rho), self._ptr(u)))
            self._cache[key] = (rho, u)
        return self._cache[key]

    def _fields(self):
        if self._state in ("fresh", "g"):
            return self._rho_host, self._u_host
        # after set_bc the reference has overwritten the wall entries (nb.py:127-131 ...)
        walls = self._state == "streamed" and len(self._bcs) > 0
        return self._macro_arrays(walls)

    def speed(self):
        """|u| with -1 on obstacle nodes -- the array plot_norm (plot.py:12-15) builds from lattice.u and
        lattice.lattice -- computed on the device from the fields of the last macro(): one plane to the
        host instead of three.  (After set_bc the wall rows of lattice.u carry the Zou-He values; use
        lattice.u there.)"""
        if self._state in ("fresh", "g"):
            v = np.sqrt(self._u_host[0] ** 2 + self._u_host[1] ** 2)
            v[np.where(self.lattice > 0.0)] = -1.0
            return v
        out = np.empty((self.nx, self.ny), dtype=self._np)
        m = np.ascontiguousarray(self.lattice > 0.0, dtype=np.uint8)
        C.check(self._L.lbm_get_speed(self._handle(), self._ptr(m), self._ptr(out)))
        return out

    @property
    def rho(self):
        return self._fields()[0]

    @rho.setter
    def rho(self, value):
        if self._state in ("fresh", "g"):
            self._rho_host = np.asarray(value, dtype=np.float64).reshape(self.nx, self.ny)

    @property
    def u(self):
        return self._fields()[1]

    @u.setter
    def u(self, value):
        if self._state in ("fresh", "g"):
            self._u_host = np.asarray(value, dtype=np.float64).reshape(2, self.nx, self.ny)

    # ------------------------------------------------------------------------------------
    # phases of the driver loop
    # ------------------------------------------------------------------------------------
    def _equilibrium_of(self, rho, u):
        out = np.empty((9, self.nx, self.ny), dtype=self._np)
        r = np.ascontiguousarray(rho, dtype=self._np)
        v = np.ascontiguousarray(u, dtype=self._np)
        C.check(self._L.lbm_equilibrium(self._handle(), self._ptr(r), self._ptr(v), self._ptr(out)))
        return out

    def equilibrium(self):
        """nb_equilibrium (lattice.py:193-195).  Inside the loop it is part of macro()'s update."""
        if self._state in ("fresh", "g"):
            self._g_eq_host = self._equilibrium_of(self._rho_host, self._u_host)

    def macro(self):
        """lattice.macro (lattice.py:178-189) -- executes the fused update, see module docstring."""
        h = self._handle()
        if self._state == "macro_done":
            return                      # macro() twice without collision_stream(): same fields
        if self._state == "fresh":
            raise C.LbmError(-3, "macro() before lattice.g was set")
        if self._state == "streamed":
            self._need_all_bcs()
            self._push_walls()
            self._push_links()
        C.check(self._L.lbm_step(h, 1, 0, 0, C.LBM_STEP_MACRO_LAST))
        self.updates += 1
        self._state = "macro_done"
        self._cache = {}

    def collision_stream(self):
        """nb_col_str (lattice.py:199-205): already executed by macro(); opens the BC window."""
        if self._state != "macro_done":
            raise C.LbmError(-3, "collision_stream() must follow macro()")
        self._state = "streamed"
        self._bcs = set()
        self._obstacles = []
        self._cache = {}

    def _need_all_bcs(self):
        missing = [b for b in _REQUIRED_BCS if b not in self._bcs]
        if missing:
            raise C.LbmError(-5, "the fused update needs all four walls and four corners to be "
                                 "applied every iteration (missing: %s)" % ", ".join(missing))

    def _record(self, name):
        if self._state != "streamed":
            raise C.LbmError(-3, "boundary conditions must follow collision_stream()")
        self._bcs.add(name)
        self._cache = {}

    def zou_he_left_wall_velocity(self):
        self._record("left")
        self._row[0:2 * self.ny] = self.u_left.reshape(-1)

    def zou_he_right_wall_velocity(self):
        self._record("right")
        self._switch_right(C.LBM_RIGHT_VELOCITY)
        self._row[2 * self.ny:4 * self.ny] = self.u_right.reshape(-1)

    def zou_he_right_wall_pressure(self):
        self._record("right")
        self._switch_right(C.LBM_RIGHT_PRESSURE)
        self._row[2 * self.ny:4 * self.ny] = self.u_right.reshape(-1)
        self._row[4 * self.ny + 4 * self.nx:] = self.rho_right

    def zou_he_top_wall_velocity(self):
        self._record("top")
        self._row[4 * self.ny:4 * self.ny + 2 * self.nx] = self.u_top.reshape(-1)

    def zou_he_bottom_wall_velocity(self):
        self._record("bottom")
        self._row[4 * self.ny + 2 * self.nx:4 * self.ny + 4 * self.nx] = self.u_bot.reshape(-1)

    def zou_he_bottom_left_corner(self):
        self._record("bl")

    def zou_he_top_left_corner(self):
        self._record("tl")

    def zou_he_top_right_corner(self):
        self._record("tr")

    def zou_he_bottom_right_corner(self):
        self._record("br")

    def _switch_right(self, kind):
        """Velocity or pressure variant on the right side: a per-update parameter of the library."""
        if self._right_wall != kind:
            C.check(self._L.lbm_set_right_wall(self._handle(), kind))
            self._right_wall = kind
            self._cache = {}

    def bounce_back_obstacle(self, obstacle):
        """nb_bounce_back_obstacle (lattice.py:218-222): recorded, executed by the next update."""
        if self._state != "streamed":
            raise C.LbmError(-3, "boundary conditions must follow collision_stream()")
        self._obstacles.append(obstacle)
        self._cache = {}

    def _push_walls(self):
        if self._row_dev is None or not np.array_equal(self._row, self._row_dev):
            # (pageable host memory: the copy has left the host buffer when the call returns -- no stream sync)
            C.check(self._L.lbm_set_walls(self._handle(), 1, self._ptr(self._row)))
            self._row_dev = self._row.copy()
            self._base_dev = None

    def _push_links(self):
        obs = self._obstacles
        key = (bool(self.IBB),) + tuple((id(o), id(o.boundary), len(o.boundary)) for o in obs)
        if key == self._links_key:
            return
        h = self._handle()
        if not obs:
            C.check(self._L.lbm_set_links(h, 0, None, None, None, 0))
        else:
            off = np.cumsum([0] + [len(o.boundary) for o in obs]).astype(np.int64)
            ijq = np.ascontiguousarray(np.concatenate([np.asarray(o.boundary).reshape(-1, 3) for o in obs]),
                                       dtype=np.int64)
            if self.IBB:
                ibb = np.ascontiguousarray(np.concatenate([np.asarray(o.ibb).reshape(-1) for o in obs]),
                                           dtype=np.float64)
                C.check(self._L.lbm_set_links(h, len(obs), self._ptr(off), self._ptr(ijq), self._ptr(ibb), 1))
            else:
                C.check(self._L.lbm_set_links(h, len(obs), self._ptr(off), self._ptr(ijq), None, 0))
        self._links_key = key
        self._link_obstacles = list(obs)

    def drag_lift(self, obs, R_ref, U_ref, L_ref):
        """nb_drag_lift (lattice.py:209-214, nb.py:49-73)."""
        if self._replay is not None:
            # batched run (run.py): sums of this iteration were stored by the batch's updates
            k = [i for i, o in enumerate(self._link_obstacles) if o is obs][0]
            fx, fy = float(self._replay[k, 0]), float(self._replay[k, 1])
            return (-2.0 * fx / (R_ref * L_ref * U_ref ** 2), -2.0 * fy / (R_ref * L_ref * U_ref ** 2))
        if self._state != "streamed":
            raise C.LbmError(-3, "drag_lift() must follow set_bc")
        if not any(o is obs for o in self._obstacles):
            # the reference can evaluate any obstacle; here it must have been bounced back
            self._obstacles.append(obs)
        self._push_links()
        k = [i for i, o in enumerate(self._link_obstacles) if o is obs][0]
        out = np.zeros(2 * len(self._link_obstacles))
        C.check(self._L.lbm_forces_now(self._handle(), self._ptr(out)))
        fx, fy = float(out[2 * k]), float(out[2 * k + 1])
        Cx = -2.0 * fx / (R_ref * L_ref * U_ref ** 2)
        Cy = -2.0 * fy / (R_ref * L_ref * U_ref ** 2)
        return Cx, Cy

    # ------------------------------------------------------------------------------------
    # batched execution (used by lbm_b200.run.run)
    # ------------------------------------------------------------------------------------
    def snapshot_walls(self):
        """The wall row the zou_he_* calls of one set_bc would record from the current arrays."""
        row = np.empty_like(self._row)
        nx, ny = self.nx, self.ny
        row[0:2 * ny] = self.u_left.reshape(-1)
        row[2 * ny:4 * ny] = self.u_right.reshape(-1)
        row[4 * ny:4 * ny + 2 * nx] = self.u_top.reshape(-1)
        row[4 * ny + 2 * nx:4 * ny + 4 * nx] = self.u_bot.reshape(-1)
        row[4 * ny + 4 * nx:] = self.rho_right
        return row

    def batch_updates(self, rows):
        """len(rows) fused updates in one library call; update k uses wall row rows[k].  Must follow a
        completed set_bc (state 'streamed').  Leaves the lattice in the state macro() leaves it in, with
        rho/u of the LAST update stored.  Returns the momentum-exchange sums [n, n_obs, 2] (slot k =
        iteration preceding update k)."""
        if self._state != "streamed":
            raise C.LbmError(-3, "batch_updates() must follow set_bc")
        self._need_all_bcs()
        self._push_links()
        h = self._handle()
        rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, self._row.size)
        n = rows.shape[0]
        C.check(self._L.lbm_set_walls(h, n, self._ptr(rows)))
        C.check(self._L.lbm_sync(h))
        self._row_dev = None
        self._base_dev = None
        C.check(self._L.lbm_step(h, n, 0, 1, C.LBM_STEP_MACRO_LAST))
        nobs = max(len(self._link_obstacles), 1)
        forces = np.zeros((n, nobs, 2))
        C.check(self._L.lbm_get_forces(h, 0, n, self._ptr(forces)))
        self.updates += n
        self._state = "macro_done"
        self._cache = {}
        return forces

    def batch_updates_ramp(self, base_row, scales):
        """batch_updates for wall rows of the form  velocity entries of base_row x scales[k]  (the apps' inlet
        ramp): the base row is uploaded once, a batch costs 8 bytes of host input per update (lbm_set_ramp)."""
        if self._state != "streamed":
            raise C.LbmError(-3, "batch_updates() must follow set_bc")
        self._need_all_bcs()
        self._push_links()
        h = self._handle()
        base_row = np.ascontiguousarray(base_row, dtype=np.float64).reshape(-1)
        scales = np.ascontiguousarray(scales, dtype=np.float64).reshape(-1)
        n = scales.size
        C.check(self._L.lbm_set_walls(h, 1, self._ptr(base_row)))        # (one row; the per-phase calls reuse the slot)
        self._row_dev = None
        self._base_dev = base_row.copy()
        C.check(self._L.lbm_set_ramp(h, self._ptr(scales), 0, n))
        C.check(self._L.lbm_step(h, n, 0, 1, C.LBM_STEP_MACRO_LAST))
        nobs = max(len(self._link_obstacles), 1)
        forces = np.zeros((n, nobs, 2))
        C.check(self._L.lbm_get_forces(h, 0, n, self._ptr(forces)))      # waits for the batch: `scales` may go
        C.check(self._L.lbm_set_ramp(h, None, 0, 0))                     # back to plain rows for the per-phase calls
        self.updates += n
        self._state = "macro_done"
        self._cache = {}
        return forces

    # -- the same batch in two halves: enqueue now, look at the drag/lift sums later.  run.py enqueues the NEXT batch in
    # between, so that the apps' per-iteration callbacks of one batch run on the host while the device executes the
    # next one (whole runs of the reference's small cases are otherwise the SUM of device and host time).
    def can_pipeline(self):
        return self.dtype == "f64"          # (f32 storage: lbm_get_forces adds a host-side constant, no asynchronous form)

    def batch_enqueue_ramp(self, base_row, scales):
        """batch_updates_ramp without the wait; returns a token for batch_result().  At most two batches in flight."""
        if self._state != "streamed":
            raise C.LbmError(-3, "batch_updates() must follow set_bc")
        self._need_all_bcs()
        self._push_links()
        h = self._handle()
        torch = self._torch
        scales = np.ascontiguousarray(scales, dtype=np.float64).reshape(-1)
        n = scales.size
        nobs = max(len(self._link_obstacles), 1)
        slot = self._pipe_slot = self._pipe_slot ^ 1
        pool = self._pipe_pool
        if pool[slot] is None or pool[slot][0].numel() < n or pool[slot][1].shape[1] != nobs or pool[slot][1].shape[0] < n:
            pool[slot] = (torch.empty(max(n, 64), dtype=torch.float64).pin_memory(),
                          torch.empty((max(n, 64), nobs, 2), dtype=torch.float64).pin_memory())
        ramp, forces = pool[slot]
        ramp.numpy()[:n] = scales                # (page-locked: the copies below do not wait for the stream)
        base_row = np.ascontiguousarray(base_row, dtype=np.float64).reshape(-1)
        if self._row_dev is not None or self._base_dev is None or not np.array_equal(self._base_dev, base_row):
            C.check(self._L.lbm_set_walls(h, 1, self._ptr(base_row)))    # (not again when batch follows batch)
            self._base_dev = base_row.copy()
        self._row_dev = None
        C.check(self._L.lbm_set_ramp(h, C.c_vp(ramp.data_ptr()), 0, n))
        C.check(self._L.lbm_step(h, n, 0, 1, C.LBM_STEP_MACRO_LAST))
        C.check(self._L.lbm_get_forces_async(h, 0, n, C.c_vp(forces.data_ptr())))
        ev = torch.cuda.Event()
        ev.record(self._stream)
        C.check(self._L.lbm_set_ramp(h, None, 0, 0))
        self.updates += n
        self._state = "macro_done"
        self._cache = {}
        return (ev, forces, n)

    def batch_result(self, token):
        """[n, n_obs, 2] momentum-exchange sums of an enqueued batch (waits for that batch only)."""
        ev, forces, n = token
        ev.synchronize()
        return forces.numpy()[:n].copy()

    def forces_now(self):
        """[n_obs, 2] momentum-exchange sums of the current post-collision array."""
        self._push_links()
        out = np.zeros((max(len(self._link_obstacles), 1), 2))
        C.check(self._L.lbm_forces_now(self._handle(), self._ptr(out)))
        return out

    def save_checkpoint(self, path):
        """Write the restartable state (post-collision populations, recorded wall profiles, iteration
        count) to an .npz file.  Valid after a completed set_bc."""
        if self._state != "streamed":
            raise C.LbmError(-3, "save_checkpoint() must follow set_bc")
        obs = self._obstacles
        off = np.cumsum([0] + [len(o.boundary) for o in obs]).astype(np.int64)
        ijq = (np.concatenate([np.asarray(o.boundary).reshape(-1, 3) for o in obs]).astype(np.int64)
               if obs else np.zeros((0, 3), dtype=np.int64))
        ibb = (np.concatenate([np.asarray(o.ibb, dtype=np.float64).reshape(-1) for o in obs])
               if obs else np.zeros(0))
        np.savez(path, F=self.g_up, row=self._row, updates=self.updates, right_wall=self._right_wall,
                 bcs=np.array(sorted(self._bcs)), dtype=self.dtype, link_offsets=off, link_ijq=ijq, link_ibb=ibb,
                 link_tags=np.array([getattr(o, "tag", k + 1) for k, o in enumerate(obs)], dtype=np.int64),
                 use_ibb=bool(self.IBB))

    def load_checkpoint(self, path, obstacles=None):
        """Restore a state written by save_checkpoint; continue with app.set_inlets / lattice.macro()
        of the next iteration.  The obstacle link lists are part of the checkpoint: the first update
        after loading bounces back on them exactly as the uninterrupted run does (the next set_bc
        re-records the app's own obstacle objects).  obstacles: the app's obstacle objects in bounce-back
        order, if drag_lift() is to find them by identity before the next set_bc."""
        z = np.load(path)
        F = np.ascontiguousarray(z["F"], dtype=self._np)
        if F.shape != (9, self.nx, self.ny):
            raise ValueError("checkpoint has shape %s, lattice is (9, %d, %d)" % (F.shape, self.nx, self.ny))
        C.check(self._L.lbm_set_post_collision(self._handle(), self._ptr(F)))
        self._switch_right(int(z["right_wall"]))
        self._row[:] = z["row"]
        self._row_dev = None
        self._bcs = set(str(b) for b in z["bcs"])
        self.updates = int(z["updates"])
        self._state = "streamed"
        self._cache = {}
        if "link_offsets" in z.files:
            off = z["link_offsets"]
            if obstacles is None:
                class _Obs:          # minimal obstacle record (obstacle.py:3-25): the two kernel inputs
                    pass
                obstacles = []
                for k in range(len(off) - 1):
                    o = _Obs()
                    o.boundary = z["link_ijq"][off[k]:off[k + 1]]
                    o.ibb = z["link_ibb"][off[k]:off[k + 1]]
                    o.tag = int(z["link_tags"][k])
                    obstacles.append(o)
            elif [len(o.boundary) for o in obstacles] != list(np.diff(off)):
                raise ValueError("obstacles do not match the link lists of the checkpoint")
            self.IBB = bool(z["use_ibb"])
            self._obstacles = list(obstacles)
            self._links_key = None
        elif obstacles is not None:
            self._obstacles = list(obstacles)
            self._links_key = None

    def save_state(self, slot=0):
        """Device-side copy of the current post-collision populations (for exact stop-rule rollback); slot 0 / 1:
        the pipelined driver keeps the starting states of two batches."""
        cur = C.c_vp()
        C.check(self._L.lbm_state_ptrs(self._handle(), ctypes.byref(cur), None))
        src = self._buf[0] if cur.value == self._buf[0].data_ptr() else self._buf[1]
        saved = self._saved
        if slot not in saved:
            saved[slot] = [self._torch.empty_like(src), 0]
        # on the library's stream: ordered after the updates already enqueued and before the next ones
        # (a copy on torch's current stream would race with them: the handle's stream is non-blocking)
        with self._torch.cuda.stream(self._stream):
            saved[slot][0].copy_(src)
        saved[slot][1] = self.updates

    def restore_state(self, slot=0):
        cur = C.c_vp()
        C.check(self._L.lbm_state_ptrs(self._handle(), ctypes.byref(cur), None))
        dst = self._buf[0] if cur.value == self._buf[0].data_ptr() else self._buf[1]
        with self._torch.cuda.stream(self._stream):
            dst.copy_(self._saved[slot][0])
        self.updates = self._saved[slot][1]
        self._cache = {}

    # ------------------------------------------------------------------------------------
    # host helpers of the reference class that the apps call
    # ------------------------------------------------------------------------------------
    def get_coords(self, i, j):
        """lattice.py:379-387."""
        dx = (self.x_max - self.x_min) / (self.nx - 1)
        dy = (self.y_max - self.y_min) / (self.ny - 1)
        return [self.x_min + i * dx, self.y_min + j * dy]

    def add_obstacle(self, obstacle):
        """lattice.add_obstacle (lattice.py:290-375): host preprocessing, see geometry.py."""
        from . import geometry
        print('# Obstacle ', str(obstacle.tag))
        area, bnd, ibb = geometry.add_obstacle(self, obstacle)
        print('# ' + str(int(np.count_nonzero(self.lattice == obstacle.tag))) + ' locations in obstacle')
        print('# ' + str(bnd.shape[0]) + ' locations on boundary')
        print('# Area = ' + '{:f}'.format(area))
        print('')
        return area, bnd, ibb

    def is_inside(self, poly, pt):
        """lattice.is_inside (lattice.py:391-414) for one point."""
        from . import geometry
        return bool(geometry.inside_polygon(np.asarray(poly), np.array([pt[0]]), np.array([pt[1]]))[0])

    def generate_image(self, obstacles):
        """Output writer of the reference (lattice.py:418-436): out of scope, see DESIGN.md."""
        return None
