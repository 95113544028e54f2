"""Thin object wrapper of one C-ABI handle (include/lbm_b200.h) for batched runs, bench.py and
the slab driver.  Device buffers come from PyTorch; nothing here touches the CPU oracle."""
import ctypes

import numpy as np

from . import _capi as C


class Solver:
    def __init__(self, nx, ny, tau=None, om_p=None, om_m=None, dtype="f64", arith="fused",
                 right_wall="velocity", device=0, x0=0, nxl=None, stream=None, own_buffers=False):
        import torch
        if not torch.cuda.is_available():
            raise C.LbmError(-2, "no CUDA device: lbm_b200 has no CPU fallback")
        self._torch = torch
        self._L = C.lib()
        if om_p is None:
            # TRT rates from tau, lattice.py:127-131
            tau_m = 0.25 / (tau - 0.5) + 0.5
            om_p, om_m = 1.0 / tau, 1.0 / tau_m
        self.nx, self.ny, self.x0 = int(nx), int(ny), int(x0)
        self.nxl = int(nx if nxl is None else nxl)
        self.dtype = dtype
        self.np_dtype = np.float64 if dtype == "f64" else np.float32
        cfg = C.LbmCfg(nx=self.nx, ny=self.ny, x0=self.x0, nxl=self.nxl, om_p=om_p, om_m=om_m,
                       dtype=C.LBM_F64 if dtype == "f64" else C.LBM_F32,
                       arith=C.LBM_ARITH_STRICT if arith == "strict" else C.LBM_ARITH_FUSED,
                       right_wall=C.LBM_RIGHT_PRESSURE if right_wall == "pressure" else C.LBM_RIGHT_VELOCITY,
                       device=device)
        h = C.c_vp()
        C.check(self._L.lbm_create(ctypes.byref(cfg), ctypes.byref(h)))
        self._h = h
        self.layout = C.LbmLayout()
        C.check(self._L.lbm_get_layout(h, ctypes.byref(self.layout)))
        lay = self.layout
        self.device = torch.device("cuda", device)
        tdt = torch.float64 if dtype == "f64" else torch.float32
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        if s.cuda_stream == 0:           # legacy default stream: no stream capture (CUDA graphs of batches)
            s = torch.cuda.Stream(device=self.device)
        self.stream = s
        C.check(self._L.lbm_set_stream(h, C.c_vp(s.cuda_stream)))
        if own_buffers:
            # the library allocates the two population buffers itself (cudaMalloc): needed for the peer
            # halo exchange, whose CUDA IPC handles must describe whole allocations
            self.buffers = None
        else:
            self.buffers = [torch.empty(lay.elems, dtype=tdt, device=self.device) for _ in range(2)]
            C.check(self._L.lbm_bind_state(h, C.c_vp(self.buffers[0].data_ptr()),
                                           C.c_vp(self.buffers[1].data_ptr()), lay.elems * lay.elem_size))
        self.row_len = int(self._L.lbm_wall_row_len(h))

    # -- lifetime ---------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None:
            self._L.lbm_destroy(self._h)
            self._h = None
            self.buffers = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state ------------------------------------------------------------------------
    @staticmethod
    def _p(a):
        return C.c_vp(a.ctypes.data)

    def set_populations(self, g):
        g = np.ascontiguousarray(g, dtype=self.np_dtype)
        assert g.shape == (9, self.nxl, self.ny)
        C.check(self._L.lbm_set_populations(self._h, self._p(g)))

    def set_post_collision(self, F):
        F = np.ascontiguousarray(F, dtype=self.np_dtype)
        assert F.shape == (9, self.nxl, self.ny)
        C.check(self._L.lbm_set_post_collision(self._h, self._p(F)))

    def init_equilibrium(self, rho=1.0, ux=0.0, uy=0.0):
        C.check(self._L.lbm_init_equilibrium(self._h, rho, ux, uy))

    def set_links(self, obstacles, use_ibb=True):
        if not obstacles:
            C.check(self._L.lbm_set_links(self._h, 0, None, None, None, 0))
            return
        off = np.cumsum([0] + [len(o.boundary) for o in obstacles]).astype(np.int64)
        ijq = np.ascontiguousarray(np.concatenate([np.asarray(o.boundary).reshape(-1, 3) for o in obstacles]), dtype=np.int64)
        ibb = np.ascontiguousarray(np.concatenate([np.asarray(o.ibb).reshape(-1) for o in obstacles]), dtype=np.float64)
        self.n_obs = len(obstacles)
        C.check(self._L.lbm_set_links(self._h, len(obstacles), self._p(off), self._p(ijq),
                                      self._p(ibb) if use_ibb else None, 1 if use_ibb else 0))

    def wall_row(self, u_left=None, u_right=None, u_top=None, u_bot=None, rho_right=None):
        """Pack one wall row: [u_left(2,ny) | u_right(2,ny) | u_top(2,nx) | u_bot(2,nx) | rho_right(ny)]."""
        nx, ny = self.nx, self.ny
        row = np.zeros(self.row_len)
        for arr, a, n in ((u_left, 0, 2 * ny), (u_right, 2 * ny, 2 * ny), (u_top, 4 * ny, 2 * nx),
                          (u_bot, 4 * ny + 2 * nx, 2 * nx), (rho_right, 4 * ny + 4 * nx, ny)):
            if arr is not None:
                row[a:a + n] = np.asarray(arr, dtype=np.float64).reshape(-1)
        return row

    def set_walls(self, rows):
        """rows: (n_rows, row_len) float64 array (numpy) or pinned torch tensor."""
        if hasattr(rows, "data_ptr"):
            n = rows.shape[0] if rows.dim() == 2 else 1
            C.check(self._L.lbm_set_walls(self._h, n, C.c_vp(rows.data_ptr())))
            self._walls_keepalive = rows
        else:
            rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, self.row_len)
            C.check(self._L.lbm_set_walls(self._h, rows.shape[0], self._p(rows)))
            self._walls_keepalive = rows

    def set_wall_profiles(self, u_left=None, u_right=None, u_top=None, u_bot=None, rho_right=None):
        """Base wall profiles (one table row) from the reference's five arrays; None = zeros."""
        def p(a, n):
            if a is None:
                return None, None
            a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
            assert a.size == n
            return a, self._p(a)
        keep = [p(u_left, 2 * self.ny), p(u_right, 2 * self.ny), p(u_top, 2 * self.nx), p(u_bot, 2 * self.nx),
                p(rho_right, self.ny)]
        C.check(self._L.lbm_set_wall_profiles(self._h, *[k[1] for k in keep]))

    def set_ramp(self, ret, it0=0):
        """Per-iteration scale of the wall velocity profiles: `row` arguments become iteration indices
        it0 .. it0+len(ret)-1.  ret: float64 numpy array or pinned torch tensor (asynchronous copy); None
        removes the table."""
        if ret is None:
            C.check(self._L.lbm_set_ramp(self._h, None, 0, 0))
            self._ramp_keepalive = None
        elif hasattr(ret, "data_ptr"):
            C.check(self._L.lbm_set_ramp(self._h, C.c_vp(ret.data_ptr()), int(it0), int(ret.numel())))
            self._ramp_keepalive = ret
        else:
            ret = np.ascontiguousarray(ret, dtype=np.float64).reshape(-1)
            C.check(self._L.lbm_set_ramp(self._h, self._p(ret), int(it0), ret.size))
            self._ramp_keepalive = ret

    # -- peer halo exchange (slab runs) --------------------------------------------------
    def peer_export(self):
        info = C.LbmPeerInfo()
        C.check(self._L.lbm_peer_export(self._h, ctypes.byref(info)))
        return bytes(info)

    def peer_attach(self, side, info_bytes):
        info = C.LbmPeerInfo.from_buffer_copy(info_bytes)
        C.check(self._L.lbm_peer_attach(self._h, int(side), ctypes.byref(info)))

    def peer_detach(self):
        C.check(self._L.lbm_peer_detach(self._h))

    def peer_push(self, which=1):
        C.check(self._L.lbm_peer_push(self._h, int(which)))

    def peer_signal(self):
        C.check(self._L.lbm_peer_signal(self._h))

    def set_right_wall(self, kind):
        C.check(self._L.lbm_set_right_wall(self._h, C.LBM_RIGHT_PRESSURE if kind == "pressure" else C.LBM_RIGHT_VELOCITY))

    # -- stepping ---------------------------------------------------------------------
    def step(self, n=1, first_row=0, row_stride=0, macro_last=False):
        C.check(self._L.lbm_step(self._h, n, first_row, row_stride, C.LBM_STEP_MACRO_LAST if macro_last else 0))

    def step_columns(self, xa, xb, row=0, slot=0):
        C.check(self._L.lbm_step_columns(self._h, xa, xb, row, slot, 0))

    def step2_columns(self, xa, xb, row1=0, row2=0):
        C.check(self._L.lbm_step2_columns(self._h, xa, xb, row1, row2))

    def stepn_columns(self, xa, xb, rows):
        """len(rows) = 2..4 consecutive updates of columns [xa, xb) in one wavefront launch."""
        r = (C.c_i64 * len(rows))(*[int(x) for x in rows])
        C.check(self._L.lbm_stepn_columns(self._h, xa, xb, len(rows), r))

    def can_stepn(self):
        """Whether multi-update launches are available in the present state (obstacle band clear of the slab edges)."""
        return bool(self._L.lbm_can_stepn(self._h))

    def set_temporal_depth(self, depth):
        C.check(self._L.lbm_set_temporal_depth(self._h, int(depth)))

    def set_tuning(self, key, value):
        C.check(self._L.lbm_set_tuning(self._h, key.encode(), int(value)))

    def set_temporal_blocking(self, enable):
        C.check(self._L.lbm_set_temporal_blocking(self._h, int(enable)))

    def flip(self):
        C.check(self._L.lbm_flip(self._h))

    def apply_bc(self, row=0):
        C.check(self._L.lbm_apply_bc(self._h, row))

    def sync(self):
        C.check(self._L.lbm_sync(self._h))

    def last_step_ms(self):
        ms = ctypes.c_float()
        C.check(self._L.lbm_last_step_ms(self._h, ctypes.byref(ms)))
        return float(ms.value)

    def checksum(self):
        """Wrap-around 64-bit sum of the bit patterns of the current population array (owned cells)."""
        v = ctypes.c_uint64()
        C.check(self._L.lbm_state_checksum(self._h, ctypes.byref(v)))
        return int(v.value)

    @property
    def launches(self):
        return int(self._L.lbm_launch_count(self._h))

    # -- results ----------------------------------------------------------------------
    def forces(self, first, n):
        nobs = max(getattr(self, "n_obs", 0), 1)
        out = np.zeros((n, nobs, 2))
        C.check(self._L.lbm_get_forces(self._h, first, n, self._p(out)))
        return out

    def forces_now(self):
        nobs = max(getattr(self, "n_obs", 0), 1)
        out = np.zeros((nobs, 2))
        C.check(self._L.lbm_forces_now(self._h, self._p(out)))
        return out

    def populations(self, which="post_collision"):
        out = np.empty((9, self.nxl, self.ny), dtype=self.np_dtype)
        w = C.LBM_POP_POST_COLLISION if which == "post_collision" else C.LBM_POP_STREAMED
        C.check(self._L.lbm_get_populations(self._h, w, self._p(out)))
        return out

    def macro(self):
        rho = np.empty((self.nxl, self.ny), dtype=self.np_dtype)
        u = np.empty((2, self.nxl, self.ny), dtype=self.np_dtype)
        C.check(self._L.lbm_get_macro(self._h, self._p(rho), self._p(u)))
        return rho, u

    def speed(self, solid=None):
        """|u| of the stored macro fields (-1 where solid != 0), computed on the device."""
        out = np.empty((self.nxl, self.ny), dtype=self.np_dtype)
        m = None if solid is None else np.ascontiguousarray(np.asarray(solid) != 0, dtype=np.uint8)
        C.check(self._L.lbm_get_speed(self._h, None if m is None else self._p(m), self._p(out)))
        return out

    def probe_line(self, axis, index, row=0, out=None):
        """(rho, ux, uy) along column x=index (axis 0) or row y=index (axis 1).  out: optional pinned torch
        tensor of shape (3, n) to receive the line (a direct DMA instead of a staged copy)."""
        n = self.ny if axis == 0 else self.nxl
        if out is not None:
            assert tuple(out.shape) == (3, n) and out.is_contiguous()
            C.check(self._L.lbm_probe_line(self._h, axis, index, row, C.c_vp(out.data_ptr())))
            return out
        out = np.empty((3, n), dtype=self.np_dtype)
        C.check(self._L.lbm_probe_line(self._h, axis, index, row, self._p(out)))
        return out

    def views(self):
        """Torch views [9, nxl+2*halo, pitch] of (current, other) population buffers; column index =
        x + halo (the first and last `halo` columns are the halo)."""
        if self.buffers is None:
            raise C.LbmError(-3, "views() needs torch-owned buffers (own_buffers=False)")
        cur, oth = C.c_vp(), C.c_vp()
        C.check(self._L.lbm_state_ptrs(self._h, ctypes.byref(cur), ctypes.byref(oth)))
        i = 0 if cur.value == self.buffers[0].data_ptr() else 1
        lay = self.layout
        start = lay.origin - lay.halo * lay.pitch

        def v(buf):
            return buf[start:start + 9 * lay.plane].view(9, self.nxl + 2 * lay.halo, lay.pitch)
        return v(self.buffers[i]), v(self.buffers[i ^ 1])
