"""Host-side mirror of the reference's ``Env2DAirfoil`` with the per-action step on device.

Same contract as /root/reference/Env2DAirfoil.py: ``Env2DAirfoil(config)``, ``reset()``,
``get_state() -> Data``, ``step(action) -> (state, rew, terminal, {})``, ``calculate_reward()``,
``return_vals()``, ``set_plot_dir(dir)`` and the attributes the drivers read (``action_space.n``,
``N_CLOSEST``, ``coord_map``, ``inv_coord_map``, ``n_closest``, ``removable``, ``velocities``,
``pressures``, ``new_drags``, ``new_lifts``, ``gt_drag``, ``gt_lift``, ``gt_time``, ``steps``,
``terminal``, ``do_nothing_offset``, ``flow_solver``).

What runs where (one episode step, Env2DAirfoil.py:318-377):
  host   : action bookkeeping; ``scipy.spatial.Delaunay`` -- the same third-party Qhull call the
           reference makes (:487) -- and the all-boundary-cell filter (:496); scalar reward formula (:406-428)
  device : mesh topology, ``smooth(50)``, facet tags, removable mask (flow_solver.remesh), point
           location + P2/P1 re-interpolation of all T snapshots from the ORIGINAL mesh (:556-590),
           vertex sampling (:515-522), airfoil distances + N-closest + state graph/features (:220-315),
           drag/lift (:390-394)
Fields are plain nodal arrays instead of DOLFIN Functions: ``agent_params['u']`` = U [T, V0+E0, 2]
(P2: vertex dofs then edge-midpoint dofs, edges in lexicographic (a<b) order), ``agent_params['p']``
= P [T, V0].  The one-time Navier-Stokes solve that produces them stays in FEniCS (out of scope):
with ``u == -1`` the constructor raises unless ``agent_params['synthetic_fields']`` gives a seed.
Result-affecting quirks B1-B8 of SURVEY.md Appendix B are reproduced on purpose.
"""
from __future__ import annotations

import ctypes
import math
import os

import numpy as np
import torch
from scipy.spatial import Delaunay

from . import _lib
from .data import Data
from .flow_solver import DeviceMesh, FlowSolver, _dev
from .probes import drag_lift_device

if torch.cuda.is_available():
    device = torch.device("cuda:0")
else:
    device = torch.device("cpu")


class Discrete:
    """Stand-in for ``gym.spaces.Discrete`` (Env2DAirfoil.py:68): only ``.n`` and ``sample()`` are used."""

    def __init__(self, n):
        self.n = int(n)

    def sample(self):
        return int(np.random.randint(self.n))


class SourceField:
    """The ORIGINAL mesh M0 with its T snapshots and the uniform-grid index used for point location.

    HBM layout: U0 f64 [T][V0+E0][2] (both components of one dof adjacent -> one 16 B load),
    P0 f64 [T][V0]; grid bins hold the ids of the cells whose inflated bounding box overlaps them.
    """

    TILED_MIN_CELLS = 16384   # below this the uniform-grid kernel (one launch) has the lower latency

    def __init__(self, mesh: DeviceMesh, U0, P0, bins_per_cell=4.0, tiled=None, leaf_cells=None, micro_bins_per_cell=4.0,
                 bucket_factor=2):
        self.mesh = mesh
        self.tiled = (mesh.nc >= self.TILED_MIN_CELLS) if tiled is None else bool(tiled)
        self.auto_leaf = leaf_cells is None
        self.leaf_cells = 128 if leaf_cells is None else int(leaf_cells)
        self.micro_bins_per_cell = float(micro_bins_per_cell)
        self.bucket_factor = int(bucket_factor)
        self.tile = None
        self._ticket = None
        d = mesh.device
        self.U0 = torch.as_tensor(np.ascontiguousarray(U0, dtype=np.float64)).to(d).contiguous()
        self.P0 = torch.as_tensor(np.ascontiguousarray(P0, dtype=np.float64)).to(d).contiguous()
        self.T = int(self.U0.shape[0])
        if self.U0.shape[1] != mesh.nv + mesh.ne or self.P0.shape[1] != mesh.nv:
            raise ValueError(f"snapshot sizes {tuple(self.U0.shape)} / {tuple(self.P0.shape)} do not match the mesh "
                             f"(V0={mesh.nv}, E0={mesh.ne})")
        xy = mesh.coordinates()
        lo, hi = xy.min(0), xy.max(0)
        w, h = float(hi[0] - lo[0]), float(hi[1] - lo[1])
        cell = math.sqrt(max(w * h, 1e-300) / max(1.0, bins_per_cell * mesh.nc))
        gx = max(1, min(4096, int(math.ceil(w / cell))))
        gy = max(1, min(4096, int(math.ceil(h / cell))))
        self.h_grid = (ctypes.c_double * 6)(float(lo[0]), float(lo[1]), gx / w if w > 0 else 1.0,
                                            gy / h if h > 0 else 1.0, float(gx), float(gy))
        self.gx, self.gy = gx, gy
        nbins = gx * gy
        L = _lib.lib()
        p = _lib.ptr
        cnt = torch.empty(nbins + 1, dtype=torch.int32, device=d)
        self.bin_ptr = torch.empty(nbins + 2, dtype=torch.int32, device=d)
        with torch.cuda.device(d):
            st = _lib.stream_ptr()
            _lib.check(L.mdq_grid_count(p(mesh.coords), p(mesh.cells), mesh.nc, self.h_grid, p(cnt), st), "mdq_grid_count")
            _lib.check(L.mdq_scan_i32(p(cnt), p(self.bin_ptr), nbins, st), "mdq_scan_i32")
            total = int(self.bin_ptr[nbins].item())
            self.bin_cells = torch.empty(max(total, 1), dtype=torch.int32, device=d)
            _lib.check(L.mdq_grid_fill(p(mesh.coords), p(mesh.cells), mesh.nc, self.h_grid, p(self.bin_ptr), p(cnt),
                                       p(self.bin_cells), st), "mdq_grid_fill")
        self.n_bin_entries = total
        if self.tiled:
            self._build_tiles(U0, P0)

    def _build_tiles(self, U0, P0):
        """Leaf-packed copy of M0 for the tiled kernel (tile_index.py); built once, M0 never changes (quirk B6)."""
        from .tile_index import build_tile_index
        m, d = self.mesh, self.mesh.device
        L = _lib.lib()
        k = self.leaf_cells
        while True:
            ti = build_tile_index(m.coordinates(), m.cells_host(), m.cell_edges.cpu().numpy(), m.ne,
                                  np.asarray(U0, dtype=np.float64), np.asarray(P0, dtype=np.float64), k,
                                  bins_per_cell=self.micro_bins_per_cell, bucket_factor=self.bucket_factor)
            # the per-CTA shared memory is sized for the typical leaf (tile_index.pick_smem); an index whose leaves
            # mostly exceed one CTA's 227 KB is rebuilt with smaller leaves (an explicit leaf_cells request raises)
            if ti.smem_bytes > 0 or k <= 32:
                break
            if not self.auto_leaf:
                raise RuntimeError(f"leaf_cells={k}: fewer than 98% of the leaves fit 227 KB of shared memory")
            k //= 2
        if ti.smem_bytes <= 0:
            raise RuntimeError("tile index: leaves do not fit shared memory even at 32 cells per leaf")
        self.leaf_cells = k
        self.tile_host = ti
        dev = {k: torch.from_numpy(np.ascontiguousarray(getattr(ti, k))).to(d)
               for k in ("tree", "leaf_info", "leaf_rect", "coordsL", "UL", "PL", "gidL")}
        dev["leaf_base"] = torch.from_numpy(np.ascontiguousarray(ti.leaf_base, dtype=np.int32)).to(d)
        for k in ("cvL", "binptrL", "binsL"):      # uint16 payloads travel as raw int16 bits
            dev[k] = torch.from_numpy(np.ascontiguousarray(getattr(ti, k)).view(np.int16)).to(d)
        self._tile_dev = dev
        c = _lib.mdq_tile_index_t()
        c.n_leaves, c.depth, c.T = ti.n_leaves, ti.depth, ti.T
        c.max_nv, c.max_np2, c.max_nc, c.max_nbin, c.max_nent = ti.max_nv, ti.max_np2, ti.max_nc, ti.max_nbin, ti.max_nent
        c.u_stride, c.p_stride, c.total_cap = ti.u_stride, ti.p_stride, ti.total_cap
        c.smem_bytes = ti.smem_bytes
        for k, t in dev.items():
            setattr(c, k, t.data_ptr())
        self.tile = c
        self.tile_bytes = ti.nbytes()
        self.tile_smem = int(_lib.lib().mdq_interp_tiled_smem_bytes(ctypes.byref(c)))
        # persistent counters (zero on entry, the kernels leave them zeroed) and per-size scratch
        self._tile_counters = torch.zeros(int(_lib.lib().mdq_interp_tiled_counter_words(ctypes.byref(c))),
                                          dtype=torch.int32, device=d)
        self._tile_scratch = None

    def interpolate(self, target: DeviceMesh, tol=1e-12, probes_of=None):
        """``Function.interpolate`` of every snapshot onto ``target`` (Env2DAirfoil.py:556-568).

        Returns U [T, V+E, 2], P [T, V], cell_of [V+E] (source cell of each target dof point) and the
        device counter of points that fell outside every source cell (closest-cell extrapolation).
        ``probes_of``: the FlowSolver whose mesh ``target`` is -- the drag / lift surface integral over its airfoil facets
        is then evaluated by the last block of the same launch (``mdq_interpolate_drag_lift``) and left where
        ``probes.drag_lift_device`` finds it, so the reward costs no further launch.
        """
        m0, d = self.mesh, self.mesh.device
        npt = target.nv + target.ne
        U = torch.empty((self.T, npt, 2), dtype=torch.float64, device=d)
        P = torch.empty((self.T, target.nv), dtype=torch.float64, device=d)
        cell_of = torch.empty(npt, dtype=torch.int32, device=d)
        miss = torch.empty(1, dtype=torch.int32, device=d)
        miss_list = torch.empty(npt, dtype=torch.int32, device=d)
        L = _lib.lib()
        p = _lib.ptr
        if self.tile is not None:
            words = int(L.mdq_interp_tiled_scratch_words(ctypes.byref(self.tile), npt))
            if self._tile_scratch is None or self._tile_scratch.numel() < words:
                self._tile_scratch = torch.empty(words, dtype=torch.int32, device=d)
            with torch.cuda.device(d):
                rc = L.mdq_interpolate_tiled(p(target.coords), target.nv, p(target.edges), target.ne,
                                             ctypes.byref(self.tile), p(m0.coords), p(m0.cells), p(m0.cell_edges), m0.nv,
                                             m0.ne, m0.nc, p(self.U0), p(self.P0), float(tol), p(U), p(P), p(cell_of),
                                             p(miss), p(miss_list), p(self._tile_counters), p(self._tile_scratch),
                                             _lib.stream_ptr())
            if rc != 0:
                self._tile_counters.zero_()
            _lib.check(rc, "mdq_interpolate_tiled")
            self.last_miss_list = miss_list
            return U, P, cell_of, miss
        fs = probes_of
        if fs is not None and fs.mesh is target and self.T <= 8:
            dl = torch.empty((2, self.T), dtype=torch.float64, device=d)
            if self._ticket is None:
                self._ticket = torch.zeros(1, dtype=torch.int32, device=d)
            with torch.cuda.device(d):
                rc = L.mdq_interpolate_drag_lift(p(target.coords), target.nv, p(target.edges), target.ne, p(m0.coords),
                                                 p(m0.cells), p(m0.cell_edges), m0.nv, m0.ne, m0.nc, self.h_grid,
                                                 p(self.bin_ptr), p(self.bin_cells), float(tol), self.T, p(self.U0),
                                                 p(self.P0), p(U), p(P), p(cell_of), p(miss), p(miss_list),
                                                 p(target.cells), p(target.cell_edges), p(fs.tags), p(target.edge_cell),
                                                 float(fs.viscosity), p(dl), p(self._ticket), _lib.stream_ptr())
            if rc != 0:
                self._ticket.zero_()
            _lib.check(rc, "mdq_interpolate_drag_lift")
            fs._dl_cache = (target, U, P, dl)
            self.last_miss_list = miss_list
            return U, P, cell_of, miss
        with torch.cuda.device(d):
            rc = L.mdq_interpolate(p(target.coords), target.nv, p(target.edges), target.ne, p(m0.coords), p(m0.cells),
                                   p(m0.cell_edges), m0.nv, m0.ne, m0.nc, self.h_grid, p(self.bin_ptr), p(self.bin_cells),
                                   float(tol), self.T, p(self.U0), p(self.P0), p(U), p(P), p(cell_of), p(miss),
                                   p(miss_list), _lib.stream_ptr())
        _lib.check(rc, "mdq_interpolate")
        self.last_miss_list = miss_list
        return U, P, cell_of, miss

    def miss_distance(self, target: DeviceMesh, cell_of, miss):
        """Largest distance between a target dof point that fell outside every source cell and the closest cell
        it was evaluated in (strict mode of SURVEY.md A.7); one device scalar read-back."""
        m0, d = self.mesh, self.mesh.device
        out = torch.zeros(1, dtype=torch.float64, device=d)
        L, p = _lib.lib(), _lib.ptr
        with torch.cuda.device(d):
            rc = L.mdq_interp_miss_distance(p(target.coords), target.nv, p(target.edges), target.ne, p(m0.coords),
                                            p(m0.cells), p(cell_of), p(miss), p(self.last_miss_list), p(out),
                                            _lib.stream_ptr())
        _lib.check(rc, "mdq_interp_miss_distance")
        return math.sqrt(float(out.item()))


class Env2DAirfoil:
    """Environment to optimize the mesh around a 2D airfoil (Env2DAirfoil.py:42-602)."""

    def __init__(self, config, mesh=None, device=None):
        self.device = _dev(device)
        self.flow_solver = FlowSolver(**config["flow_config"], mesh=mesh, device=self.device)  # :51
        fs = self.flow_solver
        ap = config["agent_params"]
        self.coordinate_list = list(range(fs.mesh.nv))
        self.initial_num_node = len(self.coordinate_list)
        self.removable = np.argwhere(fs.removable)[:, 0]
        self.N_CLOSEST = ap["N_closest"]
        self.TIME_REWARD = ap["time_reward"]
        self.action_space = Discrete(self.N_CLOSEST)
        self.solver_steps = ap.get("solver_steps", 5000)
        self.episodes = ap.get("episodes", 0)
        self.timesteps = ap["timesteps"]
        self.threshold = ap["threshold"]
        self.NEGATIVE_REWARD = -1.0
        self.removed_coordinates = []
        self.do_nothing_offset = 0
        self.save_steps = ap.get("save_steps", 1000)
        self.goal_vertices = ap["goal_vertices"]
        self.plot_dir = ap.get("plot_dir", "")
        self.gt_time = np.atleast_1d(np.array(ap.get("gt_time", -1)))
        # strict interpolation (optional): a target dof point farther than this from every source cell
        # (a new edge cutting through the airfoil hole) makes the removal fail with code 2 like the reference's
        # "INTERPOLATION BROKE" (Env2DAirfoil.py:569-573).  None (default): closest-cell extrapolation always.
        st = ap.get("interp_strict_tol", None)
        self.interp_strict_tol = None if st is None else float(st)

        u, p = ap.get("u", -1), ap.get("p", -1)
        if isinstance(u, str) and isinstance(p, str):
            # the "load saved values" branch (Env2DAirfoil.py:126-133): snapshots written by set_plot_dir
            # (<plot_dir>/snapshots/save_velocities.npy [T, V0+E0, 2], save_pressures.npy [T, V0]; this package's P2
            # layout -- vertex dofs, then edge-midpoint dofs in lexicographic edge order -- not DOLFIN's dof numbering)
            u, p = np.load(u), np.load(p)
            self.dof_map = None
            if ap.get("dof_map"):
                # files written by the REFERENCE (FEniCS): u.vector().get_local() rows in DOLFIN's dof order
                # (Env2DAirfoil.py:139-150).  agent_params['dof_map'] = npz with DOLFIN's own dof coordinates (see
                # meshdqn_b200/snapshots.py) turns them into this package's layout.  NOTE: those coordinates belong to
                # the mesh the solve ran on, i.e. the smoothed M0 -- the same mesh fs.mesh holds here.
                from .snapshots import DolfinDofMap
                self.dof_map = DolfinDofMap.load(ap["dof_map"], fs.mesh.coordinates(), fs.mesh.edges.cpu().numpy())
                u, p = self.dof_map.velocities(u), self.dof_map.pressures(p)
        if isinstance(u, int) and u == -1:
            if "synthetic_fields" in ap:
                from .synthetic import synthetic_fields
                u, p = synthetic_fields(fs.mesh.coordinates(), fs.mesh.edges.cpu().numpy(),
                                        T=int(math.ceil(self.solver_steps / self.save_steps)), seed=int(ap["synthetic_fields"]))
            else:
                raise NotImplementedError(
                    "agent_params['u'] == -1 asks for the 5000-step FEniCS Navier-Stokes solve (Env2DAirfoil.py:111-125), "
                    "which stays on the host outside this package: pass the snapshots as arrays "
                    "(agent_params['u'] [T,V0+E0,2], ['p'] [T,V0]) or set agent_params['synthetic_fields'] = seed")
        self.source = SourceField(fs.mesh, u, p)                    # original_u / original_p on M0 (never updated: B6)
        self.original_u, self.original_p = self.source.U0, self.source.P0
        self.T = self.source.T
        self.U, self.P = self.source.U0, self.source.P0            # current u / p (device)
        gt = np.array(ap.get("gt_drag", -1), dtype=np.float64)
        if gt.shape == () and gt == -1:
            dl = drag_lift_device(fs, self.U, self.P).cpu().numpy()  # what the probes return during the solve (:116-121)
            self.gt_drag, self.gt_lift = dl[0].copy(), dl[1].copy()
        else:
            self.gt_drag = np.atleast_1d(gt)
            self.gt_lift = np.atleast_1d(np.array(ap.get("gt_lift", 0.0), dtype=np.float64))
        self.POLYGON = False
        self.out_of_vertices = False
        self.last = {}
        self.reset()

    # ------------------------------------------------------------------ Env2DAirfoil.py:102-168
    def reset(self):
        self._velocities = None
        self._pressures = None
        self.steps = 0
        self.num_episodes = 0
        self.terminal = False
        self._get_distance_lookup()

    def return_vals(self):
        return self.gt_drag, self.gt_time

    @property
    def velocities(self):
        """[T, V, 2] host copy of u at the mesh vertices (= the vertex dofs; Env2DAirfoil.py:515-517)."""
        if self._velocities is None:
            self._velocities = self.U[:, : self.flow_solver.mesh.nv, :].cpu().numpy()
        return self._velocities

    @property
    def pressures(self):
        """[T, V, 1] host copy of p at the mesh vertices (Env2DAirfoil.py:520-522)."""
        if self._pressures is None:
            self._pressures = self.P.cpu().numpy()[:, :, None]
        return self._pressures

    def set_plot_dir(self, plot_dir):
        """Env2DAirfoil.py:432-449: snapshot .npy files -- this package's P2 layout, or DOLFIN's dof order when the
        environment was built from DOLFIN-ordered files with a dof map."""
        self.plot_dir = plot_dir
        os.makedirs(plot_dir + "/snapshots", exist_ok=True)
        np.save(plot_dir + "/snapshots/velocities.npy", self.velocities)
        np.save(plot_dir + "/snapshots/pressures.npy", self.pressures)
        U0, P0 = self.original_u.cpu().numpy(), self.original_p.cpu().numpy()
        dm = getattr(self, "dof_map", None)
        if dm is not None:      # the environment was fed DOLFIN-ordered files: write them back in the same order
            U0, P0 = dm.dolfin_velocities(U0), dm.dolfin_pressures(P0)
        np.save(plot_dir + "/snapshots/save_velocities.npy", U0)
        np.save(plot_dir + "/snapshots/save_pressures.npy", P0)

    # ------------------------------------------------------------------ Env2DAirfoil.py:220-241
    def _get_distance_lookup(self):
        fs = self.flow_solver
        m = fs.mesh
        d = self.device
        if not self.POLYGON:
            coords = m.coordinates()
            not_removable = np.argwhere(~fs.removable)[:, 0]
            bc = coords[not_removable]
            sel = (bc[:, 0] > -0.5) & (bc[:, 0] < 3) & (bc[:, 1] > -0.5) & (bc[:, 1] < 0.5)
            self.polygon = np.ascontiguousarray(bc[sel])
            self._ring = torch.from_numpy(self.polygon).to(d)
            self.POLYGON = True
        # removable vertex list (ascending ids) stays on device
        self._rem_idx = torch.nonzero(fs.removable_dev, as_tuple=False)[:, 0].to(torch.int32)
        nrem = int(self._rem_idx.shape[0])
        self._dist = torch.empty(max(nrem, 1), dtype=torch.float64, device=d)
        if nrem:
            L = _lib.lib()
            p = _lib.ptr
            with torch.cuda.device(d):
                rc = L.mdq_polygon_distance(p(m.coords), p(self._rem_idx), nrem, p(self._ring), int(self._ring.shape[0]),
                                            p(self._dist), _lib.stream_ptr())
            _lib.check(rc, "mdq_polygon_distance")
        self._nrem = nrem

    @property
    def distance_lookup(self):
        return self._dist[: self._nrem].cpu().numpy()

    # ------------------------------------------------------------------ Env2DAirfoil.py:244-315
    def _n_closest(self):
        """Kept for API parity; the selection itself happens inside ``get_state`` on device."""
        self.get_state()

    def get_state(self):
        fs = self.flow_solver
        m = fs.mesh
        d = self.device
        N, T = self.N_CLOSEST, self.T
        self.coordinate_list = list(range(m.nv))
        self._get_distance_lookup()
        i32 = dict(dtype=torch.int32, device=d)
        n_closest = torch.empty(N, **i32)
        coord_map = torch.empty(N, **i32)
        inv_map = torch.empty(m.nv, **i32)
        x = torch.empty((N, 3 * T + 2), dtype=torch.float32, device=d)
        ecap = 3 * m.nc
        edge_buf = torch.empty((2, ecap), dtype=torch.int64, device=d)
        n_edges = torch.empty(1, **i32)
        L = _lib.lib()
        p = _lib.ptr
        with torch.cuda.device(d):
            rc = L.mdq_build_state(p(self._dist), p(self._rem_idx), self._nrem, int(self.do_nothing_offset), N,
                                   p(m.coords), m.nv, p(m.cells), m.nc, T, p(self.U), m.nv + m.ne, p(self.P),
                                   p(n_closest), p(coord_map), p(inv_map), p(x), p(edge_buf), ecap, p(n_edges),
                                   _lib.stream_ptr())
        _lib.check(rc, "mdq_build_state")
        # one small D2H: the action -> vertex map the host needs for the next _remove_vertex, and the edge count
        packed = torch.cat([n_edges, coord_map, n_closest]).cpu().numpy()
        E = int(packed[0])
        cm = packed[1:1 + N]
        self.n_closest = packed[1 + N:1 + 2 * N]
        valid = cm >= 0
        if not valid.all():
            print("OUT OF VERTICES")
            self.out_of_vertices = True
            # the reference's feature assignment raises on a short window (Env2DAirfoil.py:284-286 assigns [k, .] into
            # [N, .]); the oracle defines the state as all-zero features there and so does the device path
            x.zero_()
        self.removable = np.argwhere(fs.removable)[:, 0]
        self.coord_map = {int(k): int(v) for k, v in enumerate(cm) if v >= 0}
        self.inv_coord_map = {v: k for k, v in self.coord_map.items()}
        edge_index = edge_buf[:, :E]
        if E != ecap:
            edge_index = edge_index.contiguous()
        return Data(x=x, edge_index=edge_index, edge_attr=None)

    # ------------------------------------------------------------------ Env2DAirfoil.py:318-377
    def step(self, action):
        broken = False
        rew = None
        action = int(action)
        if action == self.action_space.n:          # no removal: shift the N-closest window (quirk B5)
            self.do_nothing_offset += 1
            removed = 0
        else:
            removed = self._remove_vertex(action)
        state = self.get_state()
        if self.out_of_vertices:
            print("OUT OF VERTICES")
            removed = 2
        if removed == 0:
            rew, broken, self.terminal = self.calculate_reward()
            if self.terminal:
                self.rew = 0.5 * self.NEGATIVE_REWARD  # quirk B7: assigns the wrong name, reward unchanged
                print("ACCURACY THRESHOLD REACHED")
            if broken:
                rew = self.NEGATIVE_REWARD
                self.terminal = True
        elif removed == 1:
            rew = self.NEGATIVE_REWARD
        elif removed == 2:
            rew = self.NEGATIVE_REWARD
            self.terminal = True
            broken = True
        self.steps += 1
        if self.steps >= self.timesteps:
            self.terminal = True
            self.episodes += 1
        return state, rew, self.terminal, {}

    # ------------------------------------------------------------------ Env2DAirfoil.py:380-428
    def calculate_reward(self):
        try:
            dl = drag_lift_device(self.flow_solver, self.U, self.P).cpu().numpy()
        except RuntimeError:
            print("\n\nSAMPLING BROKE\n\n")
            return self.NEGATIVE_REWARD, True, True
        self.new_drags, self.new_lifts = dl[0].copy(), dl[1].copy()
        drag_factor = -2 * np.log(0.5) / self.threshold
        error_val = np.linalg.norm(np.abs(self.gt_drag - self.new_drags) / np.abs(self.gt_drag))
        drag_reward = 2 * np.exp(-drag_factor * error_val) - 1
        time_reward = (self.initial_num_node - len(self.coordinate_list)) * self.TIME_REWARD      # quirk B8
        acc_thresh = any(np.abs(np.abs(self.gt_drag - self.new_drags) / self.gt_drag) > self.threshold)
        vert_thresh = self.flow_solver.mesh.nv < self.goal_vertices * self.initial_num_node
        if vert_thresh:
            print("\nMAXIMUM REMOVALS REACHED\n")
        return drag_reward + time_reward, False, bool(acc_thresh or vert_thresh)

    # ------------------------------------------------------------------ Env2DAirfoil.py:452-512
    def _remove_vertex(self, selected_coord=None):
        try:
            selected_coord = self.coord_map[selected_coord]
        except KeyError:
            print("RAN OUT OF VERTICES")
            return 2
        m = self.flow_solver.mesh
        boundary_vertices = m.boundary_vertices().copy()            # BoundaryMesh(...).entity_map(0) (:464-465)
        coords = m.coordinates()
        self.removed_coordinates.append(coords[selected_coord].copy())
        boundary_vertices[boundary_vertices > selected_coord] -= 1
        keep = np.ones(len(coords), dtype=bool)
        keep[selected_coord] = False
        del self.coordinate_list[selected_coord]
        coords = coords[keep]
        try:
            tri = Delaunay(coords)                                   # same Qhull call as the reference (:487)
        except ValueError:
            self.coordinate_list.insert(selected_coord, selected_coord)
            print("\nMESH BROKE, COULDN'T TRIANGULATE")
            return 2
        cells = tri.simplices
        is_b = np.zeros(len(coords), dtype=bool)
        is_b[boundary_vertices] = True
        cells = cells[is_b[cells].sum(axis=1) != 3]                  # drop all-boundary cells (:496)
        self.last["delaunay_cells"] = cells
        return self._check_mesh((coords, cells), selected_coord)

    # ------------------------------------------------------------------ Env2DAirfoil.py:547-602
    def _check_mesh(self, mesh, selected_coord):
        if selected_coord in self.removable:
            fs = self.flow_solver
            old = (fs.mesh, fs.tags, fs.removable_dev, fs.num_vertices, fs._removable_host)
            try:
                fs.remesh(mesh)                                          # smooth(50), tags, removable on device
                U, P, cell_of, miss = self.source.interpolate(fs.mesh, probes_of=fs)   # all T snapshots from the ORIGINAL mesh (B6); drag / lift in the same launch
                if self.interp_strict_tol is not None and \
                        self.source.miss_distance(fs.mesh, cell_of, miss) > self.interp_strict_tol:
                    raise RuntimeError("target dof point outside the source mesh by more than interp_strict_tol")
            except RuntimeError as err:
                # Env2DAirfoil.py:569-573: restore the old mesh, put the vertex back, report a broken removal.  (The
                # reference restores only flow_solver.mesh; here the tags / removable mask go back with it so the
                # environment stays consistent.)
                print("INTERPOLATION BROKE", err)
                fs.mesh, fs.tags, fs.removable_dev, fs.num_vertices, fs._removable_host = old
                self.coordinate_list.insert(selected_coord, selected_coord)
                return 2
            self.U, self.P = U, P
            self._velocities = None
            self._pressures = None
            self.last.update(cell_of=cell_of, miss=miss)
            self.removable = np.argwhere(fs.removable)[:, 0]
            return 0
        self.coordinate_list.insert(selected_coord, selected_coord)
        print("\nMESH BROKE. SKIPPING VERTEX REMOVAL\n")
        return 2
