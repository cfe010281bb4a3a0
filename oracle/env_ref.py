"""oracle/env_ref.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py; parity unpinned).

CPU restatement of the reference's per-action environment step:

* ``FlowSolverRef``  -- mesh services of /root/reference/flow_solver.py:47-78,194-266,341-359
                        (load, smooth(50), mark_boundaries, removable, remesh; no NS solve)
* ``Env2DAirfoilRef``-- /root/reference/Env2DAirfoil.py:47-164 (init/reset), :220-315 (state),
                        :318-428 (step / reward), :452-512 (_remove_vertex), :547-602 (_check_mesh)

Result-affecting quirks B1-B8 of SURVEY.md Appendix B are replicated on purpose.
scipy.spatial.Delaunay is the same third-party Qhull the reference calls
(Env2DAirfoil.py:487).  Fields are plain arrays (U [T,V0+E0,2] P2 nodal values,
P [T,V0]) instead of DOLFIN Functions.
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.spatial import Delaunay

from . import geom


class FlowSolverRef:
    def __init__(self, flow_params, geometry_params, solver_params, mesh=None):
        self.mu = float(flow_params["mu"])
        self.smooth = bool(solver_params.get("smooth", False))
        if mesh is None:
            from meshdqn_b200.xdmf import read_xdmf_mesh  # host input plumbing only
            mesh = read_xdmf_mesh(geometry_params["mesh"])
        coords, cells = mesh
        self.remesh(np.array(coords, dtype=np.float64), np.array(cells, dtype=np.int32))

    def remesh(self, coords, cells):
        self.topo = geom.Topology(cells, len(coords))
        self.coords = geom.smooth(coords, self.topo, 50) if self.smooth else np.array(coords, dtype=np.float64)
        self.cells = self.topo.cells
        self.tags = geom.facet_tags(self.coords, self.topo)
        self.removable = geom.removable_mask(self.coords, self.topo)
        self.num_vertices = len(self.coords)

    def drag_lift(self, U, P):
        return geom.drag_lift(self.coords, self.topo, self.tags, U, P, self.mu)


class Env2DAirfoilRef:
    def __init__(self, config, mesh=None):
        self.flow_solver = FlowSolverRef(**config["flow_config"], mesh=mesh)
        fs = self.flow_solver
        ap = config["agent_params"]
        self.initial_num_node = fs.num_vertices
        self.coordinate_list = list(range(self.initial_num_node))
        self.removable = np.argwhere(fs.removable)[:, 0]
        self.N_CLOSEST = ap["N_closest"]
        self.TIME_REWARD = ap["time_reward"]
        self.n_actions = self.N_CLOSEST
        self.timesteps = ap["timesteps"]
        self.threshold = ap["threshold"]
        self.goal_vertices = ap["goal_vertices"]
        self.NEGATIVE_REWARD = -1.0
        self.do_nothing_offset = 0
        self.removed_coordinates = []
        st = ap.get("interp_strict_tol", None)          # optional strict mode of SURVEY.md A.7 (Env2DAirfoil.py:569-573)
        self.interp_strict_tol = None if st is None else float(st)
        # source mesh M0 and original fields (never updated: quirk B6)
        self.coords0 = fs.coords.copy()
        self.topo0 = fs.topo
        self.U0 = np.ascontiguousarray(ap["u"], dtype=np.float64)
        self.P0 = np.ascontiguousarray(ap["p"], dtype=np.float64)
        self.T = self.U0.shape[0]
        self.U = self.U0.copy()
        self.P = self.P0.copy()
        gt = np.array(ap.get("gt_drag", -1), dtype=np.float64)
        if gt.shape == () and gt == -1:
            self.gt_drag, self.gt_lift = fs.drag_lift(self.U0, self.P0)
        else:
            self.gt_drag = np.atleast_1d(gt)
            self.gt_lift = np.atleast_1d(np.array(ap.get("gt_lift", 0.0), dtype=np.float64))
        self.polygon = None
        self.out_of_vertices = False
        self.last = {}
        self.reset()

    def reset(self):
        nv = self.flow_solver.num_vertices
        self.velocities = self.U[:, :nv, :].copy()
        self.pressures = self.P[:, :nv, None].copy()
        self.steps = 0
        self.terminal = False
        self._get_distance_lookup()

    # ---- Env2DAirfoil.py:220-241 ----
    def _get_distance_lookup(self):
        fs = self.flow_solver
        coords = fs.coords
        if self.polygon is None:
            nr = np.argwhere(~fs.removable)[:, 0]
            bc = coords[nr]
            m = (bc[:, 0] > -0.5) & (bc[:, 0] < 3) & (bc[:, 1] > -0.5) & (bc[:, 1] < 0.5)
            self.polygon = bc[m].copy()
        self.distance_lookup = geom.polygon_distance(coords[self.removable], self.polygon)

    # ---- Env2DAirfoil.py:293-315 ----
    def _n_closest(self):
        fs = self.flow_solver
        self.coordinate_list = list(range(fs.num_vertices))
        self.removable = np.argwhere(fs.removable)[:, 0]
        self._get_distance_lookup()
        order = np.argsort(self.distance_lookup, kind="stable")
        self.n_closest = order[self.do_nothing_offset:self.N_CLOSEST + self.do_nothing_offset]
        if len(self.n_closest) < self.N_CLOSEST:
            self.out_of_vertices = True
        mapping = self.removable[self.n_closest]
        self.coord_map = dict(zip(range(len(self.n_closest)), mapping.tolist()))
        self.inv_coord_map = dict(zip(mapping.tolist(), range(len(self.n_closest))))

    # ---- Env2DAirfoil.py:244-290 ----
    def get_state(self):
        from meshdqn_b200.data import Data  # plain container, no compute

        fs = self.flow_solver
        self._n_closest()
        vals = np.array(list(self.coord_map.values())).astype(int)
        cells = fs.cells
        good = np.argwhere(np.all(np.isin(cells, vals), axis=1))[:, 0]
        ei = []
        for c in good:
            i1, i2, i3 = (self.inv_coord_map[int(v)] for v in cells[c])
            ei += [[i1, i2], [i1, i3], [i2, i3]]
        edge_index = torch.tensor(ei, dtype=torch.long).reshape(-1, 2).T.contiguous()
        T = self.velocities.shape[0]
        N = self.N_CLOSEST
        x = torch.zeros((N, 3 * T + 2), dtype=torch.float)
        nc = self.n_closest
        if len(nc) == N:
            x[:, :2] = torch.from_numpy(fs.coords[nc])                                  # quirk B1
            x[:, 2:2 * T + 2] = torch.from_numpy(self.velocities[:, nc, :].reshape(N, -1))  # quirk B2
            x[:, 2 * T + 2:] = torch.from_numpy(self.pressures[:, nc][:, :, 0].T.copy())
        return Data(x=x, edge_index=edge_index)

    # ---- Env2DAirfoil.py:318-377 ----
    def step(self, action):
        broken = False
        rew = None
        if action == self.n_actions:
            self.do_nothing_offset += 1
            removed = 0
        else:
            removed = self._remove_vertex(action)
        state = self.get_state()
        if self.out_of_vertices:
            removed = 2
        if removed == 0:
            rew, broken, self.terminal = self.calculate_reward()
            if broken:
                rew = self.NEGATIVE_REWARD
                self.terminal = True
        elif removed == 1:
            rew = self.NEGATIVE_REWARD
        elif removed == 2:
            rew = self.NEGATIVE_REWARD
            self.terminal = True
        self.steps += 1
        if self.steps >= self.timesteps:
            self.terminal = True
        return state, rew, self.terminal, {}

    # ---- Env2DAirfoil.py:380-428 ----
    def calculate_reward(self):
        fs = self.flow_solver
        self.new_drags, self.new_lifts = fs.drag_lift(self.U, self.P)
        drag_factor = -2 * np.log(0.5) / self.threshold
        error_val = np.linalg.norm(np.abs(self.gt_drag - self.new_drags) / np.abs(self.gt_drag))
        drag_reward = 2 * np.exp(-drag_factor * error_val) - 1
        time_reward = (self.initial_num_node - len(self.coordinate_list)) * self.TIME_REWARD
        acc_thresh = any(np.abs(np.abs(self.gt_drag - self.new_drags) / self.gt_drag) > self.threshold)
        vert_thresh = fs.num_vertices < self.goal_vertices * self.initial_num_node
        return drag_reward + time_reward, False, bool(acc_thresh or vert_thresh)

    # ---- Env2DAirfoil.py:452-512 ----
    def _remove_vertex(self, action):
        try:
            selected = self.coord_map[action]
        except KeyError:
            return 2
        fs = self.flow_solver
        bverts = fs.topo.boundary_vertices.copy()
        coords = fs.coords
        self.removed_coordinates.append(coords[selected].copy())
        bverts[bverts > selected] -= 1
        keep = np.ones(len(coords), dtype=bool)
        keep[selected] = False
        del self.coordinate_list[selected]
        coords = coords[keep]
        try:
            tri = Delaunay(coords)
        except ValueError:
            return 2
        cells = tri.simplices
        cells = cells[np.sum(np.isin(cells, bverts), axis=1) != 3]
        self.last["delaunay_cells"] = cells.copy()
        return self._check_mesh(coords, cells)

    # ---- Env2DAirfoil.py:547-602 ----
    def _check_mesh(self, coords, cells):
        fs = self.flow_solver
        old = dict(fs.__dict__)
        fs.remesh(coords, cells)
        pts = fs.topo.p2_points(fs.coords)
        cell_of, nmiss, d2 = geom.locate(pts, self.coords0, self.topo0.cells)
        # farthest target dof point from the source mesh (0 when every point was located); a negative tolerance breaks always
        if self.interp_strict_tol is not None and (float(np.sqrt(d2.max())) if nmiss else 0.0) > self.interp_strict_tol:
            fs.__dict__.update(old)                     # "INTERPOLATION BROKE": old mesh back, vertex back, code 2
            return 2
        u, p = geom.eval_fields(pts, fs.num_vertices, cell_of, self.coords0, self.topo0, self.U0, self.P0)
        self.U, self.P = u, p
        self.last.update(cell_of=cell_of, nmiss=nmiss, miss_d2=d2, pts=pts)
        self.velocities = u[:, :fs.num_vertices, :].copy()
        self.pressures = p[:, :, None].copy()
        self.removable = np.argwhere(fs.removable)[:, 0]
        return 0
