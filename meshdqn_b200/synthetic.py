"""Seeded synthetic inputs: velocity/pressure snapshots and refined airfoil-channel meshes.

The reference's fields come from a 5000-step FEniCS Navier-Stokes solve
(/root/reference/Env2DAirfoil.py:111-125, flow_solver.py:362-396) that stays
outside this repo; no solution ships with the reference.  Tests and bench
therefore use analytic, smooth, phase-shifted P2/P1 nodal fields with the
channel parabola of flow_solver.py:33-44 as the base flow, and
`synthetic_airfoil_mesh` builds the ~250k / ~1M-triangle meshes BASELINE.json
names (vertex order as in the shipped fixtures: corners, airfoil ring in curve
order, outer wall points, interior).
Input generation only -- not part of the timed hot path.
"""
from __future__ import annotations

import numpy as np


def field_values(pts, T=5, seed=0):
    """Analytic snapshots at points [n,2] -> (u [T,n,2], p [T,n]) float64."""
    rng = np.random.RandomState(seed)
    ph0 = rng.uniform(0, 2 * np.pi)
    amp = rng.uniform(0.05, 0.15, size=3)
    x, y = pts[:, 0], pts[:, 1]
    bot, top = -0.5, 0.5
    H = top - bot
    par = -4 * 1.5 * (y - bot) * (y - top) / H / H
    cx, cy = 0.5, 0.0
    g = 1.0 - np.exp(-(((x - cx) / 0.45) ** 2 + ((y - cy) / 0.12) ** 2))
    u = np.empty((T, len(pts), 2))
    p = np.empty((T, len(pts)))
    for t in range(T):
        ph = ph0 + 0.4 * t
        u[t, :, 0] = par * g * (1.0 + amp[0] * np.sin(2 * np.pi * (x - 0.2 * t) + ph))
        u[t, :, 1] = amp[1] * np.sin(np.pi * x + ph) * np.cos(np.pi * y) * g * 2.0
        p[t] = 0.3 * (3.0 - x) + 0.5 * np.exp(-((x - cx) ** 2 + (y - cy) ** 2) / 0.1) * np.cos(ph) \
            + amp[2] * np.sin(2 * np.pi * y + ph)
    return u, p


def synthetic_fields(coords, edges, T=5, seed=0):
    """Nodal P2 vector / P1 scalar coefficients on a mesh.

    Returns ``U [T, V+E, 2]`` (vertex dofs then edge-midpoint dofs) and ``P [T, V]``.
    """
    mid = 0.5 * coords[edges[:, 0]] + 0.5 * coords[edges[:, 1]]
    pts = np.concatenate([coords, mid], axis=0)
    u, p = field_values(pts, T, seed)
    return np.ascontiguousarray(u), np.ascontiguousarray(p[:, : len(coords)])


def _naca_ring(n, chord=1.0, t=0.12, m=0.0, pc=0.4, x0=0.0, y0=0.0):
    """Closed NACA 4-digit contour, n points in curve order (upper TE->LE, lower LE->TE).

    Symmetric (convex) by default: every Delaunay cell with three ring vertices then lies inside the
    airfoil, so the reference's all-boundary-cell rule removes exactly the hole."""
    nu = n // 2 + 1
    beta = np.linspace(0.0, np.pi, nu)
    xc = 0.5 * (1 + np.cos(beta))  # 1 -> 0
    def thick(xx):
        return 5 * t * (0.2969 * np.sqrt(xx) - 0.1260 * xx - 0.3516 * xx ** 2 + 0.2843 * xx ** 3 - 0.1036 * xx ** 4)
    def camber(xx):
        return np.where(xx < pc, m / pc ** 2 * (2 * pc * xx - xx ** 2), m / (1 - pc) ** 2 * ((1 - 2 * pc) + 2 * pc * xx - xx ** 2))
    up = np.stack([xc, camber(xc) + thick(xc)], 1)
    xl = xc[::-1][1:-1] if n % 2 == 0 else xc[::-1][1:]
    xl = xl[: n - nu]
    lo = np.stack([xl, camber(xl) - thick(xl)], 1)
    ring = np.concatenate([up, lo], 0)[:n]
    # re-sample uniformly in arc length (cosine clustering would create micron-sized edges at 1M cells)
    closed = np.concatenate([ring, ring[:1]], 0)
    seg = np.sqrt((np.diff(closed, axis=0) ** 2).sum(1))
    s_acc = np.concatenate([[0.0], np.cumsum(seg)])
    s_new = np.linspace(0.0, s_acc[-1], n + 1)[:-1]
    ring = np.stack([np.interp(s_new, s_acc, closed[:, 0]), np.interp(s_new, s_acc, closed[:, 1])], 1)
    ring[:, 0] = ring[:, 0] * chord + x0
    ring[:, 1] = ring[:, 1] * chord + y0
    return ring


def _morton_order(q, lo, hi, bits=16):
    """Z-curve rank of points [n,2] (locality-preserving numbering, as mesh generators produce)."""
    s = ((q - lo) / (hi - lo) * ((1 << bits) - 1)).astype(np.uint64)
    def spread(v):
        v = (v | (v << np.uint64(16))) & np.uint64(0x0000FFFF0000FFFF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF00FF00FF)
        v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
        v = (v | (v << np.uint64(2))) & np.uint64(0x3333333333333333)
        v = (v | (v << np.uint64(1))) & np.uint64(0x5555555555555555)
        return v
    return np.argsort(spread(s[:, 0]) | (spread(s[:, 1]) << np.uint64(1)), kind="stable")


def synthetic_airfoil_mesh(n_triangles=250_000, seed=0, n_airfoil=None, order="random"):
    """Graded Delaunay mesh of the channel [-0.5,3]x[-0.5,0.5] minus a NACA-style hole.

    Cells whose three vertices are all boundary points are dropped, the same rule the
    reference applies after re-triangulating (Env2DAirfoil.py:496).  Returns
    ``(coords f64 [V,2], cells i32 [C,3], n_ring)``.  ``order="morton"`` numbers the interior vertices along a
    Z-curve (the locality a front/quadtree mesh generator gives); ``"random"`` keeps the sampling order.
    """
    from scipy.spatial import Delaunay

    rng = np.random.RandomState(seed)
    nv_target = max(64, n_triangles // 2)
    n_af = n_airfoil or int(max(40, 6.0 * np.sqrt(nv_target)))
    ring = _naca_ring(n_af, chord=1.0, x0=0.0, y0=0.0)
    # outer boundary points
    h_wall = max(3.5 / (2.0 * np.sqrt(nv_target)), 1e-4)
    nx = int(np.ceil(3.5 / h_wall))
    ny = int(np.ceil(1.0 / h_wall))
    xs = np.linspace(-0.5, 3.0, nx + 1)[1:-1]
    ys = np.linspace(-0.5, 0.5, ny + 1)[1:-1]
    corners = np.array([[-0.5, -0.5], [3.0, -0.5], [3.0, 0.5], [-0.5, 0.5]])
    walls = np.concatenate([
        np.stack([xs, np.full_like(xs, -0.5)], 1), np.stack([xs, np.full_like(xs, 0.5)], 1),
        np.stack([np.full_like(ys, -0.5), ys], 1), np.stack([np.full_like(ys, 3.0), ys], 1)])
    nb = 4 + len(ring) + len(walls)
    # one buffer layer of interior points facing every wall edge, so that no Delaunay cell has three
    # outer-boundary vertices (the reference's drop rule would otherwise notch the corners)
    xm = 0.5 * (np.linspace(-0.5, 3.0, nx + 1)[1:] + np.linspace(-0.5, 3.0, nx + 1)[:-1])
    ym = 0.5 * (np.linspace(-0.5, 0.5, ny + 1)[1:] + np.linspace(-0.5, 0.5, ny + 1)[:-1])
    off = 0.75 * h_wall
    ym_in = ym[(ym > -0.5 + 1.2 * off) & (ym < 0.5 - 1.2 * off)]
    buffer_pts = np.concatenate([
        np.stack([xm, np.full_like(xm, -0.5 + off)], 1), np.stack([xm, np.full_like(xm, 0.5 - off)], 1),
        np.stack([np.full_like(ym_in, -0.5 + off), ym_in], 1), np.stack([np.full_like(ym_in, 3.0 - off), ym_in], 1)])
    n_int = max(16, nv_target - nb - len(buffer_pts))
    # graded interior cloud: rejection-sample density ~ 1/h^2, h = h0 + a*dist(chord)
    pts = []
    got = 0
    h0, a = 0.05, 0.6
    acc_rate = 0.05
    # polygon test helpers (ring is a simple polygon)
    def inside_ring(q):
        xq, yq = q[:, 0], q[:, 1]
        ins = np.zeros(len(q), dtype=bool)
        ax, ay = ring[:, 0], ring[:, 1]
        bx, by = np.roll(ax, -1), np.roll(ay, -1)
        for k in range(len(ring)):
            cond = (ay[k] > yq) != (by[k] > yq)
            with np.errstate(divide="ignore", invalid="ignore"):
                xi = ax[k] + (yq - ay[k]) * (bx[k] - ax[k]) / (by[k] - ay[k])
            ins ^= cond & (xq < xi)
        return ins
    while got < n_int:
        m = min(int((n_int - got) / acc_rate * 1.3) + 1024, 8_000_000)
        q = np.stack([rng.uniform(-0.5, 3.0, m), rng.uniform(-0.5, 0.5, m)], 1)
        dx = np.clip(q[:, 0], 0.0, 1.0) - q[:, 0]
        d = np.sqrt(dx ** 2 + q[:, 1] ** 2)
        h = h0 + a * d
        acc = rng.uniform(0, 1, m) < (h0 / h) ** 2
        q = q[acc]
        # keep points clear of the hole and of the boundaries
        xq = np.clip(q[:, 0], 0.0, 1.0)
        yt = 5 * 0.12 * (0.2969 * np.sqrt(xq) - 0.1260 * xq - 0.3516 * xq ** 2 + 0.2843 * xq ** 3 - 0.1036 * xq ** 4)
        band = 4.0 * (2.05 / n_af)
        near = (q[:, 0] > -band) & (q[:, 0] < 1.0 + band) & (np.abs(q[:, 1]) < yt + band)
        bad = np.zeros(len(q), dtype=bool)
        if near.any():
            qn = q[near]
            insn = inside_ring(qn)
            # clearance: distance to the ring segments must exceed 0.7 ring spacings, so no point sits in
            # the diametral circle of a ring edge and the contour stays in the Delaunay triangulation
            ax, ay = ring[:, 0], ring[:, 1]
            bx, by = np.roll(ax, -1), np.roll(ay, -1)
            ex, ey = bx - ax, by - ay
            l2 = ex * ex + ey * ey
            dmin = np.full(len(qn), np.inf)
            for lo in range(0, len(qn), 4096):
                qq = qn[lo:lo + 4096]
                tt = np.clip(((qq[:, None, 0] - ax) * ex + (qq[:, None, 1] - ay) * ey) / l2, 0.0, 1.0)
                dd = (ax + tt * ex - qq[:, None, 0]) ** 2 + (ay + tt * ey - qq[:, None, 1]) ** 2
                dmin[lo:lo + 4096] = np.sqrt(dd.min(1))
            bad[np.nonzero(near)[0]] = insn | (dmin < 0.7 * (2.05 / n_af))
        bad |= (q[:, 0] < -0.5 + 1.5 * h_wall) | (q[:, 0] > 3.0 - 1.5 * h_wall) | (np.abs(q[:, 1]) > 0.5 - 1.5 * h_wall)
        q = q[~bad]
        acc_rate = max(1e-3, len(q) / m)
        pts.append(q)
        got += len(q)
    interior = np.concatenate([buffer_pts] + pts, 0)[:n_int + len(buffer_pts)]
    if order == "morton":
        interior = interior[_morton_order(interior, np.array([-0.5, -0.5]), np.array([3.0, 0.5]))]
    elif order != "random":
        raise ValueError("order must be 'random' or 'morton'")
    coords = np.concatenate([corners, ring, walls, interior], 0)
    tri = Delaunay(coords)
    cells = tri.simplices
    is_b = np.zeros(len(coords), dtype=bool)
    is_b[:nb] = True
    cells = cells[is_b[cells].sum(1) != 3]
    cells = np.sort(cells, axis=1).astype(np.int32)
    return np.ascontiguousarray(coords), np.ascontiguousarray(cells), len(ring)


def reference_config(N_closest=180, timesteps=10000, threshold=0.001, smooth=True, **extra):
    """The reference's YAML (/root/reference/configs/ray_ys930.yaml:1-39) as a dict; mesh and fields are injected."""
    cfg = {
        "flow_config": {
            "flow_params": {"mu": 1e-3, "rho": 1.0, "inflow": "constant"},
            "geometry_params": {"mesh": None},
            "solver_params": {"dt": 0.001, "solver_type": "lu", "smooth": smooth},
        },
        "agent_params": dict(solver_steps=5000, episodes=1000000, timesteps=timesteps, threshold=threshold,
                             N_closest=N_closest, gt_drag=-1, gt_time=-1, u=-1, p=-1, do_nothing=True, time_reward=0.005,
                             smoothing=True, save_steps=1000, goal_vertices=0.95, plot_dir=""),
        "optimizer": {"lr": 1e-5, "weight_decay": 1e-6, "batch_size": 32},
        "epsilon": {"decay": 10000, "start": 1.0, "end": 0.01, "gamma": 1.0},
    }
    cfg["agent_params"].update(extra)
    return cfg


def fixture_environment_inputs(mesh_npz, device, T=5, seed=0):
    """(coords, cells, U, P) for an environment on a shipped fixture mesh: the mesh is smoothed on the DEVICE
    (``DeviceMesh.smooth(50)``, the kernel the environment itself uses) and the analytic snapshots are sampled on the
    smoothed dof points -- the product-side twin of the oracle-smoothed fields the tests build (same bits: the device
    smoother is bit-identical to the oracle's)."""
    from .flow_solver import DeviceMesh
    z = np.load(mesh_npz)
    coords, cells = z["coords"].astype(np.float64), z["cells"].astype(np.int32)
    m = DeviceMesh(coords, cells, device)
    m.smooth(50)
    U, P = synthetic_fields(m.coordinates(), m.edges.cpu().numpy(), T, seed)
    return coords, cells, U, P
