"""Convert the reference's shipped XDMF/HDF5 meshes into small .npz fixtures.

Run in the build container (where /root/reference exists):
    python tools/make_mesh_fixtures.py
The GPU box has no /root/reference, so tests and bench read tests/golden/*.npz.
Source: /root/reference/xdmf_files/{ys930_0.15000,ah93w145_0.14000}_triangle.{xdmf,h5}
(the meshes named by configs/ray_ys930.yaml:7 and configs/ray_ah93w145.yaml:7).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshdqn_b200.xdmf import read_xdmf_mesh  # noqa: E402

REF = "/root/reference/xdmf_files"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

for name, short in (("ys930_0.15000_triangle", "ys930"), ("ah93w145_0.14000_triangle", "ah93w145")):
    coords, cells = read_xdmf_mesh(os.path.join(REF, name + ".xdmf"))
    np.savez_compressed(os.path.join(OUT, f"mesh_{short}.npz"), coords=coords, cells=cells)
    print(short, coords.shape, cells.shape)
