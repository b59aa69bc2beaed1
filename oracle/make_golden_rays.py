"""Golden vectors of the reference's ``get_rays`` (nerf/utils.py:145-279).  TEST INFRASTRUCTURE ONLY.

``nerf/utils.py`` cannot be imported here (tensorboardX, lpips, torch_ema, ... are not installed), so the two
functions needed - ``custom_meshgrid`` and ``get_rays`` - are taken out of the reference file with ``ast`` and executed
unmodified in a namespace that holds torch / numpy / packaging.version.   python -m oracle.make_golden_rays
-> tests/golden/ref_rays.npz
"""
import ast
import os

import numpy as np
import torch
from packaging import version as pver

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE = os.environ.get("SANERF_REFERENCE", "/root/reference")


def load_reference_get_rays():
    src = open(os.path.join(REFERENCE, "nerf", "utils.py")).read()
    tree = ast.parse(src)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("custom_meshgrid", "get_rays")]
    assert len(wanted) == 2
    ns = {"torch": torch, "np": np, "pver": pver}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), "reference:nerf/utils.py", "exec"), ns)
    return ns["get_rays"]


def random_pose(g):
    q, _ = np.linalg.qr(g.standard_normal((3, 3)))
    pose = np.eye(4, dtype=np.float32)
    pose[:3, :3] = q
    pose[:3, 3] = g.standard_normal(3) * 0.5
    return pose


def main():
    get_rays = load_reference_get_rays()
    g = np.random.default_rng(0)
    out = {}
    # 1. full image, one pose, ndarray intrinsics (the GUI / test path, utils.py:1647-1712)
    H, W = 12, 20
    pose = random_pose(g)
    intr = np.array([23.5, 21.0, 9.7, 6.2], dtype=np.float32)
    r = get_rays(torch.from_numpy(pose)[None], intr, H, W, -1)
    out.update(full_pose=pose, full_intr=intr, full_hw=np.array([H, W]), full_o=r["rays_o"].numpy(), full_d=r["rays_d"].numpy(),
               full_inds_coarse=r["inds_coarse"].numpy())
    # 2. given pixel coordinates, one pose per ray, tensor intrinsics (random-image-batch training, colmap_provider.py)
    N = 37
    poses = np.stack([random_pose(g) for _ in range(N)])
    intrs = np.stack([intr + g.standard_normal(4).astype(np.float32) for _ in range(N)]).astype(np.float32)
    coords = np.stack([g.integers(0, H, N), g.integers(0, W, N)], -1)
    r = get_rays(torch.from_numpy(poses), torch.from_numpy(intrs), H, W, N, coords=torch.from_numpy(coords))
    out.update(co_poses=poses, co_intr=intrs, co_coords=coords, co_o=r["rays_o"].numpy(), co_d=r["rays_d"].numpy(),
               co_i=r["i"].numpy(), co_j=r["j"].numpy(), co_inds_coarse=r["inds_coarse"].numpy())
    path = os.path.join(ROOT, "tests", "golden", "ref_rays.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
