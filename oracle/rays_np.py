"""numpy restatement of the reference's ray generation (``get_rays``, nerf/utils.py:145-279) for given pixel indices.
TEST INFRASTRUCTURE ONLY - never imported by the product path.  Pinned by tests/golden/ref_rays.npz
(oracle/make_golden_rays.py runs the reference's own function)."""
import numpy as np


def get_rays(poses, intrinsics, H, W, inds=None, incoherent_mask_size=128):
    """poses [1 or N,4,4] cam2world, intrinsics [4] or [N,4] = (fx, fy, cx, cy); inds [N] flat pixel indices (row * W + col)
    or None for the whole image.  Returns rays_o, rays_d [N,3] (directions NOT normalised, utils.py:246-248), i, j, inds_coarse."""
    poses = np.asarray(poses, dtype=np.float32)
    intr = np.asarray(intrinsics, dtype=np.float32)
    if inds is None:
        inds = np.arange(H * W)
    inds = np.asarray(inds, dtype=np.int64)
    # utils.py:166-171: i runs over columns, j over rows, pixel centres
    i = (inds % W).astype(np.float32) + np.float32(0.5)
    j = (inds // W).astype(np.float32) + np.float32(0.5)
    fx, fy, cx, cy = (intr[..., k] for k in range(4))
    xs = (i - cx) / fx                                    # utils.py:243-246
    ys = -(j - cy) / fy
    zs = -np.ones_like(i)
    d = np.stack([xs, ys, zs], -1).astype(np.float32)
    R = poses[:, :3, :3]
    rays_d = np.einsum("nm,nkm->nk", d.astype(np.float64), np.broadcast_to(R, (len(d), 3, 3)).astype(np.float64)).astype(np.float32)
    rays_o = np.broadcast_to(poses[:, :3, 3], rays_d.shape).astype(np.float32)              # utils.py:255
    ix, iy = inds // W, inds % W                          # utils.py:265-271
    coarse = (ix * (incoherent_mask_size / H)).astype(np.int64) * incoherent_mask_size + (iy * (incoherent_mask_size / W)).astype(np.int64)
    return rays_o, rays_d, i.astype(np.int64), j.astype(np.int64), coarse
