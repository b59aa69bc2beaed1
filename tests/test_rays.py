"""Ray generation: oracle vs the reference's own get_rays (golden), CUDA kernel vs oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import rays_np

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_rays.npz"))


def test_oracle_full_image_matches_reference():
    H, W = GOLD["full_hw"]
    o, d, _, _, coarse = rays_np.get_rays(GOLD["full_pose"][None], GOLD["full_intr"], int(H), int(W))
    np.testing.assert_allclose(d, GOLD["full_d"], rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(o, GOLD["full_o"])
    np.testing.assert_array_equal(coarse, GOLD["full_inds_coarse"])


def test_oracle_coords_per_ray_pose_matches_reference():
    H, W = GOLD["full_hw"]
    inds = GOLD["co_coords"][:, 0] * int(W) + GOLD["co_coords"][:, 1]
    o, d, i, j, coarse = rays_np.get_rays(GOLD["co_poses"], GOLD["co_intr"], int(H), int(W), inds)
    np.testing.assert_allclose(d, GOLD["co_d"], rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(o, GOLD["co_o"])
    np.testing.assert_array_equal(i, GOLD["co_i"])
    np.testing.assert_array_equal(j, GOLD["co_j"])
    np.testing.assert_array_equal(coarse, GOLD["co_inds_coarse"])


@pytest.mark.gpu
def test_cuda_get_rays_matches_oracle_and_golden(cuda):
    from nerf.utils import get_rays, sam_decoder_features
    H, W = (int(v) for v in GOLD["full_hw"])
    r = get_rays(torch.from_numpy(GOLD["full_pose"])[None].cuda(), GOLD["full_intr"], H, W, -1)
    np.testing.assert_allclose(r["rays_d"].cpu().numpy(), GOLD["full_d"], rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(r["rays_o"].cpu().numpy(), GOLD["full_o"])
    np.testing.assert_array_equal(r["inds_coarse"].cpu().numpy(), GOLD["full_inds_coarse"])
    r = get_rays(torch.from_numpy(GOLD["co_poses"]).cuda(), torch.from_numpy(GOLD["co_intr"]).cuda(), H, W, 37,
                 coords=torch.from_numpy(GOLD["co_coords"]).cuda())
    np.testing.assert_allclose(r["rays_d"].cpu().numpy(), GOLD["co_d"], rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(r["rays_o"].cpu().numpy(), GOLD["co_o"])
    np.testing.assert_array_equal(r["i"].cpu().numpy(), GOLD["co_i"])
    np.testing.assert_array_equal(r["j"].cpu().numpy(), GOLD["co_j"])
    # a 512 x 512 frame against the oracle, and random pixels stay inside the image
    pose = torch.from_numpy(GOLD["full_pose"])[None].cuda()
    intr = np.array([443.4, 443.4, 256.0, 256.0], dtype=np.float32)
    big = get_rays(pose, intr, 512, 512, -1)
    o, d, *_ = rays_np.get_rays(GOLD["full_pose"][None], intr, 512, 512)
    np.testing.assert_allclose(big["rays_d"].cpu().numpy(), d, rtol=1e-6, atol=1e-6)
    rnd = get_rays(pose, intr, 512, 512, 1000, random_sample=True)
    assert rnd["rays_d"].shape == (1000, 3) and int(rnd["i"].max()) < 512 and int(rnd["j"].min()) >= 0
    # feature-map post-processing: longer side -> 64, zero padding
    f = sam_decoder_features(torch.randn(48, 64, 256, device="cuda"))
    assert f.shape == (1, 256, 64, 64) and float(f[:, :, 48:].abs().max()) == 0.0
    with pytest.raises(RuntimeError):
        get_rays(torch.from_numpy(GOLD["full_pose"])[None], GOLD["full_intr"], H, W, -1)     # CPU poses: no fallback
