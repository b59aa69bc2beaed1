"""Ray generation and feature-map post-processing either side of the render path (reference: ``nerf/utils.py``).

Only the pieces SURVEY §8 f3 names are mirrored here: ``get_rays`` (utils.py:145-279) for the whole image, for given
pixel coordinates and for uniform random pixels - the modes the shipped training / GUI recipes use
(``scripts/train_rgb.sh`` with ``--random_image_batch``, ``test_gui``) - and the resize + pad that turns a rendered
``[h, w, 256]`` feature map into the SAM decoder's ``[1, 256, 64, 64]`` input (utils.py:1421-1428).  The Trainer, its
metrics, checkpoint I/O and the mask-stage patch / incoherent-region samplers stay out of scope.
"""
import numpy as np
import torch
import torch.nn.functional as F

from sanerf_b200 import _lib


def get_rays(poses, intrinsics, H, W, N=-1, patch_size=1, coords=None, device="cpu", incoherent_mask=None,
             include_incoherent_region=False, incoherent_mask_size=128, random_sample=False):
    """poses [1 or N,4,4] cam2world, intrinsics [4] ndarray or [1 or N,4] tensor -> dict(rays_o, rays_d [N,3], (i, j,)
    inds_coarse).  Directions are not normalised (utils.py:246-248).  On a CUDA device one kernel writes both tensors."""
    dev = poses.device if torch.is_tensor(poses) and (poses.is_cuda or device == "cpu") else torch.device(device)
    results = {}
    if N > 0:
        if coords is not None:
            inds = (coords[:, 0] * W + coords[:, 1]).to(dev).long()
        elif not random_sample:
            # utils.py:176-234: patch sampling / multinomial sampling from the incoherent-region mask
            raise NotImplementedError("patch / incoherent-region pixel samplers belong to the mask stage (out of scope): "
                                      "pass coords or random_sample=True")
        else:                                             # utils.py:236-238: uniform random pixels, may repeat
            inds = torch.randint(0, H * W, size=[N], device=dev)
        results["i"] = (inds % W).long()
        results["j"] = (inds // W).long()
    else:
        inds = None
    n_rays = H * W if inds is None else inds.numel()
    poses_t = torch.as_tensor(poses, dtype=torch.float32, device=dev).reshape(-1, 4, 4).contiguous()
    if isinstance(intrinsics, np.ndarray):
        intr_t = torch.from_numpy(np.asarray(intrinsics, dtype=np.float32)).to(dev).reshape(-1, 4)
    else:
        intr_t = torch.as_tensor(intrinsics, dtype=torch.float32, device=dev).reshape(-1, 4)
    intr_t = intr_t.contiguous()
    if poses_t.shape[0] not in (1, n_rays) or intr_t.shape[0] not in (1, n_rays):
        raise RuntimeError("get_rays: one pose / intrinsics row for all rays or one per ray")
    if dev.type == "cuda":
        rays_o = torch.empty(n_rays, 3, device=dev)
        rays_d = torch.empty(n_rays, 3, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev), _lib.stats.span("generate_rays", N=n_rays):
            rc = lib.sanerf_generate_rays(poses_t.data_ptr(), 0 if poses_t.shape[0] == 1 else 16, intr_t.data_ptr(),
                                          0 if intr_t.shape[0] == 1 else 4, _lib.ptr(inds), W, n_rays, rays_o.data_ptr(),
                                          rays_d.data_ptr(), _lib.current_stream(dev))
        _lib.check(rc, "generate_rays")
    else:
        raise RuntimeError("get_rays: poses must live on a CUDA device (no CPU fallback)")
    results["rays_o"], results["rays_d"] = rays_o, rays_d
    flat = torch.arange(H * W, device=dev) if inds is None else inds
    ix, iy = flat // W, flat % W                          # utils.py:265-271
    results["inds_coarse"] = ((ix * (incoherent_mask_size / H)).long() * incoherent_mask_size
                              + (iy * (incoherent_mask_size / W)).long()).long()
    return results


def sam_decoder_features(samvit):
    """Rendered feature map [h, w, 256] (renderer output) or [1, 256, h, w] -> the SAM mask decoder's [1, 256, 64, 64] input:
    bilinear resize of the longer side to 64, zero padding of the other (utils.py:1421-1428)."""
    feats = samvit.permute(2, 0, 1).unsqueeze(0) if samvit.dim() == 3 else samvit
    h, w = feats.shape[2:]
    ratio = 64 / w if w > h else 64 / h
    feats = F.interpolate(feats, (int(h * ratio), int(w * ratio)), mode="bilinear", align_corners=False)
    return F.pad(feats, (0, 64 - feats.shape[3], 0, 64 - feats.shape[2]), mode="constant", value=0)
