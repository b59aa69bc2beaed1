"""Child process of tests/test_gpu_dropin.py (never imported by pytest itself).

Runs the REFERENCE's own, unmodified Python modules (staged in oracle/_ref_py by oracle/stage_ref_py.py:
gridencoder/grid.py, shencoder/sphere_harmonics.py, freqencoder/freq.py, encoding.py, activation.py, nerf/network.py,
nerf/renderer.py) in a fresh interpreter whose module path resolves the compiled-backend names the reference imports —
``_gridencoder``, ``_shencoder``, ``_freqencoder`` (grid.py:9-12 and siblings) — either

  --backend ours : to THIS repository's shims (segment-anything-nerf_b200/_gridencoder.py ...), i.e. INTEGRATION.md path (a)
  --backend ref  : to the reference CUDA extensions rebuilt in oracle/_ref (the true reference result on this GPU)

and writes the outputs and gradients of one training-mode ``render()`` of the reference's NeRFNetwork from a given state_dict.
"""
import argparse
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, "segment-anything-nerf_b200")
REF_PY = os.path.join(ROOT, "oracle", "_ref_py")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", choices=["ours", "ref"], required=True)
    ap.add_argument("--state", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--with-sam", type=int, default=0)
    ap.add_argument("--rays", type=int, default=512)
    ap.add_argument("--hw", type=int, default=16)
    args = ap.parse_args()

    import torch

    # From THIS repository the child takes only the three compiled-backend names (+ the sanerf_b200 ctypes binding they
    # call): they are imported first, then the shim directory leaves the module path again, so that ``gridencoder``,
    # ``shencoder``, ``freqencoder``, ``encoding``, ``activation`` and ``nerf`` can only be the reference's files.
    sys.path.insert(0, PKG)
    import _freqencoder  # noqa: F401
    import _gridencoder  # noqa: F401
    import _shencoder  # noqa: F401
    sys.path.remove(PKG)
    for name in [m for m in sys.modules if m.split(".")[0] in ("gridencoder", "shencoder", "freqencoder", "nerf",
                                                               "encoding", "activation")]:
        raise AssertionError(f"{name} imported too early")
    sys.path[:0] = [REF_PY, ROOT]

    from oracle import build_ref
    from oracle import render_torch as R

    for name in ("mcubes", "trimesh"):
        sys.modules.setdefault(name, types.ModuleType(name))
    dl = types.ModuleType("torch_efficient_distloss")          # un-vendored third-party package (requirements.txt:22)
    dl.eff_distloss = R.eff_distloss
    sys.modules["torch_efficient_distloss"] = dl
    if args.backend == "ref":
        for ext in ("gridencoder", "shencoder", "freqencoder"):
            sys.modules[f"_{ext}"] = build_ref.load(ext)

    import warnings
    warnings.filterwarnings("ignore")
    import gridencoder.grid as ref_grid
    import shencoder.sphere_harmonics as ref_sh
    from nerf import network as ref_network

    for mod in (ref_grid, ref_sh, ref_network):
        assert os.path.realpath(mod.__file__).startswith(os.path.realpath(REF_PY)), mod.__file__
    backend_file = os.path.realpath(getattr(ref_grid._backend, "__file__", ""))
    if args.backend == "ours":
        assert backend_file == os.path.realpath(os.path.join(PKG, "_gridencoder.py")), backend_file
        assert os.path.realpath(ref_sh._backend.__file__) == os.path.realpath(os.path.join(PKG, "_shencoder.py"))
    else:
        assert backend_file.startswith(os.path.realpath(os.path.join(ROOT, "oracle", "_ref"))), backend_file

    opt = types.SimpleNamespace(bound=128, contract=True, min_near=0.2, density_thresh=10, num_steps=[128, 64, 32],
                                background="last_sample", with_sam=bool(args.with_sam), with_mask=False, sum_after_mlp=False,
                                sam_use_view_direction=True, mask_mlp_type="default", adaptive_mlp_type="rgb", n_inst=2,
                                redundant_instance=0, lambda_proposal=1.0, lambda_distort=0.02, max_ray_batch=16384)
    model = ref_network.NeRFNetwork(opt)
    state = torch.load(args.state, map_location="cpu")
    missing, unexpected = model.load_state_dict(state["model"], strict=False)
    assert not unexpected, unexpected
    assert all("aabb" in m for m in missing), missing
    model = model.cuda().train()

    n = args.rays if not args.with_sam else args.hw * args.hw
    g = torch.Generator().manual_seed(int(state["ray_seed"]))
    o = (torch.rand(n, 3, generator=g) - 0.5).cuda()
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda()
    out = {}
    if not args.with_sam:
        gt = torch.rand(n, 3, generator=g).cuda()
        res = model.render(o, d, staged=False, bg_color=1, perturb=False, update_proposal=True)
        loss = torch.nn.functional.mse_loss(res["image"], gt, reduction="none").mean()          # nerf/utils.py:921-930
        loss = loss + opt.lambda_proposal * res["proposal_loss"] + opt.lambda_distort * res["distort_loss"]
        for k in ("image", "depth", "weights_sum", "weights", "proposal_loss", "distort_loss"):
            out[k] = res[k].detach().cpu()
    else:
        target = torch.randn(1, 256, args.hw, args.hw, generator=g).cuda()
        res = model.render(o, d, staged=False, bg_color=1, perturb=False, update_proposal=False, return_feats=1,
                           H=args.hw, W=args.hw)
        pred = res["samvit"].permute(2, 0, 1).unsqueeze(0)                                       # nerf/utils.py:1100-1106
        loss = torch.nn.functional.mse_loss(pred, target)
        for k in ("image", "depth", "samvit"):
            out[k] = res[k].detach().cpu()
    loss.backward()
    out["loss"] = loss.detach().cpu()
    out["grads"] = {k: p.grad.detach().cpu() for k, p in model.named_parameters() if p.grad is not None}
    out["keys"] = list(model.state_dict().keys())
    torch.save(out, args.out)


if __name__ == "__main__":
    main()
