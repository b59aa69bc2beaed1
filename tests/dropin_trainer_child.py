"""Child process of tests/test_gpu_dropin.py::test_reference_trainer_*: the REFERENCE's own, unmodified ``Trainer``
(nerf/utils.py:534-2166, staged in oracle/_ref_py) driving THIS repository's operator modules (INTEGRATION.md path (b)):

  --mode train : ``Trainer(..., model=NeRFNetwork(opt), optimizer=torch.optim.Adam(model.get_params(lr), eps=1e-15),
                 ema_decay=0.95, lr_scheduler=LambdaLR, scheduler_update_every_step=True)`` exactly as main.py:296-319 builds it,
                 ``trainer.train_one_epoch(loader)`` over K synthetic ray batches (train_step, backward, post_train_step,
                 optimizer / scheduler step, the per-epoch EMA update; lambda_tv = 0: with eps = 1e-15 any TV gradient moves
                 every sampled row by a full step, which the comparison trainer does not apply), then the reference's own
                 ``trainer.save_checkpoint(full=True)``
  --mode load  : the reference's own ``trainer.load_checkpoint(path)`` of a file written by sanerf_b200.checkpoint

Third-party modules the reference imports at module level and that are not installed here are stubbed: tensorboardX,
imageio, matplotlib, torchmetrics, trimesh, mcubes, lpips (unused on this path), torch_efficient_distloss and torch_ema
(restated: oracle/render_torch.eff_distloss, oracle/ema_ref.py).
"""
import argparse
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, "segment-anything-nerf_b200")
REF_PY = os.path.join(ROOT, "oracle", "_ref_py")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["train", "load"], required=True)
    ap.add_argument("--state", required=True)          # reference-format checkpoint to start from / to load
    ap.add_argument("--out", required=True)
    ap.add_argument("--workspace", required=True)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--rays", type=int, default=512)
    args = ap.parse_args()

    sys.path[:0] = [PKG, ROOT]                          # this repository's operator modules: gridencoder, nerf.network, ...
    import torch

    from oracle import ema_ref
    from oracle import render_torch as R

    _stub("tensorboardX"); _stub("imageio"); _stub("trimesh"); _stub("mcubes"); _stub("lpips")
    cm = types.SimpleNamespace(get_cmap=lambda name, n: (lambda i: (0.5, 0.5, 0.5, 1.0)))
    plt = _stub("matplotlib.pyplot", cm=cm)
    _stub("matplotlib", pyplot=plt)
    tmf = _stub("torchmetrics.functional", structural_similarity_index_measure=None, ssim=None)
    _stub("torchmetrics", functional=tmf)
    _stub("torch_ema", ExponentialMovingAverage=ema_ref.ExponentialMovingAverage)
    _stub("torch_efficient_distloss", eff_distloss=R.eff_distloss)

    import warnings
    warnings.filterwarnings("ignore")
    spec = importlib.util.spec_from_file_location("reference_nerf_utils", os.path.join(REF_PY, "nerf", "utils.py"))
    ref_utils = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_utils)                  # the reference's Trainer, unmodified

    from nerf.network import NeRFNetwork                # THIS repository's model (reference surface)
    import nerf.network as our_network
    assert os.path.realpath(our_network.__file__).startswith(os.path.realpath(PKG))

    opt = types.SimpleNamespace(
        bound=128, contract=True, min_near=0.2, density_thresh=10, num_steps=[128, 64, 32], background="last_sample",
        with_sam=False, with_mask=False, sum_after_mlp=False, sam_use_view_direction=True, mask_mlp_type="default",
        lambda_proposal=1.0, lambda_distort=0.02, lambda_entropy=0.0, lambda_tv=0.0, lambda_wd=0.0, max_ray_batch=16384,
        num_rays=args.rays, num_points=2 ** 18, adaptive_num_rays=False, lr=1e-2, iters=20000, fp16=False, use_wandb=False,
        cache_size=0, cache_interval=4, error_map=False, fused=True)
    dev = torch.device("cuda", 0)
    ck = torch.load(args.state, map_location="cpu")
    model = NeRFNetwork(opt)
    if args.mode == "train":
        model.load_state_dict(ck["model"], strict=False)
    criterion = torch.nn.MSELoss(reduction="none")      # main.py:239
    optimizer = torch.optim.Adam(model.get_params(opt.lr), eps=1e-15)                              # main.py:296
    scheduler = torch.optim.lr_scheduler.LambdaLR(optimizer, lambda it: 0.1 ** min(it / opt.iters, 1))   # main.py:312-313
    trainer = ref_utils.Trainer("ngp", opt, model, device=dev, workspace=args.workspace, optimizer=optimizer,
                                criterion=criterion, ema_decay=0.95, fp16=False, lr_scheduler=scheduler,
                                scheduler_update_every_step=True, use_checkpoint="scratch", eval_interval=1, save_interval=1,
                                mute=True)
    out = {}
    if args.mode == "train":
        g = torch.Generator().manual_seed(int(ck["ray_seed"]))
        batches = []
        for _ in range(args.steps):
            o = (torch.rand(args.rays, 3, generator=g) - 0.5).to(dev)
            d = torch.nn.functional.normalize(torch.randn(args.rays, 3, generator=g), dim=-1).to(dev)
            rgb = torch.rand(args.rays, 3, generator=g).to(dev)
            batches.append({"rays_o": o, "rays_d": d, "images": rgb, "index": [0], "H": 1, "W": args.rays})

        class Loader(list):
            batch_size = 1
            _data = types.SimpleNamespace(epoch=0, global_step=0)

        torch.manual_seed(1234)                         # the jitter stream (torch.rand in the renderer)
        trainer.epoch = 1
        trainer.train_one_epoch(Loader(batches))
        trainer.save_checkpoint(full=True)              # the reference's own writer -> workspace/checkpoints/ngp_ep0001.pth
        out["avg_loss"] = trainer.stats["loss"][-1]
        out["global_step"] = trainer.global_step
        out["params"] = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        out["ema"] = [s.detach().cpu() for s in trainer.ema.shadow_params]
        out["ckpt"] = os.path.join(args.workspace, "checkpoints", "ngp_ep0001.pth")
        out["lr"] = optimizer.param_groups[0]["lr"]
    else:
        logs = []
        trainer.log = lambda *a, **k: logs.append(" ".join(str(x) for x in a))
        trainer.load_checkpoint(args.state)             # the reference's own reader on OUR file
        out["logs"] = logs
        out["global_step"], out["epoch"] = trainer.global_step, trainer.epoch
        out["params"] = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        out["opt_state"] = {i: {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in st.items()}
                            for i, st in enumerate(optimizer.state_dict()["state"].values())}
        out["opt_state_keys"] = sorted(optimizer.state_dict()["state"].keys())
        out["lr"] = optimizer.param_groups[0]["lr"]
        out["ema"] = [s.detach().cpu() for s in trainer.ema.shadow_params]
        out["ema_updates"] = trainer.ema.num_updates
    torch.save(out, args.out)


if __name__ == "__main__":
    main()
