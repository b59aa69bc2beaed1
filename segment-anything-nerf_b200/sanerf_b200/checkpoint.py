"""Checkpoints in the reference's on-disk format, and the stage-2 warm start.

The reference's ``Trainer.save_checkpoint`` (nerf/utils.py:2041-2100) writes ``torch.save({'epoch', 'global_step',
'stats', 'model': model.state_dict()[, 'optimizer', 'lr_scheduler', 'scaler', 'ema']})``; ``load_checkpoint``
(:2102-2166) accepts that dict or a bare state_dict and loads the model with ``strict=False``.  ``NeRFNetwork`` here keeps
the reference's sub-module names, so the ``'model'`` entry is interchangeable in both directions; the optimizer entry is
written in ``torch.optim.Adam.state_dict()`` layout (per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``) from the flat
buffers of ``FusedAdam`` and read back into them; ``'ema'`` follows ``torch_ema.ExponentialMovingAverage.state_dict()``.

The optimizer entry mirrors what ``torch.optim.Adam(model.get_params(lr), eps=1e-15).state_dict()`` holds in the reference
(main.py:296 with network.py:278-308): one param group per sub-module in ``get_params`` order, parameter indices running
through the groups, state only for parameters that were updated — so the reference's own ``Trainer.load_checkpoint`` restores
it into its torch optimizer, and a file written by the reference's ``save_checkpoint(full=True)`` restores ``FusedAdam``
(``tests/test_gpu_dropin.py`` does both with the reference's unmodified Trainer).

``warm_start`` is main.py:255-262: load a stage-1 checkpoint non-strictly and freeze every parameter it provided.
"""
from __future__ import annotations

import torch


def _flush(trainer):
    if trainer is not None and hasattr(trainer, "flush"):
        trainer.flush()                       # a deferred table update must land before parameters are read or replaced


def _param_groups(model, opt):
    """([indices per group], [parameters in index order]) as ``torch.optim.Adam(model.get_params(lr))`` numbers them
    (network.py:278-308: grid, grid_mlp, view_mlp, prop_encoders, prop_mlp[, s_grid, samvit_mlp])."""
    groups, order = [], []
    for g in model.get_params(opt.lr):
        ps = list(g["params"])
        groups.append(list(range(len(order), len(order) + len(ps))))
        order += ps
    return groups, order


def checkpoint_state(model, trainer=None, epoch=0, stats=None, full=False):
    """The dict ``Trainer.save_checkpoint`` builds (nerf/utils.py:2046-2063)."""
    _flush(trainer)
    state = {"epoch": int(epoch), "global_step": int(getattr(trainer, "global_step", 0)),
             "stats": stats if stats is not None else {"loss": [], "valid_loss": [], "results": [], "checkpoints": [],
                                                        "best_result": None},
             "model": {k: v.detach().clone() for k, v in model.state_dict().items()}}
    if full and trainer is not None:
        opt = trainer.optimizer
        if opt.sharded:
            opt.gather_sharded_state()                    # multi-GPU: Adam moments are rank-sharded (collective)
        step = float(opt.step_count.item())
        lr_now = float(opt.dyn[0].item())
        groups, order = _param_groups(model, opt)
        per_param = {}
        for i, p in enumerate(order):
            if id(p) not in opt.ranges or step == 0:
                continue                                  # frozen / never updated: torch's Adam holds no state for it
            a, _ = opt.ranges[id(p)]
            n = p.numel()
            per_param[i] = {"step": torch.tensor(step), "exp_avg": opt.exp_avg[a:a + n].view_as(p).clone(),
                            "exp_avg_sq": opt.exp_avg_sq[a:a + n].view_as(p).clone()}
        state["optimizer"] = {"state": per_param,
                              "param_groups": [{"lr": lr_now, "betas": tuple(opt.betas), "eps": opt.eps, "weight_decay": 0,
                                                "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                                                "differentiable": False, "fused": None, "decoupled_weight_decay": False,
                                                "initial_lr": opt.lr, "params": idx} for idx in groups]}
        state["lr_scheduler"] = {"base_lrs": [opt.lr] * len(groups), "last_epoch": int(step), "verbose": False,
                                 "_step_count": int(step) + 1, "_get_lr_called_within_step": False,
                                 "_last_lr": [lr_now] * len(groups), "lr_lambdas": [None] * len(groups)}
        state["scaler"] = {}                              # torch.cuda.amp.GradScaler(enabled=False).state_dict() (main.py:222: fp16 off)
        if opt.ema is not None:
            state["ema"] = opt.ema_state_dict()
    return state


def save_checkpoint(path, model, trainer=None, epoch=0, stats=None, full=False):
    torch.save(checkpoint_state(model, trainer, epoch, stats, full), path)


def load_checkpoint(checkpoint, model, trainer=None, model_only=False, map_location=None):
    """``Trainer.load_checkpoint`` (nerf/utils.py:2102-2166): ``checkpoint`` is a path or an already loaded dict, either
    the full layout or a bare state_dict.  Returns (missing_keys, unexpected_keys)."""
    _flush(trainer)
    ckpt = torch.load(checkpoint, map_location=map_location or "cpu") if isinstance(checkpoint, (str, bytes)) or \
        hasattr(checkpoint, "read") else checkpoint
    if "model" not in ckpt:
        res = model.load_state_dict(ckpt)                 # parameters are views of the flat buffer: copied in place
        return list(res.missing_keys), list(res.unexpected_keys)
    res = model.load_state_dict(ckpt["model"], strict=False)
    if trainer is not None and trainer.optimizer.ema is not None and "ema" in ckpt:
        trainer.optimizer.load_ema_state_dict(ckpt["ema"])
    if model_only or trainer is None:
        return list(res.missing_keys), list(res.unexpected_keys)
    trainer.global_step = int(ckpt.get("global_step", 0))
    opt = trainer.optimizer
    if "optimizer" in ckpt:
        st = ckpt["optimizer"]["state"]
        _, order = _param_groups(model, opt)
        step = 0
        for i, p in enumerate(order):
            if i not in st or id(p) not in opt.ranges:
                continue
            a, _ = opt.ranges[id(p)]
            n = p.numel()
            opt.exp_avg[a:a + n].copy_(st[i]["exp_avg"].reshape(-1))
            opt.exp_avg_sq[a:a + n].copy_(st[i]["exp_avg_sq"].reshape(-1))
            step = max(step, int(float(st[i]["step"])))
        opt.step_count.fill_(step)
    return list(res.missing_keys), list(res.unexpected_keys)


def warm_start(model, init_ckpt, map_location=None):
    """main.py:255-262: ``model.load_state_dict(torch.load(init_ckpt)['model'], strict=False)`` and
    ``requires_grad = False`` for every parameter whose key the checkpoint holds.  Returns the frozen keys."""
    ckpt = torch.load(init_ckpt, map_location=map_location or "cpu") if isinstance(init_ckpt, (str, bytes)) or \
        hasattr(init_ckpt, "read") else init_ckpt
    model_dict = ckpt["model"]
    model.load_state_dict(model_dict, strict=False)
    frozen = []
    for k, v in model.named_parameters():
        if k in model_dict:
            v.requires_grad = False
            frozen.append(k)
    return frozen
