"""Oracle (test infrastructure, see oracle/__init__.py): CPU restatement of the reference's MIM training step.

Reference: /root/reference/Self-supervised_segmentation (abbrev. SSS)
  SSS/mim.py:153-182        train_one_epoch: forward, loss.sum().backward(), clip_grad_norm_(5.0), optimizer.step(),
                            lr_scheduler.step_update
  SSS/optimizer.py:14-33    get_pretrain_param_groups (no weight decay for 1-D parameters, biases, skip list)
  SSS/optimizer.py:73-75    torch.optim.AdamW(eps, betas, lr, weight_decay)
  SSS/lr_scheduler.py:26-35 timm CosineLRScheduler(t_initial, lr_min, warmup_lr_init, warmup_t, cycle_limit=1)
Third-party arithmetic restated here: torch.nn.utils.clip_grad_norm_ and torch.optim.AdamW (PyTorch, pinned 1.13.1 by
the reference; container 2.11 -- same algorithm), timm 0.6.12 CosineLRScheduler.  Gradients come from torch autograd
over the functional forward of oracle/vit_oracle.py (itself pinned against the reference's model.py).
Pinned by tests/golden/mim_train_tiny.npz (oracle/make_golden.py --only train: the reference's own MIM module, optimizer
grouping and torch's clip / AdamW run for two steps).
"""
from __future__ import annotations

import math

import torch

from . import vit_oracle as VO

# MIM.no_weight_decay() (model.py:79-83) asks the encoder, which defines no such method: the skip list is empty and the
# 3-D cls_token / pos_embed / mask_token are decayed like weights.
NO_DECAY_SKIP = ()


def full_name(key: str) -> str:
    """state-dict key of the encoder / decoder -> the parameter name MIM.named_parameters() yields."""
    return key if key.startswith("decoder.") else "encoder." + key


def has_weight_decay(key: str, shape) -> bool:
    """SSS/optimizer.py:21-27."""
    name = full_name(key)
    return not (len(shape) == 1 or name.endswith(".bias") or name in NO_DECAY_SKIP)


def mim_loss_and_grads(params: dict, cfg: VO.ViTConfig, x: torch.Tensor, mask: torch.Tensor):
    """params: encoder state-dict keys + 'decoder.0.weight' [Cp^2, D, 1, 1] + 'decoder.0.bias'.  -> (loss, {key: grad})."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    sd = {k: v for k, v in leaves.items() if not k.startswith("decoder.")}
    loss, _, _ = VO.mim_forward(sd, cfg, leaves["decoder.0.weight"], leaves["decoder.0.bias"], x, mask)
    loss.sum().backward()                                            # mim.py:174
    return loss.detach(), {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}


def clip_grad_norm(grads: dict, max_norm: float):
    """torch.nn.utils.clip_grad_norm_ (L2): total = || (||g_i||) ||, coef = clamp(max_norm / (total + 1e-6), max=1)."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return total, {k: g * coef for k, g in grads.items()}


def adamw_update(p, g, m, v, step: int, lr: float, beta1: float, beta2: float, eps: float, weight_decay: float):
    """torch.optim.AdamW, single tensor, amsgrad off: returns (p, m, v)."""
    p = p * (1.0 - lr * weight_decay)
    m = beta1 * m + (1.0 - beta1) * g
    v = beta2 * v + (1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


class TrainState:
    def __init__(self, params: dict):
        self.params = {k: v.detach().clone() for k, v in params.items()}
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.step = 0


def train_step(state: TrainState, cfg: VO.ViTConfig, x, mask, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05,
               clip_grad=5.0):
    """One iteration of SSS/mim.py:172-179 (no accumulation).  -> (loss, grad_norm, clipped grads)."""
    loss, grads = mim_loss_and_grads(state.params, cfg, x, mask)
    total, grads = clip_grad_norm(grads, clip_grad) if clip_grad else (torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float(), grads)
    state.step += 1
    for k in state.params:
        wd = weight_decay if has_weight_decay(k, state.params[k].shape) else 0.0
        state.params[k], state.m[k], state.v[k] = adamw_update(state.params[k], grads[k], state.m[k], state.v[k], state.step, lr,
                                                               betas[0], betas[1], eps, wd)
    return loss, total, grads


def cosine_lr(t: int, base_lr: float, t_initial: int, lr_min: float, warmup_t: int, warmup_lr_init: float) -> float:
    """timm 0.6.12 CosineLRScheduler._get_lr with cycle_limit=1, warmup_prefix=False, t_in_epochs=False."""
    if t < warmup_t:
        return warmup_lr_init + t * (base_lr - warmup_lr_init) / warmup_t
    i = t // t_initial
    if i >= 1:
        return lr_min
    return lr_min + 0.5 * (base_lr - lr_min) * (1 + math.cos(math.pi * (t - t_initial * i) / t_initial))
