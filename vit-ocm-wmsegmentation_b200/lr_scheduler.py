"""Drop-in for the reference's ``lr_scheduler.py`` (SSS/lr_scheduler.py:18-62): per-update learning-rate schedules with
linear warm-up, driven by ``step_update(num_updates)`` (SSS/mim.py:179).  Host-side arithmetic only (timm, which the
reference imports its schedulers from, is not a dependency here; the cosine formula is timm 0.6.12's CosineLRScheduler
with cycle_limit=1, t_in_epochs=False)."""
from __future__ import annotations

import math


class _WarmupScheduler:
    def __init__(self, optimizer, warmup_t=0, warmup_lr_init=0.0):
        self.optimizer = optimizer
        for g in optimizer.param_groups:
            g.setdefault("initial_lr", g["lr"])
        self.base_values = [g["initial_lr"] for g in optimizer.param_groups]
        self.warmup_t = warmup_t
        self.warmup_lr_init = warmup_lr_init
        if warmup_t:
            self.warmup_steps = [(v - warmup_lr_init) / warmup_t for v in self.base_values]
            self._set(self._get_lr(0))
        else:
            self.warmup_steps = [1 for _ in self.base_values]

    def _set(self, values):
        for g, v in zip(self.optimizer.param_groups, values):
            g["lr"] = v

    def _after_warmup(self, t):
        raise NotImplementedError

    def _get_lr(self, t):
        if t < self.warmup_t:
            return [self.warmup_lr_init + t * s for s in self.warmup_steps]
        return self._after_warmup(t)

    def step_update(self, num_updates: int):
        self._set(self._get_lr(num_updates))

    def step(self, epoch: int):      # t_in_epochs=False: epochs do not move the rate
        return None

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if k != "optimizer"}

    def load_state_dict(self, sd):
        self.__dict__.update(sd)


class CosineLRScheduler(_WarmupScheduler):
    def __init__(self, optimizer, t_initial, lr_min=0.0, warmup_t=0, warmup_lr_init=0.0, cycle_limit=1):
        self.t_initial, self.lr_min, self.cycle_limit = t_initial, lr_min, cycle_limit
        super().__init__(optimizer, warmup_t, warmup_lr_init)

    def _after_warmup(self, t):
        i = t // self.t_initial
        t_curr = t - self.t_initial * i
        if i < self.cycle_limit:
            return [self.lr_min + 0.5 * (v - self.lr_min) * (1 + math.cos(math.pi * t_curr / self.t_initial)) for v in self.base_values]
        return [self.lr_min for _ in self.base_values]


class LinearLRScheduler(_WarmupScheduler):
    """SSS/lr_scheduler.py:65-116."""

    def __init__(self, optimizer, t_initial, lr_min_rate, warmup_t=0, warmup_lr_init=0.0):
        self.t_initial, self.lr_min_rate = t_initial, lr_min_rate
        super().__init__(optimizer, warmup_t, warmup_lr_init)

    def _after_warmup(self, t):
        t = t - self.warmup_t
        total_t = self.t_initial - self.warmup_t
        return [v - ((v - v * self.lr_min_rate) * (t / total_t)) for v in self.base_values]


def build_scheduler(config, optimizer, n_iter_per_epoch):
    """SSS/lr_scheduler.py:18-62 ('cosine' -- the reference default -- and 'linear')."""
    num_steps = int(config.TRAIN.EPOCHS * n_iter_per_epoch)
    warmup_steps = int(config.TRAIN.WARMUP_EPOCHS * n_iter_per_epoch)
    name = config.TRAIN.LR_SCHEDULER.NAME
    if name == 'cosine':
        return CosineLRScheduler(optimizer, t_initial=num_steps, lr_min=config.TRAIN.MIN_LR, warmup_lr_init=config.TRAIN.WARMUP_LR,
                                 warmup_t=warmup_steps, cycle_limit=1)
    if name == 'linear':
        return LinearLRScheduler(optimizer, t_initial=num_steps, lr_min_rate=0.01, warmup_lr_init=config.TRAIN.WARMUP_LR,
                                 warmup_t=warmup_steps)
    raise NotImplementedError(f"vitocm: lr scheduler '{name}' is not mirrored (cosine, linear are)")
