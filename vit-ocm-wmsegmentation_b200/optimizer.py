"""Drop-in for the reference's ``optimizer.py`` (SSS/optimizer.py) on B200: parameter grouping + AdamW for MIM
pre-training, with the update itself in libvitocm (one fused kernel over the flat parameter / gradient / moment
buffers of ``model.MIM``: gradient clipping coefficient, decoupled weight decay, Adam moments, bias correction).

    optimizer = build_pretrain_optimizer(config, model, logger)        # SSS/mim.py:106
    ...
    loss.sum().backward()
    grad_norm = clip_grad_norm_(model.parameters(), config.TRAIN.CLIP_GRAD)   # this module's, or torch's
    optimizer.step()

``clip_grad_norm_`` here measures the norm (fp64 sum of squares on the device, no host sync) and scales the flat gradient in
place by min(1, max_norm / (norm + 1e-6)) -- ``.grad`` holds the clipped values from then on, exactly as after
``torch.nn.utils.clip_grad_norm_``; ``optimizer.step()`` then is ``torch.optim.AdamW``'s update on whatever ``.grad`` holds.
In data-parallel training the order is backward -> ``MIM.all_reduce_grads()`` -> ``clip_grad_norm_`` -> ``step``.
"""
from __future__ import annotations

import weakref

import torch

from . import _lib
from ._lib import check, cur_stream, ptr


def check_keywords_in_name(name, keywords=()):
    return any(k in name for k in keywords)


def get_pretrain_param_groups(model, logger=None, skip_list=(), skip_keywords=()):
    """SSS/optimizer.py:14-33: no weight decay for 1-D parameters, biases, and the skip list / keywords."""
    has_decay, no_decay, has_decay_name, no_decay_name = [], [], [], []
    for name, param in model.named_parameters():
        if not param.requires_grad:
            continue
        if len(param.shape) == 1 or name.endswith(".bias") or (name in skip_list) or check_keywords_in_name(name, skip_keywords):
            no_decay.append(param)
            no_decay_name.append(name)
        else:
            has_decay.append(param)
            has_decay_name.append(name)
    if logger is not None:
        logger.info(f'No decay params: {no_decay_name}')
        logger.info(f'Has decay params: {has_decay_name}')
    return [{'params': has_decay}, {'params': no_decay, 'weight_decay': 0.}]


_PARAM_OWNER: dict = {}     # id(parameter) -> weakref to the FusedAdamW that owns its flat buffers


def _unwrap(model):
    return model.module if hasattr(model, "module") else model


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction, eps outside the square root) executed by
    ``vitocm_adamw_step`` on the flat buffers of a ``model.MIM``.  ``param_groups`` keeps the reference's two groups so
    that LR schedulers that write ``group['lr']`` work unchanged (both groups must carry the same lr)."""

    def __init__(self, model, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05, skip_list=(), skip_keywords=()):
        mim = _unwrap(model)
        mim.flatten_parameters()
        groups = get_pretrain_param_groups(mim, None, skip_list, skip_keywords)
        # the extra keys are torch.optim.AdamW's own defaults: a saved state dict then carries complete groups for the stock optimizer
        super().__init__(groups, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False, foreach=None,
                                      capturable=False, differentiable=False, fused=None))
        self.mim = mim
        self.decay = mim.decay_flags(skip_list, skip_keywords)
        self.exp_avg = torch.zeros_like(mim._pflat)
        self.exp_avg_sq = torch.zeros_like(mim._pflat)
        self.steps = 0
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=mim._pflat.device)
        mim._fused_optimizer = self        # lets clip_grad_norm_(model, ...) find its optimizer
        for _, p in mim._param_list:       # ... and clip_grad_norm_(model.parameters(), ...), the reference's call form
            _PARAM_OWNER[id(p)] = weakref.ref(self)

    def measure_grad_norm(self, max_norm: float = 0.0) -> torch.Tensor:
        """Global L2 norm of the flat gradient (device tensor, no host sync); with max_norm > 0 the gradient is clipped IN
        PLACE right away, like torch.nn.utils.clip_grad_norm_ (SSS/mim.py:165-176 clips ``.grad`` on every micro-step):
        whatever happens to ``.grad`` between this call and ``step()`` -- another micro-step's backward, an all-reduce --
        acts on the clipped values, and ``step()`` never sees a stale coefficient.  Order in data-parallel training:
        backward -> all_reduce_grads -> clip_grad_norm_ -> step (the clip must see the reduced gradient)."""
        g = self.mim._gflat
        lib = _lib.load_library()
        check(lib.vitocm_grad_sumsq(ptr(g), g.numel(), ptr(self._sumsq), cur_stream()))
        norm = self._sumsq.sqrt().to(torch.float32)[0]
        if max_norm is not None and float(max_norm) > 0:
            check(lib.vitocm_grad_clip(ptr(g), g.numel(), float(max_norm), ptr(self._sumsq), cur_stream()))
        return norm

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        mim = self.mim
        if mim._pflat is None or any(p.grad is None for _, p in mim._param_list):
            raise _lib.VitocmError("FusedAdamW.step: no gradients (call loss.backward() first)")
        g0, g1 = self.param_groups[0], self.param_groups[1]
        if g0["lr"] != g1["lr"]:
            raise _lib.VitocmError("FusedAdamW: both parameter groups must carry the same learning rate")
        self.steps += 1
        b1, b2 = g0["betas"]
        n = mim._pflat.numel()
        check(_lib.load_library().vitocm_adamw_step(ptr(mim._pflat), ptr(mim._gflat), ptr(self.exp_avg), ptr(self.exp_avg_sq), ptr(self.decay), n,
                                                    float(g0["lr"]), float(b1), float(b2), float(g0["eps"]), float(g0["weight_decay"]),
                                                    self.steps, 0.0, float(grad_scale), None, cur_stream()))
        mim.encoder.refresh_engine()      # asynchronous bf16 repack of the updated masters (vitocm_refresh_weights)
        return None

    # ------------------------------------------------------------------ checkpoint interchange (SSS/utils.py:375-385)
    def _flat_slices(self):
        """id(parameter) -> (offset, numel, shape) inside the flat buffers."""
        return {id(p): (o, p.numel(), tuple(p.shape)) for (_, p), o in zip(self.mim._param_list, self.mim._flat_offsets)}

    def state_dict(self):
        """The layout ``torch.optim.AdamW.state_dict()`` has for the same two parameter groups (per-parameter ``step`` /
        ``exp_avg`` / ``exp_avg_sq`` indexed in group order), cut out of the flat moment buffers: a checkpoint written here
        resumes under the reference's stock optimizer and the other way round."""
        sd = super().state_dict()
        state, idx, where = {}, 0, self._flat_slices()
        for grp in self.param_groups:
            for p in grp["params"]:
                if self.steps > 0:
                    o, n, shape = where[id(p)]
                    state[idx] = {"step": torch.tensor(float(self.steps)),
                                  "exp_avg": self.exp_avg[o:o + n].view(shape).clone(),
                                  "exp_avg_sq": self.exp_avg_sq[o:o + n].view(shape).clone()}
                idx += 1
        sd["state"] = state
        return sd

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        groups = state_dict["param_groups"]
        if len(groups) != len(self.param_groups) or any(len(a["params"]) != len(b["params"]) for a, b in zip(groups, self.param_groups)):
            raise ValueError("loaded state dict has different parameter groups")
        for mine, theirs in zip(self.param_groups, groups):
            mine.update({k: v for k, v in theirs.items() if k != "params"})
        where, idx, steps = self._flat_slices(), 0, set()
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        for mine, theirs in zip(self.param_groups, groups):
            for p, key in zip(mine["params"], theirs["params"]):
                st = state_dict["state"].get(key)
                if st is not None:
                    o, n, _ = where[id(p)]
                    self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
                    self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                    steps.add(int(st["step"]))
                idx += 1
        if len(steps) > 1:
            raise ValueError("FusedAdamW keeps one step counter: the loaded per-parameter steps differ")
        self.steps = steps.pop() if steps else 0

    def zero_grad(self, set_to_none: bool = False):
        """One memset of the flat gradient buffer; the ``.grad`` views stay attached."""
        mim = self.mim
        if mim._gflat is not None:
            mim._gflat.zero_()
            for (_, p), g in zip(mim._param_list, mim._grad_views):
                p.grad = g


def clip_grad_norm_(parameters, max_norm, optimizer: FusedAdamW | None = None):
    """``torch.nn.utils.clip_grad_norm_`` for the fused path: clips the flat gradient in place and returns the total norm
    (device tensor, no sync).  ``parameters`` may be the model, ``model.parameters()`` (any iterable of its parameters -- the
    reference's call form, SSS/mim.py:176), a single parameter, or a FusedAdamW; the optimizer is found through the model /
    the parameters it was built for.  All parameters of the model are clipped together (the reference always passes all)."""
    opt = optimizer
    if opt is None and isinstance(parameters, FusedAdamW):
        opt = parameters
    if opt is None and hasattr(parameters, "parameters"):
        opt = getattr(_unwrap(parameters), "_fused_optimizer", None)
    if opt is None:
        first = parameters if isinstance(parameters, torch.Tensor) else next(iter(parameters), None)
        ref = _PARAM_OWNER.get(id(first)) if first is not None else None
        opt = ref() if ref is not None else None
    if opt is None:
        raise _lib.VitocmError("clip_grad_norm_: pass the FusedAdamW optimizer, the model it was built for, or that model's parameters")
    return opt.measure_grad_norm(max_norm)


def build_pretrain_optimizer(args, model, logger=None):
    """SSS/optimizer.py:47-78 (AdamW branch; lr / betas / eps / weight decay from args.TRAIN)."""
    mim = _unwrap(model)
    skip = mim.no_weight_decay() if hasattr(mim, 'no_weight_decay') else {}
    skip_keywords = mim.no_weight_decay_keywords() if hasattr(mim, 'no_weight_decay_keywords') else {}
    name = args.TRAIN.OPTIMIZER.NAME.lower()
    if name != 'adamw':
        raise NotImplementedError("vitocm: only the AdamW branch of build_pretrain_optimizer (the reference's default) is fused")
    opt = FusedAdamW(mim, lr=args.TRAIN.BASE_LR, betas=tuple(args.TRAIN.OPTIMIZER.BETAS), eps=args.TRAIN.OPTIMIZER.EPS,
                     weight_decay=args.TRAIN.WEIGHT_DECAY, skip_list=skip, skip_keywords=skip_keywords)
    mim._fused_optimizer = opt
    if logger is not None:
        logger.info(opt)
    return opt
