"""The caller of the hot path: the data-parallel training / validation step of
/root/reference MSFNO/Models/train.py (`Trainer.train_epoch_ddp` :201-298, `model_forward` :318-339, `validation`
:533-654, `ready_model` :366-380) and the parameter freezing of `FourCastNetv2_filmed.load_model`
(sfno/model.py:1011-1023), reduced to what drives the kernels:

  * every parameter outside `film_gen` is frozen (or outside `grad_layers` with retrain_film);
  * one process per GPU, DistributedDataParallel(broadcast_buffers=False) over NCCL -- the gradient all-reduce covers
    the trainable FiLM head only (<= 7 M floats);
  * gradient accumulation with `model.no_sync()` on the non-updating micro-batches, multi-step (autoregressive)
    training with a discount factor, exactly the loop structure of train.py:213-290;
  * validation with the reference's two collectives: the scalar loss (SUM / world, train.py:556-557) and the
    per-variable MSE vector [C] (train.py:566-567).

What is NOT carried over: the dist.barrier() before every log / memory / time-limit line (train.py:143,204,316,361,
387,745,755,777,819,828) -- ten host-blocking collectives per iteration that serialise the ranks for no data
dependency -- wandb, zarr / GRIB output and checkpoint plumbing (out of the hot-path scope)."""
import contextlib

import torch
import torch.distributed as dist
import torch.nn as nn


def freeze_backbone(model, retrain_film=False, grad_layers=()):
    """sfno/model.py:1011-1023: with retrain_film only parameters whose name contains one of grad_layers stay trainable,
    otherwise only the FiLM generator's.  Returns the trainable parameters."""
    for name, p in model.named_parameters():
        if retrain_film:
            p.requires_grad = any(layer in name for layer in grad_layers)
        else:
            p.requires_grad = "film_gen" in name
    return [p for p in model.parameters() if p.requires_grad]


class TrainerConfig:
    """The cfg attributes the loops below read (names as in main.py's argparse namespace)."""

    def __init__(self, **kw):
        self.ddp = False
        self.rank = 0
        self.world_size = 1
        self.model_version = "film"          # "film": model(x, cond, scale); anything else: model(x)
        self.accumulation_steps = 0
        self.multi_step_training = 0
        self.training_step_skip = 0
        self.discount_factor = 1.0
        self.multi_step_validation = 0
        self.validation_step_skip = 0
        self.validation_epochs = 1
        self.advanced_logging = True
        self.learning_rate = 1e-4
        self.__dict__.update(kw)


class Trainer:
    """data loaders yield `data` with data[step] = (fields [B,C,H,W], conditioning) for step = 0 .. multi_step + 1, the
    layout of the reference's datasets (data.py:21-231)."""

    def __init__(self, model, cfg, loss_fn=None, device=None, normalise=None):
        self.cfg = cfg
        self.device = device if device is not None else next(model.parameters()).device
        self.net = model
        self.model = model
        self.loss_fn = loss_fn if loss_fn is not None else nn.MSELoss()
        self.valid_loss_fn = self.loss_fn
        self.loss_fn_pervar = nn.MSELoss(reduction="none")
        self.normalise = normalise if normalise is not None else (lambda t: t)
        self.scale = 1.0        # FiLM scale (train.py:46); validation ramps it by 0.002 up to 1 (train.py:640-641)
        self.iter = 0
        self.optimizer = None

    # train.py:366-380
    def ready_model(self, retrain_film=False, grad_layers=()):
        params = freeze_backbone(self.net, retrain_film, grad_layers)
        self.net.train()
        if self.cfg.ddp:
            kw = {"device_ids": [self.device.index]} if self.device.type == "cuda" else {}
            self.model = nn.parallel.DistributedDataParallel(self.net, broadcast_buffers=False, **kw)
        self.optimizer = torch.optim.Adam(params, lr=self.cfg.learning_rate)
        return self.model

    # train.py:318-339
    def model_forward(self, input, data, step, return_gt=True):
        gt = None
        if return_gt:
            gt = self.normalise(data[step + 1][0]).to(self.device, non_blocking=True)
        if self.cfg.model_version == "film":
            cond = data[step][1].to(self.device, non_blocking=True)
            return self.model(input, cond, self.scale), gt
        return self.model(input), gt

    def _micro_batch(self, data):
        """forward (multi-step) + backward of one micro-batch; returns the detached loss (no host sync)."""
        cfg = self.cfg
        loss = 0
        output = None
        for step in range(cfg.multi_step_training + 1):
            input = self.normalise(data[step][0]).to(self.device, non_blocking=True) if step == 0 else output
            want_gt = step % (cfg.training_step_skip + 1) == 0
            output, gt = self.model_forward(input, data, step, return_gt=want_gt)
            if want_gt:
                loss = loss + (self.loss_fn(output, gt) / (cfg.multi_step_training + 1) / (cfg.accumulation_steps + 1)
                               * cfg.discount_factor ** step)
        loss.backward()
        return loss.detach()

    # train.py:201-290 without the per-iteration barriers and .item() host syncs
    def train_epoch(self, loader, max_iters=None):
        losses = []
        for i, data in enumerate(loader):
            update = (i + 1) % (self.cfg.accumulation_steps + 1) == 0
            sync_ctx = contextlib.nullcontext() if (update or not self.cfg.ddp) else self.model.no_sync()
            with sync_ctx:
                loss = self._micro_batch(data)
            if update:
                self.optimizer.step()
                self.model.zero_grad(set_to_none=True)
                self.iter += 1
                losses.append(loss)
                if max_iters is not None and self.iter >= max_iters:
                    break
        return torch.stack(losses) if losses else torch.zeros(0)

    # train.py:533-654
    @torch.no_grad()
    def validation(self, loader):
        cfg = self.cfg
        was_training = self.net.training
        self.model.eval()
        loss_list = [[] for _ in range(cfg.multi_step_validation + 1)]
        pervar_list = [[] for _ in range(cfg.multi_step_validation + 1)]
        for val_idx, data in enumerate(loader):
            output = None
            for step in range(cfg.multi_step_validation + 1):
                input = self.normalise(data[step][0]).to(self.device, non_blocking=True) if step == 0 else output
                want_gt = step % (cfg.validation_step_skip + 1) == 0
                output, gt = self.model_forward(input, data, step, return_gt=want_gt)
                if not want_gt:
                    continue
                val = self.valid_loss_fn(output, gt)
                if cfg.ddp:
                    dist.all_reduce(val, dist.ReduceOp.SUM)
                    val = val / dist.get_world_size()
                loss_list[step].append(val)
                if cfg.advanced_logging:
                    per_var = self.loss_fn_pervar(output, gt).mean(dim=(0, 2, 3))
                    if cfg.ddp:
                        dist.all_reduce(per_var, dist.ReduceOp.SUM)
                        per_var = per_var / dist.get_world_size()
                    pervar_list[step].append(per_var)
            if val_idx == cfg.validation_epochs - 1:
                break
        if was_training:
            self.model.train()
        self.scale = min(1.0, self.scale + 0.002)
        loss = torch.stack([torch.stack(l).mean() for l in loss_list if l])
        pervar = torch.stack([torch.stack(l).mean(dim=0) for l in pervar_list if l]) if any(pervar_list) else None
        return loss, pervar
