"""Host side of the training step around the model (first slice): the pieces of `main.py:367-512`
that are not the network itself, over the C ABI.

  * `rec_losses`            the L1 reconstruction losses Eq.(9)-(10) of main.py:404-440 (same names: total, rec_D1, rec_D2)
                            and, fused into the same pass, the gradients w.r.t. every predicted frame
  * `Adam`                  torch.optim.Adam(params, lr, betas=(0.9, 0.999), weight_decay) of main.py:179-180: same state
                            (`exp_avg`, `exp_avg_sq`, `step`), same update, one `demfi_adam_step` launch per parameter tensor
  * `allreduce_gradients`   the one exchange step of data-parallel training: the 260 gradient tensors
                            (29.6 MB) flattened into one bucket, summed over ranks with `torch.distributed.all_reduce` (NCCL over
                            NVLink on the GPUs; gloo in the CPU test) and averaged

No CPU fallback for the kernels: `rec_losses` and `Adam.step` raise on CPU tensors.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

from . import _abi as A

_WS: Dict[int, torch.Tensor] = {}


def _ws(dev, nbytes):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    t = _WS.get(key)
    if t is None or t.numel() * 8 < nbytes:
        t = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=dev)
        _WS[key] = t
    return t


def _l1_term(pred: torch.Tensor, target: torch.Tensor, grad_scale: float, want_grad: bool, out: torch.Tensor):
    if not (pred.is_cuda and target.is_cuda):
        raise A.DemfiError("demfi_b200.train.rec_losses runs on the GPU only (no CPU fallback)")
    if pred.shape != target.shape:
        raise ValueError("prediction and ground truth must have the same shape")
    pred, target = pred.detach().contiguous().float(), target.detach().contiguous().float()
    n = pred.numel()
    lib = A.lib()
    need = lib.demfi_l1_sum_workspace(n)
    ws = _ws(pred.device, need)
    grad = torch.empty_like(pred) if want_grad else None
    with torch.cuda.device(pred.device):
        A.check(lib.demfi_l1_sum(pred.data_ptr(), target.data_ptr(), n, grad_scale / n, grad.data_ptr() if want_grad else None,
                                 ws.data_ptr(), ws.numel() * 8, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "demfi_l1_sum")
    return grad


def rec_losses(sharps_prime: Sequence[torch.Tensor], sharps_final: Sequence[Sequence[torch.Tensor]], S0_GT, S1_GT, St_GT,
               rec_D1_lambda: float = 1.0, rec_D2_lambda: float = 1.0, with_grads: bool = False):
    """(total_loss, rec_D1_loss, rec_D2_loss) as Python floats -- `main.py:404-440` -- and, with_grads, also
    (grads_prime[3], grads_final[N][3]) = d total / d each predicted frame."""
    gts = [S0_GT, S1_GT, St_GT]
    dev = sharps_prime[0].device
    n_terms = 3 * (1 + len(sharps_final))
    sums = torch.zeros(n_terms, dtype=torch.float64, device=dev)
    g_prime: List[torch.Tensor] = []
    g_final: List[List[torch.Tensor]] = []
    k = 0
    for idx in range(3):
        g_prime.append(_l1_term(sharps_prime[idx], gts[idx], rec_D1_lambda / 3.0, with_grads, sums[k:k + 1]))
        k += 1
    for tri in sharps_final:
        g_final.append([])
        for idx in range(3):
            g_final[-1].append(_l1_term(tri[idx], gts[idx], rec_D2_lambda / 3.0, with_grads, sums[k:k + 1]))
            k += 1
    s = sums.cpu().tolist()                       # one read-back for all terms
    numel = [t.numel() for t in gts]
    means = [s[i] / numel[i % 3] for i in range(n_terms)]
    rec_D1 = rec_D1_lambda * (means[0] + means[1] + means[2]) / 3.0
    rec_D2 = sum(rec_D2_lambda * (means[3 + 3 * i] + means[4 + 3 * i] + means[5 + 3 * i]) / 3.0 for i in range(len(sharps_final)))
    if with_grads:
        return rec_D1 + rec_D2, rec_D1, rec_D2, g_prime, g_final
    return rec_D1 + rec_D2, rec_D1, rec_D2


class Adam:
    """torch.optim.Adam for fp32 CUDA parameters through `demfi_adam_step` (no amsgrad, no foreach grouping)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params]
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.state: Dict[int, dict] = {}
        self.param_groups = [{"params": self.params, "lr": lr}]    # `optimizer.param_groups[0]['lr']`, main.py:385

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self):
        lib = A.lib()
        lr = self.param_groups[0]["lr"]
        for i, p in enumerate(self.params):
            if p.grad is None:
                continue
            if not p.is_cuda:
                raise A.DemfiError("demfi_b200.train.Adam updates CUDA parameters only (no CPU fallback)")
            st = self.state.setdefault(i, {"step": 0, "exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)})
            st["step"] += 1
            g = p.grad.contiguous()
            assert p.is_contiguous() and p.dtype == torch.float32 and g.dtype == torch.float32
            with torch.cuda.device(p.device):
                A.check(lib.demfi_adam_step(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(),
                                            lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, st["step"],
                                            torch.cuda.current_stream().cuda_stream), "demfi_adam_step")
            # the kernel wrote through the raw pointer: tell torch (autograd's saved-tensor checks, and DeMFInet's own
            # "weights changed -> repack" test, which reads p._version)
            torch.autograd.graph.increment_version(p)


def _adam_state_dict(self):
    """`optimizer.state_dict()` in torch.optim.Adam's layout (main.py:270 saves it, :214 loads it): a checkpoint written by the
    reference loads here and the other way round.  `step` is a float32 scalar tensor as in torch >= 1.12 (a plain int from the
    pinned torch 1.7.1 is accepted on load)."""
    state = {i: {"step": torch.tensor(float(st["step"])), "exp_avg": st["exp_avg"], "exp_avg_sq": st["exp_avg_sq"]}
             for i, st in self.state.items()}
    group = {"lr": self.param_groups[0]["lr"], "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay,
             "amsgrad": False, "params": list(range(len(self.params)))}
    return {"state": state, "param_groups": [group]}


def _adam_load_state_dict(self, sd):
    g = sd["param_groups"][0]
    if g.get("amsgrad", False):
        raise NotImplementedError("amsgrad state is not supported (main.py:179 does not use it)")
    if len(g["params"]) != len(self.params):
        raise ValueError("loaded state dict has a different number of parameters")
    self.param_groups[0]["lr"] = g["lr"]
    self.betas, self.eps, self.weight_decay = tuple(g["betas"]), g["eps"], g["weight_decay"]
    self.state = {}
    for k, st in sd["state"].items():
        p = self.params[int(k)]
        if tuple(st["exp_avg"].shape) != tuple(p.shape):
            raise ValueError(f"optimizer state {k} has shape {tuple(st['exp_avg'].shape)}, the parameter {tuple(p.shape)}")
        self.state[int(k)] = {"step": int(float(st["step"])),
                              "exp_avg": st["exp_avg"].detach().to(p.device, torch.float32).clone(),
                              "exp_avg_sq": st["exp_avg_sq"].detach().to(p.device, torch.float32).clone()}


Adam.state_dict = _adam_state_dict
Adam.load_state_dict = _adam_load_state_dict


def allreduce_gradients(params, world_size: int = None, group=None) -> int:
    """Average the gradients over the ranks of the job: one flat fp32 bucket, one `all_reduce(SUM)`, scaled by 1/world and
    copied back.  Returns the number of bytes exchanged per rank (the bucket size).  A no-op without torch.distributed."""
    import torch.distributed as dist
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = world_size or dist.get_world_size(group)
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.mul_(1.0 / world)
    o = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[o:o + n].view_as(g))
        o += n
    return flat.numel() * 4


class MultiStepLR:
    """`optim.lr_scheduler.MultiStepLR(optimizer, milestones, gamma)` of main.py:186 / :511 for `train.Adam` (whose
    `param_groups[0]['lr']` it scales): after the e-th `step()` the rate is lr0 * gamma ** (number of milestones <= e)."""

    def __init__(self, optimizer, milestones, gamma=0.1):
        self.optimizer, self.milestones, self.gamma = optimizer, sorted(milestones), gamma
        self.base_lr = optimizer.param_groups[0]["lr"]
        self.last_epoch = 0

    def step(self):
        self.last_epoch += 1
        k = sum(1 for m in self.milestones if m <= self.last_epoch)
        self.optimizer.param_groups[0]["lr"] = self.base_lr * self.gamma ** k

    def get_last_lr(self):
        return [self.optimizer.param_groups[0]["lr"]]

    def state_dict(self):
        """the keys main.py:215-217 touches: last_epoch, milestones (a Counter in torch), gamma, base_lrs"""
        from collections import Counter
        return {"last_epoch": self.last_epoch, "milestones": Counter(self.milestones), "gamma": self.gamma, "base_lrs": [self.base_lr]}

    def load_state_dict(self, sd):
        self.last_epoch, self.gamma = sd["last_epoch"], sd["gamma"]
        self.milestones = sorted(sd["milestones"].elements()) if hasattr(sd["milestones"], "elements") else sorted(sd["milestones"])
        self.base_lr = sd.get("base_lrs", [self.base_lr])[0]


def split_training_batch(frames: torch.Tensor):
    """main.py:387-390: a training sample is [B, C, 9, H, W] = the four blurry inputs (B0, B1, B-1, B2), the sharp frame at t, and
    the sharp frames at 0, 1, -1, 2 -> (input_frames [B,C,4,H,W], S0_GT, S1_GT, frameT)."""
    input_frames = frames[:, :, :4]
    frameT = frames[:, :, 4]
    input_frames_GT = frames[:, :, -4:]
    return input_frames, input_frames_GT[:, :, 0], input_frames_GT[:, :, 1], frameT


def train_step(model, frames: torch.Tensor, t_value: torch.Tensor, optimizer, N_trn: int, rec_D1_lambda: float = 1.0,
               rec_D2_lambda: float = 1.0, ops=None, loss_fn=None):
    """One iteration of the loop body of `train()` (main.py:386-448): zero_grad, differentiable forward, Eq.(9)-(10) losses,
    backward, gradient all-reduce over the ranks (a no-op in a single process), optimizer step.  Returns (total, rec_D1, rec_D2)
    as floats, like the `.item()` calls of main.py:451-453.

    ops / loss_fn: the operator set of `train_net.forward_train` and the loss implementation; the defaults are this repository's
    kernels (`train_net.KernelOps`, `rec_losses`).  Tests pass torch stand-ins to check this host logic without a GPU."""
    from . import train_net
    input_frames, S0_GT, S1_GT, frameT = split_training_batch(frames)
    optimizer.zero_grad()
    res = train_net.forward_train(model, input_frames, t_value, N_trn, ops=ops or train_net.KernelOps)
    outs = list(res[0]) + [s for tri in res[1] for s in tri]
    if loss_fn is None:
        total, d1, d2, g_prime, g_final = rec_losses(res[0], res[1], S0_GT, S1_GT, frameT, rec_D1_lambda, rec_D2_lambda, with_grads=True)
        torch.autograd.backward(outs, list(g_prime) + [g for tri in g_final for g in tri])
    else:
        total_t, d1_t, d2_t = loss_fn(res[0], res[1], S0_GT, S1_GT, frameT, rec_D1_lambda, rec_D2_lambda)
        total_t.backward()
        total, d1, d2 = float(total_t.detach()), float(d1_t.detach()), float(d2_t.detach())
    allreduce_gradients([p for p in model.parameters() if p.grad is not None])
    optimizer.step()
    return total, d1, d2
