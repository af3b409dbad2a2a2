"""Minimal PPO learner step for the collected rollouts (clipped surrogate, clipped value loss,
entropy bonus), data-parallel across GPUs: each rank learns on its own env shard and the flat
gradient is summed with one NCCL all-reduce per minibatch - the only collective of the whole path
(reference hyper-parameters: src/algorithms/ippo.py:145-160, hysteretic
weighting: src/algorithms/learners/hysteretic_learner.py:39-42)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from .. import _capi
from .collector import Rollout
from .policy import ActorCritic


class _FusedPPOObjective(torch.autograd.Function):
    """Clipped surrogate + clipped value loss through K6 (``marlsc_ppo_loss``): one kernel computes the two loss terms and
    the gradients with respect to the action means, ``log_std`` and the values; backward only scales them."""

    @staticmethod
    def forward(ctx, mean, log_std, value, actions, logp_old, adv, targets, floor, clip, vf_clip, vf_coeff, beta):
        S = mean.shape[-1]
        mean_c, value_c = mean.contiguous(), value.contiguous()
        n = value_c.numel()
        if mean_c.numel() != n * S or mean_c.dtype != torch.float32 or not mean_c.is_cuda:
            raise ValueError("mean must be a float32 CUDA tensor [..., S] matching value [...]")
        args = [t.detach().to(torch.float32).contiguous() for t in (actions, logp_old, adv, targets)]
        if args[0].numel() != n * S or any(a.numel() != n for a in args[1:]):
            raise ValueError("actions / logp_old / adv / targets do not match the minibatch shape")
        ls = log_std.detach().contiguous()
        g_mean = torch.empty_like(mean_c)
        g_value = torch.empty_like(value_c)
        sums = torch.empty(2 + S, dtype=torch.float64, device=mean.device)
        _capi.check(_capi.lib().marlsc_ppo_loss(
            mean_c.data_ptr(), args[0].data_ptr(), ls.data_ptr(), float(floor), args[1].data_ptr(), args[2].data_ptr(),
            value_c.data_ptr(), args[3].data_ptr(), n, S, float(clip), float(vf_clip), float(vf_coeff),
            -1.0 if beta is None else float(beta), g_mean.data_ptr(), g_value.data_ptr(), sums.data_ptr(),
            torch.cuda.current_stream(mean.device).cuda_stream))
        g_ls = (sums[2:] * (ls >= floor)).to(torch.float32)          # clamp(min=floor) passes the gradient where it does not bind
        ctx.save_for_backward(g_mean, g_ls, g_value)
        ctx.shapes = (mean.shape, value.shape)
        policy = (-sums[0] / n).to(torch.float32)
        vf = (sums[1] / n).to(torch.float32)
        return policy + vf_coeff * vf, policy, vf

    @staticmethod
    def backward(ctx, g, _gp, _gv):
        g_mean, g_ls, g_value = ctx.saved_tensors
        return (g_mean.reshape(ctx.shapes[0]) * g, g_ls * g, g_value.reshape(ctx.shapes[1]) * g) + (None,) * 9


class PPOLearner:
    def __init__(self, policy: ActorCritic, lr: float = 5e-4, clip_param: float = 0.2, vf_clip_param: float = 10.0,
                 vf_loss_coeff: float = 1.0, entropy_coeff: float = 0.01, grad_clip: Optional[float] = None,
                 hysteretic_beta: Optional[float] = None, fused_loss: Optional[bool] = None):
        self.policy = policy
        # K6 on CUDA parameters unless asked otherwise; loss_reference() keeps the plain PyTorch form
        self.fused = next(policy.parameters()).is_cuda if fused_loss is None else bool(fused_loss)
        self.opt = torch.optim.Adam(policy.parameters(), lr=lr)
        self.clip, self.vf_clip, self.vf_coeff, self.ent_coeff = clip_param, vf_clip_param, vf_loss_coeff, entropy_coeff
        self.grad_clip, self.beta = grad_clip, hysteretic_beta
        self.params = [p for p in policy.parameters() if p.requires_grad]

    @classmethod
    def from_algorithm_config(cls, policy: ActorCritic, algo_config) -> "PPOLearner":
        sp, sh = algo_config.algorithm_specific, algo_config.shared
        lr = sh.learning_rate if isinstance(sh.learning_rate, (int, float)) else sh.learning_rate[0][1]
        return cls(policy, lr=lr, clip_param=sp.clip_param, vf_clip_param=sp.vf_clip_param, vf_loss_coeff=sp.vf_loss_coeff,
                   entropy_coeff=sp.entropy_coeff, grad_clip=sp.grad_clip, hysteretic_beta=getattr(sp, "hysteretic_beta", None))

    def loss(self, obs, actions, logp_old, adv, targets) -> Dict[str, torch.Tensor]:
        if not self.fused:
            return self.loss_reference(obs, actions, logp_old, adv, targets)
        pol = self.policy
        partial, policy, vf = _FusedPPOObjective.apply(pol.action_mean(obs), pol.log_std, pol.value(obs), actions, logp_old, adv,
                                                        targets, pol.logstd_floor, self.clip, self.vf_clip, self.vf_coeff, self.beta)
        ent = pol.entropy()
        return dict(total=partial - self.ent_coeff * ent, policy=policy.detach(), vf=vf.detach(), entropy=ent)

    def loss_reference(self, obs, actions, logp_old, adv, targets) -> Dict[str, torch.Tensor]:
        """The same objective in plain PyTorch (autograd): the numerical reference of K6."""
        pol = self.policy
        mean = pol.action_mean(obs)
        logp = pol.log_prob(mean, actions)
        ratio = (logp - logp_old).exp()
        if self.beta is not None:
            adv = torch.where(adv < 0, adv * self.beta, adv)
        surr = torch.minimum(ratio * adv, ratio.clamp(1 - self.clip, 1 + self.clip) * adv)
        v = pol.value(obs)
        vf = torch.clamp((v - targets) ** 2, max=self.vf_clip)
        ent = pol.entropy()
        total = -surr.mean() + self.vf_coeff * vf.mean() - self.ent_coeff * ent
        return dict(total=total, policy=-surr.mean(), vf=vf.mean(), entropy=ent)

    def all_reduce_grads(self) -> int:
        """Sum gradients over ranks with a single collective on a flat buffer; returns the bytes reduced."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return 0
        flat = torch.cat([p.grad.reshape(-1) for p in self.params])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat /= dist.get_world_size()
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad.copy_(flat[off:off + n].view_as(p))
            off += n
        return flat.numel() * flat.element_size()

    def minibatch_step(self, ro: Rollout, t_slice: slice, e_slice: slice) -> Dict[str, float]:
        obs = ro.obs[t_slice, e_slice].flatten(0, 1)           # [B, W, D]
        out = self.loss(obs, ro.actions[t_slice, e_slice].flatten(0, 1), ro.logp[t_slice, e_slice].flatten(0, 1),
                        ro.advantages[t_slice, e_slice].flatten(0, 1), ro.targets[t_slice, e_slice].flatten(0, 1))
        self.opt.zero_grad(set_to_none=False)
        out["total"].backward()
        self.all_reduce_grads()
        if self.grad_clip:
            torch.nn.utils.clip_grad_norm_(self.params, self.grad_clip)
        self.opt.step()
        return {k: float(v.detach()) for k, v in out.items()}
