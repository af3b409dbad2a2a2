"""Minimal PPO learner step for the collected rollouts (clipped surrogate, clipped value loss,
entropy bonus), data-parallel across GPUs: each rank learns on its own env shard and the flat
gradient is summed with one NCCL all-reduce per minibatch - the only collective of the whole path
(reference hyper-parameters: src/algorithms/ippo.py:145-160, hysteretic
weighting: src/algorithms/learners/hysteretic_learner.py:39-42)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from .collector import Rollout
from .policy import ActorCritic


class PPOLearner:
    def __init__(self, policy: ActorCritic, lr: float = 5e-4, clip_param: float = 0.2, vf_clip_param: float = 10.0,
                 vf_loss_coeff: float = 1.0, entropy_coeff: float = 0.01, grad_clip: Optional[float] = None,
                 hysteretic_beta: Optional[float] = None):
        self.policy = policy
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
