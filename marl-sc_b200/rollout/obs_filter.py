"""Running mean / std observation filter for ``obs_normalization: "meanstd"``.

In the reference this mode leaves the env's observations untouched and adds RLlib's ``MeanStdFilter`` connector
in front of the module (reference: src/algorithms/ippo.py:173-176); the env itself emits raw features
(multi_env.py:700-702 only handles the fixed-statistics modes). ``MeanStdFilter`` keeps running per-feature moments
over everything sampled so far and returns ``(x - mean) / (std + 1e-8)``. This is the batched device counterpart: one
parallel-Welford merge per call (Chan et al.), statistics per observation column shared by all agents that share a
policy.
"""
from __future__ import annotations

import torch


class MeanStdFilter:
    def __init__(self, obs_dim: int, device, eps: float = 1e-8, update: bool = True):
        self.count = torch.zeros((), dtype=torch.float64, device=device)
        self.mean = torch.zeros(obs_dim, dtype=torch.float64, device=device)
        self.m2 = torch.zeros(obs_dim, dtype=torch.float64, device=device)
        self.eps, self.update = eps, update

    @torch.no_grad()
    def __call__(self, obs: torch.Tensor) -> torch.Tensor:
        """Update the running moments with ``obs [..., obs_dim]`` (when ``update``) and normalise it in place."""
        flat = obs.reshape(-1, obs.shape[-1])
        if self.update:
            n = float(flat.shape[0])
            bm = flat.mean(0, dtype=torch.float64)
            bm2 = ((flat.to(torch.float64) - bm) ** 2).sum(0)
            tot = self.count + n
            delta = bm - self.mean
            self.mean += delta * (n / tot)
            self.m2 += bm2 + delta * delta * (self.count * n / tot)
            self.count = tot
        std = torch.sqrt(self.m2 / torch.clamp(self.count - 1.0, min=1.0))
        flat.sub_(self.mean.to(obs.dtype)).div_((std + self.eps).to(obs.dtype))
        return obs

    def state_dict(self):
        return dict(count=self.count.clone(), mean=self.mean.clone(), m2=self.m2.clone())

    def load_state_dict(self, sd):
        self.count, self.mean, self.m2 = sd["count"].clone(), sd["mean"].clone(), sd["m2"].clone()
