"""CUDA-graph replay of an episode segment for launch-bound shapes.

At the reference's default shape (3 warehouses x 2 SKUs) a step of 4,096 environments moves 1.8 MB: the kernels take
a few microseconds and the Python / ctypes / launch path around them dominates. With pre-sampled
actions and demand - BASELINE config 2, ``run_baselines.py``-style evaluation of fixed policies, parity replays - the
whole ``reset + T steps`` sequence is captured once into a CUDA graph and replayed with one launch.

The timestep of every step and the buffers are baked into the graph: a replay recomputes the same episode from whatever
the caller has written into ``actions`` (and the demand tensors) in the meantime.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from ..envs import BatchedInventoryEnv


class GraphedEpisode:
    def __init__(self, env: BatchedInventoryEnv, actions: torch.Tensor, orders: Sequence, init_inventory: Optional[torch.Tensor] = None):
        """``actions`` [T,E,W,S] float32 on the device (updated in place between replays), ``orders`` a length-T sequence of
        ``DeviceOrders`` / ``DeviceLines`` (their tensors are read at replay time). Captures ``reset`` + T steps."""
        T = actions.shape[0]
        if len(orders) != T:
            raise ValueError("one demand entry per step")
        if T > env.episode_length:
            raise ValueError("segment longer than the episode")
        self.env, self.T = env, T
        self.actions, self.orders = actions, list(orders)
        E, W = env.num_envs, env.n_warehouses
        dev = env.device
        self.rewards = torch.empty((T, E, W), device=dev)
        self.obs = torch.empty((T + 1, E, W, env.obs_dim), device=dev)
        # the start inventory is part of the graph's inputs: drawn once here when the config asks for a random one (write
        # new values into ``self.init`` between replays to vary it)
        self.init = (env._initial_inventory()[0] if init_inventory is None
                     else init_inventory.to(device=dev, dtype=torch.int32).contiguous())
        self._run()                                  # eager pass: sizes the library's workspaces, sets kernel attributes
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            with torch.cuda.graph(self.graph, stream=side):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)

    def _run(self) -> None:
        env = self.env
        env.reset(init_inventory=self.init, obs_out=self.obs[0])
        for t in range(self.T):
            env.step(self.actions[t], orders=self.orders[t], obs_out=self.obs[t + 1], rewards_out=self.rewards[t])

    def replay(self) -> torch.Tensor:
        """One graph launch = reset + T steps; returns the rewards buffer [T,E,W] (valid once the stream reaches it)."""
        self.graph.replay()
        self.env.timestep = self.T
        return self.rewards
