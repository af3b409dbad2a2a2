"""world_size-2 gloo test of the multi-GPU host logic: env sharding by rank and the learner's
bucketed gradient all-reduce issued from the backward hooks (the only collective of the path)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import marlsc_b200  # noqa: F401


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from marlsc_b200.rollout import ActorCritic, PPOLearner, shard_envs
    torch.manual_seed(0)                      # identical initial weights on every rank
    pol = ActorCritic(6, 3, 2, actor_hidden=(16,), critic_hidden=(16,))
    # two buckets (bucket_bytes below the model size) so that the per-bucket path is exercised
    learner = PPOLearner(pol, lr=1e-2, bucket_bytes=300)
    assert len(learner.buckets.buckets) >= 2
    shard = shard_envs(10, rank, world)
    g = torch.Generator().manual_seed(100 + rank)            # different data per rank
    obs = torch.randn(len(shard), 3, 6, generator=g)
    act = torch.randn(len(shard), 3, 2, generator=g)
    batch = (obs, act, torch.zeros(len(shard), 3), torch.randn(len(shard), 3, generator=g), torch.randn(len(shard), 3, generator=g))
    # local gradient without any collective (a twin module with the same weights, no hooks)
    twin = ActorCritic(6, 3, 2, actor_hidden=(16,), critic_hidden=(16,))
    twin.load_state_dict(pol.state_dict())
    ref = PPOLearner.__new__(PPOLearner)
    ref.__dict__.update(policy=twin, fused=False, clip=learner.clip, vf_clip=learner.vf_clip, vf_coeff=learner.vf_coeff,
                        ent_coeff=learner.ent_coeff, beta=None, use_kl=False, kl_coeff=0.0)
    ref.loss(*batch)["total"].backward()
    local = torch.cat([p.grad.reshape(-1) for p in twin.parameters()])
    # the learner's own path: zero the buckets, backward (hooks launch one all-reduce per bucket), finish
    learner.buckets.zero()
    loss = learner.loss(*batch)["total"]
    loss.backward()
    learner.buckets.finish()
    nbytes = learner.all_reduce_grads()
    reduced = torch.cat([p.grad.reshape(-1) for p in learner.params])
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(reduced, torch.stack(gathered).mean(0), atol=1e-6)
    assert nbytes == local.numel() * 4
    learner.opt.step()
    w = torch.cat([p.detach().reshape(-1) for p in learner.params])
    ws = [torch.zeros_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    assert torch.equal(ws[0], ws[1]), "ranks diverged after the synchronised step"
    if rank == 0:
        out.put((list(shard), float(loss)))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_and_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    shard, _ = q.get(timeout=5)
    assert shard == [0, 1, 2, 3, 4]
