"""Multi-GPU plumbing: environments are independent, so they shard contiguously across ranks
(one process per GPU) with NO data-path collective.  The only optional exchange is the per-step
{reward, done} vector (8 B/env) when ONE policy process consumes all environments
(the role of VectorizedEnvironment::step's outputs, examples/ars/ars_vectorized_environment.h:214-291)."""
import torch
import torch.distributed as dist


def shard_range(total_envs, rank, world):
    """Contiguous [lo, hi) block of environments owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(total_envs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_reward_done(reward, done, total_envs, world):
    """All-gather the local reward/done shards (1-D tensors) into full [total_envs] vectors.
    Works with NCCL (GPU tensors) and gloo (CPU tensors); shards may differ in length by one."""
    if world == 1:
        return reward, done
    base, rem = divmod(total_envs, world)
    cap = base + (1 if rem else 0)
    buf = torch.zeros((2, cap), dtype=torch.float32, device=reward.device)
    buf[0, :reward.numel()] = reward
    buf[1, :done.numel()] = done
    out = torch.empty((world, 2, cap), dtype=torch.float32, device=reward.device)
    dist.all_gather_into_tensor(out.view(-1), buf.view(-1))
    sizes = [base + (1 if r < rem else 0) for r in range(world)]
    r = torch.cat([out[k, 0, :sizes[k]] for k in range(world)])
    d = torch.cat([out[k, 1, :sizes[k]] for k in range(world)])
    return r, d
