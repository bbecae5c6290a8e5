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


class RewardDoneExchange:
    """The one collective of the path: per step, every rank contributes {reward, done} of its shard (8 B per
    environment) and receives everyone's.  No staging: `reward` and `done` are the two rows of the registered send buffer,
    handed to the step kernel as its output pointers, so the kernel's own stores fill the NCCL send buffer; `gather()` is
    one all_gather_into_tensor on preallocated tensors - nothing is allocated or stacked per step, so the call can be
    captured into the same CUDA graph as the step kernels.  With `depth` = 2 the send buffers alternate, which lets the
    gather of step k run on a side stream under step k + 1 (reward / done feed the bookkeeping of a centralized policy,
    not the next action, so they may arrive one step late)."""

    def __init__(self, n_stride, world, device, depth=1):
        self.world, self.ns, self.depth = world, n_stride, depth
        self.send = torch.zeros((depth, 2, n_stride), dtype=torch.float32, device=device)
        self.recv = torch.zeros((depth, world, 2, n_stride), dtype=torch.float32, device=device)
        self.comm = torch.cuda.Stream(device=device) if (depth > 1 and torch.device(device).type == "cuda") else None
        self._pending = [None] * depth

    def reward(self, k=0):
        return self.send[k % self.depth, 0]

    def done(self, k=0):
        return self.send[k % self.depth, 1]

    def gather(self, k=0):
        """All-gather of slot k (in stream order behind the step that filled it)."""
        if self.world == 1:
            self.recv[k % self.depth, 0].copy_(self.send[k % self.depth])
            return
        j = k % self.depth
        if self.comm is None:
            dist.all_gather_into_tensor(self.recv[j].view(-1), self.send[j].view(-1))
            return
        cur = torch.cuda.current_stream()
        self.comm.wait_stream(cur)                       # behind the step that wrote slot j
        with torch.cuda.stream(self.comm):
            dist.all_gather_into_tensor(self.recv[j].view(-1), self.send[j].view(-1))
            ev = torch.cuda.Event()
            ev.record(self.comm)
        self._pending[j] = ev

    def before_step(self, k):
        """Call before the step that overwrites slot k: its previous gather must have read the buffer."""
        ev = self._pending[k % self.depth] if self.comm is not None else None
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
            self._pending[k % self.depth] = None

    def join(self):
        if self.comm is not None:
            torch.cuda.current_stream().wait_stream(self.comm)
            self._pending = [None] * self.depth

    def reset(self):
        """Forget events of gathers already joined (call after join() + a synchronisation, e.g. before stream capture:
        an event recorded outside a capture cannot be waited on inside it)."""
        self._pending = [None] * self.depth

    def full(self, sizes, k=0):
        """(reward, done) of all environments, rank-major, trimmed to the shard sizes."""
        r = self.recv[k % self.depth]
        return (torch.cat([r[i, 0, :sizes[i]] for i in range(self.world)]),
                torch.cat([r[i, 1, :sizes[i]] for i in range(self.world)]))
