"""Rollout collection on device + the only collective of the design.

RolloutCollector replaces the env side of RLlib's sampling loop
(SURVEY.md 3.1 step 4): T steps of N environments into [T, N, ...] device
tensors, obs read straight from the library's observation ring (obs_ring = T+1
makes the ring itself the rollout buffer -- no copy).  Environments are
independent, so multi-GPU = one process per GPU, each with its own shard;
``allreduce_episode_stats`` is the optional per-rollout NCCL reduction of the
episode statistics (<= 8 float64) -- nothing on the step path communicates.
"""
import torch


class RolloutCollector:
    def __init__(self, env, policy=None, horizon=32, seed=0):
        """policy: callable(obs dict of [N, ...] tensors) -> int32 actions [N, k] or
        flat [N]; None = the device random policy (agent/random/*.py)."""
        self.env, self.policy, self.T, self.seed = env, policy, int(horizon), int(seed)
        N, dev = env.num_envs, env.device
        self.actions = torch.zeros((self.T, N, env.spec.action_width), dtype=torch.int32, device=dev)
        self.rewards = torch.zeros((self.T, N), dtype=torch.float64, device=dev)
        self.dones = torch.zeros((self.T, N), dtype=torch.uint8, device=dev)
        ring = next(iter(env.obs_ring_tensors.values())).shape[0]
        self.zero_copy = ring >= self.T + 1
        self.obs = None if self.zero_copy else {
            k: torch.zeros((self.T + 1,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev)
            for k, v in env.obs_ring_tensors.items()}
        self._step = 0

    def collect(self, reset=True):
        env = self.env
        obs = env.reset() if reset or self._step == 0 else env._obs()
        if self.policy is None and self.zero_copy:
            # the device random policy needs no host in the loop: one C call for all T steps
            # (one kernel launch for the pin-less kinds), results logged per step
            ring = next(iter(env.obs_ring_tensors.values())).shape[0]
            s0 = env.lib.pcb_current_slice(env._h)
            env.rollout_random(self.T, self.seed, actions_log=self.actions, reward_log=self.rewards,
                               done_log=self.dones)
            self._step += self.T
            idx = (s0 + torch.arange(self.T + 1, device=env.device)) % ring
            return {"obs": {k: v[idx] for k, v in env.obs_ring_tensors.items()},
                    "actions": self.actions, "rewards": self.rewards, "dones": self.dones}
        slices = []
        for t in range(self.T):
            s = env.lib.pcb_current_slice(env._h)
            slices.append(s)
            if not self.zero_copy:
                for k, v in obs.items():
                    self.obs[k][t].copy_(v)
            if self.policy is None:
                a = env.sample_random_actions(self.seed, out=self.actions[t])
            else:
                a = self.policy(obs)
                if a.dim() == 1:
                    a = env.decode_flat_actions(a)
                self.actions[t].copy_(a)
                a = self.actions[t]
            obs, r, d, _ = env.step(a)
            self.rewards[t].copy_(r)
            self.dones[t].copy_(d)
            self._step += 1
        slices.append(env.lib.pcb_current_slice(env._h))
        if self.zero_copy:
            idx = torch.tensor(slices, device=env.device)
            obs_out = {k: v[idx] for k, v in env.obs_ring_tensors.items()}
        else:
            for k, v in obs.items():
                self.obs[k][self.T].copy_(v)
            obs_out = self.obs
        return {"obs": obs_out, "actions": self.actions, "rewards": self.rewards,
                "dones": self.dones}


def shard_range(total, rank, world):
    """Environment ids [lo, hi) owned by `rank`: contiguous, sizes differ by at most one
    (env ids [g*N/G, (g+1)*N/G) on GPU g)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_stats(stats, device=None):
    """Sum a dict of float statistics over all ranks (NCCL over NVLink with the nccl
    backend -- then `device` must be the rank's GPU; gloo on CPU in the tests)."""
    import torch.distributed as dist

    keys = sorted(stats)
    backend = dist.get_backend() if dist.is_initialized() else None
    dev = device if backend == "nccl" else torch.device("cpu")
    t = torch.tensor([float(stats[k]) for k in keys], dtype=torch.float64, device=dev)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: float(v) for k, v in zip(keys, t.tolist())}


def allreduce_episode_stats(env, reset=False):
    """C1: the per-rollout reduction of the device-side episode statistics."""
    return allreduce_stats(env.episode_stats(reset=reset), env.device)
