"""A minimal torch PPO loop on top of PcbVecEnv.

Not the reference's policy networks (those are TF2/RLlib and out of scope): just
enough policy to drive the rollout collector end to end -- the observation keys
are consumed straight from the library's DLPack tensors, the action mask is
applied to the logits exactly as the reference's models do
(``logits += max(log(action_mask), float32.min)``, agent/models/square_model.py:119-141),
GAE and the clipped PPO objective run on device.  The env side of every step is
``pcb_step`` (CUDA kernels); nothing here touches the dynamics.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class MaskedPlacementPolicy(nn.Module):
    """Conv encoder over the board planes -> logits over the flat action space
    (O*H*W, the FlatteningActionWrapper layout) + value.  Planes: ``grid`` and, for the
    spatial env, the ``pin_grid`` channels; the next component's (h, w, area ratio) is
    broadcast as extra planes (RectanglePinModel feeds the same features densely)."""

    def __init__(self, env, hidden=32):
        super().__init__()
        s = env.spec
        self.kind, self.H, self.W, self.O = s.kind, s.height, s.width, s.n_orient
        planes = 1 + (s.max_num_nets + 1 if s.kind == "rectangle_spatial_pin" else 0)
        planes += 0 if s.kind == "square" else 3
        self.body = nn.Sequential(
            nn.Conv2d(planes, hidden, 3, padding=1), nn.ReLU(),
            nn.Conv2d(hidden, hidden, 3, padding=1), nn.ReLU())
        self.logit_head = nn.Conv2d(hidden, self.O, 1)
        self.value_head = nn.Linear(hidden, 1)

    def planes(self, obs):
        x = [obs["grid"].float().unsqueeze(1)]
        if self.kind == "rectangle_spatial_pin":
            x.append(obs["pin_grid"].float().permute(0, 3, 1, 2))
        if self.kind != "square":
            pm = obs["placement_mask"]
            if self.kind == "rectangle":  # next component = first one not placed yet
                cur = (pm.long().sum(1)).clamp(max=pm.shape[1] - 1)
            else:
                cur = (pm == 3).float().argmax(1)
            feat = obs["all_components_feature"][torch.arange(pm.shape[0], device=pm.device), cur]
            f = torch.stack((feat[:, 0] / self.H, feat[:, 1] / self.W, feat[:, 4]), 1)
            x.append(f[:, :, None, None].expand(-1, -1, self.H, self.W))
        return torch.cat(x, 1)

    def forward(self, obs):
        h = self.body(self.planes(obs))
        logits = self.logit_head(h).reshape(h.shape[0], -1)
        mask = obs["action_mask"].reshape(h.shape[0], -1).float()
        logits = logits + torch.clamp(torch.log(mask), min=torch.finfo(torch.float32).min)
        value = self.value_head(h.mean((2, 3))).squeeze(1)
        return logits, value

    @torch.no_grad()
    def act(self, obs):
        logits, value = self(obs)
        dist = torch.distributions.Categorical(logits=logits)
        a = dist.sample()
        return a, dist.log_prob(a), value


def compute_gae(rewards, values, dones, last_value, gamma=0.99, lam=0.95):
    """rewards/values/dones: [T, N] (done[t] = episode ended with step t; with auto-reset
    the value of the next state then belongs to a new episode and is cut)."""
    T = rewards.shape[0]
    adv = torch.zeros_like(rewards)
    nxt, run = last_value, torch.zeros_like(last_value)
    for t in range(T - 1, -1, -1):
        live = 1.0 - dones[t].to(rewards.dtype)
        delta = rewards[t] + gamma * nxt * live - values[t]
        run = delta + gamma * lam * live * run
        adv[t] = run
        nxt = values[t]
    return adv, adv + values


class PPOTrainer:
    """collect(T steps of N envs) -> GAE -> a few clipped-PPO minibatch epochs."""

    def __init__(self, env, policy=None, horizon=16, lr=3e-4, clip=0.2, epochs=2, minibatches=4):
        self.env, self.T = env, int(horizon)
        self.policy = (policy or MaskedPlacementPolicy(env)).to(env.device)
        self.opt = torch.optim.Adam(self.policy.parameters(), lr=lr)
        self.clip, self.epochs, self.minibatches = clip, epochs, minibatches
        self.obs = env.reset()

    def collect(self):
        env, T, N = self.env, self.T, self.env.num_envs
        keys = list(self.obs.keys())
        buf = {k: torch.empty((T,) + tuple(self.obs[k].shape), dtype=self.obs[k].dtype,
                              device=env.device) for k in keys}
        acts = torch.empty((T, N), dtype=torch.long, device=env.device)
        logp = torch.empty((T, N), device=env.device)
        vals = torch.empty((T, N), device=env.device)
        rews = torch.empty((T, N), device=env.device)
        dones = torch.empty((T, N), dtype=torch.bool, device=env.device)
        obs = self.obs
        for t in range(T):
            for k in keys:
                buf[k][t].copy_(obs[k])
            a, lp, v = self.policy.act(obs)
            acts[t], logp[t], vals[t] = a, lp, v
            obs, r, d, _ = env.step(a.int())  # flat actions, decoded in the step kernel
            rews[t], dones[t] = r.float(), d.bool()
        self.obs = obs
        with torch.no_grad():
            _, last_v = self.policy(obs)
        adv, ret = compute_gae(rews, vals, dones, last_v)
        return dict(obs=buf, actions=acts, logp=logp, values=vals, rewards=rews, dones=dones,
                    adv=adv, ret=ret)

    def update(self, batch):
        T, N = batch["actions"].shape
        flat = lambda x: x.reshape((T * N,) + tuple(x.shape[2:]))  # noqa: E731
        obs = {k: flat(v) for k, v in batch["obs"].items()}
        acts, old_lp = flat(batch["actions"]), flat(batch["logp"])
        adv, ret = flat(batch["adv"]), flat(batch["ret"])
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        losses = []
        for _ in range(self.epochs):
            perm = torch.randperm(T * N, device=acts.device)
            for idx in perm.chunk(self.minibatches):
                logits, value = self.policy({k: v[idx] for k, v in obs.items()})
                dist = torch.distributions.Categorical(logits=logits)
                ratio = torch.exp(dist.log_prob(acts[idx]) - old_lp[idx])
                pg = -torch.min(ratio * adv[idx],
                                torch.clamp(ratio, 1 - self.clip, 1 + self.clip) * adv[idx]).mean()
                loss = pg + 0.5 * F.mse_loss(value, ret[idx]) - 0.01 * dist.entropy().mean()
                self.opt.zero_grad(set_to_none=True)
                loss.backward()
                self.opt.step()
                losses.append(float(loss.detach()))
        return sum(losses) / len(losses)

    def train(self, iterations=1):
        out = []
        for _ in range(iterations):
            batch = self.collect()
            loss = self.update(batch)
            st = self.env.episode_stats(reset=True)
            out.append(dict(loss=loss, episodes=st["episodes"],
                            mean_return=st["sum_return"] / max(st["episodes"], 1.0)))
        return out
