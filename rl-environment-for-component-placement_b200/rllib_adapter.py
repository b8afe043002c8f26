"""RLlib ``VectorEnv``-shaped adapter.

Implements the method surface RLlib 2.2's ``VectorEnv`` expects (``vector_reset``,
``reset_at``, ``vector_step``, ``get_sub_environments``) on top of PcbVecEnv with numpy
observations in the reference dtypes, so ``experiments/PPO/PPO.py`` can register a
creator that returns one vectorised env per rollout worker.  ray is not installed in
this image: the class is duck-typed (no ray import) and exercised by the GPU tests
through those methods only -- **untested against RLlib itself**.
"""
import numpy as np
import torch

from .vec_env import OBS_KEYS, PcbVecEnv


class VectorEnvAdapter:
    def __init__(self, env_config, num_envs, device=0, seed=0, flatten=True):
        self.vec = PcbVecEnv(env_config, num_envs, device=device, autoreset=False,
                             bank_size=max(4 * num_envs, 64), seed=seed)
        self.num_envs = num_envs
        self.flatten = flatten  # FlatteningActionMaskObservationWrapper + flat Discrete actions
        self._dtype = np.float32 if self.vec.kind == "square" else np.float64
        self._obs = None

    def _split(self, obs):
        host = {k: obs[k].cpu().numpy().astype(self._dtype) for k in OBS_KEYS[self.vec.kind]}
        if self.flatten:
            host["action_mask"] = host["action_mask"].reshape(self.num_envs, -1)
        self._obs = [{k: v[i] for k, v in host.items()} for i in range(self.num_envs)]
        return self._obs

    def vector_reset(self):
        return self._split(self.vec.reset())

    def reset_at(self, index=None):
        mask = torch.zeros(self.num_envs, dtype=torch.uint8)
        mask[0 if index is None else index] = 1
        return self._split(self.vec.reset(mask=mask))[0 if index is None else index]

    def vector_step(self, actions):
        a = torch.as_tensor(np.asarray(actions), dtype=torch.int32)
        obs, rew, done, info = self.vec.step(a)
        torch.cuda.synchronize(self.vec.device)
        obs_list = self._split(obs)
        rew, done = rew.cpu().numpy(), done.cpu().numpy().astype(bool)
        wl, ni = info["wirelength"].cpu().numpy(), info["num_intersections"].cpu().numpy()
        pin = self.vec.kind in ("rectangle_pin", "rectangle_spatial_pin")
        infos = [({"wirelength": float(wl[i]), "num_intersections": float(ni[i])}
                  if pin and done[i] else {}) for i in range(self.num_envs)]
        return obs_list, [float(r) for r in rew], [bool(d) for d in done], infos

    def get_sub_environments(self):
        return []

    def close(self):
        self.vec.close()
