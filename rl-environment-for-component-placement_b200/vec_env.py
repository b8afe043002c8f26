"""PcbVecEnv -- the batched, device-resident replacement of the reference's
``DummyPlacementEnv`` classes (one object = N environments on one B200).

API mirrors gym-0.22 as the reference uses it (``reset() -> obs``,
``step(a) -> (obs, reward, done, info)``; dummy_env_rectangular_pin.py:1544-1710)
with a leading batch dimension; every tensor is a torch view (DLPack) of memory
owned by libpcbgpu, nothing is copied or computed on the host.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .config import make_spec
from .instances import InstanceBank

OBS_KEYS = {
    "square": ("grid", "action_mask"),
    "rectangle": ("grid", "action_mask", "all_components_feature", "component_mask",
                  "placement_mask"),
    "rectangle_pin": ("grid", "action_mask", "all_components_feature", "placement_mask",
                      "all_pins_num_feature", "all_pins_cat_feature"),
    "rectangle_spatial_pin": ("grid", "action_mask", "all_components_feature", "placement_mask",
                              "all_pins_num_feature", "all_pins_cat_feature", "pin_grid",
                              "component_grid"),
}
RESULT_KEYS = ("reward", "done", "wirelength", "num_intersections", "episode_return",
               "episode_length")


def _host_buffer(x, dtype, what):
    """Address of a host buffer handed to the C ABI as a raw pointer: the dtype, device and
    layout are checked here because the library cannot (a float32 reward buffer or a strided
    action array would be silently misread or overrun)."""
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            raise ValueError(f"{what} must be a host (CPU) buffer")
        ok, have = x.dtype == getattr(torch, dtype) and x.is_contiguous(), x.dtype
        addr = x.data_ptr()
    else:
        x = np.asarray(x) if not isinstance(x, np.ndarray) else x
        ok = x.dtype == np.dtype(dtype) and x.flags["C_CONTIGUOUS"] and (
            what == "actions" or x.flags["WRITEABLE"])
        have, addr = x.dtype, x.__array_interface__["data"][0]
    if not ok:
        raise ValueError(f"{what} must be a C-contiguous {dtype} buffer, got {have}")
    return addr


class HostStepper:
    """See PcbVecEnv.host_stepper."""

    def __init__(self, env, actions, reward=None, done=None):
        n = env.num_envs
        if tuple(actions.shape[1:]) not in ((n,), (n, env.spec.action_width)):
            raise ValueError("actions must have shape [T, N] or [T, N, k]")
        self._packed = reward is None and done is None
        if self._packed:
            # one packed pinned block [reward f64 x N | done u8 x N]: a single device-to-host
            # copy per step (pcb_step_host_packed)
            self.result = torch.empty(n * 9, dtype=torch.uint8, pin_memory=True)
            reward = self.result[: n * 8].view(torch.float64)
            done = self.result[n * 8:]
        if tuple(reward.shape) != (n,) or tuple(done.shape) != (n,):
            raise ValueError("reward and done must have shape [N]")
        self.reward, self.done = reward, done
        self._keep = (actions, reward, done)  # the buffers must outlive the stepper
        self._env, self._h = env, env._h
        self._fn = env.lib.pcb_step_host_packed if self._packed else env.lib.pcb_step_host
        self._fmt = 1 if len(actions.shape) == 2 else 0
        row = n * (1 if self._fmt else env.spec.action_width) * 4
        self._a0, self._row = _host_buffer(actions, "int32", "actions"), row
        self._r, self._d = _host_buffer(reward, "float64", "reward"), _host_buffer(done, "uint8", "done")
        self._steps = int(actions.shape[0])
        self._stream = torch.cuda.current_stream(env.device).cuda_stream

    def __call__(self, t):
        if not 0 <= t < self._steps:
            raise IndexError(t)
        if self._packed:
            rc = self._fn(self._h, self._a0 + t * self._row, self._fmt, self._r, self._stream)
        else:
            rc = self._fn(self._h, self._a0 + t * self._row, self._fmt, self._r, self._d, self._stream)
        if rc < 0:
            self._env._check(rc)


class PcbVecEnv:
    """N independent placement environments stepping in CUDA kernels.

    env_config : dict with the keys of the reference's ``env_config`` JSON blocks
                 (agent/config/*.json), ``type`` in {square, rectangle,
                 rectangle_pin, rectangle_spatial_pin} or a model-type alias.
    instances  : list of instances (oracle ``Instance`` objects / dicts) or an
                 ``InstanceBank``; default: ``bank_size`` instances from the
                 library's generator seeded with ``seed``.
    obs_dtype  : "compact" (default, the measured contract: u8 / i16 / f32 tensors that are
                 views of the library's memory) or "reference": every observation is cast on
                 the device to the reference's dtype (float64; float32 for the square env) and
                 the area ratio is the exact float64 quotient -- equal to the reference's
                 arrays value for value, at 8x the bytes.
    autoreset  : True  -> a finished env restarts on its next bank slot and the
                          returned observation is the first of the new episode;
                 False -> the terminal observation is returned as the reference
                          returns it and the env stays as it is until ``reset``.
    """

    def __init__(self, env_config, num_envs, device=0, autoreset=True, obs_ring=1,
                 instances=None, bank_size=None, seed=0, norm_uses_fma=True, force_generic=False,
                 generate_on_device=False, specialize=False, obs_dtype="compact"):
        if not torch.cuda.is_available():
            raise RuntimeError("pcbgpu needs a CUDA device; there is no CPU fallback")
        if obs_dtype not in ("compact", "reference"):
            raise ValueError("obs_dtype must be 'compact' or 'reference'")
        self.obs_dtype = obs_dtype
        self.spec = make_spec(env_config, autoreset, obs_ring, norm_uses_fma)
        # specialize=True: kernels compiled for exactly this configuration (nvcc, cached);
        # otherwise built-in shapes for the canonical configs, runtime-shaped kernels else
        self.lib = _lib.load(_lib.specialized_library(self.spec) if specialize else None)
        if force_generic:  # use the runtime-shaped kernels even for a canonical shape (tests)
            self.spec.c.reserved0 = 1
        self.kind = self.spec.kind
        self.num_envs = int(num_envs)
        self.device = torch.device("cuda", int(device))
        self.autoreset = bool(autoreset)
        self._h = ctypes.c_void_p()
        self._check(self.lib.pcb_create(ctypes.byref(self.spec.c), int(device), self.num_envs,
                                       ctypes.byref(self._h)))
        self._ring = {k: _lib.tensor(self._h, k, self.lib) for k in OBS_KEYS[self.kind]}
        self._res = {k: _lib.tensor(self._h, k, self.lib) for k in RESULT_KEYS}
        # per-slice / per-half view dicts are built once: indexing a tensor costs ~1 us per key
        # on the host, which is what a step of a small batch costs on the device
        self._obs_views = {}
        self._res_views = [{k: v[h] for k, v in self._res.items()} for h in (0, 1)]
        self._info_views = [{k: r[k] for k in ("wirelength", "num_intersections", "episode_return",
                                               "episode_length")} for r in self._res_views]
        self._actions = torch.zeros((self.num_envs, self.spec.action_width), dtype=torch.int32,
                                    device=self.device)
        if self.kind != "square":
            if instances is not None:
                self.set_instances(instances)
            else:
                self.generate_instances(bank_size or self.num_envs, seed, on_device=generate_on_device)

    def _check(self, rc):
        return _lib.check(rc, self.lib)

    # -- lifecycle -----------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._ring, self._res, self._obs_views, self._res_views, self._info_views = {}, {}, {}, [], []
            self.lib.pcb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- instances -----------------------------------------------------------
    def set_instances(self, instances):
        bank = instances if isinstance(instances, InstanceBank) else InstanceBank.from_list(
            self.spec, list(instances))
        s = bank.as_struct()
        self._check(self.lib.pcb_upload_instances(self._h, ctypes.byref(s)))
        self.bank_size = bank.count

    def generate_instances(self, count, seed=0, on_device=False):
        """Fill the bank with `count` generated instances: on the host (C++, uploaded) or
        with the device generator kernel (asynchronous, in place -- call it between steps
        to keep the environments on fresh instances)."""
        arr = (ctypes.c_int32 * 12)(*self.spec.generator_ranges())
        if on_device:
            self._check(self.lib.pcb_generate_instances_device(
                self._h, arr, ctypes.c_uint64(seed), ctypes.c_int64(count), self._stream()))
        else:
            self._check(self.lib.pcb_generate_instances(self._h, arr, ctypes.c_uint64(seed),
                                                       ctypes.c_int64(count)))
        self.bank_size = int(count)

    # -- gym surface ---------------------------------------------------------
    def _obs(self):
        s = self.lib.pcb_current_slice(self._h)
        obs = self._obs_views.get(s)
        if obs is None:
            obs = self._obs_views[s] = {k: v[s] for k, v in self._ring.items()}
        return dict(obs) if self.obs_dtype == "compact" else self._to_reference(obs)

    def _to_reference(self, obs):
        """Compact device tensors -> the reference's dtypes (copies, on the device)."""
        dt = torch.float32 if self.kind == "square" else torch.float64
        out = {k: v.to(dt) for k, v in obs.items()}
        f = out.get("all_components_feature")
        if f is not None:  # h*w/(H*W) as the reference divides it (rect:53-71), not float32 widened
            f[..., 4] = (f[..., 0] * f[..., 1]) / float(self.spec.height * self.spec.width)
        return out

    @property
    def single_observation_space(self):
        """{key: (shape, dtype, low, high)} of one env's observation -- the Box bounds the
        reference declares (square:56-72, rect:196-230, pin:504-542, spatial:519-575).  A
        ``gym.spaces.Dict`` is returned instead when gym / gymnasium is importable."""
        dt = ("float32" if self.kind == "square" else "float64") if self.obs_dtype == "reference" else None
        desc = {}
        for k, v in self._ring.items():
            desc[k] = (tuple(v.shape[2:]), dt or str(v.dtype).replace("torch.", ""), -1.0,
                       float(max(self.spec.height, self.spec.width, self.spec.max_num_components,
                                 self.spec.max_pins, 3)))
        for mod in ("gymnasium", "gym"):
            try:
                spaces = __import__(mod).spaces
                return spaces.Dict({k: spaces.Box(lo, hi, shape=sh, dtype=np.dtype(d))
                                    for k, (sh, d, lo, hi) in desc.items()})
            except ImportError:
                continue
        return desc

    @property
    def single_action_space(self):
        """Sizes of the Tuple(Discrete...) action of one env: (H, W) for the square env,
        (O, H, W) otherwise (square:44-46, rect:186-194, pin:493-502); the flat Discrete of the
        wrappers has prod(sizes) actions."""
        s = self.spec
        return (s.height, s.width) if self.kind == "square" else (s.n_orient, s.height, s.width)

    def reset(self, mask=None):
        """reset() of every env (mask None) or of the envs where ``mask`` is set."""
        ptr = None
        if mask is not None:
            mask = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
            ptr = ctypes.c_void_p(mask.data_ptr())
        self._check(self.lib.pcb_reset(self._h, ptr, self._stream()))
        return self._obs()

    def step_async(self, actions):
        a = actions
        if not (isinstance(a, torch.Tensor) and a.is_cuda and a.dtype == torch.int32 and a.is_contiguous()):
            a = torch.as_tensor(actions, device=self.device).to(torch.int32).contiguous()
        if a.dim() == 1:
            if a.numel() != self.num_envs:
                raise ValueError("flat actions must have shape [num_envs]")
            fmt = 1
        else:
            if tuple(a.shape) != (self.num_envs, self.spec.action_width):
                raise ValueError(f"actions must have shape [{self.num_envs}, "
                                 f"{self.spec.action_width}] or [{self.num_envs}]")
            fmt = 0
        self._last_actions = a  # keep alive until the kernels ran
        self._check(self.lib.pcb_step(self._h, ctypes.c_void_p(a.data_ptr()), fmt, self._stream()))

    def results(self):
        """reward, done and the info tensors of the step that just ran ([N] views)."""
        return dict(self._res_views[self.lib.pcb_current_result_half(self._h)])

    def step(self, actions):
        self.step_async(actions)
        h = self.lib.pcb_current_result_half(self._h)
        r = self._res_views[h]
        return self._obs(), r["reward"], r["done"], dict(self._info_views[h])

    def sample_random_actions(self, seed=0, out=None):
        """agent/random/random_policy_*.py on device: uniform over the valid actions."""
        out = self._actions if out is None else out
        self._check(self.lib.pcb_sample_random_actions(
            self._h, ctypes.c_uint64(seed), ctypes.c_void_p(out.data_ptr()), self._stream()))
        return out

    def rollout_random(self, n_steps, seed=0, actions_log=None, reward_log=None, done_log=None):
        """n_steps of {device random policy; step} from one C call (simulate() of
        agent/random/*.py).  Optional device logs: actions_log int32 [n_steps, N, k],
        reward_log float64 [n_steps, N], done_log uint8 [n_steps, N]."""
        def ptr(t, dtype, shape):
            if t is None:
                return None
            if t.dtype != dtype or tuple(t.shape) != shape or not t.is_contiguous() or not t.is_cuda:
                raise ValueError(f"log tensor must be a contiguous cuda {dtype} tensor of shape {shape}")
            return ctypes.c_void_p(t.data_ptr())

        T, N = int(n_steps), self.num_envs
        self._check(self.lib.pcb_rollout_random_logged(
            self._h, ctypes.c_uint64(seed), T,
            ptr(actions_log, torch.int32, (T, N, self.spec.action_width)),
            ptr(reward_log, torch.float64, (T, N)), ptr(done_log, torch.uint8, (T, N)),
            self._stream()))

    def reward_overlap(self, on):
        """rollout_random: run the reward kernel of step t alongside the step kernel of t+1."""
        self._check(self.lib.pcb_set_reward_overlap(self._h, 1 if on else 0))

    def graph_rollout(self, on):
        """rollout_random of the per-step kinds as CUDA-graph replays (default on) or plain
        launches; same results."""
        self._check(self.lib.pcb_set_graph_rollout(self._h, 1 if on else 0))

    @property
    def graph_replays(self):
        return int(self.lib.pcb_graph_replays(self._h))

    def fused_rollout(self, on):
        """Pin-less kinds: rollout_random as one multi-step launch (default) or per-step
        launches.  Returns True when the fused kernel will be used."""
        return bool(self._check(self.lib.pcb_set_fused_rollout(self._h, 1 if on else 0)))

    def fused_rollout_times(self):
        """(ms, launches, steps) of the fused rollout kernel since the last call."""
        ms, n, k = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
        self._check(self.lib.pcb_fused_rollout_times(self._h, ctypes.byref(ms), ctypes.byref(n),
                                                    ctypes.byref(k)))
        return ms.value, int(n.value), int(k.value)

    def kernel_timing(self, on):
        self._check(self.lib.pcb_enable_kernel_timing(self._h, 1 if on else 0))

    def kernel_times(self):
        """{kernel: (total ms, launches)} since the last call; needs kernel_timing(True)."""
        ms = (ctypes.c_double * 4)()
        n = (ctypes.c_int64 * 4)()
        self._check(self.lib.pcb_kernel_times(self._h, ms, n))
        return {k: (ms[i], int(n[i])) for i, k in enumerate(("step", "reward", "reset", "sample"))}

    def step_host(self, actions_np, reward_np, done_np):
        """step() through host buffers (pcb_step_host): H2D actions, D2H reward/done.
        actions_np int32 [N] (flat) or [N, k]; reward_np float64 [N]; done_np uint8 [N];
        C-contiguous, ideally pinned."""
        fmt = 1 if actions_np.ndim == 1 else 0
        n = self.num_envs
        if tuple(actions_np.shape) not in ((n,), (n, self.spec.action_width)):
            raise ValueError(f"actions must have shape [{n}] or [{n}, {self.spec.action_width}]")
        if tuple(reward_np.shape) != (n,) or tuple(done_np.shape) != (n,):
            raise ValueError("reward and done must have shape [num_envs]")
        a = _host_buffer(actions_np, "int32", "actions")
        r = _host_buffer(reward_np, "float64", "reward")
        d = _host_buffer(done_np, "uint8", "done")
        self._check(self.lib.pcb_step_host(self._h, a, fmt, r, d, self._stream()))

    def host_stepper(self, actions, reward=None, done=None):
        """Pre-binds host buffers for repeated pcb_step_host calls: ``actions`` [T, N] (flat) or
        [T, N, k] int32, ``reward`` [N] float64, ``done`` [N] uint8 (numpy arrays or CPU
        tensors, C-contiguous, ideally pinned).  The returned callable ``f(t)`` steps with row t
        of ``actions``; per call it costs one ctypes call (~2 us of host time instead of ~8: the
        host time between two synchronous steps is time the GPU idles).  Without ``reward`` /
        ``done`` the stepper owns one packed pinned result block (``f.reward``, ``f.done`` are
        views of it) and every step needs a single device-to-host copy."""
        return HostStepper(self, actions, reward, done)

    def steps_host(self, actions, reward, done):
        """T steps from host buffers in one call (pcb_steps_host): ``actions`` int32 [T, N]
        (flat) or [T, N, k]; ``reward`` float64 [T, N] and ``done`` uint8 [T, N] receive every
        step's results.  Each step's action row is copied up and its results are copied down
        while neighbouring steps run; the call returns when everything has arrived.  Same
        results as T calls of ``step_host``."""
        n, T = self.num_envs, int(actions.shape[0])
        if tuple(actions.shape[1:]) not in ((n,), (n, self.spec.action_width)):
            raise ValueError("actions must have shape [T, N] or [T, N, k]")
        if tuple(reward.shape) != (T, n) or tuple(done.shape) != (T, n):
            raise ValueError("reward and done must have shape [T, N]")
        fmt = 1 if len(actions.shape) == 2 else 0
        self._check(self.lib.pcb_steps_host(
            self._h, _host_buffer(actions, "int32", "actions"), T, fmt,
            _host_buffer(reward, "float64", "reward"), _host_buffer(done, "uint8", "done"),
            self._stream()))

    # -- views ---------------------------------------------------------------
    @property
    def action_mask(self):
        return self._obs()["action_mask"]

    @property
    def action_mask_flat(self):
        """FlatteningActionMaskObservationWrapper*: C-order flatten per env."""
        return self.action_mask.reshape(self.num_envs, -1)

    def decode_flat_actions(self, a):
        """FlatteningActionWrapper*.action (env_wrappers.py:80-98, wrapper_square.py:54-69)."""
        a = torch.as_tensor(a, device=self.device).long()
        H, W = self.spec.height, self.spec.width
        if self.kind == "square":
            return torch.stack((a // W, a % W), 1).int()
        o, rem = a // (H * W), a % (H * W)
        return torch.stack((o, rem // W, rem % W), 1).int()

    def occupancy(self):
        """[N, 64] uint64 bitboard rows (bit j of row r = grid[r, j])."""
        rows = np.zeros((self.num_envs, 64), np.uint64)
        self._check(self.lib.pcb_get_occupancy(self._h, ctypes.c_void_p(rows.ctypes.data),
                                              self._stream()))
        return rows

    # -- checkpoint / inspection ---------------------------------------------
    def state_layout(self):
        out = (ctypes.c_int32 * 8)()
        self._check(self.lib.pcb_state_layout(self._h, out))
        keys = ("stride", "off_comps", "off_pins", "off_board", "max_comps", "max_pins",
                "word_bytes", "rows")
        return dict(zip(keys, list(out)))

    def state_dict(self):
        """The compact per-env state records + step counter (the env has no other mutable
        state; the instance bank is an input and is not included)."""
        lay = self.state_layout()
        buf = np.zeros((self.num_envs, lay["stride"]), np.uint8)
        step = ctypes.c_uint64()
        self._check(self.lib.pcb_get_state(self._h, ctypes.c_void_p(buf.ctypes.data),
                                          ctypes.byref(step), self._stream()))
        return {"state": buf, "step_no": int(step.value), "kind": self.kind}

    def load_state_dict(self, sd):
        buf = np.ascontiguousarray(sd["state"], np.uint8)
        if buf.shape != (self.num_envs, self.state_layout()["stride"]) or sd["kind"] != self.kind:
            raise ValueError("state_dict does not match this environment")
        self._check(self.lib.pcb_set_state(self._h, ctypes.c_void_p(buf.ctypes.data),
                                          ctypes.c_uint64(sd["step_no"]), self._stream()))

    def decode_state(self, sd=None):
        """Per env: components (h, w, x, y; -1 = not placed) and pins (component,
        relative_x/y after rotation, absolute_x/y, net, pin_id) parsed from the state
        records -- what the reference keeps in env.components / env.pins."""
        sd = sd or self.state_dict()
        lay, raw = self.state_layout(), sd["state"]
        hdr = raw[:, :16]
        n_comp, n_pins, cur = hdr[:, 0], hdr[:, 2].astype(int) | (hdr[:, 3].astype(int) << 8), hdr[:, 4]
        out = []
        for i in range(self.num_envs):
            c = raw[i, lay["off_comps"]:lay["off_comps"] + 4 * int(n_comp[i])].reshape(-1, 4).astype(int)
            c[:, 2:][c[:, 2:] == 255] = -1
            p = raw[i, lay["off_pins"]:lay["off_pins"] + 8 * int(n_pins[i])].reshape(-1, 8).astype(int)
            pins = []
            for q in p:
                comp, rx, ry, net, pid = q[0], q[1], q[2], q[3], q[4] | (q[5] << 8)
                x, y = c[comp, 2], c[comp, 3]
                pins.append(dict(component_id=comp, relative_x=rx, relative_y=ry, net_id=net, pin_id=pid,
                                 absolute_x=-1 if x < 0 else x + rx, absolute_y=-1 if x < 0 else y + ry))
            out.append(dict(components=[dict(h=r[0], w=r[1], x=r[2], y=r[3]) for r in c],
                            pins=pins, current=int(cur[i])))
        return out

    def episode_stats(self, reset=False):
        out = (ctypes.c_double * 8)()
        self._check(self.lib.pcb_episode_stats(self._h, out, 1 if reset else 0, self._stream()))
        keys = ("episodes", "sum_return", "sum_length", "sum_wirelength", "sum_intersections",
                "steps")
        return {k: out[i] for i, k in enumerate(keys)}

    def check_guards(self):
        """Bytes overwritten in the guard bands around the library's device buffers (needs
        PCB_GUARD=1 in the environment when the env is created); 0 = none."""
        return int(self._check(self.lib.pcb_check_guards(self._h, self._stream())))

    @property
    def shape_id(self):
        """0 = runtime-shaped kernels, 1..99 = built-in compile-time shape, 100 = specialised"""
        return int(self.lib.pcb_shape_id(self._h))

    @property
    def launch_count(self):
        return int(self.lib.pcb_launch_count(self._h))

    @property
    def obs_ring_tensors(self):
        return self._ring
