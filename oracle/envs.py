"""The four placement environments restated on numpy (oracle; test infra).

SquareEnv  <- environment/dummy_env_square.py
RectEnv    <- environment/dummy_env_rectangular.py
PinEnv     <- environment/dummy_env_rectangular_pin.py and, with
              ``spatial=True``, environment/dummy_env_rectangular_pin_spatial.py

Same gym-0.22 surface as the reference: ``reset() -> obs`` and
``step(action) -> (obs, reward, done, info)`` with observation dicts of fresh
float64 arrays (float32 for the square env), attributes ``grid`` /
``action_mask`` readable by a random policy.  ``reset(instance=...)`` injects
an instance instead of drawing one (the reference's tests do the same by
overwriting ``env.components`` after ``reset()``).
"""
import numpy as np
from scipy import signal

from . import reward as rw
from .instance import Instance, generate_pin, generate_rect


def _mask_one(grid, ph, pw):
    """compute_action_mask_orientation for a (ph, pw) footprint:
    dummy_env_rectangular.py:526-567 / dummy_env_rectangular_pin.py:1808-1851."""
    H, W = grid.shape
    m = np.ones((H, W))
    m[[i for i in range(H - 1, H - ph, -1) if i >= 0], :] = 0
    m[:, [j for j in range(W - 1, W - pw, -1) if j >= 0]] = 0
    free = signal.convolve2d(grid, np.ones((ph, pw)), mode="valid") == 0
    m[: H - ph + 1, : W - pw + 1] = np.where(free, 1, 0)
    return m


class SquareEnv:
    """dummy_env_square.py:37-255."""

    def __init__(self, height, width, component_n):
        self.height, self.width, self.component_n = height, width, component_n
        if height < 0 or width < 0:
            raise ValueError("Grid size must be greater than 0.")
        if component_n > height or component_n > width:
            raise ValueError("Component size must be less than or equal to the grid size.")
        self.mask_shape = (height, width)

    def _obs(self):
        return {"grid": self.grid.copy(), "action_mask": self.action_mask.copy()}

    def reset(self, *a, **k):  # :74-113
        n = self.component_n
        self.grid = np.zeros((self.height, self.width), dtype=np.float32)
        self.action_mask = np.ones((self.height, self.width), dtype=np.float32)
        if n > 1:
            self.action_mask[max(self.height - n + 1, 0):, :] = 0
            self.action_mask[:, max(self.width - n + 1, 0):] = 0
        return self._obs()

    def validate_action(self, x, y):  # :155-173
        try:
            return self.action_mask[x, y] == 1
        except IndexError:
            return False

    def step(self, action, verbose=False):  # :115-153
        x, y = action
        if not self.validate_action(x, y):
            return self._obs(), 0.0, True, {}
        n = self.component_n
        self.grid[x:x + n, y:y + n] = 1
        # update_action_mask (:187-244): the four sub-rectangles together are
        # rows [x-n+1, x+n) x cols [y-n+1, y+n), clipped at 0.
        self.action_mask[max(0, x - n + 1):x + n, max(0, y - n + 1):y + n] = 0
        done = bool(np.all(self.action_mask == 0))
        return self._obs(), 1.0, done, {}


class RectEnv:
    """dummy_env_rectangular.py:152-599."""

    KEYS = ("grid", "action_mask", "all_components_feature", "component_mask", "placement_mask")

    def __init__(self, height, width, min_component_w, max_component_w, min_component_h,
                 max_component_h, max_num_components, min_num_components):
        self.height, self.width, self.area = height, width, height * width
        self.cfg = dict(min_component_w=min_component_w, max_component_w=max_component_w,
                        min_component_h=min_component_h, max_component_h=max_component_h,
                        max_num_components=max_num_components,
                        min_num_components=min_num_components)
        self.max_num_components = max_num_components
        if height < 0 or width < 0:
            raise ValueError("Grid size must be greater than 0.")
        if max_component_w > height or max_component_h > width:  # sic, :241
            raise ValueError("Component size must be less than or equal to the grid size.")
        if min_component_w < 1 or min_component_h < 1:
            raise ValueError("Component size must be greater than 0.")
        if max_num_components < 1 or max_num_components > self.area:
            raise ValueError("Number of components must be greater than 0"
                             "and less than or equal to the grid area.")
        self.mask_shape = (2, height, width)

    def _obs(self):
        return {k: getattr(self, k).copy() for k in self.KEYS}

    def _feature(self, c):  # Component.calculate_feature, :60-79
        x, y = self.pos[c]
        h, w = int(self.inst.comp_h[c]), int(self.inst.comp_w[c])
        return np.array([h, w, x, y, (h * w) / self.area])

    def compute_action_mask(self, c):  # :569-584
        h, w = int(self.inst.comp_h[c]), int(self.inst.comp_w[c])
        m = np.zeros((2, self.height, self.width))
        m[0] = _mask_one(self.grid, h, w)
        m[1] = _mask_one(self.grid, w, h)
        return m

    def reset(self, *a, instance=None, **k):  # :310-351
        C = self.max_num_components
        self.grid = np.zeros((self.height, self.width))
        self.inst = instance if instance is not None else generate_rect(self.cfg)
        n = self.inst.n_comp
        self.pos = [(-1.0, -1.0)] * n
        self.cur = 0
        self.all_components_feature = np.zeros((C, 5))
        self.action_mask = self.compute_action_mask(0)
        self.placement_mask = np.zeros(C)
        self.component_mask = np.zeros(C)
        self.component_mask[:n] = 1.0
        for c in range(n):
            self.all_components_feature[c, :] = self._feature(c)
        return self._obs()

    def validate_action(self, o, x, y):  # :434-456
        try:
            return self.action_mask[o, x, y] == 1
        except IndexError:
            return False

    def step(self, action, verbose=False):  # :353-432
        o, x, y = action
        if not self.validate_action(o, x, y):
            return self._obs(), 0.0, True, {}
        c = self.cur
        h, w = int(self.inst.comp_h[c]), int(self.inst.comp_w[c])
        ph, pw = (h, w) if o == 0 else (w, h)
        self.grid[x:x + ph, y:y + pw] = 1
        self.pos[c] = (x, y)
        self.all_components_feature[c, :] = self._feature(c)
        self.placement_mask[c] = 1.0
        if c + 1 < self.inst.n_comp:
            self.cur = c + 1
            self.action_mask = self.compute_action_mask(self.cur)
            done = bool(np.all(self.action_mask == 0))  # :586-599
        else:
            self.cur = -1
            self.action_mask = np.zeros((2, self.height, self.width))
            done = True
        return self._obs(), 1.0, done, {}


class PinEnv:
    """dummy_env_rectangular_pin.py:396-1886; spatial=True switches to the
    observation layout and constants of dummy_env_rectangular_pin_spatial.py."""

    def __init__(self, height, width, net_distribution, pin_spread, min_component_w,
                 max_component_w, min_component_h, max_component_h, max_num_components,
                 min_num_components, min_num_nets, max_num_nets, max_num_pins_per_net,
                 min_num_pins_per_net=2, reward_type="both", reward_beam_width=2,
                 weight_wirelength=0.5, weight_num_intersections=0.5, spatial=False):
        self.spatial = spatial
        self.height, self.width, self.area = height, width, height * width
        self.cfg = dict(net_distribution=net_distribution, pin_spread=pin_spread,
                        min_component_w=min_component_w, max_component_w=max_component_w,
                        min_component_h=min_component_h, max_component_h=max_component_h,
                        max_num_components=max_num_components,
                        min_num_components=min_num_components, min_num_nets=min_num_nets,
                        max_num_nets=max_num_nets, max_num_pins_per_net=max_num_pins_per_net,
                        min_num_pins_per_net=min_num_pins_per_net)
        self.max_num_components = max_num_components
        self.max_num_nets = max_num_nets
        self.max_num_pins_per_component = max_component_h * max_component_w  # :481
        self.reward_type = reward_type
        self.reward_beam_width = reward_beam_width
        self.weight_wirelength = weight_wirelength
        self.weight_num_intersections = weight_num_intersections
        self._validate()
        diag = rw.distance((0, 0), (height, width))  # :761-783
        self.max_wirelength = 0.5 * diag * (max_num_nets * max_num_pins_per_net)
        bound = 0.5 * (max_num_pins_per_net ** 2) * max_num_nets * (max_num_nets - 1)
        if spatial:
            self.max_wirelength = self.max_wirelength / (height + width)  # spatial :746
            self.max_num_intersections = bound  # float, spatial :785-791
        else:
            self.max_num_intersections = int(bound)  # :822-830
        self.reward_intersection = -1
        self.reward_wirelength = -1.0
        self.mask_shape = (4, height, width)
        self.keys = ["grid", "action_mask", "all_components_feature", "placement_mask",
                     "all_pins_num_feature", "all_pins_cat_feature"]
        if spatial:
            self.keys += ["pin_grid", "component_grid"]

    def _validate(self):
        c = self.cfg
        H, W = self.height, self.width
        if H < 0 or W < 0:
            raise ValueError("Grid size must be greater than 0.")
        if self.spatial:  # spatial :572-601
            if c["max_component_w"] > H or c["max_component_h"] > W:
                raise ValueError("Component size must be less than or equal to the grid size.")
        elif c["max_component_w"] > W or c["max_component_h"] > H:  # :600
            raise ValueError("Component size must be less than or equal to the grid size.")
        if c["min_component_w"] < 1 or c["min_component_h"] < 1:
            raise ValueError("Component size must be greater than 0.")
        if c["max_num_components"] < 1 or c["max_num_components"] > H * W:
            raise ValueError("Number of components must be greater than 0"
                             "and less than or equal to the grid area.")
        bw = self.reward_beam_width
        if self.spatial:
            if self.reward_type not in ["beam", "centroid", "both"]:
                raise ValueError("Invalid type of routing for the reward")
            if bw < 2 or bw > c["max_num_pins_per_net"]:
                raise ValueError("The beam width must be an integer greater than 2 and at "
                                 "most the maximum number of pins per net.")
            if not isinstance(bw, int):
                raise ValueError("The beam width must be an integer.")
            if not isinstance(self.weight_wirelength, float):
                raise ValueError("The weight of wirelength must be a float.")
            if self.weight_wirelength < 0:
                raise ValueError("The weight of wirelength must be greater than 0.")
            return
        if c["min_num_pins_per_net"] > c["max_num_pins_per_net"]:  # :618-641
            raise ValueError("min_num_pins_per_net must not be greater than max num pins per net")
        if c["min_num_pins_per_net"] < 2:
            raise ValueError("min num pins per net must be less than or equal to 2.")
        if (c["min_num_pins_per_net"] * c["min_num_nets"]
                > c["min_component_w"] * c["min_component_h"] * c["min_num_components"]):
            raise ValueError("min_num_pins_per_net * min_num_nets must be less than or equal "
                             "to the total minimum area covered by the components")
        if bw < 1 or not isinstance(bw, int):
            raise ValueError("Beam width must be a positive integer.")
        if self.reward_type not in ["beam", "centroid", "both"]:
            raise ValueError("Reward type must be either 'beam', 'centroid', or 'both'.")

    # ---- observation pieces -------------------------------------------------
    def _obs(self):
        return {k: getattr(self, k).copy() for k in self.keys}

    def _comp_feature(self, c):  # Component.calculate_feature :217-239 / spatial :203-239
        h, w = int(self.inst.comp_h[c]), int(self.inst.comp_w[c])
        x, y = self.pos[c]
        f = np.array([h, w, x, y, (h * w) / self.area])
        if self.spatial:
            ids = np.full(self.max_num_pins_per_component, -1)
            mine = self.inst.pin_id[self._comp_pins[c]]
            ids[: len(mine)] = mine
            f = np.append(f, ids)
        return f

    def _write_pin_rows(self):
        """update_all_pins_feature, :1521-1542 / spatial :1469-1485: component
        order, then that component's pins in env.pins order; later rows
        overwrite earlier ones when (component, pin_id) repeats."""
        I = self.inst
        for c in range(I.n_comp):
            for p in self._comp_pins[c]:
                row = [self.rel[p][0], self.rel[p][1], self.abs[p][0], self.abs[p][1]]
                if self.spatial:
                    self.all_pins_num_feature[I.pin_id[p], :] = row
                    self.all_pins_cat_feature[I.pin_id[p], :] = [I.pin_net[p], c]
                else:
                    self.all_pins_num_feature[c, I.pin_id[p], :] = row
                    self.all_pins_cat_feature[c, I.pin_id[p], :] = [I.pin_net[p]]

    def _draw_pins(self):  # spatial :1663-1675
        code = self.grid.copy()
        I = self.inst
        for p in range(I.n_pins):
            ax, ay = self.abs[p]
            if ax == -1 or ay == -1:
                continue
            code[ax, ay] = I.pin_net[p] + 2
        self.pin_grid = np.eye(self.max_num_nets + 2)[code.astype(int)][:, :, 1:]

    def _draw_components(self):  # spatial :1677-1697 (reset only, un-rotated coords)
        I = self.inst
        g = np.zeros((I.n_comp, self.cfg["max_component_h"], self.cfg["max_component_w"],
                      self.max_num_nets + 1))
        for p in range(I.n_pins):
            g[I.pin_comp[p], I.pin_rx[p], I.pin_ry[p], I.pin_net[p] + 1] = 1
        g[:, :, :, 0] = 1
        self.component_grid = g

    def compute_action_mask(self, c):  # :1853-1870
        h, w = int(self.inst.comp_h[c]), int(self.inst.comp_w[c])
        m = np.zeros((4, self.height, self.width))
        m[0] = _mask_one(self.grid, h, w)
        m[1] = _mask_one(self.grid, w, h)
        m[2] = m[0]
        m[3] = m[1]
        return m

    # ---- gym surface --------------------------------------------------------
    def reset(self, *a, instance=None, **k):  # :1544-1597 / spatial :1487-1549
        C, P = self.max_num_components, self.max_num_pins_per_component
        self.grid = np.zeros((self.height, self.width))
        self.inst = I = instance if instance is not None else generate_pin(self.cfg, self.spatial)
        self.pos = [(-1, -1)] * I.n_comp
        # component.pins / net_pins[net]: pin indices in env.pins order (pin:1029-1035, 1166-1169)
        self._comp_pins = [[] for _ in range(I.n_comp)]
        self._net_pins = [[] for _ in range(I.n_nets)]
        for p in range(I.n_pins):
            self._comp_pins[int(I.pin_comp[p])].append(p)
            self._net_pins[int(I.pin_net[p])].append(p)
        self.rel = [(int(I.pin_rx[p]), int(I.pin_ry[p])) for p in range(I.n_pins)]
        self.abs = [(-1, -1)] * I.n_pins
        if self.spatial:
            self.pin_grid = np.zeros((self.height, self.width, self.max_num_nets + 1))
            self._draw_components()
            self.all_components_feature = np.zeros((C, 5 + P))
            self.all_pins_num_feature = np.zeros((C * P + 1, 4))
            self.all_pins_cat_feature = np.zeros((C * P + 1, 2))
            self.all_pins_cat_feature[-1, :] = -1
        else:
            self.all_components_feature = np.zeros((C, 5))
            self.all_pins_num_feature = np.zeros((C, P, 4))
            self.all_pins_cat_feature = np.zeros((C, P, 1))
        self.cur = 0
        self.action_mask = self.compute_action_mask(0)
        self.placement_mask = np.ones(C)
        self.placement_mask[I.n_comp:] = 0.0
        self.placement_mask[0] = 3.0
        for c in range(I.n_comp):
            self.all_components_feature[c, :] = self._comp_feature(c)
        self._write_pin_rows()
        return self._obs()

    def validate_action(self, o, x, y):  # :1712-1736
        try:
            return self.action_mask[o, x, y] == 1
        except IndexError:
            return False

    def _place(self, c, o, x, y):
        """update_grid :1738-1765 + Component.place_component :156-204."""
        I = self.inst
        h, w = int(I.comp_h[c]), int(I.comp_w[c])
        ph, pw = (h, w) if o in (0, 2) else (w, h)
        self.grid[x:x + ph, y:y + pw] = 1
        self.pos[c] = (x, y)
        for p in self._comp_pins[c]:
            rx, ry = self.rel[p]
            if o == 1:
                rx, ry = ry, h - rx - 1
            elif o == 2:
                rx, ry = h - rx - 1, w - ry - 1
            elif o == 3:
                rx, ry = w - ry - 1, rx
            self.rel[p] = (rx, ry)
            self.abs[p] = (x + rx, y + ry)

    def step(self, action, verbose=False):  # :1599-1710 / spatial :1551-1661
        o, x, y = action
        if not self.validate_action(o, x, y):
            if self.spatial:
                self._draw_pins()
            reward = self.find_reward()
            return self._obs(), reward, True, self._info()
        c = self.cur
        self._place(c, o, x, y)
        self.all_components_feature[c, :] = self._comp_feature(c)
        self._write_pin_rows()
        if self.spatial:
            self._draw_pins()
        self.placement_mask[c] = 2.0
        if c + 1 < self.inst.n_comp:
            self.cur = c + 1
            self.placement_mask[self.cur] = 3.0
            self.action_mask = self.compute_action_mask(self.cur)
            done = bool(np.all(self.action_mask == 0))  # :1872-1886
        else:
            self.cur = -1
            self.action_mask = np.zeros((4, self.height, self.width))
            done = True
        if not done:
            return self._obs(), 0.0, False, {}
        reward = self.find_reward()
        return self._obs(), reward, True, self._info()

    def _info(self):
        return {"wirelength": self.reward_wirelength,
                "num_intersections": self.reward_intersection}

    # ---- reward -------------------------------------------------------------
    def net_positions(self):  # get_net_pin_positions :1266-1280
        return [[self.abs[p] for p in pins] for pins in self._net_pins]

    def normalizers(self):  # :879-896
        c = self.cfg
        by_area = (np.mean([c["min_component_h"], c["max_component_h"]])
                   * np.mean([c["min_component_w"], c["max_component_w"]])
                   * np.mean([c["min_num_components"], c["max_num_components"]]))
        by_net = (np.mean([c["min_num_pins_per_net"], c["max_num_pins_per_net"]])
                  * np.mean([c["min_num_nets"], c["max_num_nets"]]))
        return self.height + self.width, np.min([by_area, by_net])

    def penalty(self):  # :899-909
        wl_norm, ni_norm = self.normalizers()
        return (-self.weight_wirelength * (self.max_wirelength / wl_norm)
                - self.weight_num_intersections * (self.max_num_intersections / ni_norm))

    def find_reward(self):  # :832-975
        wl_norm, ni_norm = self.normalizers()
        if self.cur != -1:
            self.reward_intersection = self.max_num_intersections
            self.reward_wirelength = self.max_wirelength
            return self.penalty()
        nets = self.net_positions()
        if self.reward_type == "beam":
            route = rw.route_beam(nets, self.reward_beam_width)
            crossings = rw.count_crossings(route)
        elif self.reward_type == "centroid":
            route = rw.route_centroid(nets)
            crossings = rw.count_crossings(route)
        else:  # "both": fewer crossings wins, beam on ties (:643-661, 951-974)
            routes = [rw.route_beam(nets, self.reward_beam_width), rw.route_centroid(nets)]
            counts = [rw.count_crossings(r) for r in routes]
            crossings = min(counts)
            route = routes[counts.index(crossings)]
        self.raw_crossings = crossings
        wl = rw.wirelength(route) / wl_norm
        ni = crossings / ni_norm
        self.reward_intersection = ni
        self.reward_wirelength = wl
        return -1 * (self.weight_wirelength * wl + self.weight_num_intersections * ni)


_PIN_ARGS = ("height", "width", "net_distribution", "pin_spread", "min_component_w",
             "max_component_w", "min_component_h", "max_component_h", "max_num_components",
             "min_num_components", "min_num_nets", "max_num_nets", "max_num_pins_per_net",
             "min_num_pins_per_net", "reward_type", "reward_beam_width", "weight_wirelength")


def make_oracle_env(env_config):
    """utils/agent/utils.py:317-391 (init_env) on the oracle classes.  Like the
    reference, ``weight_num_intersections`` is NOT forwarded unless the caller
    passes ``forward_weight_num_intersections=True`` in the dict."""
    kind = env_config["type"]
    if kind == "square":
        return SquareEnv(env_config["height"], env_config["width"], env_config["component_n"])
    if kind in ("rectangle", "rectangle_factorized"):
        return RectEnv(*[env_config[k] for k in (
            "height", "width", "min_component_w", "max_component_w", "min_component_h",
            "max_component_h", "max_num_components", "min_num_components")])
    kw = {}
    if env_config.get("forward_weight_num_intersections"):
        kw["weight_num_intersections"] = env_config["weight_num_intersections"]
    return PinEnv(*[env_config[k] for k in _PIN_ARGS],
                  spatial=(kind == "rectangle_spatial_pin"), **kw)
