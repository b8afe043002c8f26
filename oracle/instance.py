"""Instance container + the reference's instance generators (oracle; test infra).

An *instance* is what ``reset()`` draws in the reference: the component list
and, for the pin variants, the pins with their component / net / relative
coordinates.  The reference draws it from two process-global Mersenne-Twister
streams (``np.random.*`` legacy RandomState and Python ``random``); the
generators below issue the same calls in the same order,
so after ``np.random.seed(s); random.seed(s)`` they return the instance the
reference's ``reset()`` would build.
"""
import random
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Instance:
    """Flat description of one placement problem.

    comp_h/comp_w : per component (comp_id == index).
    pin_*         : per pin, in ``env.pins`` order == net-major order
                    (dummy_env_rectangular_pin.py:1166-1169).  ``pin_id`` is the
                    feature-row id: per (net, component) allocation index in the
                    pin env (:1256-1258), global creation index in the spatial
                    env (spatial :977-981, no reassignment).
    """

    comp_h: np.ndarray
    comp_w: np.ndarray
    pin_comp: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))
    pin_rx: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))
    pin_ry: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))
    pin_net: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))
    pin_id: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))
    n_nets: int = 0

    @property
    def n_comp(self):
        return len(self.comp_h)

    @property
    def n_pins(self):
        return len(self.pin_comp)

    def to_dict(self):
        return {
            "comp_h": np.asarray(self.comp_h, np.int64),
            "comp_w": np.asarray(self.comp_w, np.int64),
            "pin_comp": np.asarray(self.pin_comp, np.int64),
            "pin_rx": np.asarray(self.pin_rx, np.int64),
            "pin_ry": np.asarray(self.pin_ry, np.int64),
            "pin_net": np.asarray(self.pin_net, np.int64),
            "pin_id": np.asarray(self.pin_id, np.int64),
            "n_nets": np.asarray(self.n_nets, np.int64),
        }

    @staticmethod
    def from_dict(d):
        return Instance(
            np.asarray(d["comp_h"], np.int64),
            np.asarray(d["comp_w"], np.int64),
            np.asarray(d["pin_comp"], np.int64),
            np.asarray(d["pin_rx"], np.int64),
            np.asarray(d["pin_ry"], np.int64),
            np.asarray(d["pin_net"], np.int64),
            np.asarray(d["pin_id"], np.int64),
            int(d["n_nets"]),
        )


def _draw_components(cfg):
    """n, then (h, w) per component -- dummy_env_rectangular.py:260-271,
    dummy_env_rectangular_pin.py:1039-1041, 992-997."""
    n = np.random.randint(cfg["min_num_components"], cfg["max_num_components"] + 1)
    hs, ws = [], []
    for _ in range(n):
        hs.append(np.random.randint(cfg["min_component_h"], cfg["max_component_h"] + 1))
        ws.append(np.random.randint(cfg["min_component_w"], cfg["max_component_w"] + 1))
    return np.array(hs, np.int64), np.array(ws, np.int64)


def generate_rect(cfg):
    """dummy_env_rectangular.py:253-273."""
    h, w = _draw_components(cfg)
    return Instance(h, w)


def _truncated_multinomial(n, m, p, k):
    """dummy_env_rectangular_pin.py:258-295: m sequential single draws, a net
    that reached k extra pins gets probability 0 and the rest renormalise."""
    if k < 1 or k > m:
        raise ValueError("Invalid value of k.")
    if not np.isclose(np.sum(p), 1):
        raise ValueError("Probabilities should sum to 1.")
    got = np.zeros(n, dtype=int)
    for _ in range(m):
        q = p * (got < k)
        q /= np.sum(q)
        got += np.random.multinomial(1, q)
    return got.tolist()


def generate_pin(cfg, spatial=False):
    """dummy_env_rectangular_pin.py:1006-1035 (generate_instances) and the
    helpers it calls; ``spatial=True`` follows the two generator differences of
    dummy_env_rectangular_pin_spatial.py."""
    comp_h, comp_w = _draw_components(cfg)
    n_comp = len(comp_h)
    area = comp_h * comp_w
    total_area = int(area.sum())

    # nets and pin count: :1043-1065
    n_nets = np.random.randint(cfg["min_num_nets"], cfg["max_num_nets"] + 1)
    n_nets = min(n_nets, int(total_area / 2))
    lo, hi = cfg["min_num_pins_per_net"], cfg["max_num_pins_per_net"]
    total_pins = np.random.randint(lo * n_nets, hi * n_nets + 1)
    total_pins = min(total_pins, total_area)

    # pins -> nets: :1067-1127.  Pins are identified by creation index.
    nd = max(0, min(9, cfg["net_distribution"]))
    logits = np.random.normal(1 / n_nets, 1 / (nd + 1), n_nets)
    probs = np.exp(logits) / np.sum(np.exp(logits))
    net_of = np.full(total_pins, -1, np.int64)
    net_pins = []  # per net: creation indices in net_pins[net] order
    nxt = 0
    for net in range(n_nets):
        net_of[nxt:nxt + lo] = net
        net_pins.append(list(range(nxt, nxt + lo)))
        nxt += lo
    extra = total_pins - lo * n_nets
    if hi > lo and extra > 0:
        alloc = _truncated_multinomial(n_nets, extra, probs, min(hi - lo, extra))
        for net in range(n_nets):
            if alloc[net] > 0:
                net_of[nxt:nxt + alloc[net]] = net
                net_pins[net].extend(range(nxt, nxt + alloc[net]))
                nxt += alloc[net]

    # pins -> components: :1129-1264
    ps = max(0, min(9, cfg["pin_spread"]))
    if spatial:
        k0 = min(int((ps / 10) * n_comp) + 1, n_comp)  # spatial :1102-1104
    else:
        k0 = min(max(int(((ps + 1) / 10) * n_comp), 1), n_comp)  # :1148-1151
    space = [(c, int(area[c])) for c in range(n_comp)]  # ordered like the dict
    comp_of = np.full(total_pins, -1, np.int64)
    row_id = np.arange(total_pins, dtype=np.int64)  # spatial keeps creation index
    for net in range(n_nets):
        todo = len(net_pins[net])
        space = sorted(space, key=lambda it: it[1], reverse=True)  # stable, :1203-1209
        k = k0 - 1
        cap = 0
        while cap < todo:  # :1212-1222
            k += 1
            cap = sum(s for _, s in space[:k])
        done = 0
        while todo > 0:  # :1225-1262
            chosen = space[:k]
            tot = sum(s for _, s in chosen)
            counts = np.random.multinomial(todo, np.array([s / tot for _, s in chosen]))
            for i, ((c, s), cnt) in enumerate(zip(chosen, counts)):
                cnt = min(int(cnt), space[i][1])
                space[i] = (c, space[i][1] - cnt)
                for j in range(cnt):
                    p = net_pins[net][done + j]
                    comp_of[p] = c
                    if not spatial:
                        row_id[p] = j  # :1256-1258
                done += cnt
                todo -= cnt

    # env.pins order = net-major: :1166-1169
    order = [p for net in range(n_nets) for p in net_pins[net]]
    pin_comp = comp_of[order]
    pin_net = net_of[order]
    pin_id = row_id[order]

    # relative coordinates: per component, per pin in component.pins order
    # (= env.pins order filtered), Python-random pick without replacement from
    # the row-major cell list: :1029-1035, 1462-1498
    pin_rx = np.full(total_pins, -1, np.int64)
    pin_ry = np.full(total_pins, -1, np.int64)
    for c in range(n_comp):
        cells = [(x, y) for x in range(int(comp_h[c])) for y in range(int(comp_w[c]))]
        for i in np.flatnonzero(pin_comp == c):
            cell = random.choice(cells)
            cells.remove(cell)
            pin_rx[i], pin_ry[i] = cell
    return Instance(comp_h, comp_w, pin_comp, pin_rx, pin_ry, pin_net, pin_id, int(n_nets))
