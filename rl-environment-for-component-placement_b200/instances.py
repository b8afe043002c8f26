"""Host-side instance bank: what ``reset()`` draws in the reference
(generate_instances, dummy_env_rectangular_pin.py:1006-1035), as the struct of
arrays that pcb_upload_instances takes (include/pcb_env.h, pcb_instance_bank_t)."""
import ctypes

import numpy as np

from ._lib import PcbInstanceBank


class InstanceBank:
    def __init__(self, spec, count):
        self.spec, self.count = spec, int(count)
        C, M = spec.max_num_components, max(spec.max_pins, 1)
        self.n_comp = np.zeros(count, np.int32)
        self.comp_h = np.zeros((count, C), np.uint8)
        self.comp_w = np.zeros((count, C), np.uint8)
        self.n_pins = np.zeros(count, np.int32)
        self.n_nets = np.zeros(count, np.int32)
        self.pin_comp = np.zeros((count, M), np.uint8)
        self.pin_rx = np.zeros((count, M), np.uint8)
        self.pin_ry = np.zeros((count, M), np.uint8)
        self.pin_net = np.zeros((count, M), np.uint8)
        self.pin_id = np.zeros((count, M), np.uint16)

    def set(self, i, inst):
        """inst: object or dict with comp_h, comp_w and (pin kinds) pin_comp, pin_rx,
        pin_ry, pin_net, pin_id, n_nets -- pins in env.pins (net-major) order."""
        get = (lambda k: inst[k]) if isinstance(inst, dict) else (lambda k: getattr(inst, k))
        h, w = np.asarray(get("comp_h")), np.asarray(get("comp_w"))
        n = len(h)
        if n > self.comp_h.shape[1]:
            raise ValueError(f"instance has {n} components, max {self.comp_h.shape[1]}")
        self.n_comp[i] = n
        self.comp_h[i, :n], self.comp_w[i, :n] = h, w
        if self.spec.kind in ("square", "rectangle"):
            return
        pc = np.asarray(get("pin_comp"))
        m = len(pc)
        if m > self.pin_comp.shape[1]:
            raise ValueError(f"instance has {m} pins, max {self.pin_comp.shape[1]}")
        self.n_pins[i], self.n_nets[i] = m, int(get("n_nets"))
        self.pin_comp[i, :m] = pc
        self.pin_rx[i, :m] = np.asarray(get("pin_rx"))
        self.pin_ry[i, :m] = np.asarray(get("pin_ry"))
        self.pin_net[i, :m] = np.asarray(get("pin_net"))
        self.pin_id[i, :m] = np.asarray(get("pin_id"))

    @classmethod
    def from_list(cls, spec, instances):
        bank = cls(spec, len(instances))
        for i, inst in enumerate(instances):
            bank.set(i, inst)
        return bank

    def as_struct(self):
        b = PcbInstanceBank()
        b.count = self.count
        for name in ("n_comp", "comp_h", "comp_w", "n_pins", "n_nets", "pin_comp", "pin_rx",
                     "pin_ry", "pin_net", "pin_id"):
            arr = np.ascontiguousarray(getattr(self, name))
            setattr(self, name, arr)
            setattr(b, name, arr.ctypes.data_as(ctypes.c_void_p))
        return b


def generate_bank(spec, count, seed=0):
    """``count`` instances from the library's host generator (pcb_generate_host) as an
    InstanceBank -- no GPU needed."""
    from . import _lib

    lib = _lib.load()
    bank = InstanceBank(spec, count)
    arr = (ctypes.c_int32 * 12)(*spec.generator_ranges())
    ptr = lambda a: ctypes.c_void_p(a.ctypes.data)  # noqa: E731
    from .config import KINDS

    _lib.check(lib.pcb_generate_host(
        KINDS[spec.kind], spec.max_num_components, max(spec.max_pins, 1), spec.max_num_nets, arr,
        ctypes.c_uint64(seed), ctypes.c_int64(count), ptr(bank.n_comp), ptr(bank.comp_h),
        ptr(bank.comp_w), ptr(bank.n_pins), ptr(bank.n_nets), ptr(bank.pin_comp), ptr(bank.pin_rx),
        ptr(bank.pin_ry), ptr(bank.pin_net), ptr(bank.pin_id)))
    return bank
