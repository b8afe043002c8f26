"""INTEGRATION.md is executable: the binding a maintainer of the reference would add is cut
out of the document and run.  On CPU (with the unmodified reference present) its
config_from_env must produce exactly the pcb_config_t that pcbgpu builds from the same
env_config; on the GPU its GpuVectorEnv must step like PcbVecEnv."""
import ctypes
import os
import re
import types

import numpy as np
import pytest

from ref_harness import CONFIGS, load_reference, make_reference_env

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def binding_module():
    import pcbgpu
    from pcbgpu import _lib

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(import ctypes, numpy as np, torch.*?)```", text, re.S).group(1)
    code = code.replace('ctypes.CDLL("libpcbgpu.so")', f"ctypes.CDLL({_lib.LIB_PATH!r})")
    mod = types.ModuleType("pcbgpu_binding")
    exec(compile(code, "INTEGRATION.md", "exec"), mod.__dict__)
    return mod, pcbgpu


def template_env(cfg):
    """What the binding reads from a reference env, as a plain namespace (GPU box: no reference)."""
    import pcbgpu

    spec = pcbgpu.make_spec(cfg)
    ns = types.SimpleNamespace(**{k: v for k, v in cfg.items() if k != "type"})
    ns.max_num_pins_per_component = cfg["max_component_h"] * cfg["max_component_w"]
    ns.max_wirelength, ns.max_num_intersections = spec.max_wirelength, spec.max_num_intersections
    return ns


FIELDS = ("kind", "height", "width", "max_num_components", "max_component_h", "max_component_w",
          "max_num_nets", "max_pins", "reward_type", "reward_beam_width", "weight_wirelength",
          "weight_num_intersections", "wirelength_norm", "intersections_norm", "penalty_reward",
          "penalty_wirelength", "penalty_intersections")


@pytest.mark.parametrize("name", ["C", "C_both", "S_pin", "REF_pin"])
def test_binding_config_equals_make_spec(name):
    mods = load_reference()
    mod, pcbgpu = binding_module()
    cfg = CONFIGS[name]
    env = make_reference_env(mods, cfg) if mods else template_env(cfg)
    got, want = mod.config_from_env(env), pcbgpu.make_spec(cfg).c
    for f in FIELDS:
        assert getattr(got, f) == getattr(want, f), (name, f, getattr(got, f), getattr(want, f))
    assert ctypes.sizeof(got) == ctypes.sizeof(want)


@pytest.mark.gpu
def test_binding_steps_like_the_package():
    import torch

    from parity import draw_instance

    mod, pcbgpu = binding_module()
    cfg = CONFIGS["C"]
    insts = [draw_instance(cfg, s) for s in range(16)]
    spec = pcbgpu.make_spec(cfg, True, 1, True)
    bank_arrays = pcbgpu.InstanceBank.from_list(spec, insts)  # owns the numpy arrays
    bank = bank_arrays.as_struct()
    a = mod.GpuVectorEnv(template_env(cfg), 16, bank=bank)
    b = pcbgpu.PcbVecEnv(cfg, 16, instances=insts)
    oa, ob = a.reset(), b.reset()
    for t in range(12):
        act = b.sample_random_actions(seed=3).clone()
        oa, ra, da, ia = a.step(act)
        ob, rb, db, ib = b.step(act)
        torch.cuda.synchronize()
        for k in oa:
            assert torch.equal(oa[k], ob[k]), (t, k)
        assert torch.equal(ra, rb) and torch.equal(da, db)
        assert torch.equal(ia["wirelength"], ib["wirelength"])
    assert float(np.abs(rb.cpu().numpy()).sum()) >= 0.0
    mod.lib.pcb_destroy(a.h)
    b.close()
