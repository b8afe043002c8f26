"""CPU oracle for the RL-PCB placement environments -- TEST INFRASTRUCTURE.

A numpy restatement of the reset/step dynamics of the four
``DummyPlacementEnv`` variants of PBozmarov/RL-Environment-for-Component-
Placement (environment/dummy_env_{square,rectangular,rectangular_pin,
rectangular_pin_spatial}.py) and of the flatten wrappers.  Every function
cites the reference file:line it follows.

Who may use it: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- always as the
checker or the timed CPU baseline, never as part of the product path.  The
product package (``pcbgpu``) must not import anything from here.

Parity status: PINNED.  The oracle is checked (tests/test_oracle_*.py)
 * against the reference's own known-answer tests,
 * against the unmodified reference imported from /root/reference behind the
   gym stub in tests/refstub (differential, seeded; runs only where the
   reference is present, i.e. in the build container),
 * against golden trajectories the reference produced, committed under
   tests/golden/ together with tests/golden/make_golden.py.
"""
from .instance import Instance, generate_rect, generate_pin  # noqa: F401
from .envs import SquareEnv, RectEnv, PinEnv, make_oracle_env  # noqa: F401
from .flat import flat_to_tuple, tuple_to_flat, flatten_mask  # noqa: F401
