"""CPU stand-in for Hp1Engine over the host shadow of the HP1 device code (tests/csrc/host_shadow_hp1.inc): same attributes,
same AgxHp1Config / AgxHp1Buffers structs (host pointers), same four calls.  Test infrastructure only."""
import ctypes as C

import numpy as np
import torch

from aerial_gym_simulator_b200 import _lib
from aerial_gym_simulator_b200.hp1 import build_config

from . import _shadow


class ShadowHp1Engine:
    def __init__(self, spec, num_envs, *, physics_steps=1, episode_len_steps=500, seed=0, env_id_offset=0, device_rng_reset=True,
                 strict_stale_obs=True, materialize_derived=True, per_env_params="all", debug_wrench=False, coop_reset=False):
        self.lib = _shadow.load()
        self.spec, self.N, self.M, self.coop_reset = spec, int(num_envs), spec.num_motors, bool(coop_reset)
        self.cfg = build_config(spec, self.N, physics_steps=physics_steps, episode_len_steps=episode_len_steps, seed=seed,
                                env_id_offset=env_id_offset, device_rng_reset=device_rng_reset, strict_stale_obs=strict_stale_obs)
        N, M = self.N, self.M
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt)
        self.root_state = z(N, 13)
        self.root_state[:, 6] = 1.0
        self.motor_thrust, self.sim_steps, self.target_position = z(N, M), z(N, dt=torch.int32), z(N, 3)
        self.obs, self.reward = z(N, 13), z(N)
        self.terminations, self.truncations, self.reset_mask = z(N, dt=torch.bool), z(N, dt=torch.bool), z(N, dt=torch.bool)
        self.any_reset, self.episode_count = z(32, dt=torch.int32), z(N, dt=torch.int32)
        self.bounds_min = torch.tensor(spec.bounds_lower_range[0], dtype=torch.float32).expand(N, -1).clone()
        self.bounds_max = torch.tensor(spec.bounds_upper_range[0], dtype=torch.float32).expand(N, -1).clone()
        allp = per_env_params == "all"
        nd = lambda r: allp or (np.asarray(r[0]) != np.asarray(r[1])).any()
        full = lambda v, *s: torch.full(s, float(v), dtype=torch.float32)
        self.tau_inc = full(spec.tau_inc_range[0], N, M) if nd(spec.tau_inc_range) else None
        self.tau_dec = full(spec.tau_dec_range[0], N, M) if nd(spec.tau_dec_range) else None
        self.k_thrust = full(spec.k_thrust_range[0], N, M) if (spec.use_rps and nd(spec.k_thrust_range)) else None
        gains = allp or spec.randomize_params
        gk = lambda a: torch.tensor([a[i] for i in range(3)], dtype=torch.float32).expand(N, -1).clone()
        self.K_pos, self.K_vel = (gk(self.cfg.K_pos), gk(self.cfg.K_vel)) if gains else (None, None)
        self.K_rot, self.K_angvel = (gk(self.cfg.K_rot), gk(self.cfg.K_angvel)) if gains else (None, None)
        md = materialize_derived
        self.euler, self.vehicle_orientation = (z(N, 3), z(N, 4)) if md else (None, None)
        self.vehicle_linvel, self.body_linvel, self.body_angvel = (z(N, 3), z(N, 3), z(N, 3)) if md else (None, None, None)
        self.body_wrench = z(N, 6) if debug_wrench else None
        self._buf = _lib.AgxHp1Buffers()
        for name in _lib._HP1_BUF_FIELDS:
            if name not in ("actions", "disturbance", "dist_counter", "dist_offset_", "publish_ctr"):
                t = getattr(self, name, None)
                setattr(self._buf, name, None if t is None else t.data_ptr())

    def _set_inputs(self, actions, disturbance, physics_steps, dist_counter=None, dist_offset=0):
        assert actions.dtype == torch.float32 and actions.is_contiguous() and actions.shape == (self.N, self.cfg.num_actions)
        self._keep = (actions, disturbance, dist_counter)
        self._buf.actions = actions.data_ptr()
        self._buf.disturbance = None if disturbance is None else disturbance.contiguous().data_ptr()
        # as Hp1Engine.physics_step: the in-kernel draw only when no explicit [N,6] tensor is given
        self._buf.dist_counter = None if (dist_counter is None or disturbance is not None) else dist_counter.data_ptr()
        self._buf.dist_offset = int(dist_offset)
        if physics_steps is not None:
            self.cfg.physics_steps = int(physics_steps)

    def physics_step(self, actions, disturbance=None, physics_steps=None, dist_counter=None, dist_offset=0):
        self._set_inputs(actions, disturbance, physics_steps, dist_counter, dist_offset)
        assert self.lib.shadow_hp1_physics_step(C.byref(self.cfg), C.byref(self._buf)) == 0

    def position_task_step(self, actions, disturbance=None, physics_steps=None):
        self._set_inputs(actions, disturbance, physics_steps)
        assert self.lib.shadow_hp1_position_task_step(C.byref(self.cfg), C.byref(self._buf), int(self.coop_reset)) == 0

    def reset(self, mask, draws=None):
        d = None
        if draws is not None:
            d = _lib.AgxHp1ResetDraws()
            self._draws = draws
            for k in _lib._HP1_DRAW_FIELDS:
                t = draws.get(k)
                assert t is None or (t.dtype == torch.float32 and t.is_contiguous())
                setattr(d, k, None if t is None else t.data_ptr())
        assert mask.dtype == torch.bool and mask.shape == (self.N,)
        assert self.lib.shadow_hp1_reset(C.byref(self.cfg), C.byref(self._buf), C.c_void_p(mask.data_ptr()),
                                         C.byref(d) if d is not None else None) == 0

    def refresh(self):
        assert self.lib.shadow_hp1_refresh(C.byref(self.cfg), C.byref(self._buf)) == 0
