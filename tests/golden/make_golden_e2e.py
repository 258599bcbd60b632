"""Golden vectors for the two motor-command position tasks (sim2real_end_to_end, sim2real_px4), produced by RUNNING THE
REFERENCE'S OWN CODE on CPU (this container only):

    python tests/golden/make_golden_e2e.py

compute_reward / exp_func / exp_penalty_func (module level) and process_obs_for_task (method, run on a stand-in ``self``) are pulled
out of the two task files with ``ast`` and executed unchanged.  process_obs_for_task needs four pytorch3d functions; pytorch3d is not
installable here, so they come from oracle/e2e_task_oracle.py (restatements of pytorch3d's published algorithms) -- the obs fixture
therefore pins the TASK's arithmetic around those calls, not pytorch3d itself.  torch.normal draws are recorded."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _ref_loader  # noqa: E402
from make_golden_aux import _funcs_from, rand_unit_quat  # noqa: E402

from aerial_gym.utils import math as ref_math  # noqa: E402
from oracle import e2e_task_oracle as P3D  # noqa: E402  (pytorch3d restatements only)

TASKS = {"end_to_end": "aerial_gym/task/position_setpoint_task_sim2real_end_to_end/position_setpoint_task_sim2real_end_to_end.py",
         "px4": "aerial_gym/task/position_setpoint_task_sim2real_px4/position_setpoint_task_sim2real_px4.py"}
CRASH = {"end_to_end": 1.5, "px4": 6.5}


def main(n=96, seed=61):
    out = {}
    g = torch.Generator().manual_seed(seed)
    pos = torch.randn(n, 3, generator=g) * 0.6
    pos[:6] *= 6.0  # beyond both crash distances
    target = torch.zeros(n, 3)
    q = rand_unit_quat(n, g)
    q[:30] = ref_math.quat_from_euler_xyz(torch.randn(30, generator=g) * 0.2, torch.randn(30, generator=g) * 0.2, torch.randn(30, generator=g) * 0.3)
    linvel, body_angvel = torch.randn(n, 3, generator=g) * 0.8, torch.randn(n, 3, generator=g) * 0.8
    prev_pos_err = (target - pos) + 0.05 * torch.randn(n, 3, generator=g)
    crashes = torch.zeros(n, dtype=torch.bool)
    crashes[::17] = True
    out.update(pos=pos.numpy(), quat=q.numpy(), linvel=linvel.numpy(), body_angvel=body_angvel.numpy(), prev_pos_error=prev_pos_err.numpy(),
               crashes_in=crashes.numpy())
    for tag, rel in TASKS.items():
        path = os.path.join(_ref_loader.REF_ROOT, rel)
        ns = {"torch": torch}
        for k in dir(ref_math):
            if not k.startswith("_"):
                ns[k] = getattr(ref_math, k)
        _funcs_from(path, {"exp_func", "exp_penalty_func", "compute_reward"}, ns)
        hover = {"end_to_end": 9.81 * 0.372 / 4, "px4": 9.81 * 1.6559999883174896 / 4}[tag]
        act = hover + 0.3 * torch.randn(n, 4, generator=g)
        prev_act = act + 0.1 * torch.randn(n, 4, generator=g)
        cr = crashes.clone()
        rew, cr_out = ns["compute_reward"](target - pos, q, linvel, body_angvel, cr, act.clone(), prev_act, prev_pos_err, CRASH[tag])
        out.update({f"{tag}_actions": act.numpy(), f"{tag}_prev_actions": prev_act.numpy(), f"{tag}_reward": rew.numpy(),
                    f"{tag}_crashes_out": cr_out.numpy(), f"{tag}_crash_dist": np.float32(CRASH[tag])})
        # ---- process_obs_for_task on a stand-in self, torch.normal recorded -------------------------------------------------------
        draws = []
        gg = torch.Generator().manual_seed(seed + 7)

        def normal(mean, std):
            d = torch.randn(mean.shape, generator=gg) * std + mean
            draws.append(d.clone())
            return d
        fake_torch = types.SimpleNamespace(normal=normal, zeros_like=torch.zeros_like, pi=torch.pi)
        ns2 = dict(ns, torch=fake_torch, euler_angles_to_matrix=P3D.euler_angles_to_matrix, matrix_to_rotation_6d=P3D.matrix_to_rotation_6d,
                   quaternion_to_matrix=P3D.quaternion_to_matrix, matrix_to_euler_angles=P3D.matrix_to_euler_angles)
        cls = "PositionSetpointTaskSim2RealEndToEnd" if tag == "end_to_end" else "PositionSetpointTaskSim2RealPX4"
        _funcs_from(path, {"process_obs_for_task"}, ns2, in_class=cls)
        obs = torch.full((n, 15), 7.0)
        me = types.SimpleNamespace(obs_dict={"robot_position": pos, "robot_orientation": q, "robot_linvel": linvel, "robot_body_angvel": body_angvel},
                                   target_position=target, task_obs={"observations": obs}, rewards=rew, terminations=cr_out, truncations=cr_out)
        ns2["process_obs_for_task"](me)
        assert len(draws) == 4 and [tuple(d.shape) for d in draws] == [(n, 3)] * 4
        out.update({f"{tag}_obs": obs.numpy(), f"{tag}_noise": torch.cat(draws, dim=1).numpy()})
    np.savez_compressed(os.path.join(HERE, "e2e_task_epilogue.npz"), **out)
    print("wrote e2e_task_epilogue.npz")


if __name__ == "__main__":
    main()
