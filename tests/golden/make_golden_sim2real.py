"""Golden vectors for the two setpoint-command sim2real position tasks, produced by RUNNING THE REFERENCE'S OWN CODE on CPU (this
container only):   python tests/golden/make_golden_sim2real.py

compute_reward + its helpers (module level) and compute_rewards_and_crashes / process_obs_for_task (methods, run on a stand-in
``self``) are pulled out of the two task files with ``ast`` and executed unchanged; torch.randn_like draws are recorded."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_loader  # noqa: E402
from make_golden_aux import _funcs_from, rand_unit_quat  # noqa: E402

from aerial_gym.utils import math as ref_math  # noqa: E402

TASKS = {"vel": ("aerial_gym/task/position_setpoint_task_sim2real/position_setpoint_task_sim2real.py", "PositionSetpointTaskSim2Real"),
         "acc": ("aerial_gym/task/position_setpoint_task_acceleration_sim2real/position_setpoint_task_acceleration_sim2real.py",
                 "PositionSetpointTaskAccelerationSim2Real")}


def main(n=96, seed=71):
    g = torch.Generator().manual_seed(seed)
    pos = torch.randn(n, 3, generator=g) * 1.5
    pos[:5] *= 8.0  # beyond the 10 m crash radius
    target = torch.zeros(n, 3)
    q = rand_unit_quat(n, g)
    q[::2] *= -1.0  # both signs of qw
    veh_q = ref_math.vehicle_frame_quat_from_quat(rand_unit_quat(n, g))  # stale: independent of q
    blv, bav = torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g)
    prev_dist = (target - pos).norm(dim=1) + 0.02 * torch.randn(n, generator=g)
    act, prev_act = torch.rand(n, 4, generator=g) * 2 - 1, torch.rand(n, 4, generator=g) * 2 - 1
    robot_actions = torch.rand(n, 4, generator=g) * 2 - 1
    crashes = torch.zeros(n, dtype=torch.bool)
    crashes[::19] = True
    out = dict(pos=pos.numpy(), quat=q.numpy(), vehicle_orientation=veh_q.numpy(), body_linvel=blv.numpy(), body_angvel=bav.numpy(),
               prev_dist=prev_dist.numpy(), actions=act.numpy(), prev_actions=prev_act.numpy(), robot_actions=robot_actions.numpy(),
               crashes_in=crashes.numpy())
    for tag, (rel, cls) in TASKS.items():
        path = os.path.join(_ref_loader.REF_ROOT, rel)
        ns = {"torch": torch}
        for k in dir(ref_math):
            if not k.startswith("_"):
                ns[k] = getattr(ref_math, k)
        _funcs_from(path, {"exp_func", "abs_exp_func", "exp_penalty_func", "abs_exp_penalty_func", "compute_reward"}, ns)
        _funcs_from(path, {"compute_rewards_and_crashes", "process_obs_for_task"}, ns, in_class=cls)
        od = {"robot_position": pos, "robot_body_linvel": blv, "robot_vehicle_orientation": veh_q, "robot_orientation": q.clone(),
              "robot_body_angvel": bav, "crashes": crashes.clone(), "robot_actions": robot_actions}
        me = types.SimpleNamespace(obs_dict=od, target_position=target, device="cpu", prev_dist=prev_dist, actions=act.clone(),
                                   prev_actions=prev_act.clone(), actions_vehicle_frame=torch.zeros(n, 4),
                                   prev_actions_vehicle_frame=prev_act.clone(), task_config=types.SimpleNamespace(reward_parameters={}))
        rew, cr = ns["compute_rewards_and_crashes"](me, od)
        out[f"{tag}_reward"], out[f"{tag}_crashes_out"] = rew.numpy().copy(), cr.numpy().copy()
        if tag == "acc":
            out["acc_actions_vehicle_frame"] = me.actions_vehicle_frame.numpy().copy()
        draws, gg = [], torch.Generator().manual_seed(seed + 3)

        def randn_like(t):
            d = torch.randn(t.shape, generator=gg)
            draws.append(d.clone())
            return d
        fake = types.SimpleNamespace(randn_like=randn_like, sign=torch.sign)
        ns2 = dict(ns, torch=fake)
        _funcs_from(path, {"process_obs_for_task"}, ns2, in_class=cls)
        obs = torch.full((n, 17), 7.0)
        od2 = dict(od, robot_orientation=q.clone())
        me2 = types.SimpleNamespace(obs_dict=od2, target_position=target, task_obs={"observations": obs}, rewards=rew, terminations=cr, truncations=cr)
        ns2["process_obs_for_task"](me2)
        assert [tuple(d.shape) for d in draws] == [(n, 3)] * 4
        out.update({f"{tag}_obs": obs.numpy(), f"{tag}_noise": torch.cat(draws, dim=1).numpy(), f"{tag}_quat_after": od2["robot_orientation"].numpy()})
    np.savez_compressed(os.path.join(HERE, "sim2real_task_epilogue.npz"), **out)
    print("wrote sim2real_task_epilogue.npz")


if __name__ == "__main__":
    main()
