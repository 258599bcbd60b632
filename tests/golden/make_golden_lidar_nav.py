"""Golden vectors for the LiDARNavigationTask epilogue, produced by RUNNING THE REFERENCE'S OWN CODE on CPU
(this container only):

    python tests/golden/make_golden_lidar_nav.py

The module task/lidar_navigation_task/lidar_navigation_task.py cannot be imported (isaacgym, gymnasium, SimBuilder), so
its functions are pulled out of the file with ``ast`` and executed unchanged:
  * module level: exponential_reward_function, exponential_penalty_function, compute_reward (decorators dropped);
  * methods of LiDARNavigationTask, run on a stand-in ``self``: process_image_observation,
    add_noise_to_downsampled_lidar_data, compute_rewards_and_crashes, process_obs_for_task;
  * reward_parameters is read from config/task_config/lidar_navigation_task_config.py.
torch.rand_like draws of process_obs_for_task are recorded next to the outputs; the task's own lidar noise
(torch.bernoulli / masked rand) runs under torch.manual_seed and its input and output are both recorded."""
import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_loader  # noqa: E402
from make_golden_aux import _funcs_from, _Recorder, rand_unit_quat  # noqa: E402  (also installs the reference loader)

from aerial_gym.utils import math as ref_math  # noqa: E402

RADAR_PATH = os.path.join(_ref_loader.REF_ROOT, "aerial_gym/task/radar_navigation_task/radar_navigation_task.py")
TASK_PATH = os.path.join(_ref_loader.REF_ROOT, "aerial_gym/task/lidar_navigation_task/lidar_navigation_task.py")
CFG_PATH = os.path.join(_ref_loader.REF_ROOT, "aerial_gym/config/task_config/lidar_navigation_task_config.py")
NOISE_SEED = 1234


def reward_parameters():
    tree = ast.parse(open(CFG_PATH).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "task_config")
    node = next(n for n in cls.body if isinstance(n, ast.Assign) and n.targets[0].id == "reward_parameters")
    return ast.literal_eval(node.value)


def _namespace():
    ns = {"torch": torch}
    for k in dir(ref_math):
        if not k.startswith("_"):
            ns[k] = getattr(ref_math, k)
    _funcs_from(TASK_PATH, {"exponential_reward_function", "exponential_penalty_function", "compute_reward"}, ns)
    ns["erf"], ns["epf"] = ns["exponential_reward_function"], ns["exponential_penalty_function"]  # module-level aliases, :518-519
    _funcs_from(TASK_PATH, {"process_image_observation", "add_noise_to_downsampled_lidar_data", "compute_rewards_and_crashes",
                            "process_obs_for_task"}, ns, in_class="LiDARNavigationTask")
    return ns


def gen_pool(ns, out, seed=31, n=4, H=48, W=120):
    g = torch.Generator().manual_seed(seed)
    pos = torch.randn(n, 3, generator=g) * 2.0
    linvel = torch.randn(n, 3, generator=g) * 1.5
    linvel[1] = 0.0  # a hovering robot: every pixel takes the default time to collision
    d = torch.randn(n, H, W, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    rng = torch.rand(n, H, W, 1, generator=g) * 13.0 + 0.05           # 0.05 .. 13.05 m: both clipping branches are hit
    rng[2, :, :, :] = torch.rand(H, W, 1, generator=g) * 0.8 + 0.25   # an env with something very close everywhere
    far = torch.rand(H, W, 1, generator=g) * 3.0 + 10.5               # open space (every return beyond max range) ...
    far[10:30, 20:80] = torch.rand(20, 60, 1, generator=g) * 8.0 + 1.0  # ... except one obstacle; windows straddle its edge
    rng[3] = far
    pc = pos.view(n, 1, 1, 3) + d * rng
    recorded = {}

    def noisy(me_, x):
        recorded["pre"] = x.clone()
        torch.manual_seed(NOISE_SEED)
        y = ns["add_noise_to_downsampled_lidar_data"](me_, x)
        recorded["post"] = y.clone()
        return y

    me = types.SimpleNamespace(
        obs_dict={"depth_range_pixels": pc.unsqueeze(1).clone(), "robot_position": pos, "robot_linvel": linvel},
        world_dir_vectors=torch.ones(n, H, W, 3), num_envs=n, device="cpu", time_to_collision=torch.zeros(n),
        downsampled_lidar_data=torch.zeros(n, (H // 3) * (W // 6)))
    me.add_noise_to_downsampled_lidar_data = lambda x: noisy(me, x)
    ns["process_image_observation"](me)
    out.update(pool_pointcloud=pc.numpy(), pool_robot_position=pos.numpy(), pool_robot_linvel=linvel.numpy(),
               pool_image_ds=recorded["pre"].numpy(), pool_image_noisy=recorded["post"].numpy(),
               pool_time_to_collision=me.time_to_collision.numpy().copy(), pool_downsampled_lidar_data=me.downsampled_lidar_data.numpy().copy(),
               pool_noise_seed=np.int64(NOISE_SEED))


def gen_reward_obs(ns, out, seed=41, n=96):
    g = torch.Generator().manual_seed(seed)
    params = reward_parameters()
    pt = {k: torch.tensor(v) for k, v in params.items()}
    pos = torch.randn(n, 3, generator=g) * 3.0
    target = torch.randn(n, 3, generator=g) * 3.0
    target[:8] = pos[:8] + 0.3 * torch.randn(8, 3, generator=g)  # dist < 1: the stable-at-goal branch
    q = rand_unit_quat(n, g)
    veh_q = ref_math.vehicle_frame_quat_from_quat(q)
    euler = torch.rand(n, 3, generator=g) * 2 * np.pi              # get_euler_xyz_tensor range [0, 2 pi)
    target_yaw = (torch.rand(n, generator=g) * 2 - 1) * np.pi
    target_yaw[:4] = ref_math.ssa(euler[:4, 2])                    # aligned with the target yaw
    veh_linvel = torch.randn(n, 3, generator=g) * 1.5
    veh_linvel[8:12] *= 3.0                                        # > 3 m/s: velocity magnitude penalty
    veh_linvel[12] = 0.0
    body_angvel = torch.randn(n, 3, generator=g)
    blv = torch.randn(n, 3, generator=g)
    prev_err = torch.randn(n, 3, generator=g) * 3.0
    crashes = torch.zeros(n, dtype=torch.bool)
    crashes[::11] = True
    act = torch.rand(n, 4, generator=g) * 2 - 1
    prev_act = act + 0.3 * torch.randn(n, 4, generator=g)
    robot_actions = torch.rand(n, 4, generator=g) * 2 - 1
    ttc = torch.rand(n, generator=g) * 10.0
    ttc[:6] = torch.tensor([0.0, 0.05, 0.3, 1.0, 10.0, 0.6])
    lidar = torch.rand(n, 320, generator=g) * 5.0
    rns = dict(ns)  # RadarNavigationTask: its own module-level compute_reward (one term differs) + its noise method
    _funcs_from(RADAR_PATH, {"exponential_reward_function", "exponential_penalty_function", "compute_reward"}, rns)
    rns["erf"], rns["epf"] = rns["exponential_reward_function"], rns["exponential_penalty_function"]
    _funcs_from(RADAR_PATH, {"compute_rewards_and_crashes", "add_noise_to_downsampled_lidar_data"}, rns, in_class="RadarNavigationTask")
    for tag, frac in (("c0", 0.0), ("c1", 0.4444444477558136)):
        mr = types.SimpleNamespace(
            obs_dict={"robot_position": pos, "robot_vehicle_orientation": veh_q, "robot_orientation": q, "robot_euler_angles": euler,
                      "robot_vehicle_linvel": veh_linvel, "robot_body_angvel": body_angvel, "crashes": crashes.clone()},
            target_position=target, device="cpu", pos_error_vehicle_frame_prev=torch.zeros(n, 3),
            pos_error_vehicle_frame=prev_err.clone(), target_yaw=target_yaw, current_action=act, prev_action=prev_act,
            time_to_collision=ttc, curriculum_progress_fraction=frac, task_config=types.SimpleNamespace(reward_parameters=pt))
        out[f"radar_reward_{tag}"] = rns["compute_rewards_and_crashes"](mr, mr.obs_dict)[0].numpy().copy()
    ds = torch.rand(6, 16, 20, generator=g) * 9.0 + 0.3
    torch.manual_seed(NOISE_SEED)
    out["radar_noise_in"] = ds.numpy().copy()
    out["radar_noise_out"] = rns["add_noise_to_downsampled_lidar_data"](types.SimpleNamespace(device="cpu"), ds.clone()).numpy()
    for tag, frac in (("c0", 0.0), ("c1", 0.4444444477558136)):
        me = types.SimpleNamespace(
            obs_dict={"robot_position": pos, "robot_vehicle_orientation": veh_q, "robot_orientation": q, "robot_euler_angles": euler,
                      "robot_vehicle_linvel": veh_linvel, "robot_body_angvel": body_angvel, "crashes": crashes.clone()},
            target_position=target, device="cpu", pos_error_vehicle_frame_prev=torch.zeros(n, 3),
            pos_error_vehicle_frame=prev_err.clone(), target_yaw=target_yaw, current_action=act, prev_action=prev_act,
            time_to_collision=ttc, curriculum_progress_fraction=frac, task_config=types.SimpleNamespace(reward_parameters=pt))
        rew, cr = ns["compute_rewards_and_crashes"](me, me.obs_dict)
        out[f"reward_{tag}"] = rew.numpy().copy()
        out[f"frac_{tag}"] = np.float32(frac)
        assert torch.equal(me.pos_error_vehicle_frame_prev, prev_err)
    out.update(pos=pos.numpy(), target=target.numpy(), vehicle_orientation=veh_q.numpy(), euler=euler.numpy(), target_yaw=target_yaw.numpy(),
               vehicle_linvel=veh_linvel.numpy(), body_angvel=body_angvel.numpy(), body_linvel=blv.numpy(), prev_pos_error=prev_err.numpy(),
               crashes=crashes.numpy(), actions=act.numpy(), prev_actions=prev_act.numpy(), robot_actions=robot_actions.numpy(),
               time_to_collision=ttc.numpy(), pos_error=me.pos_error_vehicle_frame.numpy().copy(),  # (lidar_obs = obs[:, 17:])
               param_names=np.array(list(params.keys())), param_values=np.array(list(params.values()), dtype=np.float64))

    # ---- process_obs_for_task with recorded rand_like draws ------------------------------------------
    rec = _Recorder(torch.Generator().manual_seed(seed + 1))
    fake_torch = types.SimpleNamespace(**{k: getattr(torch, k) for k in ("norm", "zeros", "zeros_like")})
    fake_torch.rand_like = rec.rand_like
    ns2 = dict(ns)
    ns2["torch"] = fake_torch
    _funcs_from(TASK_PATH, {"process_obs_for_task"}, ns2, in_class="LiDARNavigationTask")
    obs = torch.full((n, 17 + 320), 7.0)
    me = types.SimpleNamespace(
        obs_dict={"robot_vehicle_orientation": veh_q, "robot_position": pos, "robot_euler_angles": euler, "robot_body_linvel": blv,
                  "robot_body_angvel": body_angvel, "robot_actions": robot_actions},
        target_position=target, target_yaw=target_yaw, task_obs={"observations": obs}, downsampled_lidar_data=lidar)
    ns2["process_obs_for_task"](me)
    out.update(obs=obs.numpy(), obs_draw_vec=rec.draws[0].numpy(), obs_draw_euler=rec.draws[1].numpy())


if __name__ == "__main__":
    ns = _namespace()
    out = {}
    gen_pool(ns, out)
    gen_reward_obs(ns, out)
    np.savez_compressed(os.path.join(HERE, "lidar_nav_task_epilogue.npz"), **out)
    print("wrote lidar_nav_task_epilogue.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim > 0})
