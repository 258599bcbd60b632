"""Golden vectors for the SURVEY 8(f) rows built in round 1 -- navigation-task epilogue (f3) and IMU (f4)
-- produced by RUNNING THE REFERENCE'S OWN CODE on CPU (this container only):

    python tests/golden/make_golden_aux.py

* navigation_task.compute_reward (+ the two exponential helpers) and the body of
  NavigationTask.process_obs_for_task are pulled out of the reference file with ``ast`` (importing the
  module would drag in isaacgym / gymnasium / the VAE weights) and executed unchanged;
* sensors/imu_sensor.py imports only torch + utils.math: the reference IMUSensor class itself is
  instantiated and stepped.
Random draws are recorded next to the outputs (torch.rand_like / torch.randn are wrapped), so the
oracle and the kernel can be fed the very same numbers."""
import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_loader  # noqa: E402

_ref_loader.install()
from aerial_gym.utils import math as ref_math  # noqa: E402

NAV_PATH = os.path.join(_ref_loader.REF_ROOT, "aerial_gym/task/navigation_task/navigation_task.py")
NAV_CFG_PATH = os.path.join(_ref_loader.REF_ROOT, "aerial_gym/config/task_config/navigation_task_config.py")


def _funcs_from(path, wanted, ns, in_class=None):
    tree = ast.parse(open(path).read())
    body = tree.body
    if in_class:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == in_class).body
    for node in body:
        if isinstance(node, ast.FunctionDef) and node.name in wanted:
            node.decorator_list = []
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns


def rand_unit_quat(n, g):
    q = torch.randn(n, 4, generator=g)
    return q / q.norm(dim=-1, keepdim=True)


class _Recorder:
    """wraps torch.rand_like / torch.randn so the draws are recorded in call order"""

    def __init__(self, g):
        self.g, self.draws = g, []
        self._randn, self._rand = torch.randn, torch.rand  # the real functions (torch.randn is patched below)

    def rand_like(self, t):
        d = self._rand(t.shape, generator=self.g)
        self.draws.append(d.clone())
        return d

    def randn(self, shape, device=None):
        d = self._randn(shape, generator=self.g)
        self.draws.append(d.clone())
        return d


def reward_parameters():
    ns = {"torch": torch, "AERIAL_GYM_DIRECTORY": "/nonexistent"}
    tree = ast.parse(open(NAV_CFG_PATH).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "task_config")
    node = next(n for n in cls.body if isinstance(n, ast.Assign) and n.targets[0].id == "reward_parameters")
    return ast.literal_eval(ast.unparse(node.value).replace("1.0 / 3.5", repr(1.0 / 3.5)))


def gen_nav(seed=11, n=96):
    g = torch.Generator().manual_seed(seed)
    ns = {"torch": torch}
    for k in dir(ref_math):
        if not k.startswith("_"):
            ns[k] = getattr(ref_math, k)
    _funcs_from(NAV_PATH, {"exponential_reward_function", "exponential_penalty_function", "compute_reward"}, ns)
    params = reward_parameters()
    pt = {k: torch.tensor(v) for k, v in params.items()}
    pos = torch.randn(n, 3, generator=g) * 3.0
    target = torch.randn(n, 3, generator=g) * 3.0
    target[:5] = pos[:5] + 0.05 * torch.randn(5, 3, generator=g)  # very close to the goal
    veh_q = ref_math.vehicle_frame_quat_from_quat(rand_unit_quat(n, g))
    prev_err = torch.randn(n, 3, generator=g) * 3.0
    crashes = torch.zeros(n, dtype=torch.bool)
    crashes[::13] = True
    act = torch.rand(n, 4, generator=g) * 2 - 1
    prev_act = act + 0.3 * torch.randn(n, 4, generator=g)
    out = {}
    for tag, frac in (("c0", 0.0), ("c1", 0.4285714328289032)):
        err = ref_math.quat_rotate_inverse(veh_q, target - pos)
        rew, cr = ns["compute_reward"](err, prev_err, crashes.clone(), act, prev_act, frac, pt)
        out[f"reward_{tag}"] = rew.numpy()
        out[f"frac_{tag}"] = np.float32(frac)
    out.update(pos=pos.numpy(), target=target.numpy(), vehicle_orientation=veh_q.numpy(), prev_pos_error=prev_err.numpy(),
               crashes=crashes.numpy(), actions=act.numpy(), prev_actions=prev_act.numpy(), pos_error=err.numpy(),
               param_names=np.array(list(params.keys())), param_values=np.array(list(params.values()), dtype=np.float64))

    # ---- process_obs_for_task, executed with a stand-in `self` -------------------------------------
    rec = _Recorder(torch.Generator().manual_seed(seed + 1))
    fake_torch = types.SimpleNamespace(**{k: getattr(torch, k) for k in ("norm", "zeros", "zeros_like")})
    fake_torch.rand_like = rec.rand_like
    ns2 = dict(ns)
    ns2["torch"] = fake_torch
    _funcs_from(NAV_PATH, {"process_obs_for_task"}, ns2, in_class="NavigationTask")
    euler = (torch.rand(n, 3, generator=g) * 2 * np.pi)  # get_euler_xyz_tensor range [0, 2pi)
    blv, bav = torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g)
    robot_actions = torch.rand(n, 4, generator=g) * 2 - 1
    obs = torch.full((n, 17 + 64), 7.0)
    me = types.SimpleNamespace(
        obs_dict={"robot_vehicle_orientation": veh_q, "robot_position": pos, "robot_euler_angles": euler, "robot_body_linvel": blv,
                  "robot_body_angvel": bav, "robot_actions": robot_actions},
        target_position=target, task_obs={"observations": obs},
        task_config=types.SimpleNamespace(vae_config=types.SimpleNamespace(use_vae=False)), image_latents=None)
    ns2["process_obs_for_task"](me)
    out.update(euler=euler.numpy(), body_linvel=blv.numpy(), body_angvel=bav.numpy(), robot_actions=robot_actions.numpy(),
               obs=obs.numpy(), obs_draw_vec=rec.draws[0].numpy(), obs_draw_euler=rec.draws[1].numpy())
    np.savez_compressed(os.path.join(HERE, "nav_task_epilogue.npz"), **out)
    print("wrote nav_task_epilogue.npz")


def gen_imu(seed=23, n=64):
    from aerial_gym.sensors import imu_sensor as ref_imu  # torch + utils.math only
    from aerial_gym.config.sensor_config.imu_config.base_imu_config import BaseImuConfig

    out = {}
    for tag, world_frame, gcomp in (("body", False, False), ("world", True, False), ("gcomp", False, True)):
        class Cfg(BaseImuConfig):
            pass
        Cfg.world_frame, Cfg.gravity_compensation = world_frame, gcomp
        g = torch.Generator().manual_seed(seed)
        gtd = {
            "robot_position": torch.randn(n, 3, generator=g), "robot_orientation": rand_unit_quat(n, g),
            "gravity": torch.tensor([0.0, 0.0, -9.81]).expand(n, -1), "dt": 0.01, "robot_mass": torch.rand(n, generator=g) + 0.5,
            "robot_linvel": torch.randn(n, 3, generator=g), "robot_angvel": torch.randn(n, 3, generator=g),
            "robot_body_angvel": torch.randn(n, 3, generator=g), "robot_body_linvel": torch.randn(n, 3, generator=g),
            "robot_euler_angles": torch.randn(n, 3, generator=g), "force_sensor_tensor": torch.randn(n, 6, generator=g) * 4.0,
        }
        gtd["force_sensor_tensor"][:3, 0:3] *= 60.0  # drive a few measurements into the clamp
        rec = _Recorder(torch.Generator().manual_seed(seed + 5))
        real_randn, real_rand_like = torch.randn, torch.rand_like
        torch.manual_seed(seed + 9)  # sensor_quats / reset use the global generator (values are recorded below)
        sensor = ref_imu.IMUSensor(Cfg, n, "cpu")
        sensor.init_tensors(gtd)
        sensor.reset()
        out[f"{tag}_sensor_quats"] = sensor.sensor_quats.numpy().copy()
        out[f"{tag}_bias0"] = sensor.bias.numpy().copy()
        torch.randn = lambda shape, device=None: rec.randn(shape)
        try:
            meas = []
            for _ in range(3):
                sensor.update()
                meas.append(sensor.imu_meas.numpy().copy())
        finally:
            torch.randn, torch.rand_like = real_randn, real_rand_like
        out[f"{tag}_meas"] = np.stack(meas)
        out[f"{tag}_draws"] = np.stack([d.numpy() for d in rec.draws])  # noise0, bias0, noise1, bias1, ...
        out[f"{tag}_bias_end"] = sensor.bias.numpy().copy()
        for k in ("robot_orientation", "robot_mass", "robot_body_angvel", "force_sensor_tensor"):
            out[f"{tag}_{k}"] = gtd[k].numpy()
        out[f"{tag}_cfg"] = np.array([float(world_frame), float(gcomp)])
    out["bias_std"] = np.array(BaseImuConfig.bias_std)
    out["imu_noise_std"] = np.array(BaseImuConfig.imu_noise_std)
    out["max_measurement_value"] = np.array(BaseImuConfig.max_measurement_value)
    np.savez_compressed(os.path.join(HERE, "imu_sensor.npz"), **out)
    print("wrote imu_sensor.npz")


if __name__ == "__main__":
    gen_nav()
    gen_imu()
