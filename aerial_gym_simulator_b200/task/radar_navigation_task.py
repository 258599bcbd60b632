"""RadarNavigationTask (task/radar_navigation_task/radar_navigation_task.py): LiDARNavigationTask on the lmf2 with the 48 x 120 "fake
radar" point cloud in env_with_obstacles.  The reference subclasses the LiDAR task and overrides three things: the noise put on the
pooled range image (80 % of the pixels invalidated), process_image_observation / step (textually the LiDAR task's), and its own copy
of compute_reward, which differs in ONE term (x-velocity penalty on clamp(vx, max=0) instead of clamp(vx, min=0), :251) -- selected
here by AgxLidarNavRewardParams.radar_variant."""
import torch

from ..utils.math import torch_rand_float_tensor
from .lidar_navigation_task import LiDARNavigationTask


def add_noise_to_downsampled_radar_data(ds):
    """radar_navigation_task.py:7-21, same torch calls in the same order.  In place."""
    noise_mask = torch.bernoulli(0.03 * torch.ones_like(ds))
    sel = noise_mask == 1
    ds[sel] += torch_rand_float_tensor(0.2 * torch.ones_like(noise_mask[sel]), 10.0 * torch.ones_like(noise_mask[sel]))
    invalid_points_mask = torch.bernoulli(0.8 * torch.ones_like(ds))
    ds[invalid_points_mask == 1] = -1.0
    return ds


class RadarNavigationTask(LiDARNavigationTask):
    def __init__(self, task_config, seed=None, num_envs=None, headless=None, device=None, use_warp=None):
        super().__init__(task_config, seed=seed, num_envs=num_envs, headless=headless, device=device, use_warp=use_warp)
        self._params.radar_variant = 1

    def add_noise_to_downsampled_lidar_data(self, ds_lidar_data):
        return add_noise_to_downsampled_radar_data(ds_lidar_data)
