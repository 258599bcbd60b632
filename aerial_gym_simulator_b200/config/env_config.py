"""config/env_config/{empty_env,env_with_obstacles,env_with_lidar_nav_obstacles}.py"""
from . import asset_config as _ac
from .asset_config import (back_wall, bottom_wall, front_wall, left_wall, object_asset_params, panel_asset_params,
                           right_wall, top_wall)


class BaseEnvCfg:  # config/env_config/base_env_config.py
    pass


class EmptyEnvCfg:
    class env:
        num_envs = 3
        num_env_actions = 0
        env_spacing = 1.0
        num_physics_steps_per_env_step_mean = 1
        num_physics_steps_per_env_step_std = 0
        render_viewer_every_n_steps = 10
        collision_force_threshold = 0.010
        manual_camera_trigger = False
        reset_on_collision = True
        create_ground_plane = False
        sample_timestep_for_latency = True
        perturb_observations = True
        keep_same_env_for_num_episodes = 1
        write_to_sim_at_every_timestep = False
        use_warp = False
        e_s = env_spacing
        lower_bound_min = [-e_s, -e_s, -e_s]
        lower_bound_max = [-e_s, -e_s, -e_s]
        upper_bound_min = [e_s, e_s, e_s]
        upper_bound_max = [e_s, e_s, e_s]

    class env_config:
        include_asset_type = {}
        asset_type_to_dict_map = {}


class EnvWithObstaclesCfg:
    class env:
        num_envs = 64
        num_env_actions = 4
        env_spacing = 5.0
        num_physics_steps_per_env_step_mean = 10
        num_physics_steps_per_env_step_std = 0
        render_viewer_every_n_steps = 1
        reset_on_collision = True
        collision_force_threshold = 0.05
        create_ground_plane = False
        sample_timestep_for_latency = True
        perturb_observations = True
        keep_same_env_for_num_episodes = 1
        write_to_sim_at_every_timestep = False
        use_warp = True
        lower_bound_min = [-2.0, -4.0, -3.0]
        lower_bound_max = [-1.0, -2.5, -2.0]
        upper_bound_min = [9.0, 2.5, 2.0]
        upper_bound_max = [10.0, 4.0, 3.0]

    class env_config:
        include_asset_type = {
            "panels": True, "tiles": False, "thin": False, "trees": False, "objects": True,
            "left_wall": True, "right_wall": True, "back_wall": True, "front_wall": True,
            "top_wall": True, "bottom_wall": True,
        }
        asset_type_to_dict_map = {
            "panels": panel_asset_params, "thin": _ac.thin_asset_params, "trees": _ac.tree_asset_params, "objects": object_asset_params,
            "left_wall": left_wall, "right_wall": right_wall, "back_wall": back_wall,
            "front_wall": front_wall, "bottom_wall": bottom_wall, "top_wall": top_wall, "tiles": _ac.tile_asset_params,
        }


class EnvWithLidarNavObstaclesCfg(EnvWithObstaclesCfg):
    """config/env_config/env_with_lidar_nav_obstacles.py: bigger bounds, the lidar_nav asset set."""
    class env(EnvWithObstaclesCfg.env):
        lower_bound_min = [-7.50, -7.50, -5.0]
        lower_bound_max = [-5.0, -5.0, -3.0]
        upper_bound_min = [5.0, 5.0, 3.0]
        upper_bound_max = [7.5, 7.5, 5.0]

    class env_config:
        include_asset_type = dict(EnvWithObstaclesCfg.env_config.include_asset_type)
        asset_type_to_dict_map = {
            "panels": _ac.lidar_nav_panel_asset_params, "thin": _ac.thin_asset_params, "trees": _ac.lidar_nav_tree_asset_params,
            "objects": _ac.lidar_nav_object_asset_params,
            "left_wall": _ac.lidar_nav_left_wall, "right_wall": _ac.lidar_nav_right_wall, "back_wall": _ac.lidar_nav_back_wall,
            "front_wall": _ac.lidar_nav_front_wall, "bottom_wall": _ac.lidar_nav_bottom_wall, "top_wall": _ac.lidar_nav_top_wall,
            "tiles": _ac.lidar_nav_tile_asset_params,
        }


class DynamicEnvironmentCfg(EnvWithObstaclesCfg):
    """config/env_config/dynamic_environment.py: 40 free-floating objects (no panels, no walls) whose twist is set through
    env.step(actions, env_actions=[N,40,6]) (examples/dynamic_env_example.py:33-45)."""
    class env(EnvWithObstaclesCfg.env):
        num_env_actions = 6
        create_ground_plane = True
        write_to_sim_at_every_timestep = True
        lower_bound_min = [-2.0, -4.0, 0.0]
        lower_bound_max = [-1.0, -2.5, 0.0]
        upper_bound_min = [9.0, 2.5, 4.0]
        upper_bound_max = [10.0, 4.0, 5.0]

    class env_config:
        include_asset_type = {k: (k == "objects") for k in EnvWithObstaclesCfg.env_config.include_asset_type if k != "tiles"}
        asset_type_to_dict_map = {
            "panels": _ac.dynamic_panel_asset_params, "thin": _ac.dynamic_thin_asset_params, "trees": _ac.dynamic_tree_asset_params,
            "objects": _ac.dynamic_object_asset_params, "left_wall": _ac.dynamic_left_wall, "right_wall": _ac.dynamic_right_wall,
            "back_wall": _ac.dynamic_back_wall, "front_wall": _ac.dynamic_front_wall, "bottom_wall": _ac.dynamic_bottom_wall,
            "top_wall": _ac.dynamic_top_wall,
        }


class EnvCfg2Ms(EmptyEnvCfg):
    """config/env_config/env_config_2ms.py: the empty env with 5 physics steps per env step (used with the 2 ms sim config)"""
    class env(EmptyEnvCfg.env):
        num_physics_steps_per_env_step_mean = 5
        render_viewer_every_n_steps = 2


class ForestEnvCfg(EnvWithObstaclesCfg):
    """config/env_config/forest_env.py: a tree (cylinders, per-link segmentation), 35 objects and the floor in a 10 x 10 x 4 m box"""
    class env(EnvWithObstaclesCfg.env):
        collision_force_threshold = 0.005
        lower_bound_min = [-5.0, -5.0, -1.0]
        lower_bound_max = [-5.0, -5.0, -1.0]
        upper_bound_min = [5.0, 5.0, 3.0]
        upper_bound_max = [5.0, 5.0, 3.0]

    class env_config:
        include_asset_type = {"trees": True, "objects": True, "bottom_wall": True}
        asset_type_to_dict_map = {"trees": _ac.tree_asset_params, "objects": _ac.object_asset_params, "bottom_wall": _ac.bottom_wall}
