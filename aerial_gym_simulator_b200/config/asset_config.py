"""config/asset_config/env_object_config.py -- obstacle asset parameter classes (the box-shaped
families the navigation envs ship: panels, objects, six walls)."""
import numpy as np

from . import RESOURCES_DIRECTORY

PI = np.pi
_ENV_ASSETS = f"{RESOURCES_DIRECTORY}/models/environment_assets"

THIN_SEMANTIC_ID, TREE_SEMANTIC_ID, OBJECT_SEMANTIC_ID, PANEL_SEMANTIC_ID = 1, 2, 3, 20
FRONT_WALL_SEMANTIC_ID, BACK_WALL_SEMANTIC_ID, LEFT_WALL_SEMANTIC_ID = 9, 10, 11
RIGHT_WALL_SEMANTIC_ID, BOTTOM_WALL_SEMANTIC_ID, TOP_WALL_SEMANTIC_ID = 12, 13, 14


def _ratio(pos_lo, pos_hi, eul_lo=(0, 0, 0), eul_hi=(0, 0, 0)):
    lo = list(pos_lo) + list(eul_lo) + [1.0] + [0.0] * 6
    hi = list(pos_hi) + list(eul_hi) + [1.0] + [0.0] * 6
    return lo, hi


class asset_state_params:
    num_assets = 1
    asset_folder = _ENV_ASSETS
    file = None  # None -> files are picked at random from the folder
    min_position_ratio = [0.5, 0.5, 0.5]
    max_position_ratio = [0.5, 0.5, 0.5]
    collision_mask = 1
    disable_gravity = False
    replace_cylinder_with_capsule = True   # PhysX asset options, kept for config parity (unused: no PhysX here)
    flip_visual_attachments = True
    density = 0.001
    angular_damping = 0.1
    linear_damping = 0.1
    max_angular_velocity = 100.0
    max_linear_velocity = 100.0
    armature = 0.001
    place_force_sensor = False
    force_sensor_parent_link = "base_link"
    force_sensor_transform = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0]
    collapse_fixed_joints = True
    fix_base_link = True
    specific_filepath = None
    color = None
    keep_in_env = False
    body_semantic_label = 0
    link_semantic_label = 0
    per_link_semantic = False
    semantic_masked_links = {}
    semantic_id = -1
    use_collision_mesh_instead_of_visual = False
    min_state_ratio, max_state_ratio = _ratio((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))


class BaseAssetParams:
    """config/asset_config/base_asset.py: the older base class (lighter damping and density); kept for import parity"""
    num_assets = 1
    asset_folder = _ENV_ASSETS
    file = None
    min_position_ratio, max_position_ratio = [0.5, 0.5, 0.5], [0.5, 0.5, 0.5]
    collision_mask = 1
    disable_gravity = False
    replace_cylinder_with_capsule = True
    flip_visual_attachments = True
    density = 0.000001
    angular_damping = linear_damping = 0.0001
    max_angular_velocity = max_linear_velocity = 100.0
    armature = 0.001
    collapse_fixed_joints = True
    fix_base_link = True
    color = None
    keep_in_env = False
    body_semantic_label = link_semantic_label = 0
    per_link_semantic = False
    semantic_masked_links = {}
    place_force_sensor = False
    force_sensor_parent_link = "base_link"
    force_sensor_transform = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0]
    use_collision_mesh_instead_of_visual = False


class panel_asset_params(asset_state_params):
    num_assets = 3
    asset_folder = f"{_ENV_ASSETS}/panels"
    min_position_ratio, max_position_ratio = [0.3, 0.05, 0.05], [0.85, 0.95, 0.95]
    specified_position = [-1000.0, -1000.0, -1000.0]  # unused by the reference as well
    min_euler_angles, max_euler_angles = [0.0, 0.0, -PI / 3.0], [0.0, 0.0, PI / 3.0]
    min_state_ratio, max_state_ratio = _ratio((0.3, 0.05, 0.05), (0.85, 0.95, 0.95), (0, 0, -PI / 3.0), (0, 0, PI / 3.0))
    keep_in_env = True
    semantic_id = -1  # assigned incrementally per instance
    color = [170, 66, 66]


class object_asset_params(asset_state_params):
    num_assets = 35
    asset_folder = f"{_ENV_ASSETS}/objects"
    min_state_ratio, max_state_ratio = _ratio((0.30, 0.05, 0.05), (0.85, 0.9, 0.9), (-PI, -PI, -PI), (PI, PI, PI))
    keep_in_env = False
    semantic_id = -1
    color = None  # (the colour line is commented out in env_object_config.py:313)


def _wall(fname, ratio, sem, keep_in_env=True, color=(100, 200, 210)):
    lo, hi = _ratio(ratio, ratio)
    return type(fname.replace(".urdf", ""), (asset_state_params,), dict(
        num_assets=1, asset_folder=f"{_ENV_ASSETS}/walls", file=fname, min_state_ratio=lo, max_state_ratio=hi, keep_in_env=keep_in_env,
        specific_filepath="cube.urdf",
        semantic_id=sem, color=list(color)))


left_wall = _wall("left_wall.urdf", (0.5, 1.0, 0.5), LEFT_WALL_SEMANTIC_ID)
right_wall = _wall("right_wall.urdf", (0.5, 0.0, 0.5), RIGHT_WALL_SEMANTIC_ID)
top_wall = _wall("top_wall.urdf", (0.5, 0.5, 1.0), TOP_WALL_SEMANTIC_ID)
bottom_wall = _wall("bottom_wall.urdf", (0.5, 0.5, 0.0), BOTTOM_WALL_SEMANTIC_ID, color=(100, 150, 150))
front_wall = _wall("front_wall.urdf", (1.0, 0.5, 0.5), FRONT_WALL_SEMANTIC_ID)
back_wall = _wall("back_wall.urdf", (0.0, 0.5, 0.5), BACK_WALL_SEMANTIC_ID)


# ---- config/asset_config/lidar_nav_env_config.py: the denser scene of env_with_lidar_nav_obstacles.  Same asset families;
# 15 panels + 70 objects + 6 walls = 91 boxes, wider placement ratios, and NOTHING is kept in the env when the curriculum
# level is below the asset's slot (keep_in_env = False everywhere, walls included: lidar_nav_env_config.py:115,355-592).
class lidar_nav_panel_asset_params(panel_asset_params):
    num_assets = 15
    min_position_ratio, max_position_ratio = [0.1, 0.0, 0.0], [1.0, 1.0, 1.0]
    min_state_ratio, max_state_ratio = _ratio((0.35, 0.0, 0.0), (1.0, 1.0, 1.0), (0, 0, -PI / 3.0), (0, 0, PI / 3.0))
    keep_in_env = False


class lidar_nav_object_asset_params(object_asset_params):
    num_assets = 70
    min_state_ratio, max_state_ratio = _ratio((0.30, 0.0, 0.0), (1.0, 1.0, 1.0), (-PI, -PI, -PI), (PI, PI, PI))
    keep_in_env = False


lidar_nav_left_wall = _wall("left_wall.urdf", (0.5, 1.0, 0.5), LEFT_WALL_SEMANTIC_ID, keep_in_env=False)
lidar_nav_right_wall = _wall("right_wall.urdf", (0.5, 0.0, 0.5), RIGHT_WALL_SEMANTIC_ID, keep_in_env=False)
lidar_nav_top_wall = _wall("top_wall.urdf", (0.5, 0.5, 1.0), TOP_WALL_SEMANTIC_ID, keep_in_env=False)
lidar_nav_bottom_wall = _wall("bottom_wall.urdf", (0.5, 0.5, 0.0), BOTTOM_WALL_SEMANTIC_ID, keep_in_env=False, color=(100, 150, 150))
lidar_nav_front_wall = _wall("front_wall.urdf", (1.0, 0.5, 0.5), FRONT_WALL_SEMANTIC_ID, keep_in_env=False)
lidar_nav_back_wall = _wall("back_wall.urdf", (0.0, 0.5, 0.5), BACK_WALL_SEMANTIC_ID, keep_in_env=False)


# ---- config/asset_config/dynamic_env_object_config.py: free-floating objects moved by env_actions ("dynamic_env")
class dynamic_object_asset_params(object_asset_params):
    num_assets = 40          # dynamic_env_object_config.py:274
    disable_gravity = True   # :28
    fix_base_link = False    # :41


class tree_asset_params(asset_state_params):
    """config/asset_config/env_object_config.py:225-270: a trunk + branches made of cylinders, one segmentation id per link"""
    num_assets = 1
    asset_folder = f"{_ENV_ASSETS}/trees"
    min_state_ratio, max_state_ratio = _ratio((0.1, 0.1, 0.0), (0.9, 0.9, 0.0), (0, -PI / 6.0, -PI), (0, PI / 6.0, PI))
    collapse_fixed_joints = True
    per_link_semantic = True
    keep_in_env = True
    semantic_id = -1
    color = [70, 200, 100]


class thin_asset_params(asset_state_params):
    """config/asset_config/env_object_config.py:181-222: thin rods (one box each, 1000 files); off by default (num_assets = 0)"""
    num_assets = 0
    asset_folder = f"{_ENV_ASSETS}/thin"
    min_state_ratio, max_state_ratio = _ratio((0.3, 0.05, 0.05), (0.85, 0.95, 0.95), (-PI, -PI, -PI), (PI, PI, PI))
    collapse_fixed_joints = True
    per_link_semantic = False
    semantic_id = -1
    color = [170, 66, 66]


class tile_asset_params(asset_state_params):
    """config/asset_config/env_object_config.py:123-178.  The reference ships no `tile_meshes` folder either: the class exists so that
    the env configs' asset_type_to_dict_map is complete; "tiles" is False in every include_asset_type."""
    num_assets = 1
    asset_folder = f"{_ENV_ASSETS}/tile_meshes"
    min_position_ratio, max_position_ratio = [0.3, 0.05, 0.05], [0.85, 0.95, 0.95]
    specified_position = [-1000.0, -1000.0, -1000.0]
    min_euler_angles, max_euler_angles = [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]
    min_state_ratio, max_state_ratio = _ratio((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))
    keep_in_env = True
    collapse_fixed_joints = True
    per_link_semantic = False
    semantic_id = -1


def _variant(base, prefix, **fields):
    return type(prefix + base.__name__, (base,), dict(fields))


# lidar_nav_env_config.py keeps its own tile / tree classes: as env_object_config's, but not kept when the curriculum thins the scene
lidar_nav_tile_asset_params = _variant(tile_asset_params, "lidar_nav_", keep_in_env=False)
lidar_nav_tree_asset_params = _variant(tree_asset_params, "lidar_nav_", keep_in_env=False)
# dynamic_env_object_config.py: EVERY class there floats (gravity off, base link free), and it has six trees
_FLOATING = dict(disable_gravity=True, fix_base_link=False)
dynamic_asset_state_params = _variant(asset_state_params, "dynamic_", **_FLOATING)
dynamic_panel_asset_params = _variant(panel_asset_params, "dynamic_", **_FLOATING)
dynamic_thin_asset_params = _variant(thin_asset_params, "dynamic_", **_FLOATING)
dynamic_tree_asset_params = _variant(tree_asset_params, "dynamic_", num_assets=6, **_FLOATING)
dynamic_left_wall, dynamic_right_wall, dynamic_top_wall, dynamic_bottom_wall, dynamic_front_wall, dynamic_back_wall = (
    _variant(w, "dynamic_", **_FLOATING) for w in (left_wall, right_wall, top_wall, bottom_wall, front_wall, back_wall))
