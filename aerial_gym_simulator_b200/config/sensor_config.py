"""config/sensor_config/{base_sensor_config, camera_config/base_depth_camera_config,
lidar_config/{base_lidar_config,osdome_64_config}}.py"""


class BaseSensorConfig:
    num_sensors = 1
    randomize_placement = False
    min_translation = [0.07, -0.06, 0.01]
    max_translation = [0.12, 0.03, 0.04]
    min_euler_rotation_deg = [-5.0, -5.0, -5.0]
    max_euler_rotation_deg = [5.0, 5.0, 5.0]


class BaseDepthCameraConfig(BaseSensorConfig):
    num_sensors = 1
    sensor_type = "camera"
    height, width = 135, 240
    horizontal_fov_deg = 87.000
    max_range, min_range = 10.0, 0.2
    calculate_depth = True  # depth image; False -> range image
    return_pointcloud = False
    pointcloud_in_world_frame = False
    segmentation_camera = True
    euler_frame_rot_deg = [-90.0, 0, -90.0]  # optical axis (+z) onto robot +x
    normalize_range = True
    far_out_of_range_value = max_range if normalize_range else -1.0
    near_out_of_range_value = -max_range if normalize_range else -1.0
    randomize_placement = True
    nominal_position = [0.10, 0.0, 0.03]
    nominal_orientation_euler_deg = [0.0, 0.0, 0.0]
    use_collision_geometry = False

    class sensor_noise:
        enable_sensor_noise = False
        pixel_dropout_prob = 0.01
        pixel_std_dev_multiplier = 0.01


class BaseLidarConfig(BaseSensorConfig):
    num_sensors = 1
    sensor_type = "lidar"
    height, width = 128, 512
    horizontal_fov_deg_min, horizontal_fov_deg_max = -180, 180
    vertical_fov_deg_min, vertical_fov_deg_max = -45, +45
    max_range, min_range = 10.0, 0.2
    return_pointcloud = False
    pointcloud_in_world_frame = False
    segmentation_camera = True
    euler_frame_rot_deg = [0.0, 0.0, 0.0]
    normalize_range = True
    far_out_of_range_value = max_range if normalize_range else -1.0
    near_out_of_range_value = -max_range if normalize_range else -1.0
    randomize_placement = True
    nominal_position = [0.10, 0.0, 0.03]
    nominal_orientation_euler_deg = [0.0, 0.0, 0.0]

    class sensor_noise:
        enable_sensor_noise = True
        std_a = 0.00001
        std_b = 0.00001
        std_c = 0.00001
        mean_offset = -0.05
        pixel_dropout_prob = 0.0


class OSDome_64_Config(BaseLidarConfig):
    height, width = 64, 512
    vertical_fov_deg_min, vertical_fov_deg_max = 0, 90
    max_range, min_range = 20.0, 0.5
    # quirk kept: far/near_out_of_range_value are inherited from BaseLidarConfig (computed with ITS
    # max_range = 10), so an out-of-range return normalises to 10/20 = 0.5 (osdome_64_config.py:4-13)
    randomize_placement = False
    min_translation = [0.0, 0.0, 0.0]
    max_translation = [0.0, 0.0, 0.0]
    min_euler_rotation_deg = [0.0, 0.0, 0.0]
    max_euler_rotation_deg = [0.0, 0.0, 0.0]

    class sensor_noise:
        enable_sensor_noise = False
        std_a = 0.00038089
        std_b = -0.00343351
        std_c = 0.01553284
        mean_offset = -0.025
        pixel_dropout_prob = 0.0


class RSLidar_Airy_Config(BaseLidarConfig):
    """config/sensor_config/lidar_config/rslidar_airy_config.py: 48 x 120 hemispherical LiDAR that returns a WORLD-frame
    point cloud (what LiDARNavigationTask.process_image_observation consumes)."""
    height, width = 48, 120
    vertical_fov_deg_min, vertical_fov_deg_max = 0, 90
    max_range, min_range = 10.0, 0.2
    return_pointcloud = True
    segmentation_camera = False
    normalize_range = False
    pointcloud_in_world_frame = True
    randomize_placement = True  # a degenerate range: the mount is fixed, but the draws are still made
    min_translation = [-0.05, 0.0, 0.0]
    max_translation = [-0.05, 0.0, 0.0]
    min_euler_rotation_deg = [0.0, -90.0, 0.0]
    max_euler_rotation_deg = [0.0, -90.0, 0.0]

    class sensor_noise:
        enable_sensor_noise = False
        std_a = 0.00038089
        std_b = -0.00343351
        std_c = 0.01553284
        mean_offset = -0.025
        pixel_dropout_prob = 0.0


class DepthCamera64x48Config(BaseDepthCameraConfig):
    """BASELINE.json north-star sensor: 64x48 depth + segmentation."""
    height, width = 48, 64


class BaseImuConfig:
    """config/sensor_config/imu_config/base_imu_config.py (values of a sample VN100)."""
    num_sensors = 1
    sensor_type = "imu"
    world_frame = False
    enable_noise = True
    enable_bias = True
    bias_std = [9.782812831313576e-07] * 3 + [2.6541629581345176e-05] * 3          # accel x3, gyro x3
    imu_noise_std = [0.001688956233495657] * 3 + [0.0010679343003532472] * 3
    max_measurement_value = [100.0, 100.0, 100.0, 10.0, 10.0, 10.0]
    max_bias_init_value = [1.0e-03] * 6
    gravity_compensation = False
    randomize_placement = False
    min_euler_rotation_deg = [-2.0, -2.0, -2.0]
    max_euler_rotation_deg = [2.0, 2.0, 2.0]
