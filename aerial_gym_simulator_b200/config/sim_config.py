"""config/sim_config/{base_sim_config,base_sim_headless_config,sim_config_2ms,sim_config_4ms}.py"""


class BaseSimConfig:
    class viewer:  # kept for attribute compatibility; this build is headless (no Isaac Gym viewer)
        headless = False
        ref_env = 0
        camera_position = [-5, -5, 4]
        lookat = [0, 0, 0]
        camera_orientation_euler_deg = [0, 0, 0]
        camera_follow_type = "FOLLOW_TRANSFORM"
        width, height = 1280, 720
        max_range, min_range = 100.0, 0.1
        horizontal_fov_deg = 90
        use_collision_geometry = False
        camera_follow_transform_local_offset = [-1.0, 0.0, 0.3]
        camera_follow_position_global_offset = [-1.0, 0.0, 0.3]

    class sim:
        dt = 0.01
        substeps = 1
        gravity = [0.0, 0.0, -9.81]
        up_axis = 1
        use_gpu_pipeline = True

        class physx:  # PhysX solver knobs have no counterpart in the fused integrator; kept as data
            num_threads = 10
            solver_type = 1
            num_position_iterations = 4
            num_velocity_iterations = 1
            contact_offset = 0.002
            rest_offset = 0.001
            bounce_threshold_velocity = 0.1
            max_depenetration_velocity = 1.0
            max_gpu_contact_pairs = 2**24
            default_buffer_size_multiplier = 10
            contact_collection = 1


class BaseSimHeadlessConfig(BaseSimConfig):
    class viewer(BaseSimConfig.viewer):
        headless = True


class SimCfg2Ms(BaseSimConfig):
    """sim_config_2ms.py: a stand-alone class in the reference; only these values differ from BaseSimConfig"""
    class viewer(BaseSimConfig.viewer):
        camera_follow_transform_local_offset = [-1.0, 0.0, 0.2]
        camera_follow_position_global_offset = [-1.0, 0.0, 0.4]

    class sim(BaseSimConfig.sim):
        dt = 0.002

        class physx(BaseSimConfig.sim.physx):
            num_velocity_iterations = 2


class SimCfg4Ms(SimCfg2Ms):
    class sim(SimCfg2Ms.sim):
        dt = 0.004


class BaseSimNoGravityConfig(BaseSimConfig):
    """base_sim_no_gravity_config.py"""
    class sim(BaseSimConfig.sim):
        gravity = [0.0, 0.0, 0.0]


class CustomSimConfig(BaseSimConfig):
    """config/sim_config/custom_sim_config.py: the reference's example of a user-defined sim config (1 ms steps, gravity along +x)"""
    class sim(BaseSimConfig.sim):
        dt = 0.001
        gravity = [+1.0, 0.0, 0.0]

        class physx(BaseSimConfig.sim.physx):
            num_threads = 5
            solver_type = 1
            num_position_iterations = 10
            num_velocity_iterations = 15
            contact_offset = 0.01
            rest_offset = 0.01
            bounce_threshold_velocity = 0.5
            max_depenetration_velocity = 1.0
            max_gpu_contact_pairs = 2**20
            default_buffer_size_multiplier = 5
            contact_collection = 0
