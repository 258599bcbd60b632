"""config/controller_config/{lee_controller_config,lee_controller_config_octarotor,
fully_actuated_controller_rov,lmf2_controller_config,magpie_controller_config,no_control_config}.py"""
import numpy as np


class lee_controller_config:
    num_actions = 4
    max_inclination_angle_rad = np.pi / 3.0
    max_yaw_rate = np.pi / 3.0
    K_pos_tensor_max, K_pos_tensor_min = [3.0, 3.0, 2.0], [2.0, 2.0, 1.0]
    K_vel_tensor_max, K_vel_tensor_min = [3.0, 3.0, 3.0], [2.0, 2.0, 2.0]
    K_rot_tensor_max, K_rot_tensor_min = [1.2, 1.2, 0.6], [0.8, 0.8, 0.4]
    K_angvel_tensor_max, K_angvel_tensor_min = [0.2, 0.2, 0.2], [0.1, 0.1, 0.1]
    randomize_params = False


class lee_controller_config_octarotor(lee_controller_config):
    K_rot_tensor_max, K_rot_tensor_min = [10.2, 10.2, 5.6], [10.8, 10.8, 5.4]  # min > max as shipped
    K_angvel_tensor_max, K_angvel_tensor_min = [2.2, 2.2, 2.2], [2.1, 2.1, 2.1]
    randomize_params = True


class fully_actuated_controller_config(lee_controller_config):
    num_actions = 7
    K_pos_tensor_max, K_pos_tensor_min = [1.0, 1.0, 1.0], [1.0, 1.0, 1.0]
    K_vel_tensor_max, K_vel_tensor_min = [8.0, 8.0, 8.0], [8.0, 8.0, 8.0]
    K_rot_tensor_max, K_rot_tensor_min = [2.2, 2.2, 2.6], [2.2, 2.2, 2.6]
    K_angvel_tensor_max, K_angvel_tensor_min = [2.2, 2.2, 2.2], [2.1, 2.1, 2.1]
    randomize_params = True


class lmf2_controller_config(lee_controller_config):
    K_pos_tensor_max, K_pos_tensor_min = [2.0, 2.0, 1.0], [2.0, 2.0, 1.0]
    K_vel_tensor_max, K_vel_tensor_min = [3.3, 3.3, 1.3], [2.7, 2.7, 1.7]
    K_rot_tensor_max, K_rot_tensor_min = [1.85, 1.85, 0.4], [1.6, 1.6, 0.25]
    K_angvel_tensor_max, K_angvel_tensor_min = [0.5, 0.5, 0.09], [0.4, 0.4, 0.075]
    randomize_params = True


class no_control_config:
    num_actions = 4  # replaced by num_motors at robot construction (robots/base_robot.py:33-34)


class magpie_controller_config(lee_controller_config):
    K_pos_tensor_max, K_pos_tensor_min = [2.0, 2.0, 1.0], [2.0, 2.0, 1.0]
    K_vel_tensor_max, K_vel_tensor_min = [3.3, 3.3, 2.6], [2.7, 2.7, 2.3]
    K_rot_tensor_max, K_rot_tensor_min = [12.9453125, 12.9453125, 0.32499998807907104], [8.9453125, 8.9453125, 0.32499998807907104]
    K_angvel_tensor_max = [0.8910937666893005, 0.8910937666893005, 0.04881835892796516]
    K_angvel_tensor_min = [0.6591093766689301, 0.6591093766689301, 0.028818358927965165]
    randomize_params = True
