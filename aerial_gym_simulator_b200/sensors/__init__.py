from .imu_sensor import IMUSensor  # noqa: F401
