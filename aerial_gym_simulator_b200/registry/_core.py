"""Name -> class/config registries with the reference's method surface
(registry/{task,robot,controller,env,sim}_registry.py).  One generic table underneath."""


class _Table:
    def __init__(self, kind):
        self._kind = kind
        self._classes = {}
        self._configs = {}

    def _put(self, name, cls, cfg):
        self._classes[name] = cls
        self._configs[name] = cfg

    def _get(self, table, name):
        if name not in table:
            raise ValueError(f"{self._kind} {name} not found in {self._kind} registry. Available: {sorted(table)}")
        return table[name]


class TaskRegistry(_Table):
    """registry/task_registry.py"""

    def __init__(self):
        super().__init__("Task")
        self.task_class_registry = self._classes
        self.task_config_registry = self._configs

    def register_task(self, task_name, task_class, task_config):
        self._put(task_name, task_class, task_config)

    def get_task_class(self, task_name):
        return self._get(self._classes, task_name)

    def get_task_config(self, task_name):
        return self._get(self._configs, task_name)

    def get_task_names(self):
        return list(self._classes.keys())

    def get_task_classes(self):
        return list(self._classes.values())

    def get_task_configs(self):
        return list(self._configs.values())

    def make_task(self, task_name, seed=None, num_envs=None, headless=None, use_warp=None, args=None):
        """registry/task_registry.py:25-30, plus `args` (not in the reference): merged into task_config.args for this task, e.g.
        {"world_size": W, "rank": r} for an env-sharded task whose observation is all-gathered (task/base_task.py)"""
        cfg = self.get_task_config(task_name)
        if args:
            cfg.args = {**dict(getattr(cfg, "args", None) or {}), **dict(args)}
        return self.get_task_class(task_name)(cfg, seed=seed, num_envs=num_envs, headless=headless, use_warp=use_warp)


class RobotRegistry(_Table):
    """registry/robot_registry.py"""

    def __init__(self):
        super().__init__("Robot")
        self.robot_classes = self._classes
        self.robot_configs = self._configs

    def register(self, robot_name, robot_class, robot_config):
        self._put(robot_name, robot_class, robot_config)

    def get_robot_class(self, robot_name):
        return self._get(self._classes, robot_name)

    def get_robot_config(self, robot_name):
        return self._get(self._configs, robot_name)

    def get_robot_names(self):
        return self._classes.keys()

    def make_robot(self, robot_name, controller_name, env_config, device):
        cls, cfg = self.get_robot_class(robot_name), self.get_robot_config(robot_name)
        return cls(cfg, controller_name, env_config, device), cfg


class ControllerRegistry(_Table):
    """registry/controller_registry.py"""

    def __init__(self):
        super().__init__("Controller")
        self.controller_classes = self._classes
        self.controller_configs = self._configs

    def register_controller(self, controller_name, controller_class, controller_config):
        self._put(controller_name, controller_class, controller_config)

    def get_controller_class(self, controller_name):
        return self._get(self._classes, controller_name)

    def get_controller_names(self):
        return self._classes.keys()

    def get_controller_config(self, controller_name):
        return self._get(self._configs, controller_name)

    def make_controller(self, controller_name, num_envs, device, mode="robot"):
        cls, cfg = self.get_controller_class(controller_name), self.get_controller_config(controller_name)
        return cls(cfg, num_envs, device), cfg


class EnvConfigRegistry(_Table):
    """registry/env_registry.py"""

    def __init__(self):
        super().__init__("Env")
        self.env_configs = self._configs

    def register(self, env_name, env_config):
        self._put(env_name, None, env_config)

    def get_env_config(self, env_name):
        return self._get(self._configs, env_name)

    def get_env_names(self):
        return self._configs.keys()

    def make_env(self, env_name):
        return self.get_env_config(env_name)


class SimConfigRegistry(_Table):
    """registry/sim_registry.py"""

    def __init__(self):
        super().__init__("Sim")
        self.sim_configs = self._configs

    def register(self, sim_name, sim_config):
        self._put(sim_name, None, sim_config)

    def get_sim_config(self, sim_name):
        return self._get(self._configs, sim_name)

    def get_sim_names(self):
        return self._configs.keys()

    def make_sim(self, sim_name):
        return self.get_sim_config(sim_name)


task_registry = TaskRegistry()
robot_registry = RobotRegistry()
controller_registry = ControllerRegistry()
env_config_registry = EnvConfigRegistry()
sim_config_registry = SimConfigRegistry()
