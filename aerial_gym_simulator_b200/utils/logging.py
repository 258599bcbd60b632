"""CustomLogger(name): the console logger the reference's scripts construct at import time (utils/logging.py there).  Same surface --
a logging.Logger subclass with `.ch` (its stream handler), setLoggerLevel(level) and print_example_message() -- and the same line
layout, so log output of a script reads the same on either package."""
import logging

_LAYOUT = "[%(relativeCreated)d ms][%(name)s] - %(levelname)s : %(message)s (%(filename)s:%(lineno)d)"
_ANSI = {"DEBUG": "36", "INFO": "37", "WARNING": "33;20", "ERROR": "31;20", "CRITICAL": "31;1"}  # cyan, white, yellow, red, bold red


class _LevelColours(logging.Formatter):
    """one pre-built formatter per level name; unknown levels print uncoloured"""

    def __init__(self):
        super().__init__(_LAYOUT)
        self._by_level = {name: logging.Formatter(f"\x1b[{code}m{_LAYOUT}\x1b[0m") for name, code in _ANSI.items()}

    def format(self, record):
        chosen = self._by_level.get(record.levelname)
        return chosen.format(record) if chosen is not None else super().format(record)


class CustomLogger(logging.Logger):
    def __init__(self, logger_name, level=logging.INFO):
        super().__init__(logger_name, level)
        self.ch = logging.StreamHandler()
        self.ch.setFormatter(_LevelColours())
        self.ch.setLevel(level)
        self.addHandler(self.ch)

    def setLoggerLevel(self, level) -> None:
        for target in (self, self.ch):
            target.setLevel(level)
        # a logger constructed directly is not in logging's manager, whose cache flush setLevel relies on: drop the stale answers
        getattr(self, "_cache", {}).clear()

    def print_example_message(self):
        for method, article in (("debug", "A"), ("info", "An"), ("warning", "A"), ("error", "An"), ("critical", "A")):
            getattr(self, method)(f"{article} {method.capitalize()} message will look like this")
