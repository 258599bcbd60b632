"""utils/logging.py surface: CustomLogger(name) with setLoggerLevel."""
import logging


class _ColourFormatter(logging.Formatter):
    _C = {logging.DEBUG: "\x1b[36m", logging.INFO: "\x1b[37m", logging.WARNING: "\x1b[33;20m",
          logging.ERROR: "\x1b[31;20m", logging.CRITICAL: "\x1b[31;1m"}
    _FMT = "[%(relativeCreated)d ms][%(name)s] - %(levelname)s : %(message)s (%(filename)s:%(lineno)d)"

    def format(self, record):
        return logging.Formatter(self._C.get(record.levelno, "") + self._FMT + "\x1b[0m").format(record)


class CustomLogger(logging.Logger):
    def __init__(self, logger_name):
        super().__init__(logger_name)
        self.setLevel(logging.INFO)
        self.ch = logging.StreamHandler()
        self.ch.setLevel(logging.INFO)
        self.ch.setFormatter(_ColourFormatter())
        self.addHandler(self.ch)

    def setLoggerLevel(self, level) -> None:
        self.setLevel(level)
        self.ch.setLevel(level)

    def print_example_message(self):  # logging.py:48-53
        for level, text in (("debug", "A Debug"), ("info", "An Info"), ("warning", "A Warning"), ("error", "An Error"), ("critical", "A Critical")):
            getattr(self, level)(f"{text} message will look like this")
