"""What the two command lines share: the comma-separated click type of `/root/reference/utils.py:742-830` and a logger with
loguru's method names (`predict.py` / `analyze.py` of the reference log through loguru, which this image does not have; when
it is importable it is used as is)."""
from __future__ import annotations

import logging
import sys
from pathlib import Path

import click

LOG_LEVELS = ["TRACE", "DEBUG", "INFO", "SUCCESS", "WARNING", "ERROR", "CRITICAL"]
_LEVEL_NO = {"TRACE": 5, "DEBUG": 10, "INFO": 20, "SUCCESS": 25, "WARNING": 30, "ERROR": 40, "CRITICAL": 50}


class CommaSeparated(click.ParamType):
    """"a,b,c" -> [type_(a), type_(b), type_(c)]; with `n`, exactly n values (utils.py:742-830)."""

    name = "comma_separated"

    def __init__(self, type_: type = str, n: int | None = None) -> None:
        if n is not None and n <= 0:
            raise ValueError("n must be None or a positive integer")
        self.type, self.n = type_, n

    def convert(self, value, param, ctx):
        if value is None:
            return None
        if isinstance(value, (list, tuple)):
            return list(value)
        items = [v.strip() for v in str(value).split(",")]
        if self.n is not None and len(items) != self.n:
            self.fail(f"{value!r} does not contain exactly {self.n} comma separated values", param, ctx)
        try:
            return [self.type(v) for v in items]
        except ValueError:
            self.fail(f"{value!r} is not a comma separated list of {self.type.__name__}", param, ctx)


class _StdLogger:
    """info / success / warning / error / critical on the standard logging module."""

    def __init__(self):
        for name, no in _LEVEL_NO.items():
            logging.addLevelName(no, name)
        self._log = logging.getLogger("depth_completion_b200")
        self._log.propagate = False

    def configure(self, level: str, path: Path | None):
        for h in list(self._log.handlers):
            self._log.removeHandler(h)
        fmt = logging.Formatter("%(asctime)s | %(levelname)-8s | %(message)s")
        handlers = [logging.StreamHandler(sys.stderr)]
        if path is not None:
            path.parent.mkdir(parents=True, exist_ok=True)
            handlers.append(logging.FileHandler(path))
        for h in handlers:
            h.setFormatter(fmt)
            self._log.addHandler(h)
        self._log.setLevel(_LEVEL_NO[level])

    def __getattr__(self, name):
        no = _LEVEL_NO.get(name.upper())
        if no is None:
            raise AttributeError(name)
        return lambda msg: self._log.log(no, msg)


def get_logger(level: str = "INFO", path: Path | None = None):
    """A configured logger: loguru's when installed (same sinks as predict.py:385-394), else the shim above."""
    try:
        from loguru import logger  # noqa: PLC0415
    except ImportError:
        logger = _StdLogger()
        logger.configure(level, path)
        return logger
    logger.remove()
    logger.add(sys.stderr, level=level)
    if path is not None:
        path.parent.mkdir(parents=True, exist_ok=True)
        logger.add(path, rotation="100 MB", level=level)
    return logger
