"""Logger factory with the reference's conventions (hwer/logging.py:1-13): level from the LOGLEVEL environment
variable unless given, records tagged with pid, time, level and logger name."""
import logging as _logging
import os as _os

_RECORD = "[PID: %(process)d] [%(asctime)s] [%(levelname)s] [%(name)s]: %(message)s"


def _default_level():
    return _os.environ.get("LOGLEVEL", "INFO")


_logging.basicConfig(format=_RECORD, level=_default_level())


def getLogger(name, level=None):
    logger = _logging.getLogger(name)
    logger.setLevel(_default_level() if level is None else level)
    return logger
