"""hwer/logging.py:1-13 -- same LOGLEVEL convention and record format."""
import logging
import os

FORMAT = '[PID: %(process)d] [%(asctime)s] [%(levelname)s] [%(name)s]: %(message)s'
logging.basicConfig(format=FORMAT, level=os.environ.get("LOGLEVEL", "INFO"))


def getLogger(name, level=None):
    log = logging.getLogger(name)
    log.setLevel(level if level is not None else os.environ.get("LOGLEVEL", "INFO"))
    return log
