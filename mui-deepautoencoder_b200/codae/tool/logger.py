"""Host-side logging helpers (codae/tool/logger.py of the reference).  Out of the hot path: plain
`logging`, no matplotlib (PlotDrawer is a no-op recorder)."""
import datetime
import json
import logging
import os
import sys


def set_logging(log_file_path=None, logging_level=logging.INFO):
    """Root logger to stdout (+ log/CODAE_<date>.log when a directory is given) (logger.py:15-38)."""
    log = logging.getLogger("CODAE")
    log.setLevel(logging_level)
    log.handlers = []
    fmt = logging.Formatter("%(asctime)s [%(levelname)s] %(message)s")
    sh = logging.StreamHandler(sys.stdout)
    sh.setFormatter(fmt)
    log.addHandler(sh)
    if log_file_path is not None:
        os.makedirs(log_file_path, exist_ok=True)
        fh = logging.FileHandler(os.path.join(log_file_path, "CODAE_" + get_date() + ".log"))
        fh.setFormatter(fmt)
        log.addHandler(fh)
    return log


def get_date():
    return datetime.datetime.now().strftime("%Y_%m_%d_%H_%M_%S")


def display_info(config, nb_observation, metric_log):
    log = logging.getLogger("CODAE")
    log.info("Observations: %d", nb_observation)
    for section, values in config.items():
        log.info("%s: %s", section, values)
    metric_log["nb_observation"] = nb_observation
    metric_log["config"] = config
    return metric_log


def export_parameters_to_json(config, path):
    with open(path, "w") as f:
        json.dump(config, f, indent=1)


class PlotDrawer:
    """Records series instead of drawing them (matplotlib is not part of the hot path)."""

    def __init__(self, *args, **kwargs):
        self.series = {}

    def add(self, name, x, y):
        self.series.setdefault(name, []).append((x, y))

    def export_to_png(self, *args, **kwargs):
        return None
