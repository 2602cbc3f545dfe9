"""Singleton logger (reference: util/model_log.py:5-49) -- stdlib logging, stream handler only
unless a log file is requested."""
import logging
import os


class create_log:
    _instance = None

    def __new__(cls, type=None, experiment_type=None, version=None, log_dir=None):
        if cls._instance is None:
            inst = super().__new__(cls)
            logger = logging.getLogger("mtamrecommender_b200")
            if not logger.handlers:
                logger.setLevel(logging.INFO)
                h = logging.StreamHandler()
                h.setFormatter(logging.Formatter("%(asctime)s %(levelname)s %(message)s"))
                logger.addHandler(h)
            if log_dir and type and experiment_type:
                os.makedirs(log_dir, exist_ok=True)
                fh = logging.FileHandler(os.path.join(log_dir, f"{type}_{experiment_type}_{version}_log.txt"))
                logger.addHandler(fh)
            inst.logger = logger
            cls._instance = inst
        return cls._instance
