"""Thin stand-in for tf.Session (train_process.py:146): names the CUDA device the engine uses."""


class Session:
    def __init__(self, device="cuda:0"):
        self.device = device

    def as_default(self):
        return self

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
