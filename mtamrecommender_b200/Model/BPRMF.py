"""BPR-MF (Model/BPRMF.py:10-59)."""
from .base_model import base_model


class BPRMF(base_model):
    KIND = "BPRMF"

    def __init__(self, FLAGS, Embeding, sess):
        super().__init__(FLAGS, Embeding)
        self._build(sess)
