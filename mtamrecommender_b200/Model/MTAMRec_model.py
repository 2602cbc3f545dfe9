"""MTAM: T-GRU short-term intent + N-hop time-aware attentive memory (Model/MTAMRec_model.py:14-92)."""
from .base_model import base_model


class MTAMRec_model(base_model):
    KIND = "MTAM"

    def __init__(self, FLAGS, Embeding, sess):
        super().__init__(FLAGS, Embeding)
        self.regulation_rate = FLAGS.regulation_rate
        self._build(sess)


class MTAM(MTAMRec_model):
    pass
