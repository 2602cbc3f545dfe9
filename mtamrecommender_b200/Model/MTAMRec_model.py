"""MTAM: T-GRU short-term intent + N-hop time-aware attentive memory (Model/MTAMRec_model.py:14-92), and the sibling
MTAM_via_T_GRU (:167-204), whose hops read the T-GRU's output sequence as their memory."""
from .base_model import base_model


class MTAMRec_model(base_model):
    KIND = "MTAM"

    def __init__(self, FLAGS, Embeding, sess):
        super().__init__(FLAGS, Embeding)
        self.regulation_rate = FLAGS.regulation_rate
        self._build(sess)


class MTAM(MTAMRec_model):
    pass


class MTAM_via_T_GRU(MTAMRec_model):
    """Model/MTAMRec_model.py:167-204: user_history = the T-GRU output sequence, query = layer_norm(short-term intent)."""
    KIND = "MTAM_VIA_T_GRU"


class MTAM_no_time_aware_rnn(MTAMRec_model):
    """Model/MTAMRec_model.py:93-125: MTAM whose short-term intent comes from a plain GRU (GRU.gru_net, gru.py:60-67)."""
    KIND = "MTAM_NO_TIME_AWARE_RNN"


class MTAM_via_rnn(MTAMRec_model):
    """Model/MTAMRec_model.py:206-233: MTAM_via_T_GRU with the plain GRU."""
    KIND = "MTAM_VIA_RNN"
