"""PISTRec: N blocks of time-aware self-attention, loss without the user L2 term
(Model/PISTRec_model.py:11-74; blocks hard-code num_units=128 at :42)."""
from .base_model import base_model


class PISTRec_model(base_model):
    KIND = "PISTREC"

    def __init__(self, FLAGS, Embeding, sess):
        super().__init__(FLAGS, Embeding)
        if FLAGS.num_units != 128:
            raise ValueError("Time_Aware_self_Attention_model hard-codes num_units=128 for its blocks "
                             "(PISTRec_model.py:42): run with --num_units 128")
        self._build(sess)


class Time_Aware_self_Attention_model(PISTRec_model):
    pass
