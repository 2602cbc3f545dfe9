"""Attention baselines (Model/attention_baseline_models.py:7-84)."""
from .base_model import base_model


class Attention_Baseline_Model(base_model):
    def __init__(self, FLAGS, Embeding, sess):
        super().__init__(FLAGS, Embeding)
        self._build(sess)


class Self_Attention_Model(Attention_Baseline_Model):
    KIND = "SASREC"


class Time_Aware_Self_Attention_Model(Attention_Baseline_Model):
    KIND = "TA_SASREC"


class Ti_Self_Attention_Model(Attention_Baseline_Model):
    KIND = "TISASREC"
