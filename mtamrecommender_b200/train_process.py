"""Training driver: the drop-in for the reference's `train_process.py` (`Train_main_process`).

Reference: train_process.py:35-118 (construction: flags, data, embedding), :136-218 (model dispatch on
FLAGS.experiment_type), :237-303 (`eval_topk`: HR/NDCG over the test set, best-so-far bookkeeping, recall@k / ndgc@k
summaries), :309-397 (epoch loop: shuffle, `DataInput`, the learning-rate rule, `model.train`, periodic evaluation and
`save_model`), :401-427 (`save_model`: every 50 000 steps).

What differs, and why:
  * the raw-log ETL (`DataHandle/get_origin_data_*`, `Prepare/prepare_data_base`) is out of scope: the
    data set arrives prepared -- the reference's own list of 9-tuples, its `train_data.txt` / `test_data.txt` text form,
    or a `PackedRecords` columnar store -- together with the three vocabulary sizes;
  * the reference wraps every step in `try/except` and logs "Error" (:369-371), which hides failures; here a failing
    step raises;
  * the learning rate is plain host arithmetic (`lr_schedule`), not two `tf.train.exponential_decay` graph nodes
    evaluated with an extra `sess.run` per step.
"""
from __future__ import annotations

import math
import random
import time
from typing import Optional

import numpy as np

from .config.model_parameter import model_parameter
from .DataHandle.get_input_data import DataInput
from .DataHandle.record_store import PackedRecords
from .Embedding.Behavior_embedding_time_aware_attention import Behavior_embedding_time_aware_attention
from .session import Session
from .util import summary as tb
from .util.model_log import create_log


def exponential_decay(learning_rate: float, global_step: int, decay_steps: int, decay_rate: float,
                      staircase: bool = True) -> float:
    """tf.train.exponential_decay: lr * rate^(step/decay_steps), the exponent floored when `staircase`.  TF evaluates
    it in float32."""
    p = global_step / decay_steps
    if staircase:
        p = math.floor(p)
    return float(np.float32(learning_rate) * np.power(np.float32(decay_rate), np.float32(p)))


def lr_schedule(current_lr: float, global_step: int, flags_learning_rate: float, flags_decay_rate: float) -> float:
    """The per-step rule of train_process.py:154-159, 329-333: with the CURRENT value above 0.001 the rate follows
    lr1 = FLAGS.learning_rate * 0.99^floor(gs/100), otherwise lr2 = 0.001 * FLAGS.decay_rate^floor(gs/100).  `gs` is the
    global step (never reset); the caller resets `current_lr` to FLAGS.learning_rate at every epoch start (:324)."""
    if current_lr > 0.001:
        return exponential_decay(flags_learning_rate, global_step, 100, 0.99, True)
    return exponential_decay(0.001, global_step, 100, flags_decay_rate, True)


def model_class(experiment_type: str):
    """train_process.py:164-218.  Models outside the path this library builds raise by name."""
    from .Model.attention_baseline_models import (Self_Attention_Model, Ti_Self_Attention_Model,
                                                  Time_Aware_Self_Attention_Model)
    from .Model.BPRMF import BPRMF
    from .Model.MTAMRec_model import MTAM
    from .Model.PISTRec_model import Time_Aware_self_Attention_model
    table = {"MTAM": MTAM, "SASrec": Self_Attention_Model,
             "Time_Aware_Self_Attention_Model": Time_Aware_Self_Attention_Model,
             "Ti_Self_Attention_Model": Ti_Self_Attention_Model, "bpr": BPRMF,
             "pistrec": Time_Aware_self_Attention_model, "PISTRec": Time_Aware_self_Attention_model}
    from .Model.MTAMRec_model import MTAM_no_time_aware_rnn, MTAM_via_rnn, MTAM_via_T_GRU
    table.update({"MTAM_via_T_GRU": MTAM_via_T_GRU, "MTAM_no_time_aware_rnn": MTAM_no_time_aware_rnn,
                  "MTAM_via_rnn": MTAM_via_rnn})
    if experiment_type not in table:
        raise NotImplementedError(f"experiment_type {experiment_type!r} is outside the hot path this library builds "
                                  f"(built: {sorted(table)})")
    return table[experiment_type]


def _load(data):
    if isinstance(data, str):
        return PackedRecords.load(data) if not data.endswith(".txt") else PackedRecords.from_text(data)
    return data


class Train_main_process:
    def __init__(self, FLAGS=None, train_set=None, test_set=None, user_count: Optional[int] = None,
                 item_count: Optional[int] = None, category_count: Optional[int] = None, argv=None, device="cuda:0"):
        start_time = time.time()
        if FLAGS is None:
            mp = model_parameter(argv)
            FLAGS = mp.get_parameter(mp.flags.FLAGS.experiment_name).FLAGS
        self.FLAGS = FLAGS
        self.logger = create_log(type=FLAGS.type, experiment_type=FLAGS.experiment_type, version=FLAGS.version).logger
        self.logger.info("hello world the experiment begin")
        self.logger.info("The model parameter is :" + str(vars(FLAGS)))
        if train_set is None or test_set is None or None in (user_count, item_count, category_count):
            raise ValueError("Train_main_process needs the prepared train / test sets (lists of 9-tuples, PackedRecords or "
                             "file paths) and user / item / category counts: the raw-log ETL is not part of this library")
        self.train_set, self.test_set = _load(train_set), _load(test_set)
        self.logger.info("DataHandle Process.\tCost time: %.2fs" % (time.time() - start_time))
        self.emb = Behavior_embedding_time_aware_attention(is_training=FLAGS.is_training, user_count=user_count,
                                                           item_count=item_count, category_count=category_count,
                                                           max_length_seq=FLAGS.length_of_user_history)
        self.device = device
        self.global_step = 0
        self.one_epoch_step = 0
        self.now_epoch = 0
        self.learning_rates = []          # the rate of every step taken (inspection / tests)

    # ---- evaluation (train_process.py:237-303) ------------------------------------------------------------------
    def eval_topk(self):
        sums = np.zeros(10, np.float64)
        max_step = 0
        for _, batch_data in DataInput(self.test_set, self.FLAGS.test_batch_size):
            max_step += 1
            sums += np.asarray(self.model.metrics_topK(sess=self.sess, batch_data=batch_data,
                                                       global_step=self.global_step, topk=self.FLAGS.top_k), np.float64)
        sums /= max(max_step, 1)
        for q, k in enumerate((1, 5, 10, 30, 50)):
            hr, ndcg = float(sums[2 * q]), float(sums[2 * q + 1])
            if hr > self.best[k][0] and ndcg > self.best[k][1]:
                self.best[k] = (hr, ndcg)
            self.model.train_writer.add_summary(tb.scalars([("recall@" + str(k), hr)]), global_step=self.global_step)
            self.model.train_writer.add_summary(tb.scalars([("ndgc@" + str(k), ndcg)]), global_step=self.global_step)
            self.logger.info("Test recall rate @ %d : %.4f   ndcg @ %d: %.4f" % (k, hr, k, ndcg))
        return tuple(float(x) for x in sums)

    # ---- training (train_process.py:136-397) --------------------------------------------------------------------
    def train(self, max_steps: Optional[int] = None):
        F = self.FLAGS
        start_time = time.time()
        self.sess = Session(device=self.device)
        self.model = model_class(F.experiment_type)(F, self.emb, self.sess)
        self.logger.info("Init finish.\tCost time: %.2fs" % (time.time() - start_time))
        self.best = {k: (0.0, 0.0) for k in (1, 5, 10, 30, 50)}
        test_start = time.time()
        self.eval_topk()
        self.logger.info("End test. \tTest Cost time: %.2fs" % (time.time() - test_start))
        self.logger.info("Training....\tmax_epochs:%d\tepoch_size:%d" % (F.max_epochs, F.train_batch_size))
        start_time, avg_loss, step_loss = time.time(), 0.0, float("nan")
        rng = random.Random(1234)                     # the reference seeds `random` with 1234 (:30)
        done = False
        for epoch in range(F.max_epochs):
            order = list(range(len(self.train_set)))
            rng.shuffle(order)
            data = self.train_set.take(order) if hasattr(self.train_set, "take") else [self.train_set[i] for i in order]
            self.logger.info("tain_set:%d" % len(self.train_set))
            epoch_start_time = time.time()
            learning_rate = F.learning_rate
            # The step is queued (model.train_submit: padding, the host->device copy and the launch of step i+1 overlap
            # the device's step i) and its loss is booked one iteration later; the pipeline is drained wherever the
            # reference looks at the model (evaluation, saving, the end of an epoch).  FLAGS.pipeline_input = False
            # gives the reference's blocking model.train call per step.
            pipelined = bool(getattr(F, "pipeline_input", True)) and hasattr(self.model, "train_submit")
            pending = []                              # global steps whose results have not been booked yet

            def book(res):
                nonlocal avg_loss, step_loss
                if res is None:
                    return
                step_loss, merge = res
                self.model.train_writer.add_summary(merge, pending.pop(0))
                avg_loss += float(step_loss)

            for _, train_batch_data in DataInput(data, F.train_batch_size):
                learning_rate = lr_schedule(learning_rate, self.global_step, F.learning_rate, F.decay_rate)
                self.learning_rates.append(learning_rate)
                add_summary = bool(self.global_step % F.display_freq == 0)
                pending.append(self.global_step)
                if pipelined:
                    book(self.model.train_submit(train_batch_data, learning_rate))
                else:
                    book(self.model.train(self.sess, train_batch_data, learning_rate, add_summary, self.global_step, epoch))
                self.global_step += 1
                self.one_epoch_step += 1
                last = max_steps is not None and self.global_step >= max_steps
                if pipelined and (self.global_step % F.eval_freq == 0 or last):
                    book(self.model.train_flush())
                if self.global_step % F.eval_freq == 0:
                    self.logger.info("Epoch step is " + str(self.one_epoch_step))
                    self.logger.info("Global step is " + str(self.global_step))
                    self.logger.info("Train_loss is " + str(avg_loss / F.eval_freq))
                    self.eval_topk()
                    avg_loss = 0.0
                    self.save_model()
                if max_steps is not None and self.global_step >= max_steps:
                    done = True
                    break
            if pipelined:
                book(self.model.train_flush())
            self.logger.info("one epoch Cost time: %.2f" % (time.time() - epoch_start_time))
            self.logger.info("Global step is " + str(self.global_step))
            self.logger.info("Train_loss is " + str(step_loss))
            self.eval_topk()
            for k in (1, 5, 10, 30, 50):
                self.logger.info("Max recall rate @ %d: %.4f   ndcg @ %d: %.4f" % (k, self.best[k][0], k, self.best[k][1]))
            self.logger.info("Epoch %d DONE\tCost time: %.2f" % (self.now_epoch, time.time() - start_time))
            self.now_epoch += 1
            self.one_epoch_step = 0
            if done:
                break
        self.model.save(self.sess, self.global_step)
        self.model.train_writer.flush()
        self.logger.info("Finished")
        return step_loss

    def save_model(self):
        """train_process.py:401-427: only every 50 000 global steps."""
        if self.global_step % 50000 == 0:
            self.model.save(self.sess, self.global_step)


if __name__ == "__main__":
    import sys
    raise SystemExit("Train_main_process needs prepared data: construct it from Python with train_set / test_set / counts "
                     "(see INTEGRATION.md); flags given: " + " ".join(sys.argv[1:]))
