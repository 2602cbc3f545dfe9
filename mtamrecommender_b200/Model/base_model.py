"""Base class of all models: train step, top-k evaluation, checkpointing.
Reference: Model/base_model.py (:21-45 ctor, :124-147 save/restore, :150-167 train,
:188-242 metrics_topK / calculate_topK, :274-287 summery, :290-328 loss / gradients).

The graph building of the reference is replaced by an `Engine` (libmtam_b200.so); the surface
(`train`, `metrics_topK`, `save`, `restore`, `train_writer`) is unchanged.
"""
import json
import math
import os
import time

import numpy as np

from .. import _lib
from .._lib import GEMM_FP32, GEMM_TF32, GEMM_TF32X3
from ..engine import Engine, ModelConfig
from ..util import checkpoint as ckpt
from ..util import summary as tb
from ..util.model_log import create_log


class base_model(object):
    KIND = None

    def __init__(self, FLAGS, Embedding):
        self.FLAGS = FLAGS
        self.version = self.FLAGS.version
        if getattr(self.FLAGS, "checkpoint_path_dir", None) is not None:
            self.checkpoint_path_dir = self.FLAGS.checkpoint_path_dir
        else:
            self.checkpoint_path_dir = ("data/check_point/" + self.FLAGS.type + "_" + self.FLAGS.experiment_type +
                                        "_" + self.version)
        self.init_optimizer()
        self.embedding = Embedding
        self.logger = create_log().logger

    def init_optimizer(self):
        """base_model.py:71-80: 'adam' (every preset), 'adadelta', 'rmsprop', anything else -> plain gradient descent.
        Built here: adam and gradient descent; the other two raise (TF applies them lazily per IndexedSlices row, which
        needs slot bookkeeping no preset exercises)."""
        name = self.FLAGS.optimizer
        if name in ("adadelta", "rmsprop"):
            raise NotImplementedError(f"optimizer {name!r} is not built (built: 'adam', and gradient descent for any "
                                      "other name, as in base_model.py:79-80)")
        self.optimizer_name = "adam" if name == "adam" else "sgd"

    # ---- construction (the reference's __init__ of each model family + build_model) ----------
    def _build(self, sess, include_user_l2=True):
        F, emb = self.FLAGS, self.embedding
        self.num_units, self.num_heads, self.num_blocks = F.num_units, F.num_heads, F.num_blocks
        self.dropout_rate = F.dropout
        self.max_len = F.length_of_user_history
        (self.user_embedding, self.behavior_list_embedding_dense, self.item_list_emb, self.category_list_emb,
         self.position_list_emb, self.time_list, self.timelast_list, self.timenow_list, self.target,
         self.seq_length) = emb.get_embedding(self.num_units)
        cfg = ModelConfig(kind=self.KIND, max_batch=max(F.train_batch_size, F.test_batch_size), L=self.max_len,
                          D=self.num_units, H=self.num_heads, N=self.num_blocks, user_count=emb.user_count,
                          item_count=emb.item_count, category_count=emb.category_count, reg=F.regulation_rate,
                          clip=F.max_gradient_norm,
                          gemm_mode={"fp32": GEMM_FP32, "tf32": GEMM_TF32}.get(getattr(F, "gemm_mode", "tf32x3"), GEMM_TF32X3),
                          optimizer=self.optimizer_name,
                          # attention dropout: applied by the plain / TiSAS blocks only (multihead_attention.py:179,
                          # time_aware_attention.py:198); the time-aware kinds ignore it, as in the reference
                          dropout=float(F.dropout), dropout_seed=int(getattr(F, "dropout_seed", 1234)))
        device = getattr(sess, "device", "cuda:0") if sess is not None else "cuda:0"
        self.engine = Engine(cfg, device=device, seed=1234)
        emb.bind(self.engine)
        self.summery()
        self.init_variables(sess, self.checkpoint_path_dir)

    def init_variables(self, sess, path, var_list=None):
        """base_model.py:48-69: 'from_scratch' keeps the initialisers, 'full' restores every variable from the model's
        own checkpoint directory, 'fine_tune' restores `var_list` from FLAGS.fine_tune_load_path."""
        lt = self.FLAGS.load_type
        if lt == "full":
            self.restore(sess, path=path)
        elif lt == "fine_tune":
            self.restore(sess, path=self.FLAGS.fine_tune_load_path, variable_list=var_list)
        elif lt != "from_scratch":
            raise ValueError(f"unknown load_type {lt!r} (from_scratch, full, fine_tune)")

    def summery(self):
        """base_model.py:274-287: writers under data/tensorboard_result/<type>_<experiment_type>_<version>_<time>/.
        FLAGS.summary_dir (not a reference flag) moves the root; an empty string keeps the scalars in memory only."""
        self.merged = None
        self._pipe, self._pipe_lr = None, 0.0       # train_submit's input pipeline (engine.FeedPipeline)
        root = getattr(self.FLAGS, "summary_dir", "data/tensorboard_result")
        if root:
            stamp = time.strftime("%Y-%m-%d--%H:%M:%S", time.localtime(time.time()))
            base = os.path.join(root, f"{self.FLAGS.type}_{self.FLAGS.experiment_type}_{self.FLAGS.version}_{stamp}")
            self.train_writer = tb.FileWriter(os.path.join(base, "tensorboard_train"))
            self.eval_writer = tb.FileWriter(os.path.join(base, "tensorboard_eval"))
        else:
            self.train_writer, self.eval_writer = tb.FileWriter(None), tb.FileWriter(None)

    # ---- checkpoint -------------------------------------------------------------------------
    def save(self, sess, global_step=None, path=None, variable_list=None):
        """tf.train.Saver(var_list).save(sess, path/"model.ckpt", global_step) (base_model.py:124-138): file naming and
        the `checkpoint` state file as TF writes them, reference variable names, Adam slots `<var>/Adam`,
        `<var>/Adam_1`, `beta1_power`, `beta2_power` (util/checkpoint.py says what is and is not byte-compatible)."""
        path = path or self.checkpoint_path_dir
        prefix = os.path.join(path, "model.ckpt" + (f"-{int(global_step)}" if global_step is not None else ""))
        eng = self.engine
        blob = {}
        for k in eng.param_names():
            if variable_list is not None and k not in variable_list:
                continue
            blob[k] = eng.get_param(k)
            if eng.cfg.optimizer == "adam" and not eng.is_dead(k):
                blob[k + "/Adam"] = eng.adam_m_view(k).cpu().numpy()
                blob[k + "/Adam_1"] = eng.adam_v_view(k).cpu().numpy()
        t = eng.adam_step()
        if eng.cfg.optimizer == "adam":
            # TF keeps beta^(t+1) in these slots (they start at beta and are multiplied after every step)
            blob["beta1_power"] = np.asarray(np.float32(eng.cfg.beta1) ** (t + 1), np.float32)
            blob["beta2_power"] = np.asarray(np.float32(eng.cfg.beta2) ** (t + 1), np.float32)
        meta = {"kind": eng.cfg.kind, "adam_step": int(t), "global_step": None if global_step is None else int(global_step),
                "config": {k: getattr(eng.cfg, k) for k in ("L", "D", "H", "N", "user_count", "item_count", "category_count",
                                                             "optimizer")}}
        fn = ckpt.save(prefix, blob, meta)
        self.logger.info("model saved at %s" % fn)
        return fn

    def restore(self, sess, path, variable_list=None, graph_path=None):
        """saver.restore(sess, tf.train.latest_checkpoint(path)) (base_model.py:140-147)."""
        prefix = ckpt.latest_checkpoint(path)
        if prefix is None:
            raise FileNotFoundError(f"no checkpoint state file in {path!r}")
        blob = ckpt.load(prefix)
        meta = json.load(open(prefix + ".meta"))
        eng = self.engine
        import torch
        for k in eng.param_names():
            if variable_list is not None and k not in variable_list:
                continue
            if k not in blob:
                raise KeyError(f"variable {k!r} not in checkpoint {prefix}")
            eng.set_param(k, blob[k])
            if k + "/Adam" in blob:
                eng.adam_m_view(k).copy_(torch.from_numpy(blob[k + "/Adam"]).reshape(eng.adam_m_view(k).shape))
                eng.adam_v_view(k).copy_(torch.from_numpy(blob[k + "/Adam_1"]).reshape(eng.adam_v_view(k).shape))
        if variable_list is None:
            eng.set_adam_step(int(meta.get("adam_step", 0)))
        self.logger.info("model restored from %s" % path)

    # ---- steps ------------------------------------------------------------------------------
    def _feed(self, batch_data):
        d = self.embedding.make_feed_dic_new(batch_data=batch_data)
        return {p.key: v for p, v in d.items()}

    def train(self, sess, batch_data, learning_rate, add_summary=False, global_step=0, epoch=0):
        if hasattr(batch_data, "pack_into"):      # PackedRecords: padded straight into the pinned feed buffers
            loss = self.engine.train_step_records(batch_data, learning_rate)
        else:
            loss = self.engine.train_step(self._feed(batch_data), learning_rate)
        sc = self.engine.last_scalars
        # tf.summary.merge_all of base_model.py:324-327, evaluated in the same run as the step
        self.merged = tb.scalars((("Training Loss", sc[_lib.S_LOSS_ORIGIN]), ("normalized Training Loss", sc[_lib.S_LOSS]),
                                  ("l2_norm", sc[_lib.S_L2_NORM]), ("Learning_rate", float(learning_rate))))
        return np.float32(loss), self.merged

    def _merged(self, sc, learning_rate):
        return tb.scalars((("Training Loss", sc[_lib.S_LOSS_ORIGIN]), ("normalized Training Loss", sc[_lib.S_LOSS]),
                           ("l2_norm", sc[_lib.S_L2_NORM]), ("Learning_rate", float(learning_rate))))

    def train_submit(self, batch_data, learning_rate):
        """Pipelined `train` (engine.FeedPipeline): queues this step and returns `(loss, merged)` of the step submitted
        BEFORE it -- None for the first -- so the host's padding / copy of the next batch overlaps the device's step.
        The driver loop (train_process.py) uses it; `train` stays the blocking call of the reference
        (base_model.py:150-167).  Call `train_flush()` before evaluating, saving or reading weights."""
        if self._pipe is None:
            from ..engine import FeedPipeline
            self._pipe = FeedPipeline(self.engine)
        feed = batch_data if hasattr(batch_data, "pack_into") else self._feed(batch_data)
        sc = self._pipe.submit(feed, learning_rate)
        prev_lr, self._pipe_lr = self._pipe_lr, learning_rate
        if sc is None:
            return None
        return np.float32(sc[_lib.S_LOSS]), self._merged(sc, prev_lr)

    def train_flush(self):
        """`(loss, merged)` of the last step queued by `train_submit`, None if there is none."""
        if self._pipe is None:
            return None
        sc = self._pipe.flush()
        if sc is None:
            return None
        return np.float32(sc[_lib.S_LOSS]), self._merged(sc, self._pipe_lr)

    def metrics_topK(self, sess, batch_data, global_step, topk):
        if hasattr(batch_data, "pack_into"):
            b = self.engine.upload_records(batch_data)
        else:
            b = self.engine.upload(self._feed(batch_data))
        idx, _ = self.engine.eval_topk_device(b, 50)
        m = self.engine.hr_ndcg_device(idx, b.t["target_item_id"]).cpu().numpy()
        return tuple(float(x) for x in m)

    def calculate_topK(self, k, indices_result, result_item, global_step, length):
        """Host restatement kept for callers that hold index arrays (base_model.py:215-242)."""
        hit, nd = 0, 0.0
        for row, tgt in zip(indices_result, result_item):
            row = list(row)[:k]
            if tgt in row:
                hit += 1
                nd += math.log(2) / math.log(row.index(tgt) + 2)
        return hit / len(indices_result), (nd / length if nd else 0)
