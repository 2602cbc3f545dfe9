"""Base class of all models: train step, top-k evaluation, checkpointing.
Reference: Model/base_model.py (:21-45 ctor, :124-147 save/restore, :150-167 train,
:188-242 metrics_topK / calculate_topK, :274-287 summery, :290-328 loss / gradients).

The graph building of the reference is replaced by an `Engine` (libmtam_b200.so); the surface
(`train`, `metrics_topK`, `save`, `restore`, `train_writer`) is unchanged.
"""
import math
import os

import numpy as np

from .._lib import GEMM_FP32, GEMM_TF32X3
from ..engine import Engine, ModelConfig
from ..util.model_log import create_log


class _NullWriter:
    """train_writer / eval_writer stand-in (tf.summary.FileWriter): keeps the scalars it is given."""

    def __init__(self):
        self.events = []

    def add_summary(self, summary, global_step=None):
        self.events.append((global_step, summary))


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
        if self.FLAGS.optimizer != "adam":
            raise NotImplementedError(f"optimizer {self.FLAGS.optimizer!r}: only 'adam' (every preset) is built")

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
                          gemm_mode=GEMM_FP32 if getattr(F, "gemm_mode", "tf32x3") == "fp32" else GEMM_TF32X3)
        device = getattr(sess, "device", "cuda:0") if sess is not None else "cuda:0"
        self.engine = Engine(cfg, device=device, seed=1234)
        emb.bind(self.engine)
        self.summery()
        self.init_variables(sess, self.checkpoint_path_dir)

    def init_variables(self, sess, path, var_list=None):
        if self.FLAGS.load_type == "full":
            self.restore(sess, path=path)
        elif self.FLAGS.load_type == "fine_tune":
            self.restore(sess, path=self.FLAGS.fine_tune_load_path, variable_list=var_list)

    def summery(self):
        self.merged = None
        self.train_writer = _NullWriter()
        self.eval_writer = _NullWriter()

    # ---- checkpoint -------------------------------------------------------------------------
    def save(self, sess, global_step=None, path=None, variable_list=None):
        path = path or self.checkpoint_path_dir
        os.makedirs(path, exist_ok=True)
        fn = os.path.join(path, "model.ckpt" + (f"-{global_step}" if global_step is not None else "") + ".npz")
        eng = self.engine
        blob = {}
        for k in eng.param_names():
            if variable_list is not None and k not in variable_list:
                continue
            blob[k] = eng.get_param(k)
            blob[k + "/Adam"] = eng.adam_m_view(k).cpu().numpy()
            blob[k + "/Adam_1"] = eng.adam_v_view(k).cpu().numpy()
        blob["__adam_step__"] = np.int64(eng.adam_step())
        np.savez(fn, **blob)
        with open(os.path.join(path, "checkpoint"), "w") as f:
            f.write(os.path.basename(fn) + "\n")
        self.logger.info("model saved at %s" % fn)
        return fn

    def restore(self, sess, path, variable_list=None, graph_path=None):
        with open(os.path.join(path, "checkpoint")) as f:
            fn = os.path.join(path, f.read().strip())
        blob = np.load(fn)
        eng = self.engine
        import torch
        for k in eng.param_names():
            if variable_list is not None and k not in variable_list:
                continue
            eng.set_param(k, blob[k])
            if k + "/Adam" in blob:
                eng.adam_m_view(k).copy_(torch.from_numpy(blob[k + "/Adam"]).reshape(eng.adam_m_view(k).shape))
                eng.adam_v_view(k).copy_(torch.from_numpy(blob[k + "/Adam_1"]).reshape(eng.adam_v_view(k).shape))
        if "__adam_step__" in blob:
            eng.set_adam_step(int(blob["__adam_step__"]))
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
        return np.float32(loss), self.merged

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
