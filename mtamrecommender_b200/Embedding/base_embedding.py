"""Base class of the embedding module (reference: Embedding/base_embedding.py:7-60).

In the reference this class creates tf.placeholders and tf.get_variable tables.  Here a table is a
named block of the engine's parameter arena and a placeholder is a feed key; the arithmetic runs in
libmtam_b200.so.
"""
import numpy as np

from ..util.model_log import create_log


class Placeholder:
    """Feed key standing in for a tf.placeholder (hashable, carries the reference's name/dtype)."""

    def __init__(self, key, name, dtype):
        self.key, self.name, self.dtype = key, name, dtype

    def __repr__(self):
        return f"<placeholder {self.name}:{np.dtype(self.dtype).name}>"


class TableRef:
    """Lazy handle on an embedding table `[total_count, embedding_dim]` (init U(+-sqrt(6/dim)),
    base_embedding.py:46-60); reads through to the engine once a model is built."""

    def __init__(self, name, total_count, embedding_dim):
        self.name, self.shape = "embedding_layer/" + name, (int(total_count), int(embedding_dim))
        self._engine = None

    def bind(self, engine):
        self._engine = engine

    def numpy(self):
        if self._engine is None:
            raise RuntimeError(f"table {self.name} is not bound to a model yet")
        return self._engine.get_param(self.name)


class Base_embedding:
    def __init__(self, is_training=True, config_file=None):
        self.embedding_file_path = config_file
        self.is_training = is_training
        self.logger = create_log().logger
        self.init_placeholders()

    def padding(self, one_list, max_len):
        # base_embedding.py:23-36: np.pad result is discarded there, so only truncation has an effect
        if len(one_list) > max_len:
            one_list = one_list[:max_len]
        return one_list

    def init_placeholders(self):
        pass

    def get_embedding(self):
        pass

    def make_feed_dic(self, batch_data):
        pass

    def init_embedding_lookup_table(self, name, total_count, embedding_dim, is_training=True):
        return TableRef(name, total_count, embedding_dim)
