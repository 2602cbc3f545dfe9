"""Behaviour-sequence embedding: feed contract and host-side batch packing.
Reference: Embedding/Behavior_embedding_time_aware_attention.py (:21-46 placeholders, :62-114
get_embedding, :146-192 make_feed_dic_new).
"""
import numpy as np

from .base_embedding import Base_embedding, Placeholder


class Behavior_embedding_time_aware_attention(Base_embedding):
    def __init__(self, is_training=True, user_count=0, item_count=0, category_count=0, max_length_seq=0):
        super().__init__(is_training)
        self.user_count = user_count
        self.item_count = item_count
        self.category_count = category_count
        self.position_count = max_length_seq

    def init_placeholders(self):
        i32, f32 = np.int32, np.float32
        self.user_id = Placeholder("user_id", "user", i32)
        self.item_list = Placeholder("item_list", "item_seq", i32)
        self.category_list = Placeholder("category_list", "category_list", i32)
        self.time_list = Placeholder("time_list", "time_list", f32)
        self.timelast_list = Placeholder("timelast_list", "timelast_list", f32)
        self.timenow_list = Placeholder("timenow_list", "timenow_list", f32)
        self.position_list = Placeholder("position_list", "position_list", i32)
        self.target_item_id = Placeholder("target_item_id", "target_item_id", i32)
        self.target_item_category = Placeholder("target_item_category", "target_item_category", i32)
        self.target_item_time = Placeholder("target_item_time", "target_item_time", f32)
        self.seq_length = Placeholder("seq_length", "seq_length", i32)

    def get_embedding(self, num_units):
        """Declares the four tables `[count+3, num_units]` (:64,71,78,86) and returns the same 10-tuple
        of graph handles as the reference; the lookups themselves run inside the CUDA step."""
        mk = self.init_embedding_lookup_table
        self.user_emb_lookup_table = mk("user", self.user_count + 3, num_units, self.is_training)
        self.item_emb_lookup_table = mk("item", self.item_count + 3, num_units, self.is_training)
        self.category_emb_lookup_table = mk("category", self.category_count + 3, num_units, self.is_training)
        self.position_emb_lookup_table = mk("position", self.position_count + 3, num_units, self.is_training)
        return ("user_embedding", "behavior_list_embedding_dense", "item_list_embedding",
                "category_list_embedding", "position_list_embedding", self.time_list, self.timelast_list,
                self.timenow_list, [self.target_item_id, self.target_item_category, self.target_item_time],
                self.seq_length)

    def bind(self, engine):
        for t in ("user", "item", "category", "position"):
            getattr(self, t + "_emb_lookup_table").bind(engine)

    def make_feed_dic_new(self, batch_data):
        """Right-pads each example's six lists with 0 to position_count (:166-178).  `batch_data` is the
        reference's list of 9-tuples or a `PackedRecords` view (DataHandle/record_store.py), which is padded by
        the library's mtam_pack_records in one pass."""
        B, L = len(batch_data), int(self.position_count)
        if hasattr(batch_data, "pack_into"):
            return {getattr(self, k): v for k, v in batch_data.feed(L).items()}
        a = {"user_id": np.zeros(B, np.int32), "item_list": np.zeros((B, L), np.int32),
             "category_list": np.zeros((B, L), np.int32), "time_list": np.zeros((B, L), np.float32),
             "timelast_list": np.zeros((B, L), np.float32), "timenow_list": np.zeros((B, L), np.float32),
             "position_list": np.zeros((B, L), np.int32), "target_item_id": np.zeros(B, np.int32),
             "target_item_category": np.zeros(B, np.int32), "target_item_time": np.zeros(B, np.float32),
             "seq_length": np.zeros(B, np.int32)}
        for b, ex in enumerate(batch_data):
            n = int(ex[8])
            if n > L:
                raise ValueError(f"example length {n} exceeds max_length_seq {L}")   # np.pad would raise too
            a["user_id"][b] = ex[0]
            a["item_list"][b, :n] = ex[1]
            a["category_list"][b, :n] = ex[2]
            a["time_list"][b, :n] = ex[3]
            a["timelast_list"][b, :n] = ex[4]
            a["timenow_list"][b, :n] = ex[5]
            a["position_list"][b, :n] = ex[6]
            a["target_item_id"][b], a["target_item_category"][b], a["target_item_time"][b] = ex[7][0], ex[7][1], ex[7][2]
            a["seq_length"][b] = n
        return {getattr(self, k): v for k, v in a.items()}
