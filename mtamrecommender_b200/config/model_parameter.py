"""Flags and experiment presets (reference: config/model_parameter.py).

Same flag names and defaults as the reference's tf.flags definitions (:9-72) and the same
`model_parameter().get_parameter(experiment_name).FLAGS` access pattern (:75-396), built on
argparse so no TensorFlow is needed.
"""
import argparse

_DEFS = [
    ("version", str, "bpr"), ("checkpoint_path_dir", str, None), ("hidden_units", int, 128),
    ("num_blocks", int, 6), ("num_heads", int, 8), ("num_units", int, 128), ("dropout", float, 0.5),
    ("regulation_rate", float, 0.00005), ("itemid_embedding_size", int, 64), ("cateid_embedding_size", int, 64),
    ("concat_time_emb", bool, True), ("optimizer", str, "adam"), ("learning_rate", float, 0.001),
    ("decay_rate", float, 0.001), ("max_gradient_norm", float, 1.0), ("train_batch_size", int, 256),
    ("test_batch_size", int, 100), ("max_epochs", int, 200), ("display_freq", int, 10), ("eval_freq", int, 200),
    ("max_len", int, 150), ("global_step", int, 100), ("cuda_visible_devices", str, "0"),
    ("per_process_gpu_memory_fraction", float, 0.8), ("gap_num", int, 6), ("is_training", bool, True),
    ("type", str, "yoochoose"), ("experiment_type", str, "pistrec"), ("length_of_user_history", int, 50),
    ("length_of_item_history", int, 50), ("max_length_seq", int, 50), ("init_origin_data", bool, False),
    ("init_train_data", bool, False), ("user_count_limit", int, 10000), ("causality", str, "unidirection"),
    ("pos_embedding", str, "time"), ("test_frac", int, 5), ("mask_rate", float, 0.2),
    ("neg_sample_ratio", float, 20), ("remove_duplicate", bool, True), ("experiment_data_type", str, "item_based"),
    ("fine_tune_load_path", str, None), ("load_type", str, "from_scratch"), ("draw_pic", bool, False),
    ("top_k", int, 20), ("experiment_name", str, "data_init"),
    # not a reference flag: dense contractions on tcgen05 with the 3-term TF32 split ("tf32x3", fp32-class accuracy)
    # or on exact-fp32 FFMA tiles ("fp32")
    ("gemm_mode", str, "tf32x3"),
]

# name -> (type, num_blocks, experiment_type, version, test_batch_size)
_MODEL_PRESETS = {
    "Ti_Self_Attention_Modelb3_beauty": ("beauty", 3, "Ti_Self_Attention_Model", None, 2048),
    "STAMP_beauty": ("beauty", 6, "STAMP", None, 2048),
    "MTAM_via_rnnb6_beauty": ("beauty", 6, "MTAM", "MTAM_via_rnn6_beauty", 2048),
    "Time_Aware_Self_Attention_Modelb3_yoochoose": ("yoochoose", 3, "Time_Aware_Self_Attention_Model", None, 2048),
    "MTAMb7_elec": ("elec", 7, "MTAM", None, 2048),
    "MTAMb8_elec": ("elec", 8, "MTAM", None, 2048),
    "MTAM_with_T_SeqRecb6_yoochoose": ("yoochoose", 6, "MTAM_with_T_SeqRec", None, 2048),
    "MTAM_no_time_aware_attb7_music_256": ("music", 7, "MTAM_no_time_aware_att", None, 2048),
    "MTAM_with_T_SeqRecb7_music": ("music", 7, "MTAM_with_T_SeqRec", None, 2048),
    "MTAM_via_rnnb7_music": ("music", 7, "MTAM_via_rnn", None, 1500),
    "Time_Aware_Self_Attention_Modelb3_music": ("music", 3, "Time_Aware_Self_Attention_Model", None, 2048),
    "Time_Aware_Self_Attention_Modelb2_elec": ("elec", 2, "Time_Aware_Self_Attention_Model", None, 2048),
    "Time_Aware_Self_Attention_Modelb1_elec": ("elec", 1, "MTAM_with_T_SeqRec", None, 2048),
}
_MODEL_COMMON = dict(causality="unidirection", num_heads=1, learning_rate=0.001, decay_rate=0.995,
                     regulation_rate=0.00005, checkpoint_path_dir=None, user_count_limit=1000000,
                     init_train_data=False, init_origin_data=False, max_epochs=200, load_type="from_scratch",
                     train_batch_size=256, eval_freq=500, dropout=0.5, cuda_visible_devices="0",
                     length_of_user_history=50)


def _str2bool(v):
    return str(v).lower() in ("1", "true", "t", "yes", "y")


class _Flags:
    """Stand-in for tf.flags: `flags.FLAGS.<name>` after parsing."""

    def __init__(self, argv=None):
        p = argparse.ArgumentParser(allow_abbrev=False)
        for name, typ, default in _DEFS:
            p.add_argument("--" + name, type=_str2bool if typ is bool else typ, default=default)
        self.FLAGS, _ = p.parse_known_args(argv)


class model_parameter:
    def __init__(self, argv=None):
        self.flags = _Flags(argv if argv is not None else [])

    def get_parameter(self, type):
        F = self.flags.FLAGS
        if type == "data_init":        # note: also the default experiment_name, so it overwrites FLAGS.type
            F.type, F.init_train_data, F.init_origin_data = "taobaoapp", False, False
            F.user_count_limit, F.version, F.pos_embedding, F.test_frac = 80000, "tmall_init", "time", 100
            F.causality, F.remove_duplicate, F.gap_num, F.length_of_user_history = "unidirection", False, 15, 50
        elif type == "statistics":
            F.type, F.init_train_data, F.init_origin_data = "beauty", True, True
            F.user_count_limit, F.version, F.pos_embedding, F.test_frac = 100000000, "beauty_statistics", "time", 100
            F.causality, F.gap_num, F.length_of_item_history = "unidirection", 15, 50
        elif type in _MODEL_PRESETS:
            t, nb, et, ver, tbs = _MODEL_PRESETS[type]
            for k, v in _MODEL_COMMON.items():
                setattr(F, k, v)
            F.type, F.num_blocks, F.experiment_type, F.test_batch_size = t, nb, et, tbs
            F.version = ver or type
        return self.flags
