"""CPU tests of the host-side mirror of the reference surface (no GPU): feed packing, DataInput, flags and
presets, synthetic generator invariants."""
import math

import numpy as np

from oracle import mtam_oracle as O
from mtamrecommender_b200.DataHandle.get_input_data import DataInput
from mtamrecommender_b200.Embedding.Behavior_embedding_time_aware_attention import Behavior_embedding_time_aware_attention
from mtamrecommender_b200.config.model_parameter import model_parameter
from mtamrecommender_b200.synth import ZipfSampler, feed_to_records, synth_feed


def test_make_feed_dic_new_matches_oracle_restatement():
    cfg = O.OracleConfig(kind=O.MTAM, L=9, D=32, user_count=7, item_count=30, category_count=4)
    recs = O.synth_records(cfg, 17, 3)
    emb = Behavior_embedding_time_aware_attention(True, 7, 30, 4, 9)
    d = emb.make_feed_dic_new(recs)
    ref = O.make_feed(cfg, recs)
    assert len(d) == 11
    for p, v in d.items():
        assert v.dtype == p.dtype and np.array_equal(v, ref[p.key]), p
    assert emb.position_count == 9 and emb.item_count == 30
    handles = emb.get_embedding(32)
    assert len(handles) == 10 and emb.item_emb_lookup_table.shape == (33, 32)


def test_datainput_yields_consecutive_slices_with_short_tail():
    data = list(range(10))
    got = [(i, b) for i, b in DataInput(data, 4)]
    assert [b for _, b in got] == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]] and [i for i, _ in got] == [1, 2, 3]
    assert list(DataInput([], 4)) == []


def test_flags_defaults_and_presets():
    F = model_parameter().flags.FLAGS
    assert (F.num_units, F.num_blocks, F.num_heads, F.dropout, F.regulation_rate) == (128, 6, 8, 0.5, 0.00005)
    assert (F.learning_rate, F.decay_rate, F.max_gradient_norm, F.train_batch_size, F.test_batch_size) == (0.001, 0.001, 1.0, 256, 100)
    assert F.length_of_user_history == 50 and F.optimizer == "adam" and F.experiment_type == "pistrec"
    F = model_parameter().get_parameter("MTAMb7_elec").FLAGS
    assert (F.type, F.num_blocks, F.num_heads, F.decay_rate, F.experiment_type, F.version) == ("elec", 7, 1, 0.995, "MTAM", "MTAMb7_elec")
    F = model_parameter().get_parameter("data_init").FLAGS       # the default experiment_name overwrites FLAGS.type
    assert F.type == "taobaoapp"
    F = model_parameter(["--num_units", "64", "--type", "movielen"]).flags.FLAGS
    assert F.num_units == 64 and F.type == "movielen"


def test_synth_feed_follows_record_construction():
    L, items, cats = 12, 500, 20
    f = synth_feed(64, L, items, cats, 100, 5, ZipfSampler(items))
    n = f["seq_length"]
    assert n.min() >= 2 and n.max() <= L
    for b in range(64):
        k = n[b]
        assert f["item_list"][b, k - 1] == items + 1 and f["category_list"][b, k - 1] == cats + 1
        assert np.all(f["item_list"][b, k:] == 0) and np.all(f["time_list"][b, k:] == 0)
        assert np.array_equal(f["position_list"][b, :k], np.arange(k))
        assert f["time_list"][b, k - 1] == f["target_item_time"][b]
        assert f["timelast_list"][b, 0] == 0 and f["timelast_list"][b, k - 1] == 0 and f["timenow_list"][b, k - 1] == 0
        assert np.all(f["item_list"][b, :k - 1] < items)
        assert np.all(np.diff(f["time_list"][b, :k]) > 0)
    recs = feed_to_records(f)
    emb = Behavior_embedding_time_aware_attention(True, 100, items, cats, L)
    back = emb.make_feed_dic_new(recs)
    assert all(np.array_equal(v, f[p.key]) for p, v in back.items())


def test_calculate_topk_host_restatement():
    from mtamrecommender_b200.Model.base_model import base_model
    top = np.array([[5, 7, 9], [1, 2, 3], [4, 4, 4]])
    hr, nd = base_model.calculate_topK(None, 3, top, [9, 8, 4], 0, 3)
    assert abs(hr - 2 / 3) < 1e-12 and abs(nd - (math.log(2) / math.log(4) + 1.0) / 3) < 1e-12


def test_combine_lse_is_logsumexp_over_shards():
    """The sharded softmax's host-side combination (parallel.combine_lse): [n_shards, B] per-shard log-sum-exps -> the
    log-sum-exp over the whole catalogue, also when one shard dominates by hundreds of units."""
    import torch
    from mtamrecommender_b200.parallel import combine_lse, shard_rows
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(6, 1000, generator=g, dtype=torch.float64) * 5
    logits[0, 3] = 400.0
    S = shard_rows(1000, 8)
    per = torch.stack([torch.logsumexp(logits[:, r * S:(r + 1) * S], dim=1) for r in range(8)])
    assert torch.allclose(combine_lse(per), torch.logsumexp(logits, dim=1), rtol=1e-13, atol=0)
    assert shard_rows(10_000_003, 8) == 1_250_001 and shard_rows(7, 8) == 1


def test_bench_stdout_carries_only_the_json_line(tmp_path):
    """bench.run_cuda points file descriptor 1 at stderr while the arm runs: output of native libraries (NCCL's banner)
    and stray prints end up on stderr, the JSON line alone on stdout."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "t.py"
    script.write_text(
        "import ctypes, sys\n"
        f"sys.path.insert(0, {root!r})\n"
        "import bench\n"
        "def fake(args, w):\n"
        "    libc = ctypes.CDLL(None); libc.puts(b'banner from a C library'); libc.fflush(None)\n"
        "    print('python-level noise')\n"
        "    return '{\"ok\": 1}'\n"
        "bench._run_cuda = fake\n"
        "bench.run_cuda(None, None)\n")
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == '{"ok": 1}'
    assert "banner from a C library" in r.stderr and "python-level noise" in r.stderr
