"""CPU tests of the host-side mirror of the reference surface (no GPU): feed packing, DataInput, flags and
presets, synthetic generator invariants."""
import math

import os

import os

import numpy as np
import pytest

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


def test_lr_schedule_matches_the_reference_rule():
    """train_process.py:154-159, 324-336: lr1 = FLAGS.lr * 0.99^floor(gs/100) while the CURRENT rate is above 0.001,
    else lr2 = 0.001 * FLAGS.decay_rate^floor(gs/100); the oracle restates the same rule independently."""
    from oracle import mtam_oracle as O
    from mtamrecommender_b200.train_process import exponential_decay, lr_schedule
    assert exponential_decay(0.1, 250, 100, 0.5, True) == pytest.approx(0.025)
    assert exponential_decay(0.1, 250, 100, 0.5, False) == pytest.approx(0.1 * 0.5 ** 2.5, rel=1e-6)
    for flags_lr, decay in ((0.001, 0.995), (0.01, 0.995), (0.001, 0.001), (0.0005, 0.9)):
        cur = flags_lr
        for gs in range(0, 1200, 7):
            want = O.lr_schedule(flags_lr, decay, gs, cur)
            got = lr_schedule(cur, gs, flags_lr, decay)
            assert got == pytest.approx(want, rel=2e-6), (flags_lr, decay, gs)
            cur = got
    # the flag default decay_rate = 0.001 collapses the rate by 1000x every 100 steps
    assert lr_schedule(0.001, 100, 0.001, 0.001) == pytest.approx(1e-6, rel=1e-5)
    # a rate above 0.001 follows lr1 until lr1 itself has decayed to 0.001 or below, then switches to lr2 for good
    cur, seen_lr2 = 0.0011, False
    for gs in range(0, 2000, 50):
        cur = lr_schedule(cur, gs, 0.0011, 0.995)
        seen_lr2 |= cur <= 0.001
    assert seen_lr2


def test_checkpoint_layout_and_summary_writer(tmp_path):
    from mtamrecommender_b200.util import checkpoint as ck, summary as tb
    d = str(tmp_path / "m")
    t1 = {"a/kernel": np.arange(6, dtype=np.float32).reshape(2, 3), "a/kernel/Adam": np.ones((2, 3), np.float32),
          "beta1_power": np.asarray(0.9, np.float32)}
    for step in range(1, 8):
        ck.save(os.path.join(d, f"model.ckpt-{step}"), t1, {"adam_step": step})
    assert ck.latest_checkpoint(d).endswith("model.ckpt-7")
    state = open(os.path.join(d, "checkpoint")).read().splitlines()
    assert state[0] == 'model_checkpoint_path: "model.ckpt-7"' and len(state) == 6          # max_to_keep = 5
    assert not os.path.exists(os.path.join(d, "model.ckpt-2.index")) and os.path.exists(os.path.join(d, "model.ckpt-3.index"))
    back = ck.load(ck.latest_checkpoint(d))
    assert set(back) == set(t1) and all(np.array_equal(back[k], t1[k]) for k in t1)
    assert ck.latest_checkpoint(str(tmp_path / "none")) is None
    w = tb.FileWriter(str(tmp_path / "tb"))
    w.add_summary(tb.scalars([("Training Loss", 1.5), ("l2_norm", 2.0)]), 3)
    w.add_summary(None, 4)
    w.close()
    import json
    lines = [json.loads(x) for x in open(tmp_path / "tb" / "events.jsonl")]
    assert [(x["step"], x["tag"], x["value"]) for x in lines] == [(3, "Training Loss", 1.5), (3, "l2_norm", 2.0)]
