"""The reference-facing surface (train_process.py:99-104, 326-344): Model(FLAGS, emb, sess), model.train, model.metrics_topK,
fed by DataInput over the reference's list of 9-tuples and over the columnar record store -- same losses, same
parameters, same metrics; and the first losses equal the CPU oracle's."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _flags(**kw):
    from mtamrecommender_b200.config.model_parameter import model_parameter
    F = model_parameter([]).flags.FLAGS
    F.type, F.experiment_type, F.version = "synthetic", "MTAM", "t"
    F.num_units, F.num_blocks, F.num_heads, F.length_of_user_history = 64, 2, 1, 12
    F.train_batch_size, F.test_batch_size, F.dropout = 16, 16, 0.0
    F.summary_dir = ""          # scalars stay in memory unless a test asks for files
    for k, v in kw.items():
        setattr(F, k, v)
    return F


def _model(F, users, items, cats):
    from mtamrecommender_b200.Embedding.Behavior_embedding_time_aware_attention import Behavior_embedding_time_aware_attention
    from mtamrecommender_b200.Model.MTAMRec_model import MTAM
    from mtamrecommender_b200.session import Session
    emb = Behavior_embedding_time_aware_attention(True, users, items, cats, F.length_of_user_history)
    sess = Session("cuda:0")
    return MTAM(F, emb, sess), sess


@pytest.mark.parametrize("gemm_mode", ["fp32", "tf32x3"])
def test_train_and_eval_from_tuples_and_from_record_store(gemm_mode):
    from oracle import mtam_oracle as O
    from mtamrecommender_b200.DataHandle.get_input_data import DataInput
    from mtamrecommender_b200.DataHandle.record_store import PackedRecords
    users, items, cats = 30, 400, 7
    F = _flags(gemm_mode=gemm_mode)
    cfg = O.OracleConfig(kind=O.MTAM, L=12, D=64, H=1, N=2, user_count=users, item_count=items, category_count=cats)
    recs = O.synth_records(cfg, 40, 9)                       # 2 full batches and a short one
    rs = PackedRecords.from_records(recs)
    m1, s1 = _model(F, users, items, cats)
    m2, s2 = _model(F, users, items, cats)
    P = m1.engine.get_params()
    m2.engine.set_params(P)
    tr = O.OracleTrainer(cfg, {k: v.copy() for k, v in P.items()})
    lr = 1e-3
    for (i, b1), (_, b2) in zip(DataInput(recs, 16), DataInput(rs, 16)):
        l1, _ = m1.train(s1, b1, lr)
        l2, _ = m2.train(s2, b2, lr)
        lo = tr.train_step(O.make_feed(cfg, b1), lr)
        assert l1 == l2, (i, l1, l2)
        assert abs(float(l1) - lo) <= 2e-5 * abs(lo), (i, l1, lo)
    assert bool((m1.engine.params == m2.engine.params).all())
    a = m1.metrics_topK(s1, recs[:16], 0, 50)
    b = m2.metrics_topK(s2, rs[:16], 0, 50)
    assert a == b and len(a) == 10
    ref, _, _ = O.metrics_topk(cfg, tr.params, O.make_feed(cfg, recs[:16]))
    assert np.allclose(np.array(a), np.array(ref), atol=1e-6)


def test_train_process_driver_schedule_summaries_and_checkpoint(tmp_path):
    """The `train_process.py` drop-in (Train_main_process.train): the learning rate of every step follows
    train_process.py:154-159/324-336, the training scalars and recall@k / ndgc@k reach the train writer under the
    reference's tag names, and the final checkpoint -- TF-Saver file naming -- restores into a fresh model that then
    continues bit-identically."""
    import os
    from oracle import mtam_oracle as O
    from mtamrecommender_b200.train_process import Train_main_process
    users, items, cats = 30, 400, 7
    cfg = O.OracleConfig(kind=O.MTAM, L=12, D=64, H=1, N=2, user_count=users, item_count=items, category_count=cats)
    train, test = O.synth_records(cfg, 70, 3), O.synth_records(cfg, 20, 4)
    ck = str(tmp_path / "ck")
    F = _flags(max_epochs=2, eval_freq=3, learning_rate=0.002, decay_rate=0.5, checkpoint_path_dir=ck,
               summary_dir=str(tmp_path / "tb"))
    drv = Train_main_process(FLAGS=F, train_set=train, test_set=test, user_count=users, item_count=items, category_count=cats)
    drv.train()
    steps = drv.global_step
    assert steps == 2 * 5                                              # 70 records / 16 = 5 batches per epoch
    # oracle's restatement of the rule, driven the same way (reset at each epoch start, global step never reset)
    want, gs = [], 0
    for epoch in range(2):
        cur = F.learning_rate
        for _ in range(5):
            cur = O.lr_schedule(F.learning_rate, F.decay_rate, gs, cur)
            want.append(cur); gs += 1
    assert np.allclose(drv.learning_rates, want, rtol=1e-6)
    tags = {t for _, t, _ in drv.model.train_writer.events}
    assert {"Training Loss", "normalized Training Loss", "l2_norm", "Learning_rate", "recall@1", "ndgc@50"} <= tags
    assert os.path.exists(os.path.join(drv.model.train_writer.get_logdir(), "events.jsonl"))
    # TF-Saver layout
    assert sorted(os.listdir(ck)) == ["checkpoint", f"model.ckpt-{steps}.data-00000-of-00001", f"model.ckpt-{steps}.index",
                                      f"model.ckpt-{steps}.meta"]
    assert open(os.path.join(ck, "checkpoint")).read().startswith(f'model_checkpoint_path: "model.ckpt-{steps}"')
    F2 = _flags(load_type="full", checkpoint_path_dir=ck)
    m2, s2 = _model(F2, users, items, cats)
    assert bool((m2.engine.params == drv.model.engine.params).all()) and m2.engine.adam_step() == steps
    l1, _ = drv.model.train(drv.sess, train[:16], 1e-3)
    l2, _ = m2.train(s2, train[:16], 1e-3)
    assert l1 == l2 and bool((m2.engine.params == drv.model.engine.params).all())
    with pytest.raises(ValueError):
        _model(_flags(load_type="sideways"), users, items, cats)
    with pytest.raises(NotImplementedError):
        _model(_flags(optimizer="rmsprop"), users, items, cats)


@pytest.mark.parametrize("graph", [False, True])
def test_pipelined_training_equals_the_blocking_calls(graph):
    """engine.FeedPipeline / model.train_submit (the input pipeline of SURVEY 8f row 1: padding and the host->device
    copy of step i+1 overlap step i, the loss comes back one call later): the same losses, in the same order, and
    bit-identical weights as model.train called step by step -- from 9-tuples, from the record store, with ragged
    batch sizes (eager steps) and with the captured CUDA graph; and Train_main_process gives the same result with
    FLAGS.pipeline_input on and off."""
    from oracle import mtam_oracle as O
    from mtamrecommender_b200.DataHandle.get_input_data import DataInput
    from mtamrecommender_b200.DataHandle.record_store import PackedRecords
    from mtamrecommender_b200.train_process import Train_main_process
    users, items, cats = 30, 400, 7
    F = _flags()
    cfg = O.OracleConfig(kind=O.MTAM, L=12, D=64, H=1, N=2, user_count=users, item_count=items, category_count=cats)
    recs = O.synth_records(cfg, 16 * 7 + (0 if graph else 5), 11)
    rs = PackedRecords.from_records(recs)
    m1, s1 = _model(F, users, items, cats)
    m2, _ = _model(F, users, items, cats)
    m3, _ = _model(F, users, items, cats)
    P = m1.engine.get_params()
    m2.engine.set_params(P)
    m3.engine.set_params(P)
    if graph:
        for m in (m1, m2, m3):
            m.engine.capture_train_graph(16)
    want, got2, got3 = [], [], []
    for i, b in DataInput(recs, 16):
        want.append(float(m1.train(s1, b, 1e-3 * (1 + i))[0]))
    for n, ((i, b2), (_, b3)) in enumerate(zip(DataInput(recs, 16), DataInput(rs, 16))):
        for m, b, got in ((m2, b2, got2), (m3, b3, got3)):
            r = m.train_submit(b, 1e-3 * (1 + i))
            assert (r is None) == (n == 0)
            if r is not None:
                got.append(float(r[0]))
    got2.append(float(m2.train_flush()[0]))
    got3.append(float(m3.train_flush()[0]))
    assert m2.train_flush() is None
    assert want == got2 == got3
    assert bool((m1.engine.params == m2.engine.params).all()) and bool((m1.engine.params == m3.engine.params).all())
    assert bool((m1.engine.adam_m == m2.engine.adam_m).all())
    # a second round after the flush (the pipeline starts over), then the blocking call again: still in step
    for m in (m2, m3):
        assert m.train_submit(recs[:16], 2e-3) is None
        m.train_flush()
    m1.train(s1, recs[:16], 2e-3)
    assert float(m1.train(s1, recs[16:32], 1e-3)[0]) == float(m2.train(None, recs[16:32], 1e-3)[0])
    if graph:
        return
    res = []
    for pipelined in (True, False):
        Fd = _flags(max_epochs=2, eval_freq=3, learning_rate=0.002, decay_rate=0.5, pipeline_input=pipelined)
        drv = Train_main_process(FLAGS=Fd, train_set=recs[:70], test_set=recs[70:90], user_count=users, item_count=items,
                                 category_count=cats)
        last = drv.train()
        ev = [(s, t, v) for s, t, v in drv.model.train_writer.events if t == "normalized Training Loss"]
        res.append((float(last), drv.global_step, ev, drv.model.engine.params.clone()))
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1] == 10 and res[0][2] == res[1][2] and len(res[0][2]) == 10
    assert bool((res[0][3] == res[1][3]).all())
