"""Columnar record store (DataHandle/record_store.py + mtam_pack_records): the packed form of the reference's list of
9-tuples pads to exactly the arrays make_feed_dic_new builds (Behavior_...py:146-192), round-trips through the binary
file and the reference's text form (prepare_data_base.py:79-92), and slices like the reference's list.  Host code only."""
import numpy as np
import pytest

from oracle import mtam_oracle as O
from mtamrecommender_b200.DataHandle.get_input_data import DataInput
from mtamrecommender_b200.DataHandle.record_store import PackedRecords, convert_text
from mtamrecommender_b200.Embedding.Behavior_embedding_time_aware_attention import Behavior_embedding_time_aware_attention

CFG = O.OracleConfig(kind=O.MTAM, L=11, D=32, user_count=9, item_count=40, category_count=5)


def _records(n, seed):
    return O.synth_records(CFG, n, seed)


def _same_feed(a, b):
    assert set(a) == set(b)
    for k in a:
        assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k


def test_packed_batch_equals_make_feed_dic_new_of_the_tuples():
    recs = _records(37, 1)
    emb = Behavior_embedding_time_aware_attention(True, 9, 40, 5, CFG.L)
    ref = {p.key: v for p, v in emb.make_feed_dic_new(recs).items()}
    _same_feed(ref, O.make_feed(CFG, recs))
    rs = PackedRecords.from_records(recs)
    assert len(rs) == 37 and rs[5][0] == recs[5][0] and rs[5][1] == [int(x) for x in recs[5][1]] and rs[5][8] == recs[5][8]
    _same_feed({p.key: v for p, v in emb.make_feed_dic_new(rs).items()}, ref)
    # a slice, a shuffled view, a slice of a shuffled view
    part = {p.key: v for p, v in emb.make_feed_dic_new(recs[8:21]).items()}
    _same_feed(rs[8:21].feed(CFG.L), part)
    perm = np.random.default_rng(0).permutation(37)
    sh = rs.take(perm)
    _same_feed(sh.feed(CFG.L), {p.key: v for p, v in emb.make_feed_dic_new([recs[i] for i in perm]).items()})
    _same_feed(sh[3:9].feed(CFG.L), {p.key: v for p, v in emb.make_feed_dic_new([recs[i] for i in perm[3:9]]).items()})


def test_pack_into_larger_staging_arrays_leaves_other_rows_alone():
    recs = _records(6, 2)
    rs = PackedRecords.from_records(recs)
    L, cap = CFG.L, 10
    out = {k: np.full((cap, L) if k.endswith("_list") else (cap,), 77, v.dtype) for k, v in rs.feed(L).items()}
    assert rs.pack_into(out, L) == 6
    ref = rs.feed(L)
    for k in out:
        assert np.array_equal(out[k][:6], ref[k]) and np.all(out[k][6:] == 77)


def test_binary_file_and_text_form_round_trip(tmp_path):
    recs = _records(23, 3)
    txt = tmp_path / "train_data.txt"
    with open(txt, "w") as f:                      # the reference writes str(tuple) per line (prepare_data_base.py:334-339)
        for r in recs:                             # plain Python numbers, as the reference's ETL produces
            r = (int(r[0]), [int(x) for x in r[1]], [int(x) for x in r[2]], [float(x) for x in r[3]],
                 [float(x) for x in r[4]], [float(x) for x in r[5]], [int(x) for x in r[6]],
                 [int(r[7][0]), int(r[7][1]), float(r[7][2])], int(r[8]))
            f.write(str(r) + "\n")
    rs = convert_text(str(txt), str(tmp_path / "train.mtamrec"))
    assert len(rs) == 23
    _same_feed(rs.feed(CFG.L), O.make_feed(CFG, recs))
    again = PackedRecords.load(str(tmp_path / "train.mtamrec"), mmap=False)
    for name in rs.cols:
        assert np.array_equal(np.asarray(rs.cols[name]), again.cols[name])
    # a view saved on its own holds exactly its records
    rs[4:9].save(str(tmp_path / "part.mtamrec"))
    _same_feed(PackedRecords.load(str(tmp_path / "part.mtamrec")).feed(CFG.L), O.make_feed(CFG, recs[4:9]))
    with open(tmp_path / "bad.mtamrec", "wb") as f:
        f.write(b"nope" * 32)
    with pytest.raises(ValueError):
        PackedRecords.load(str(tmp_path / "bad.mtamrec"))


def test_datainput_over_a_record_store_matches_the_list():
    recs = _records(10, 4)
    rs = PackedRecords.from_records(recs)
    got = list(DataInput(rs, 4))
    assert [i for i, _ in got] == [1, 2, 3] and [len(b) for _, b in got] == [4, 4, 2]
    for (_, b), (_, lb) in zip(got, DataInput(recs, 4)):
        _same_feed(b.feed(CFG.L), O.make_feed(CFG, lb))


def test_errors_are_raised_not_swallowed():
    from mtamrecommender_b200._lib import MtamError
    rs = PackedRecords.from_records(_records(5, 5))
    with pytest.raises(MtamError):                 # a record longer than max_length_seq: np.pad raises in the reference
        rs.feed(2)
    with pytest.raises(ValueError):
        rs[2:2].feed(CFG.L)                        # empty batch
    with pytest.raises(MtamError):
        rs.take(np.array([0, 99])).feed(CFG.L)     # record number out of range
