"""Columnar record store: the data set format of the B200 path.

The reference keeps a data set as a Python list of 9-tuples
`(user, items, cats, times, timelast, timenow, positions, [target id, cat, time], length)`, written one
`str(tuple)` per line and read back with `eval` (Prepare/prepare_data_base.py:79-92, 304-314, 334-339), and
`make_feed_dic_new` pads the six lists of every example with one `np.pad` each
(Embedding/Behavior_embedding_time_aware_attention.py:146-192).  Here the same records live in 12 flat arrays
(CSR offsets + one column per field; `mtam_record_store` in include/mtam.h), on disk as one little-endian binary
file that is memory-mapped, and a batch is padded by `mtam_pack_records` straight into the step's pinned feed
buffers.  `PackedRecords` slices like the reference's list (`data[a:b]`, `len(data)`), so `DataInput` and
`model.train(sess, batch, lr)` take either form.

File layout (`MTAMREC1`): 64-byte header {magic[8], u64 n_records, u64 n_steps, u32 version, pad}, then the
columns in COLUMNS order, each starting at a multiple of 64 bytes.
"""
from __future__ import annotations

import ast
import ctypes as C
import os
from typing import Dict, Iterable, Optional, Sequence

import numpy as np

from .. import _lib

MAGIC = b"MTAMREC1"
VERSION = 1
# (name, dtype, per_step)
COLUMNS = (("offsets", np.int64, None), ("user_id", np.int32, False), ("target_item_id", np.int32, False),
           ("target_item_category", np.int32, False), ("target_item_time", np.float32, False),
           ("seq_length", np.int32, False), ("item", np.int32, True), ("category", np.int32, True),
           ("position", np.int32, True), ("time", np.float32, True), ("timelast", np.float32, True),
           ("timenow", np.float32, True))
FEED_OF_STEP_COLUMN = {"item": "item_list", "category": "category_list", "position": "position_list",
                       "time": "time_list", "timelast": "timelast_list", "timenow": "timenow_list"}


def _align(n: int, a: int = 64) -> int:
    return (n + a - 1) // a * a


class PackedRecords:
    """A data set (or a view of one) in columnar form.  `cols` maps COLUMNS names to 1-D arrays."""

    def __init__(self, cols: Dict[str, np.ndarray], index: Optional[np.ndarray] = None, first: int = 0,
                 count: Optional[int] = None):
        self.cols = cols
        self.n_total = int(cols["user_id"].shape[0])
        self.index = index                       # explicit record numbers (a shuffled / gathered view) or None
        self.first = int(first)
        self.count = int(count if count is not None else (len(index) if index is not None else self.n_total))
        self._c = None

    # ---- construction ---------------------------------------------------------------------
    @classmethod
    def from_records(cls, records: Iterable[Sequence]) -> "PackedRecords":
        """From the reference's in-memory form: an iterable of 9-tuples."""
        user, tid, tcat, tt, ln, lens = [], [], [], [], [], []
        step = {k: [] for k in FEED_OF_STEP_COLUMN}
        for ex in records:
            n = len(ex[1])
            for k, j in (("item", 1), ("category", 2), ("time", 3), ("timelast", 4), ("timenow", 5), ("position", 6)):
                if len(ex[j]) != n:
                    raise ValueError("the six lists of a record must have the same length")
                step[k].extend(ex[j])
            user.append(ex[0]); tid.append(ex[7][0]); tcat.append(ex[7][1]); tt.append(ex[7][2]); ln.append(ex[8])
            lens.append(n)
        off = np.zeros(len(lens) + 1, np.int64)
        np.cumsum(np.asarray(lens, np.int64), out=off[1:])
        cols = {"offsets": off, "user_id": np.asarray(user, np.int32), "target_item_id": np.asarray(tid, np.int32),
                "target_item_category": np.asarray(tcat, np.int32), "target_item_time": np.asarray(tt, np.float32),
                "seq_length": np.asarray(ln, np.int32)}
        for name, dt, per_step in COLUMNS:
            if per_step:
                cols[name] = np.asarray(step[name], dt)
        return cls(cols)

    @classmethod
    def from_text(cls, path: str, limit: Optional[int] = None) -> "PackedRecords":
        """From the reference's on-disk form: one `str(tuple)` per line (the reference `eval`s each line;
        `ast.literal_eval` accepts the same literals without executing anything)."""
        def lines():
            with open(path) as f:
                for i, line in enumerate(f):
                    if limit is not None and i >= limit:
                        break
                    line = line.strip()
                    if line:
                        yield ast.literal_eval(line)
        return cls.from_records(lines())

    # ---- disk -------------------------------------------------------------------------------
    def save(self, path: str) -> str:
        src = self.materialise()
        n, steps = src.n_total, int(src.cols["offsets"][-1])
        with open(path, "wb") as f:
            hdr = bytearray(64)
            hdr[0:8] = MAGIC
            hdr[8:16] = np.uint64(n).tobytes()
            hdr[16:24] = np.uint64(steps).tobytes()
            hdr[24:28] = np.uint32(VERSION).tobytes()
            f.write(hdr)
            for name, dt, _ in COLUMNS:
                a = np.ascontiguousarray(src.cols[name], dtype=np.dtype(dt).newbyteorder("<"))
                f.write(a.tobytes())
                f.write(b"\0" * (_align(a.nbytes) - a.nbytes))
        return path

    @classmethod
    def load(cls, path: str, mmap: bool = True) -> "PackedRecords":
        with open(path, "rb") as f:
            hdr = f.read(64)
        if len(hdr) < 64 or hdr[0:8] != MAGIC:
            raise ValueError(f"{path}: not a {MAGIC.decode()} record store")
        n = int(np.frombuffer(hdr[8:16], np.uint64)[0])
        steps = int(np.frombuffer(hdr[16:24], np.uint64)[0])
        ver = int(np.frombuffer(hdr[24:28], np.uint32)[0])
        if ver != VERSION:
            raise ValueError(f"{path}: record store version {ver}, this build reads {VERSION}")
        cols, off = {}, 64
        size = os.path.getsize(path)
        for name, dt, per_step in COLUMNS:
            cnt = n + 1 if per_step is None else (steps if per_step else n)
            nbytes = cnt * np.dtype(dt).itemsize
            if off + nbytes > size:
                raise ValueError(f"{path}: truncated (column {name})")
            if mmap:
                cols[name] = np.memmap(path, dtype=dt, mode="r", offset=off, shape=(cnt,))
            else:
                cols[name] = np.fromfile(path, dtype=dt, count=cnt, offset=off)
            off += _align(nbytes)
        if int(cols["offsets"][0]) != 0 or int(cols["offsets"][-1]) != steps:
            raise ValueError(f"{path}: corrupt offsets")
        return cls(cols)

    # ---- list-like surface --------------------------------------------------------------------
    def __len__(self) -> int:
        return self.count

    def __getitem__(self, key):
        if isinstance(key, slice):
            a, b, st = key.indices(self.count)
            if st != 1:
                return self.take(np.arange(a, b, st))
            b = max(a, b)
            if self.index is not None:
                return PackedRecords(self.cols, index=self.index[a:b])
            return PackedRecords(self.cols, first=self.first + a, count=b - a)
        i = int(key)
        if i < 0:
            i += self.count
        if not 0 <= i < self.count:
            raise IndexError(i)
        return self.record(i)

    def take(self, positions) -> "PackedRecords":
        """A view holding the records at `positions` (of this view), e.g. a shuffled epoch order."""
        positions = np.asarray(positions, np.int64)
        base = self.index[positions] if self.index is not None else positions + self.first
        return PackedRecords(self.cols, index=np.ascontiguousarray(base, np.int64))

    def shuffled(self, seed: int) -> "PackedRecords":
        return self.take(np.random.default_rng(seed).permutation(self.count))

    def _rec_no(self, i: int) -> int:
        return int(self.index[i]) if self.index is not None else self.first + i

    def record(self, i: int):
        """Record i of this view as the reference's 9-tuple."""
        r = self._rec_no(i)
        c = self.cols
        a, b = int(c["offsets"][r]), int(c["offsets"][r + 1])
        return (int(c["user_id"][r]), c["item"][a:b].tolist(), c["category"][a:b].tolist(), c["time"][a:b].tolist(),
                c["timelast"][a:b].tolist(), c["timenow"][a:b].tolist(), c["position"][a:b].tolist(),
                [int(c["target_item_id"][r]), int(c["target_item_category"][r]), float(c["target_item_time"][r])],
                int(c["seq_length"][r]))

    def materialise(self) -> "PackedRecords":
        """A store that owns exactly this view's records, in order."""
        if self.index is None and self.first == 0 and self.count == self.n_total:
            return self
        return PackedRecords.from_records(self.record(i) for i in range(self.count))

    # ---- packing ------------------------------------------------------------------------------
    def _c_store(self):
        if self._c is None:
            s = _lib.RecordStore()
            s.n_records = self.n_total
            keep = []
            for name, dt, _ in COLUMNS:
                a = self.cols[name]
                if a.dtype != np.dtype(dt) or not a.flags["C_CONTIGUOUS"]:
                    a = np.ascontiguousarray(a, dt)
                    self.cols[name] = a
                keep.append(a)
                setattr(s, name, a.ctypes.data)
            self._c = (s, keep)
        return self._c[0]

    def pack_into(self, out: Dict[str, np.ndarray], L: int) -> int:
        """make_feed_dic_new for this view: writes the 11 feed arrays (`out[key]`: C-contiguous [B,L] / [B] with
        B >= len(self)) through mtam_pack_records.  Returns the batch size."""
        B = self.count
        if B < 1:
            raise ValueError("empty batch")
        b = _lib.Batch()
        b.B = B
        for k in ("user_id", "item_list", "category_list", "position_list", "time_list", "timelast_list",
                  "timenow_list", "target_item_id", "target_item_category", "target_item_time", "seq_length"):
            a = out[k]
            want = np.float32 if k in ("time_list", "timelast_list", "timenow_list", "target_item_time") else np.int32
            if a.dtype != want or not a.flags["C_CONTIGUOUS"] or a.shape[0] < B or (a.ndim == 2 and a.shape[1] != L):
                raise ValueError(f"feed array {k}: need C-contiguous {np.dtype(want).name} with {B}+ rows"
                                 + (f" of {L}" if a.ndim == 2 else ""))
            setattr(b, k, a.ctypes.data)
        idx = self.index.ctypes.data if self.index is not None else None
        _lib.check(_lib.load().mtam_pack_records(C.byref(self._c_store()), idx, self.first, B, int(L), C.byref(b)),
                   "mtam_pack_records")
        return B

    def feed(self, L: int) -> Dict[str, np.ndarray]:
        """The 11 feed arrays of this view as fresh numpy arrays (the dict make_feed_dic_new returns, by key name)."""
        B = self.count
        i32, f32 = np.int32, np.float32
        out = {"user_id": np.empty(B, i32), "item_list": np.empty((B, L), i32), "category_list": np.empty((B, L), i32),
               "position_list": np.empty((B, L), i32), "time_list": np.empty((B, L), f32),
               "timelast_list": np.empty((B, L), f32), "timenow_list": np.empty((B, L), f32),
               "target_item_id": np.empty(B, i32), "target_item_category": np.empty(B, i32),
               "target_item_time": np.empty(B, f32), "seq_length": np.empty(B, i32)}
        self.pack_into(out, L)
        return out


def convert_text(src: str, dst: str, limit: Optional[int] = None) -> PackedRecords:
    """Reference text data set (train_data.txt / test_data.txt) -> binary record store at `dst`."""
    rs = PackedRecords.from_text(src, limit)
    rs.save(dst)
    return PackedRecords.load(dst)
