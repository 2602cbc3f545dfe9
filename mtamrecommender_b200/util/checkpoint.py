"""Checkpoint files in the layout tf.train.Saver leaves behind (Model/base_model.py:124-147).

`saver.save(sess, path/"model.ckpt", global_step=N)` writes `model.ckpt-N.index`, `model.ckpt-N.data-00000-of-00001`,
`model.ckpt-N.meta` and updates the text file `checkpoint` (`model_checkpoint_path: "model.ckpt-N"` +
`all_model_checkpoint_paths`), which `tf.train.latest_checkpoint(path)` reads back.  The same files are written here,
under the reference's variable names with Adam's slots as `<var>/Adam`, `<var>/Adam_1` and the scalars
`beta1_power`, `beta2_power`:
  * `.data-00000-of-00001`: the tensors, raw little-endian fp32, back to back (64-byte aligned) -- as in a TF bundle;
  * `.index`: JSON {name: {shape, dtype, offset, nbytes}} instead of TF's SSTable of BundleEntryProto records;
  * `.meta`: JSON of the model configuration instead of a MetaGraphDef (there is no graph).
So the naming, the directory protocol and the tensor bytes are the reference's; the two small index files are not
byte-compatible with TensorFlow's readers (an intentional break: no protobuf / SSTable code in this library).
"""
from __future__ import annotations

import json
import os
import re
from typing import Dict, Optional

import numpy as np


def _state_path(ckpt_dir: str) -> str:
    return os.path.join(ckpt_dir, "checkpoint")


def latest_checkpoint(ckpt_dir: str) -> Optional[str]:
    """tf.train.latest_checkpoint: the prefix named by `model_checkpoint_path` in <dir>/checkpoint, or None."""
    try:
        txt = open(_state_path(ckpt_dir)).read()
    except OSError:
        return None
    m = re.search(r'^model_checkpoint_path:\s*"([^"]+)"', txt, flags=re.M)
    if not m:
        return None
    p = m.group(1)
    return p if os.path.isabs(p) else os.path.join(ckpt_dir, p)


def save(prefix: str, tensors: Dict[str, np.ndarray], meta: dict, max_to_keep: int = 5) -> str:
    """Writes <prefix>.{index,data-00000-of-00001,meta} and updates the directory's `checkpoint` file."""
    d = os.path.dirname(prefix) or "."
    os.makedirs(d, exist_ok=True)
    index, off = {}, 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in sorted(tensors):
            a = np.asarray(tensors[name])
            shape = list(a.shape)                       # (ascontiguousarray turns a scalar into shape (1,))
            a = np.ascontiguousarray(a.astype("<f4") if a.dtype.kind == "f" else a.astype("<i8"))
            pad = (-off) % 64
            f.write(b"\0" * pad)
            off += pad
            index[name] = {"shape": shape, "dtype": "float32" if a.dtype.kind == "f" else "int64",
                           "offset": off, "nbytes": int(a.nbytes)}
            f.write(a.tobytes())
            off += a.nbytes
    with open(prefix + ".index", "w") as f:
        json.dump(index, f)
    with open(prefix + ".meta", "w") as f:
        json.dump(meta, f)
    base = os.path.basename(prefix)
    older = []
    try:
        older = re.findall(r'^all_model_checkpoint_paths:\s*"([^"]+)"', open(_state_path(d)).read(), flags=re.M)
    except OSError:
        pass
    keep = [p for p in older if p != base][-(max_to_keep - 1):] + [base] if max_to_keep > 1 else [base]
    for p in older:                       # tf.train.Saver(max_to_keep=5) deletes what falls off the list
        if p not in keep:
            for ext in (".index", ".data-00000-of-00001", ".meta"):
                try:
                    os.remove(os.path.join(d, p + ext))
                except OSError:
                    pass
    with open(_state_path(d), "w") as f:
        f.write(f'model_checkpoint_path: "{base}"\n')
        for p in keep:
            f.write(f'all_model_checkpoint_paths: "{p}"\n')
    return prefix


def load(prefix: str) -> Dict[str, np.ndarray]:
    index = json.load(open(prefix + ".index"))
    data = np.memmap(prefix + ".data-00000-of-00001", dtype=np.uint8, mode="r")
    out = {}
    for name, e in index.items():
        dt = "<f4" if e["dtype"] == "float32" else "<i8"
        out[name] = np.frombuffer(data, dtype=dt, count=int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1,
                                  offset=e["offset"]).reshape(e["shape"]).copy()
    return out
