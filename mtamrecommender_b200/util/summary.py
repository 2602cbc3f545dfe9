"""Scalar summaries: the stand-in for tf.Summary / tf.summary.FileWriter.

Reference: Model/base_model.py:274-287 (`summery`: merged summaries, train / eval writers under
data/tensorboard_result/<type>_<experiment_type>_<version>_<time>/tensorboard_{train,eval}), :324-327 (the four
training scalars) and train_process.py:290-297 (recall@k / ndgc@k written through `model.train_writer`).

No TensorFlow here, so the event files are JSON lines -- one object {"wall_time", "step", "tag", "value"} per scalar in
`events.jsonl` inside the writer's directory -- which TensorBoard does not read but every plotting tool does.  The
directory layout and the tag names are the reference's.
"""
from __future__ import annotations

import json
import os
import time
from typing import List, Optional


class Summary:
    """tf.Summary(value=[tf.Summary.Value(tag=..., simple_value=...)])."""

    class Value:
        def __init__(self, tag: str, simple_value: float):
            self.tag, self.simple_value = str(tag), float(simple_value)

        def __repr__(self):
            return f"Value(tag={self.tag!r}, simple_value={self.simple_value!r})"

    def __init__(self, value: Optional[List["Summary.Value"]] = None):
        self.value = list(value or [])

    def __repr__(self):
        return f"Summary({self.value!r})"


def scalars(pairs) -> Summary:
    return Summary([Summary.Value(t, v) for t, v in pairs])


class FileWriter:
    """tf.summary.FileWriter(logdir): `add_summary(summary, global_step)`.  The file is created on the first event."""

    def __init__(self, logdir: Optional[str]):
        self.logdir = logdir
        self.events = []            # (step, tag, value): kept in memory too (tests, notebooks)
        self._f = None

    def get_logdir(self):
        return self.logdir

    def add_summary(self, summary, global_step=None):
        if summary is None:
            return
        now = time.time()
        for v in getattr(summary, "value", []):
            self.events.append((global_step, v.tag, v.simple_value))
            if self.logdir is not None:
                if self._f is None:
                    os.makedirs(self.logdir, exist_ok=True)
                    self._f = open(os.path.join(self.logdir, "events.jsonl"), "a")
                self._f.write(json.dumps({"wall_time": now, "step": None if global_step is None else int(global_step),
                                          "tag": v.tag, "value": v.simple_value}) + "\n")

    def flush(self):
        if self._f is not None:
            self._f.flush()

    def close(self):
        if self._f is not None:
            self._f.close()
            self._f = None
