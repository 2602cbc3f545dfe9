"""Vectorised synthetic batches with the reference's record construction:
history of `len-1` real events followed by the mask step (item_count+1 / category_count+1,
time = target time, timelast = timenow = 0), positions 0..len-1, right-padded with 0 to L.
Reference for the construction: Prepare/prepare_data_base.py:252-314, Prepare/mask_data_process.py:245-255.
"""
from __future__ import annotations

import numpy as np


class ZipfSampler:
    def __init__(self, n: int, s: float = 1.05):
        r = np.arange(1, n + 1, dtype=np.float64)
        p = r ** (-s)
        self.cdf = np.cumsum(p / p.sum())
        self.n = n

    def sample(self, rng, size):
        return np.minimum(np.searchsorted(self.cdf, rng.random(size)), self.n - 1).astype(np.int32)


def synth_feed(B: int, L: int, item_count: int, category_count: int, user_count: int, seed: int = 1234,
               sampler: ZipfSampler | None = None, uniform_items: bool = False):
    rng = np.random.default_rng(seed)
    length = rng.integers(2, L + 1, size=B).astype(np.int32)          # seq_length in [2, L]
    h = length - 1                                                    # real history events
    ar = np.arange(L)[None, :]
    real = ar < h[:, None]
    mask_pos = ar == h[:, None]
    if uniform_items or sampler is None:
        items = rng.integers(0, item_count, size=(B, L)).astype(np.int32)
        tgt = rng.integers(0, item_count, size=B).astype(np.int32)
    else:
        items = sampler.sample(rng, (B, L))
        tgt = sampler.sample(rng, B)
    cats = (items % max(category_count, 1)).astype(np.int32)
    gaps = rng.geometric(1.0 / 24.0, size=(B, L + 1)).astype(np.float64)
    start = rng.integers(300000, 400000, size=B).astype(np.float64)
    times = start[:, None] + np.cumsum(gaps[:, :L], axis=1)
    last_t = np.take_along_axis(times, np.maximum(h - 1, 0)[:, None].astype(np.int64), 1)[:, 0]
    gap_t = np.take_along_axis(gaps, h[:, None].astype(np.int64), 1)[:, 0]
    target_time = last_t + gap_t
    timelast = np.concatenate([np.zeros((B, 1)), np.diff(times, axis=1)], axis=1)
    timenow = target_time[:, None] - times
    f = {
        "user_id": rng.integers(0, user_count, size=B).astype(np.int32),
        "item_list": np.where(real, items, 0).astype(np.int32),
        "category_list": np.where(real, cats, 0).astype(np.int32),
        "position_list": np.where(ar < length[:, None], ar, 0).astype(np.int32),
        "time_list": np.where(real, times, 0.0),
        "timelast_list": np.where(real, timelast, 0.0).astype(np.float32),
        "timenow_list": np.where(real, timenow, 0.0).astype(np.float32),
        "target_item_id": tgt,
        "target_item_category": (tgt % max(category_count, 1)).astype(np.int32),
        "target_item_time": target_time.astype(np.float32),
        "seq_length": length,
    }
    f["item_list"][mask_pos] = item_count + 1
    f["category_list"][mask_pos] = category_count + 1
    f["time_list"] = np.where(mask_pos, target_time[:, None], f["time_list"]).astype(np.float32)
    return f


def feed_to_records(feed):
    """Feed arrays -> list of the reference's 9-tuples (for the Model.train(batch_data) surface)."""
    recs = []
    for b in range(len(feed["user_id"])):
        n = int(feed["seq_length"][b])
        recs.append((int(feed["user_id"][b]), list(feed["item_list"][b, :n]), list(feed["category_list"][b, :n]),
                     list(feed["time_list"][b, :n]), list(feed["timelast_list"][b, :n]),
                     list(feed["timenow_list"][b, :n]), list(feed["position_list"][b, :n]),
                     [int(feed["target_item_id"][b]), int(feed["target_item_category"][b]),
                      float(feed["target_item_time"][b])], n))
    return recs
