// Host-side record store -> feed arrays (no device code: the target buffers are the caller's pinned staging arrays).
//
// Reference: Embedding/Behavior_embedding_time_aware_attention.py:146-192 (make_feed_dic_new): every example is a
// 9-tuple (user, items, cats, times, timelast, timenow, positions, [target id, cat, time], length) whose six lists
// are right-padded with 0 to max_length_seq by np.pad, one Python call per list per example.  The reference keeps a
// data set as a Python list of such tuples read back with eval() (Prepare/prepare_data_base.py:79-92).  Here a data
// set is a columnar store (CSR offsets + one array per field, include/mtam.h: mtam_record_store) and a batch is
// packed straight into the step's pinned feed buffers in one pass.
#include <string.h>

#include "common.cuh"
#include "../../include/mtam.h"

extern "C" int mtam_pack_records(const mtam_record_store* rs, const int64_t* index, int64_t first, int32_t B, int32_t L,
                                 const mtam_batch* out) {
  using mtam::set_error;
  if (!rs || !out || B < 1 || L < 1) return set_error(MTAM_ERR_INVALID, "mtam_pack_records: bad argument");
  if (!rs->offsets || !rs->user_id || !rs->item || !rs->category || !rs->position || !rs->time || !rs->timelast ||
      !rs->timenow || !rs->target_item_id || !rs->target_item_category || !rs->target_item_time || !rs->seq_length)
    return set_error(MTAM_ERR_INVALID, "mtam_pack_records: record store has a null column");
  if (!index && (first < 0 || first + B > rs->n_records))
    return set_error(MTAM_ERR_INVALID, "mtam_pack_records: records [%lld, %lld) outside [0, %lld)", (long long)first,
                     (long long)(first + B), (long long)rs->n_records);
  int32_t* o_user = const_cast<int32_t*>(out->user_id);
  int32_t* o_item = const_cast<int32_t*>(out->item_list);
  int32_t* o_cat = const_cast<int32_t*>(out->category_list);
  int32_t* o_pos = const_cast<int32_t*>(out->position_list);
  float* o_time = const_cast<float*>(out->time_list);
  float* o_tl = const_cast<float*>(out->timelast_list);
  float* o_tn = const_cast<float*>(out->timenow_list);
  int32_t* o_tid = const_cast<int32_t*>(out->target_item_id);
  int32_t* o_tcat = const_cast<int32_t*>(out->target_item_category);
  float* o_tt = const_cast<float*>(out->target_item_time);
  int32_t* o_len = const_cast<int32_t*>(out->seq_length);
  if (!o_user || !o_item || !o_cat || !o_pos || !o_time || !o_tl || !o_tn || !o_tid || !o_tcat || !o_tt || !o_len)
    return set_error(MTAM_ERR_INVALID, "mtam_pack_records: output batch has a null array");
  for (int32_t b = 0; b < B; ++b) {
    const int64_t r = index ? index[b] : first + b;
    if (r < 0 || r >= rs->n_records)
      return set_error(MTAM_ERR_INVALID, "mtam_pack_records: record %lld outside [0, %lld)", (long long)r,
                       (long long)rs->n_records);
    const int64_t o = rs->offsets[r];
    const int64_t n = rs->offsets[r + 1] - o;
    if (n < 0 || n > L)   // np.pad with a negative width raises in the reference
      return set_error(MTAM_ERR_INVALID, "record %lld has %lld steps, max_length_seq is %d", (long long)r, (long long)n, L);
    const size_t row = (size_t)b * L, nb = (size_t)n * 4, zb = (size_t)(L - n) * 4;
    memcpy(o_item + row, rs->item + o, nb);         memset(o_item + row + n, 0, zb);
    memcpy(o_cat + row, rs->category + o, nb);      memset(o_cat + row + n, 0, zb);
    memcpy(o_pos + row, rs->position + o, nb);      memset(o_pos + row + n, 0, zb);
    memcpy(o_time + row, rs->time + o, nb);         memset(o_time + row + n, 0, zb);
    memcpy(o_tl + row, rs->timelast + o, nb);       memset(o_tl + row + n, 0, zb);
    memcpy(o_tn + row, rs->timenow + o, nb);        memset(o_tn + row + n, 0, zb);
    o_user[b] = rs->user_id[r];
    o_tid[b] = rs->target_item_id[r];
    o_tcat[b] = rs->target_item_category[r];
    o_tt[b] = rs->target_item_time[r];
    o_len[b] = rs->seq_length[r];
  }
  return 0;
}
