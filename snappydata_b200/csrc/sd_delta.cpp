// sd_delta.cpp -- ColumnDeltaEncoder.merge (enc/ColumnDeltaEncoder.scala:348-556), the write-side counterpart of the delta
// overlay the scan kernels apply: an UPDATE's new delta is merged with what the table already holds for that column of the
// batch -- another delta (depth folding: ColumnDelta.scala:244-301 keeps <= 2 levels in play for a scan) or the full column
// (the deltas are folded into the base column when a batch is compacted).
//
//   * two-way merge of the ascending positions with duplicate elimination; on equal positions the NEW delta wins (:432-470);
//   * the result is re-encoded with the type's default encoder (ColumnEncoding.getColumnEncoder, :395: STRING -> Dictionary
//     in first-seen order of the merged stream, BOOLEAN -> BooleanBitSet, everything else Uncompressed), nullable iff either
//     input has nulls (:394); a merged delta keeps the [numBaseRows][numDeltas][positions] section after the null words
//     (writeHeader, :300-331), a merged full column has none.
//
// Host code: a delta holds at most a few thousand entries (ColumnDelta.INIT_SIZE = 100 at depth 0, ~1.3 k at depth 1 for a
// 200 k-row batch) and the reference performs this merge on the CPU inside the UPDATE; the merged buffer is then re-put
// into the device store (sd_store_put_batch) like any other delta.  Byte layout = snappydata_b200/column_format.py
// (encode_delta / encode_column), which the tests compare against.
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "sd_host.h"

namespace {

using namespace sd;

struct DV { bool isnull = false; int64_t i = 0; uint64_t raw = 0; std::string s; };

inline int32_t rd32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }

int width_of(int t) {
  switch (t) {
    case SD_BOOLEAN: case SD_BYTE: return 1;
    case SD_SHORT: return 2;
    case SD_INT: case SD_DATE: case SD_FLOAT: return 4;
    case SD_LONG: case SD_TIMESTAMP: case SD_DOUBLE: case SD_DECIMAL: return 8;
  }
  return 0;
}

// decode `n` logical entries whose null bits are `words` (bit j = entry j is NULL) and whose encoded values start at `body`
int decode_entries(const uint8_t* buf, int64_t len, int type_id, const uint8_t* words, int nwords, int64_t body, int type, int n,
                   std::vector<DV>& out) {
  out.assign((size_t)n, DV());
  int nn = 0;
  for (int j = 0; j < n; j++) {
    const int w = j >> 6;
    uint64_t word = 0;
    if (w < nwords) memcpy(&word, words + 8 * (size_t)w, 8);
    out[j].isnull = (word >> (j & 63)) & 1;
    if (!out[j].isnull) nn++;
  }
  const uint8_t* p = buf + body;
  const uint8_t* end = buf + len;
  auto need = [&](int64_t bytes) { return p + bytes <= end; };
  std::vector<std::string> dict_s;
  std::vector<uint64_t> dict_f;
  if (type_id == ENC_DICTIONARY || type_id == ENC_BIG_DICTIONARY) {
    if (!need(4)) return set_error(SD_ERR_INVALID, "delta merge: truncated dictionary");
    const int nd = rd32(p); p += 4;
    if (nd < 0) return set_error(SD_ERR_INVALID, "delta merge: bad dictionary size");
    for (int k = 0; k < nd; k++) {
      if (type == SD_STRING) {
        if (!need(4)) return set_error(SD_ERR_INVALID, "delta merge: truncated dictionary");
        const int l = rd32(p);
        if (l < 0 || !need(4 + (int64_t)l)) return set_error(SD_ERR_INVALID, "delta merge: bad dictionary entry");
        dict_s.emplace_back(reinterpret_cast<const char*>(p + 4), (size_t)l); p += 4 + l;
      } else {
        const int w = (type == SD_INT || type == SD_DATE) ? 4 : 8;
        if (!need(w)) return set_error(SD_ERR_INVALID, "delta merge: truncated dictionary");
        uint64_t v = 0; memcpy(&v, p, (size_t)w);
        if (w == 4) v = (uint64_t)(int64_t)(int32_t)v;
        dict_f.push_back(v); p += w;
      }
    }
    const int iw = type_id == ENC_DICTIONARY ? 2 : 4;
    if (!need((int64_t)nn * iw)) return set_error(SD_ERR_INVALID, "delta merge: truncated dictionary indexes");
    for (int j = 0, k = 0; j < n; j++) {
      if (out[j].isnull) continue;
      int idx;
      if (iw == 2) { int16_t x; memcpy(&x, p + 2 * (size_t)k, 2); idx = x; } else idx = rd32(p + 4 * (size_t)k);
      k++;
      if (idx < 0 || idx >= (int)(type == SD_STRING ? dict_s.size() : dict_f.size())) return set_error(SD_ERR_INVALID, "delta merge: dictionary index out of range");
      if (type == SD_STRING) out[j].s = dict_s[(size_t)idx]; else out[j].raw = dict_f[(size_t)idx];
    }
    return 0;
  }
  if (type_id == ENC_BOOLEAN_BITSET) {
    if (!need(((int64_t)nn + 63) / 64 * 8)) return set_error(SD_ERR_INVALID, "delta merge: truncated bit set");
    for (int j = 0, k = 0; j < n; j++) {
      if (out[j].isnull) continue;
      uint64_t word; memcpy(&word, p + 8 * (size_t)(k >> 6), 8);
      out[j].raw = (word >> (k & 63)) & 1;
      k++;
    }
    return 0;
  }
  if (type_id != ENC_UNCOMPRESSED) return set_error(SD_ERR_UNSUPPORTED, "delta merge: encoding %d", type_id);
  if (type == SD_STRING) {
    for (int j = 0; j < n; j++) {
      if (out[j].isnull) continue;
      if (!need(4)) return set_error(SD_ERR_INVALID, "delta merge: truncated strings");
      const int l = rd32(p);
      if (l < 0 || !need(4 + (int64_t)l)) return set_error(SD_ERR_INVALID, "delta merge: bad string length");
      out[j].s.assign(reinterpret_cast<const char*>(p + 4), (size_t)l); p += 4 + l;
    }
    return 0;
  }
  const int w = width_of(type);
  if (!w) return set_error(SD_ERR_UNSUPPORTED, "delta merge: column type %d", type);
  if (!need((int64_t)nn * w)) return set_error(SD_ERR_INVALID, "delta merge: truncated values");
  for (int j = 0, k = 0; j < n; j++) {
    if (out[j].isnull) continue;
    uint64_t v = 0; memcpy(&v, p + (size_t)k * w, (size_t)w);
    out[j].raw = v; k++;
  }
  return 0;
}

struct Parsed { int nbase = 0; std::vector<int32_t> positions; std::vector<DV> vals; bool has_nulls = false; };

// delta buffer: header + relative null words, [numBaseRows][numDeltas][positions], pad to 8, values
int parse_delta(const uint8_t* buf, int64_t len, int type, Parsed& out) {
  if (len < 16) return set_error(SD_ERR_INVALID, "delta merge: buffer too short");
  const int type_id = rd32(buf), nb = rd32(buf + 4);
  if (type_id < 0 || nb < 0 || (nb & 7) || 16 + (int64_t)nb > len) return set_error(SD_ERR_INVALID, "delta merge: bad header");
  const uint8_t* q = buf + 8 + nb;
  out.nbase = rd32(q);
  const int n = rd32(q + 4);
  if (n < 0 || 16 + (int64_t)nb + 4ll * n > len) return set_error(SD_ERR_INVALID, "delta merge: positions truncated");
  out.positions.resize((size_t)n);
  memcpy(out.positions.data(), q + 8, 4 * (size_t)n);
  const int64_t body = ((8 + nb + 8 + 4ll * n + 7) >> 3) << 3;
  out.has_nulls = nb > 0;
  return decode_entries(buf, len, type_id, buf + 8, nb >> 3, body, type, n, out.vals);
}
// full column buffer: null bits by row ordinal
int parse_column(const uint8_t* buf, int64_t len, int type, int num_rows, Parsed& out) {
  if (len < 8) return set_error(SD_ERR_INVALID, "delta merge: buffer too short");
  const int type_id = rd32(buf), nb = rd32(buf + 4);
  if (type_id < 0 || nb < 0 || (nb & 7) || 8 + (int64_t)nb > len) return set_error(SD_ERR_INVALID, "delta merge: bad header");
  out.nbase = num_rows;
  out.has_nulls = nb > 0;
  return decode_entries(buf, len, type_id, buf + 8, nb >> 3, 8 + nb, type, num_rows, out.vals);
}

void put32(std::vector<uint8_t>& b, int32_t v) { b.insert(b.end(), reinterpret_cast<uint8_t*>(&v), reinterpret_cast<uint8_t*>(&v) + 4); }

// the type's default encoder over `vals`; as_delta: the positions section sits between the null words and the values
void encode(int type, const std::vector<DV>& vals, bool as_delta, int nbase, const std::vector<int32_t>& positions, std::vector<uint8_t>& out) {
  const int n = (int)vals.size();
  std::vector<uint64_t> words(((size_t)n + 63) / 64, 0);
  for (int j = 0; j < n; j++) if (vals[j].isnull) words[j >> 6] |= 1ull << (j & 63);
  while (!words.empty() && words.back() == 0) words.pop_back();
  std::vector<uint8_t> body;
  int type_id = ENC_UNCOMPRESSED;
  if (type == SD_STRING) {
    std::unordered_map<std::string, int> ids;
    std::vector<const std::string*> order;
    std::vector<int32_t> idx;
    for (auto& v : vals) {
      if (v.isnull) continue;
      auto it = ids.find(v.s);
      if (it == ids.end()) { it = ids.emplace(v.s, (int)order.size()).first; order.push_back(&it->first); }
      idx.push_back(it->second);
    }
    const bool big = order.size() > 32767;
    type_id = big ? ENC_BIG_DICTIONARY : ENC_DICTIONARY;
    put32(body, (int32_t)order.size());
    for (auto* sp : order) { put32(body, (int32_t)sp->size()); body.insert(body.end(), sp->begin(), sp->end()); }
    for (int32_t x : idx) { if (big) put32(body, x); else { int16_t h = (int16_t)x; body.insert(body.end(), reinterpret_cast<uint8_t*>(&h), reinterpret_cast<uint8_t*>(&h) + 2); } }
  } else if (type == SD_BOOLEAN) {
    type_id = ENC_BOOLEAN_BITSET;
    std::vector<uint64_t> bits;
    int k = 0;
    for (auto& v : vals) { if (v.isnull) continue; if ((size_t)(k >> 6) >= bits.size()) bits.push_back(0); if (v.raw) bits[k >> 6] |= 1ull << (k & 63); k++; }
    body.insert(body.end(), reinterpret_cast<uint8_t*>(bits.data()), reinterpret_cast<uint8_t*>(bits.data()) + bits.size() * 8);
  } else {
    const int w = width_of(type);
    for (auto& v : vals) { if (v.isnull) continue; body.insert(body.end(), reinterpret_cast<const uint8_t*>(&v.raw), reinterpret_cast<const uint8_t*>(&v.raw) + w); }
  }
  out.clear();
  put32(out, type_id); put32(out, (int32_t)words.size() * 8);
  out.insert(out.end(), reinterpret_cast<uint8_t*>(words.data()), reinterpret_cast<uint8_t*>(words.data()) + words.size() * 8);
  if (as_delta) {
    put32(out, nbase); put32(out, (int32_t)positions.size());
    out.insert(out.end(), reinterpret_cast<const uint8_t*>(positions.data()), reinterpret_cast<const uint8_t*>(positions.data()) + positions.size() * 4);
    while (out.size() % 8) out.push_back(0);
  }
  out.insert(out.end(), body.begin(), body.end());
}

}  // namespace

extern "C" int sd_delta_merge(const sd_column* column, const void* new_delta, int64_t new_len, const void* existing, int64_t existing_len,
                              int32_t existing_is_delta, int32_t num_rows, void* out, int64_t cap, int64_t* out_len) {
  if (!column || !new_delta || !existing || !out_len) return set_error(SD_ERR_INVALID, "sd_delta_merge: null argument");
  std::vector<uint8_t> tmp1, tmp2;
  const uint8_t* b1 = reinterpret_cast<const uint8_t*>(new_delta);
  const uint8_t* b2 = reinterpret_cast<const uint8_t*>(existing);
  if (new_len >= 8 && rd32(b1) < 0) { int rc = decompress_envelope_host(b1, new_len, tmp1); if (rc) return rc; b1 = tmp1.data(); new_len = (int64_t)tmp1.size(); }
  if (existing_len >= 8 && rd32(b2) < 0) { int rc = decompress_envelope_host(b2, existing_len, tmp2); if (rc) return rc; b2 = tmp2.data(); existing_len = (int64_t)tmp2.size(); }
  Parsed left, right;
  int rc = parse_delta(b1, new_len, column->type, left);
  if (rc) return rc;
  rc = existing_is_delta ? parse_delta(b2, existing_len, column->type, right) : parse_column(b2, existing_len, column->type, existing_is_delta ? 0 : (num_rows > 0 ? num_rows : left.nbase), right);
  if (rc) return rc;
  const bool nullable = column->nullable && (left.has_nulls || right.has_nulls);
  std::vector<DV> merged;
  std::vector<int32_t> positions;
  if (existing_is_delta) {   // union of the positions, the new delta wins on equal ones
    size_t i = 0, j = 0;
    while (i < left.positions.size() || j < right.positions.size()) {
      if (j >= right.positions.size() || (i < left.positions.size() && left.positions[i] <= right.positions[j])) {
        if (j < right.positions.size() && left.positions[i] == right.positions[j]) j++;
        positions.push_back(left.positions[i]); merged.push_back(left.vals[i]); i++;
      } else { positions.push_back(right.positions[j]); merged.push_back(right.vals[j]); j++; }
    }
  } else {                    // the delta applied to the full column
    merged = right.vals;
    for (size_t i = 0; i < left.positions.size(); i++) {
      const int32_t pos = left.positions[i];
      if (pos < 0 || pos >= (int32_t)merged.size()) return set_error(SD_ERR_INVALID, "sd_delta_merge: position %d outside the %zu-row column", pos, merged.size());
      merged[(size_t)pos] = left.vals[i];
    }
  }
  if (!nullable) for (auto& v : merged) if (v.isnull) return set_error(SD_ERR_INVALID, "sd_delta_merge: NULL in a NOT NULL column");
  std::vector<uint8_t> enc;
  encode(column->type, merged, existing_is_delta != 0, left.nbase, positions, enc);
  *out_len = (int64_t)enc.size();
  if ((int64_t)enc.size() > cap) return set_error(SD_ERR_OVERFLOW, "sd_delta_merge: output needs %zu bytes", enc.size());
  memcpy(out, enc.data(), enc.size());
  return 0;
}
