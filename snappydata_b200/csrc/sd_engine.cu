// sd_engine.cu -- plan handles: stats-row batch skipping, per-batch descriptor/table preparation,
// kernel launch, partial-row emission, final merge.  Host-side counterpart of
//   ColumnTableScan.doProduce batch loop        core/execution/columnar/ColumnTableScan.scala:518-599
//   ColumnTableScan.generateStatPredicate       core/execution/columnar/ColumnTableScan.scala:820-963
//   SnappyHashAggregateExec partial output      core/execution/aggregate/SnappyHashAggregateExec.scala:1148-1178
//   CollectAggregateExec / final merge          core/execution/aggregate/CollectAggregateExec.scala:67-121
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>

#include "sd_host.h"

using namespace sd;

namespace {

thread_local int t_device = 0;

inline int32_t rd_i32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }
inline int16_t rd_i16(const uint8_t* p) { int16_t v; memcpy(&v, p, 2); return v; }
inline int64_t rd_i64(const uint8_t* p) { int64_t v; memcpy(&v, p, 8); return v; }
inline double rd_f64(const uint8_t* p) { double v; memcpy(&v, p, 8); return v; }
inline float rd_f32(const uint8_t* p) { float v; memcpy(&v, p, 4); return v; }

// ---- host values (stats rows, partial rows) -------------------------------------------------------
typedef __int128 i128;
struct HVal {
  bool isnull = false;
  int64_t i = 0;
  double d = 0;
  i128 w = 0;      // DECIMAL unscaled value (any precision); `i` mirrors it when the precision is <= 18
  std::string s;
};
i128 pow10_128(int k) { i128 r = 1; for (int j = 0; j < k; j++) r *= 10; return r; }

int cmp_str(const std::string& a, const std::string& b) {
  size_t n = std::min(a.size(), b.size());
  int c = n ? memcmp(a.data(), b.data(), n) : 0;
  return c ? c : (a.size() < b.size() ? -1 : (a.size() > b.size() ? 1 : 0));
}
int cmp_f64(double x, double y) {   // Utils.nanSafeCompareDoubles
  const bool xn = std::isnan(x), yn = std::isnan(y);
  if (xn || yn) return xn && yn ? 0 : (xn ? 1 : -1);
  return x < y ? -1 : (x > y ? 1 : 0);
}
int cmp_hval(const HVal& a, const HVal& b, int ft) {
  const int t = ft_base(ft);
  if (t == SD_STRING) return cmp_str(a.s, b.s);
  if (type_is_fp(t)) return cmp_f64(a.d, b.d);
  if (t == SD_DECIMAL) return a.w < b.w ? -1 : (a.w > b.w ? 1 : 0);
  return a.i < b.i ? -1 : (a.i > b.i ? 1 : 0);
}

// field `idx` of a Spark UnsafeRow with `nfields` fields
bool unsafe_field(const uint8_t* row, int64_t len, int nfields, int idx, int ftype, HVal* out) {
  const int64_t bits = ((nfields + 63) / 64) * 8;
  if (idx < 0 || idx >= nfields || bits + 8 * (int64_t)nfields > len) return false;
  *out = HVal();
  if (row[idx >> 3] & (1u << (idx & 7))) { out->isnull = true; return true; }
  const uint8_t* slot = row + bits + 8 * (int64_t)idx;
  const int type = ft_base(ftype);
  if (type == SD_DECIMAL) {   // UnsafeRow.getDecimal: long for precision <= 18, else BigInteger bytes (big-endian two's complement)
    if (ft_precision(ftype) <= 18) { out->i = rd_i64(slot); out->w = out->i; return true; }
    const int64_t ol = rd_i64(slot);
    const int64_t off = ol >> 32, ln = ol & 0xffffffff;
    if (off < 0 || ln > 16 || off + ln > len) return false;
    i128 v = (ln > 0 && (row[off] & 0x80)) ? -1 : 0;
    for (int64_t k = 0; k < ln; k++) v = (v << 8) | row[off + k];
    out->w = v; out->i = (int64_t)v;
    return true;
  }
  switch (type) {
    case SD_STRING: {
      const int64_t ol = rd_i64(slot);
      const int64_t off = ol >> 32, ln = ol & 0xffffffff;
      if (off < 0 || off + ln > len) return false;
      out->s.assign(reinterpret_cast<const char*>(row + off), (size_t)ln);
      break;
    }
    case SD_BOOLEAN: out->i = slot[0] != 0; break;
    case SD_BYTE: out->i = (int8_t)slot[0]; break;
    case SD_SHORT: out->i = rd_i16(slot); break;
    case SD_INT: case SD_DATE: out->i = rd_i32(slot); break;
    case SD_FLOAT: out->d = rd_f32(slot); break;
    case SD_DOUBLE: out->d = rd_f64(slot); break;
    default: out->i = rd_i64(slot); break;
  }
  return true;
}

void emit_unsafe_row(std::vector<uint8_t>& out, const std::vector<int>& types, const std::vector<HVal>& vals) {
  const int n = (int)types.size();
  const int64_t bits = ((n + 63) / 64) * 8, fixed = bits + 8 * (int64_t)n;
  int64_t var = 0;
  for (int i = 0; i < n; i++) {
    if (types[i] == SD_STRING && !vals[i].isnull) var += ((int64_t)vals[i].s.size() + 7) & ~int64_t(7);
    if (ft_base(types[i]) == SD_DECIMAL && ft_precision(types[i]) > 18) var += 16;   // always reserved (UnsafeRowWriter.write(Decimal))
  }
  const int64_t sz = fixed + var;
  const size_t base = out.size();
  out.resize(base + 8 + sz, 0);
  uint8_t* r = out.data() + base;
  memcpy(r, &sz, 8);
  r += 8;
  int64_t voff = fixed;
  for (int i = 0; i < n; i++) {
    uint8_t* slot = r + bits + 8 * (int64_t)i;
    if (ft_base(types[i]) == SD_DECIMAL) {
      const bool wide = ft_precision(types[i]) > 18;
      if (vals[i].isnull) {
        r[i >> 3] |= (uint8_t)(1u << (i & 7));
        if (wide) { const int64_t ol = voff << 32; memcpy(slot, &ol, 8); voff += 16; }   // offset kept, size 0
        continue;
      }
      if (!wide) { const int64_t v = (int64_t)vals[i].w; memcpy(slot, &v, 8); continue; }
      // BigInteger.toByteArray(): minimal big-endian two's complement
      uint8_t be[16];
      i128 v = vals[i].w;
      for (int k = 15; k >= 0; k--) { be[k] = (uint8_t)(v & 0xff); v >>= 8; }
      int first = 0;
      while (first < 15 && ((be[first] == 0x00 && !(be[first + 1] & 0x80)) || (be[first] == 0xff && (be[first + 1] & 0x80)))) first++;
      const int nb = 16 - first;
      memcpy(r + voff, be + first, (size_t)nb);
      const int64_t ol = (voff << 32) | (int64_t)nb;
      memcpy(slot, &ol, 8);
      voff += 16;
      continue;
    }
    if (vals[i].isnull) { r[i >> 3] |= (uint8_t)(1u << (i & 7)); continue; }
    switch (types[i]) {
      case SD_STRING: {
        const int64_t ol = (voff << 32) | (uint32_t)vals[i].s.size();
        memcpy(slot, &ol, 8);
        memcpy(r + voff, vals[i].s.data(), vals[i].s.size());
        voff += ((int64_t)vals[i].s.size() + 7) & ~int64_t(7);
        break;
      }
      case SD_BOOLEAN: slot[0] = vals[i].i != 0; break;
      case SD_BYTE: { int8_t v = (int8_t)vals[i].i; memcpy(slot, &v, 1); break; }
      case SD_SHORT: { int16_t v = (int16_t)vals[i].i; memcpy(slot, &v, 2); break; }
      case SD_INT: case SD_DATE: { int32_t v = (int32_t)vals[i].i; memcpy(slot, &v, 4); break; }
      case SD_FLOAT: { float v = (float)vals[i].d; memcpy(slot, &v, 4); break; }
      case SD_DOUBLE: memcpy(slot, &vals[i].d, 8); break;
      default: memcpy(slot, &vals[i].i, 8); break;
    }
  }
}

// ---- stats-row batch skipping (ColumnTableScan.generateStatPredicate, :820-963) -------------------
struct Tri { bool isnull; bool v; };
Tri tri_and(Tri a, Tri b) { if ((!a.isnull && !a.v) || (!b.isnull && !b.v)) return {false, false}; if (a.isnull || b.isnull) return {true, false}; return {false, true}; }
Tri tri_or(Tri a, Tri b) { if ((!a.isnull && a.v) || (!b.isnull && b.v)) return {false, true}; if (a.isnull || b.isnull) return {true, false}; return {false, false}; }
Tri tri_cmp(const HVal& a, const HVal& b, int t, bool le) {
  if (a.isnull || b.isnull) return {true, false};
  const int c = cmp_hval(a, b, t);
  return {false, le ? c <= 0 : c < 0};
}
HVal lit_val(const sd_literal& l, int t) {
  HVal v;
  v.isnull = l.is_null != 0;
  v.i = l.i;
  v.d = t == SD_FLOAT ? (double)(float)l.d : l.d;
  if (l.s && l.slen > 0) v.s.assign(l.s, (size_t)l.slen);
  return v;
}

struct StatEval {
  const PlanSpec& p;
  const std::vector<sd_literal>& lits;
  const uint8_t* stats; int64_t slen; int nfields; int num_rows;
  bool stat(int ord, int which, int type, HVal* out) const {   // which: 0 lower, 1 upper, 2 nullCount
    const int idx = 1 + 3 * ord + which;
    return idx < nfields && unsafe_field(stats, slen, nfields, idx, which == 2 ? (int)SD_INT : type, out);
  }
  // true when a stats filter is defined for `node` (buildFilter.isDefinedAt)
  bool eval(int node, Tri* out) const {
    const sd_expr& e = p.exprs[node];
    Tri l, r;
    switch (e.op) {
      case SD_OP_AND: {
        const bool dl = eval(e.a, &l), dr = eval(e.b, &r);
        if (!dl && !dr) return false;
        *out = dl && dr ? tri_and(l, r) : (dl ? l : r);
        return true;
      }
      case SD_OP_OR: {
        const bool dl = eval(e.a, &l), dr = eval(e.b, &r);
        if (!(dl && dr)) return false;
        *out = tri_or(l, r);
        return true;
      }
      case SD_OP_EQ: case SD_OP_LT: case SD_OP_LE: case SD_OP_GT: case SD_OP_GE: {
        const sd_expr &ea = p.exprs[e.a], &eb = p.exprs[e.b];
        const bool cl = ea.op == SD_OP_COL && eb.op == SD_OP_LIT, cr = eb.op == SD_OP_COL && ea.op == SD_OP_LIT;
        if (!cl && !cr) return false;
        const sd_expr& ec = cl ? ea : eb;
        const sd_expr& el = cl ? eb : ea;
        const int t = ec.type, ord = p.cols[ec.a].table_ordinal;
        HVal lo, hi;
        if (!stat(ord, 0, t, &lo) || !stat(ord, 1, t, &hi)) return false;
        const HVal lit = lit_val(lits[el.a], t);
        int op = e.op;
        if (cr) op = op == SD_OP_LT ? SD_OP_GT : op == SD_OP_LE ? SD_OP_GE : op == SD_OP_GT ? SD_OP_LT : op == SD_OP_GE ? SD_OP_LE : op;
        switch (op) {
          case SD_OP_EQ: *out = tri_and(tri_cmp(lo, lit, t, true), tri_cmp(lit, hi, t, true)); break;
          case SD_OP_LT: *out = tri_cmp(lo, lit, t, false); break;
          case SD_OP_LE: *out = tri_cmp(lo, lit, t, true); break;
          case SD_OP_GT: *out = tri_cmp(lit, hi, t, false); break;
          default: *out = tri_cmp(lit, hi, t, true); break;
        }
        return true;
      }
      case SD_OP_IN: {
        const sd_expr& ea = p.exprs[e.a];
        if (ea.op != SD_OP_COL || e.c > 200 || e.c < 1) return false;
        const int t = ea.type, ord = p.cols[ea.a].table_ordinal;
        HVal lo, hi, mn, mx;
        if (!stat(ord, 0, t, &lo) || !stat(ord, 1, t, &hi)) return false;
        bool have = false;
        for (int k = 0; k < e.c; k++) {   // Greatest / Least skip nulls
          if (lits[e.b + k].is_null) continue;
          const HVal v = lit_val(lits[e.b + k], t);
          if (!have) { mn = mx = v; have = true; }
          else { if (cmp_hval(v, mn, t) < 0) mn = v; if (cmp_hval(v, mx, t) > 0) mx = v; }
        }
        if (!have) mn.isnull = mx.isnull = true;
        *out = tri_and(tri_cmp(lo, mx, t, true), tri_cmp(mn, hi, t, true));
        return true;
      }
      case SD_OP_STARTSWITH: {   // StartsWithForStats (ColumnTableScan.scala:1028-1088); never NULL
        const sd_expr &ea = p.exprs[e.a], &eb = p.exprs[e.b];
        if (ea.op != SD_OP_COL || eb.op != SD_OP_LIT) return false;
        const int ord = p.cols[ea.a].table_ordinal;
        HVal lo, hi;
        if (!stat(ord, 0, SD_STRING, &lo) || !stat(ord, 1, SD_STRING, &hi)) return false;
        const sd_literal& L = lits[eb.a];
        Tri r0 = {false, true};
        if (!L.is_null) {
          std::string pat(L.s ? L.s : "", (size_t)std::max(0, L.slen)), up = pat;
          int last = (int)up.size() - 1;
          while (last >= 0 && (uint8_t)up[last] == 0xff) last--;
          if (last < 0 || lo.isnull) {
            if (!hi.isnull) r0.v = cmp_str(pat, hi.s) <= 0;
          } else {
            up[last] = (char)((uint8_t)up[last] + 1);
            r0.v = (hi.isnull || cmp_str(pat, hi.s) <= 0) && cmp_str(lo.s, up) < 0;
          }
        }
        *out = r0;
        return true;
      }
      case SD_OP_ISNULL: case SD_OP_ISNOTNULL: {
        const sd_expr& ea = p.exprs[e.a];
        if (ea.op != SD_OP_COL) return false;
        HVal nc;
        if (!stat(p.cols[ea.a].table_ordinal, 2, SD_INT, &nc)) return false;
        *out = nc.isnull ? Tri{true, false} : Tri{false, e.op == SD_OP_ISNULL ? nc.i > 0 : num_rows > nc.i};
        return true;
      }
    }
    return false;
  }
};

}  // namespace

// =====================================================================================================
struct sd_plan {
  int device = 0;
  PlanSpec spec;
  KernelEntry kernel;             // generic kernel of the plan
  KernelEntry kernel_reg;         // register-group-table variant (MODE_GROUPS, <= REG_GROUPS_MAX groups), lazily resolved
  int kernel_reg_state = 0;       // 0 unknown, 1 available, -1 unavailable
  // lazily resolved variants of the generic kernel, indexed by (NULL literal in this execution) | (scan has batches
  // that need the per-row decode / delta / delete paths) << 1; index 0 is `kernel` itself (staged paths only)
  KernelEntry variant[4];
  int variant_state[4] = {0, 0, 0, 0};
  bool force_hash = false;        // the dense group table was abandoned for the hash table during an execution
  const KernelEntry* active = nullptr;
  std::string kernel_name;
  int max_ctas_per_sm = 0;
  int chunk_rows = CHUNK_ROWS;
  size_t last_smem = (size_t)-1;
  const KernelEntry* last_kernel = nullptr;
  int num_sms = 0;
  int smem_optin = 0;
  // literals
  std::vector<sd_literal> lits;
  std::vector<std::string> lit_strs;
  bool lits_set = false;
  // bytes of the STRING literals on the device (raw-string predicates compare against them); slot k: packed offset | length
  uint8_t* d_litpool = nullptr;
  size_t litpool_cap = 0;
  bool litpool_dirty = true;
  std::vector<int64_t> lit_packed;
  // streams / events
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  // one event pair per launch of an execution: aggTime = the SUM of the launches' device times (host work between two
  // launches -- descriptor building, waiting for a store's lock -- is not kernel time)
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pairs;
  size_t ev_used = 0;
  // private store for sd_batch_submit / sd_rows_submit
  sd_store* priv = nullptr;
  // pending batches
  struct Pending { const StoredBatch* sb; };
  std::vector<Pending> pending;
  int64_t pending_bytes = 0;
  // device state
  Arena scratch;                 // descriptors, aux tables (reset per execution)
  // device state of one execution in ONE allocation so that it comes back in one copy:
  //   d_state = [8 x uint64 counters][result_cap x uint64 running result]
  uint64_t* d_state = nullptr;
  uint64_t* d_result = nullptr;  // = d_state + STATE_HDR
  uint64_t* h_pinned = nullptr;  // pinned host mirror of d_state: identities out, counters + result back
  size_t h_pinned_cap = 0;
  uint64_t* d_partials = nullptr;
  size_t partials_cap = 0;
  unsigned int* d_ticket = nullptr;
  unsigned long long* d_counters = nullptr;
  size_t result_cap = 0;         // entries
  int ngroups = 1;
  int32_t radix[MAX_KEYS] = {1, 1, 1, 1};
  bool result_init = false;
  // group key dictionaries (query-global ids per key column)
  std::vector<std::unordered_map<std::string, int>> key_ids;
  std::vector<std::vector<std::string>> key_vals;
  std::vector<int> key_null_id;
  // store-scan cache
  // A cached scan is a list of SEGMENTS, each a descriptor set over a run of the store's batches, launched one after the other
  // into the same result.  A store only ever appends batches (sd_store: never removed or moved), so when it has grown since
  // the last execution (ingest running beside the queries: BASELINE.json's hybrid configuration) only the new batches get
  // descriptors -- as a new segment -- instead of re-deriving all of them (11 us per batch: 3.3 ms per query at 300 batches).
  struct ScanSegment {
    const void* d_batches = nullptr; const int32_t* d_prefix = nullptr; int nbatches = 0; int total_chunks = 0; int needs_slow = 0;
    int64_t rows = 0, algo_bytes = 0, updated_cols = 0, deleted_batches = 0;
    size_t covered = 0;                          // batches of the store's snapshot this segment looked at (passing or skipped)
    std::vector<const StoredBatch*> batches;     // those that pass the stats check, in order
  };
  struct ScanCache {
    const sd_store* store = nullptr; int64_t version = -1; std::vector<int32_t> buckets; std::string lit_key;
    std::vector<ScanSegment> segs;
    std::vector<uint64_t> snap_uids;             // uid of every snapshot batch the segments cover, in order
    int64_t seen = 0, skipped = 0;
    int consolidations = 0;
    bool valid = false;
  } cache;
  Arena cache_arena;
  PinnedArena pinned;             // staging of descriptor uploads (valid until the next reset)
  // MODE_HASH group table + the launches of this execution (replayed after a grow)
  HashTable hash = {};
  uint32_t hash_capacity = 0;
  uint64_t* d_hash_ident = nullptr;
  bool hash_init = false;
  struct Launch { const void* d_batches; const int32_t* d_prefix; int nbatches; int total_chunks; int batch_base; int needs_slow; };
  // MODE_PROJECT output records + the batches of this execution (records carry a batch ordinal)
  uint8_t* d_out = nullptr;
  int64_t out_cap = 0;
  uint64_t* h_recs = nullptr;     // page-locked read-back staging of the projection records
  size_t h_recs_cap = 0;
  RowWriterBuffers roww;          // device row writer (sd_rows.cu): offsets, rows, string-source tables
  int64_t dev_rows_len = -1;      // >= 0: the finished rows of this execution are roww.d_rows[0, dev_rows_len) (not finished_rows)
  unsigned long long* d_out_count = nullptr;
  std::vector<const StoredBatch*> exec_batches;
  std::vector<uint8_t> finished_rows;   // rows of the last sd_plan_finish (re-served when the caller's buffer was too small)
  int64_t finished_nrows = -1;
  std::vector<Launch> launch_log;
  int64_t metrics[SD_NUM_METRICS] = {0};
  float agg_ms = 0;
  bool have_timing = false;
};

namespace {

std::string literal_key(const sd_plan* p) {
  std::string k;
  for (auto& l : p->lits) {
    k.append(reinterpret_cast<const char*>(&l.is_null), 4);
    k.append(reinterpret_cast<const char*>(&l.i), 8);
    k.append(reinterpret_cast<const char*>(&l.d), 8);
    if (l.s) k.append(l.s, (size_t)l.slen);
    k.push_back('|');
  }
  return k;
}

constexpr size_t STATE_HDR = 8;   // uint64 words in front of the running result: [0] rows scanned, [1] rows passed

int ensure_result(sd_plan* p, size_t entries) {
  if (entries > p->result_cap || !p->d_state) {
    SD_CUDA(cudaSetDevice(p->device));
    const size_t cap = std::max<size_t>(entries, 64);
    uint64_t* ns = nullptr;
    SD_CUDA(cudaMalloc(&ns, (STATE_HDR + cap) * 8));
    if (p->d_state) {   // the counters of the running execution move with the table
      SD_CUDA(cudaStreamSynchronize(p->stream));
      SD_CUDA(cudaMemcpy(ns, p->d_state, STATE_HDR * 8, cudaMemcpyDeviceToDevice));
      cudaFree(p->d_state);
    } else {
      SD_CUDA(cudaMemset(ns, 0, STATE_HDR * 8));
    }
    p->d_state = ns;
    p->d_result = ns + STATE_HDR;
    p->d_counters = reinterpret_cast<unsigned long long*>(ns);
    p->result_cap = cap;
    p->result_init = false;
  }
  if (STATE_HDR + p->result_cap > p->h_pinned_cap) {
    if (p->h_pinned) { SD_CUDA(cudaStreamSynchronize(p->stream)); cudaFreeHost(p->h_pinned); }
    p->h_pinned_cap = STATE_HDR + p->result_cap;
    SD_CUDA(cudaMallocHost(&p->h_pinned, p->h_pinned_cap * 8));
  }
  return 0;
}

// (re)initialise the running result with the slot identities for `ngroups` groups (async: the pinned
// staging buffer outlives the copy)
int init_result(sd_plan* p, int ngroups) {
  const int ns = (int)p->spec.slots.size();
  const size_t ne = (size_t)ngroups * ns;
  int rc = ensure_result(p, ne);
  if (rc) return rc;
  SD_CUDA(cudaStreamSynchronize(p->stream));   // the staging buffer may still be in use by a previous read-back
  uint64_t* hid = p->h_pinned + STATE_HDR;
  for (size_t e = 0; e < ne; e++) {
    const int op = p->spec.slots[e % ns].op;
    hid[e] = op == SLOT_MIN_I64 ? 0x7fffffffffffffffull : op == SLOT_MAX_I64 ? 0x8000000000000000ull
                   : op == SLOT_MIN_F64 ? 0x7ff8000000000000ull : op == SLOT_MAX_F64 ? 0xfff0000000000000ull : 0ull;
  }
  SD_CUDA(cudaMemcpyAsync(p->d_result, hid, ne * 8, cudaMemcpyHostToDevice, p->stream));
  p->result_init = true;
  return 0;
}

// group id of one key value in the query-global dictionary of key k
int key_id(sd_plan* p, int k, const std::string& s) {
  auto it = p->key_ids[k].find(s);
  if (it != p->key_ids[k].end()) return it->second;
  const int id = (int)p->key_vals[k].size();
  p->key_ids[k].emplace(s, id);
  p->key_vals[k].push_back(s);
  return id;
}
int key_null(sd_plan* p, int k) {
  if (p->key_null_id[k] < 0) {
    p->key_null_id[k] = (int)p->key_vals[k].size();
    p->key_vals[k].push_back(std::string());   // placeholder; identified by key_null_id
  }
  return p->key_null_id[k];
}

// when dictionaries grew between launches of one execution the dense group table is re-indexed
int remap_result(sd_plan* p, const int32_t* old_radix, int old_groups, const int32_t* new_radix, int new_groups) {
  const int ns = (int)p->spec.slots.size(), nk = (int)p->spec.keys.size();
  std::vector<uint64_t> oldh((size_t)old_groups * ns);
  SD_CUDA(cudaStreamSynchronize(p->stream));
  SD_CUDA(cudaMemcpy(oldh.data(), p->d_result, oldh.size() * 8, cudaMemcpyDeviceToHost));
  int rc = init_result(p, new_groups);
  if (rc) return rc;
  SD_CUDA(cudaStreamSynchronize(p->stream));
  std::vector<uint64_t> newh((size_t)new_groups * ns);
  SD_CUDA(cudaMemcpy(newh.data(), p->d_result, newh.size() * 8, cudaMemcpyDeviceToHost));
  for (int g = 0; g < old_groups; g++) {
    int idx[MAX_KEYS], rem = g;
    for (int k = nk - 1; k >= 0; k--) { idx[k] = rem % old_radix[k]; rem /= old_radix[k]; }
    int ng = 0;
    for (int k = 0; k < nk; k++) ng = ng * new_radix[k] + idx[k];
    memcpy(&newh[(size_t)ng * ns], &oldh[(size_t)g * ns], (size_t)ns * 8);
  }
  SD_CUDA(cudaMemcpy(p->d_result, newh.data(), newh.size() * 8, cudaMemcpyHostToDevice));
  return 0;
}

constexpr int REG_GROUPS_MAX = 8;
int resolve_kernel(const sd_plan_desc& desc, const CodegenOptions& opt, int device, KernelEntry* out, PlanSpec* spec_out);

struct BuiltScan {
  const void* d_batches = nullptr; const int32_t* d_prefix = nullptr; int nbatches = 0; int total_chunks = 0;
  int64_t rows = 0, algo_bytes = 0, updated_cols = 0, deleted_batches = 0;
  int needs_slow = 0;   // some batch needs the kernel variant with the per-row paths
  int needs_hash = 0;   // a key column of a dense-table plan is a raw (variable-width) string in some batch
};

// Build the device descriptors + per-batch tables for a list of resident batches.
// `up` is the stream the descriptors are uploaded on (from page-locked staging: the caller is never blocked); the
// scan must be ordered after it.
int build_scan(sd_plan* p, const std::vector<const StoredBatch*>& list, Arena& arena, cudaStream_t up, BuiltScan* out) {
  const PlanSpec& sp = p->spec;
  const int nc = (int)sp.cols.size();
  const size_t bstride = sizeof(DevBatch<1>) - sizeof(DevCol) + (size_t)std::max(nc, 1) * sizeof(DevCol);
  const int nt = (int)sp.tables.size();
  std::vector<uint8_t> hb(bstride * std::max<size_t>(list.size(), 1), 0);
  std::vector<int32_t> prefix(list.size() + 1, 0);
  std::vector<uint8_t> aux;
  std::vector<size_t> aux_off(list.size(), 0);
  for (size_t bi = 0; bi < list.size(); bi++) {
    const StoredBatch& sb = *list[bi];
    uint8_t* rec = hb.data() + bi * bstride;
    DevBatch<1>* hdr = reinterpret_cast<DevBatch<1>*>(rec);
    hdr->num_rows = sb.num_rows;
    hdr->num_deletes = sb.num_deletes;
    hdr->deletes = sb.dev_deletes;
    bool all_fast = sb.dev_deletes == nullptr;
    bool base_fast = true;   // every base column can take the vector path (deltas / deletes aside)
    bool simple_enc = true;  // every column has a directly addressable encoding (NULLs allowed)
    bool any_delta = false;
    if (sb.dev_deletes) out->deleted_batches++;
    DevCol* dc = reinterpret_cast<DevCol*>(rec + (sizeof(DevBatch<1>) - sizeof(DevCol)));
    for (int c = 0; c < nc; c++) {
      const int t = sb.positional ? c : sp.cols[c].table_ordinal;
      if (t < 0 || t >= (int)sb.cols.size() || !sb.cols[t].present)
        return set_error(SD_ERR_INVALID, "batch %lld: table column %d is not resident", (long long)sb.batch_id, t);
      const StoredCol& sc = sb.cols[t];
      if (!sc.unsupported.empty()) return set_error(SD_ERR_UNSUPPORTED, "column %d: %s", t, sc.unsupported.c_str());
      dc[c] = sc.dev;
      all_fast = all_fast && sc.fast;
      {
        const bool simple = (sp.kinds[c] == K_CODE && (sc.dev.enc == ENC_DICTIONARY || sc.dev.enc == ENC_BIG_DICTIONARY || sc.dev.enc == ENC_STR_RAW)) ||
                            (sp.kinds[c] != K_CODE && sc.dev.enc == ENC_UNCOMPRESSED);
        simple_enc = simple_enc && simple && (!sc.has_nulls || sp.cols[c].nullable);   // NULLs in a column the plan calls non-nullable: per-row path
        any_delta = any_delta || sc.delta[0].present || sc.delta[1].present;
      }
      base_fast = base_fast && !sc.has_nulls && ((sp.kinds[c] == K_CODE && (sc.dev.enc == ENC_DICTIONARY || sc.dev.enc == ENC_BIG_DICTIONARY || sc.dev.enc == ENC_STR_RAW)) ||
                                                 (sp.kinds[c] != K_CODE && sc.dev.enc == ENC_UNCOMPRESSED));
      out->algo_bytes += sc.algo_bytes;
      if (sc.delta[0].present || sc.delta[1].present) out->updated_cols++;
      for (int d = 0; d < 2; d++) if (sc.delta[d].present) out->algo_bytes += sc.delta[d].len;
    }
    if (sb.dev_deletes) out->algo_bytes += 12 + 4 * (int64_t)sb.num_deletes;
    hdr->flags = all_fast ? BATCH_ALL_FAST : (base_fast ? BATCH_FAST_OVERLAY : ((simple_enc && !any_delta && !sb.dev_deletes) ? BATCH_FAST_NULLS : 0));
    if (!(hdr->flags == BATCH_ALL_FAST || (hdr->flags == BATCH_FAST_NULLS && p->kernel.staged))) out->needs_slow = 1;
    // per-batch tables: [int32 offset x nt][pad 8][uint64 kpack x nt][tables]; every table is indexed by the
    // unified dictionary code; key maps of <= 8 codes are also packed one byte per code into kpack
    if (nt) {
      while (aux.size() % 16) aux.push_back(0);
      aux_off[bi] = aux.size();
      const size_t base = aux.size();
      const size_t kp_off = ((4 * (size_t)nt + 7) & ~size_t(7));
      aux.resize(base + kp_off + 8 * (size_t)nt, 0xff);
      for (int ti = 0; ti < nt; ti++) {
        const TableSpec& ts = sp.tables[ti];
        const StoredCol& sc = sb.cols[sb.positional ? ts.col : sp.cols[ts.col].table_ordinal];
        const int n = sc.dev.dict_n;                         // NULL code of a nullable column
        const bool nullable = sp.cols[ts.col].nullable != 0;
        if (sc.raw_str) {   // no code space: predicates / keys of this batch work on the bytes (the table stays empty)
          if (ts.kind == TABLE_KEYMAP) out->needs_hash = 1;   // a dense group table needs dictionary ids: this plan must use the hash table
          while (aux.size() % 8) aux.push_back(0);
          const int32_t off0 = (int32_t)(aux.size() - base);
          memcpy(aux.data() + base + 4 * (size_t)ti, &off0, 4);
          continue;
        }
        // codes: [0,n) base dictionary; n = NULL (nullable columns) or an unused placeholder when update
        // deltas appended strings; (n, ...) strings that occur only in update deltas
        const int ncodes = std::max((int)sc.dict_strings.size(), nullable ? n + 1 : n);
        while (aux.size() % 8) aux.push_back(0);
        const int32_t off = (int32_t)(aux.size() - base);
        memcpy(aux.data() + base + 4 * (size_t)ti, &off, 4);
        uint64_t kpack = ~0ull;
        bool packable = ts.kind == TABLE_KEYMAP && ncodes <= 8;
        uint8_t packed[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (ts.kind == TABLE_TRUTH) {
          for (int code = 0; code < ncodes; code++) {
            const bool isnull = code == n || code >= (int)sc.dict_strings.size();
            const std::string* s = isnull ? nullptr : &sc.dict_strings[code];
            aux.push_back((uint8_t)eval_string_predicate(sp, ts.node, s ? s->data() : nullptr, s ? (int)s->size() : 0, p->lits.data()));
          }
        } else if (ts.kind == TABLE_KEYPTR) {   // device address of every code's [len][bytes] record
          for (int code = 0; code < ncodes; code++) {
            const int64_t ptr = code < (int)sc.dict_rec_ptr.size() ? sc.dict_rec_ptr[code] : 0;
            aux.insert(aux.end(), reinterpret_cast<const uint8_t*>(&ptr), reinterpret_cast<const uint8_t*>(&ptr) + 8);
          }
        } else {
          for (int code = 0; code < ncodes; code++) {
            const bool isnull = code == n || code >= (int)sc.dict_strings.size();
            const int32_t id = isnull ? (nullable ? key_null(p, ts.key) : 0) : key_id(p, ts.key, sc.dict_strings[code]);
            aux.insert(aux.end(), reinterpret_cast<const uint8_t*>(&id), reinterpret_cast<const uint8_t*>(&id) + 4);
            if (packable) { if (id >= 255) packable = false; else packed[code] = (uint8_t)id; }
          }
          if (packable) memcpy(&kpack, packed, 8);
        }
        memcpy(aux.data() + base + kp_off + 8 * (size_t)ti, &kpack, 8);
      }
    }
    out->rows += sb.num_rows;
    prefix[bi + 1] = prefix[bi] + (sb.num_rows + p->chunk_rows - 1) / p->chunk_rows;
  }
  // NULL key ids are only materialised when a nullable key column is present in the plan
  uint8_t* d_aux = nullptr;
  if (!aux.empty()) {
    d_aux = arena.alloc(aux.size() + 16, 16);
    uint8_t* h_aux = p->pinned.alloc(aux.size());
    if (!d_aux || !h_aux) return SD_ERR_CUDA;
    memcpy(h_aux, aux.data(), aux.size());
    SD_CUDA(cudaMemcpyAsync(d_aux, h_aux, aux.size(), cudaMemcpyHostToDevice, up));
    for (size_t bi = 0; bi < list.size(); bi++) reinterpret_cast<DevBatch<1>*>(hb.data() + bi * bstride)->aux = d_aux + aux_off[bi];
  }
  uint8_t* d_b = arena.alloc(hb.size() + 16, 16);
  uint8_t* d_p = arena.alloc(prefix.size() * 4 + 16, 16);
  uint8_t* h_b = p->pinned.alloc(hb.size());
  uint8_t* h_p = p->pinned.alloc(prefix.size() * 4);
  if (!d_b || !d_p || !h_b || !h_p) return SD_ERR_CUDA;
  memcpy(h_b, hb.data(), hb.size());
  memcpy(h_p, prefix.data(), prefix.size() * 4);
  SD_CUDA(cudaMemcpyAsync(d_b, h_b, hb.size(), cudaMemcpyHostToDevice, up));
  SD_CUDA(cudaMemcpyAsync(d_p, h_p, prefix.size() * 4, cudaMemcpyHostToDevice, up));
  out->d_batches = d_b;
  out->d_prefix = reinterpret_cast<const int32_t*>(d_p);
  out->nbatches = (int)list.size();
  out->total_chunks = prefix.back();
  return 0;
}

void hash_free(sd_plan* p) {
  if (p->hash.state) cudaFree(p->hash.state);
  if (p->hash.keys) cudaFree(p->hash.keys);
  if (p->hash.knull) cudaFree(p->hash.knull);
  if (p->hash.vals) cudaFree(p->hash.vals);
  if (p->hash.overflow) cudaFree(p->hash.overflow);
  p->hash = HashTable{};
  p->hash_capacity = 0;
}

int hash_ensure(sd_plan* p, uint32_t capacity) {
  const int ns = (int)p->spec.slots.size(), nk = std::max<int>(1, (int)p->spec.keys.size());
  if (p->hash_capacity != capacity) {
    hash_free(p);
    SD_CUDA(cudaMalloc(&p->hash.state, (size_t)capacity * 4));
    SD_CUDA(cudaMalloc(&p->hash.keys, (size_t)capacity * nk * 8));
    SD_CUDA(cudaMalloc(&p->hash.knull, (size_t)capacity * 4));
    SD_CUDA(cudaMalloc(&p->hash.vals, (size_t)capacity * ns * 8));
    SD_CUDA(cudaMalloc(&p->hash.overflow, 1024));   // [0] overflow flag, [8] key count; the rest: diagnostic histograms (SD_EXP_VERIFY builds)
    SD_CUDA(cudaMemset(p->hash.overflow, 0, 1024));
    p->hash.count = p->hash.overflow + 8;
    p->hash.mask = capacity - 1;
    p->hash.max_probe = std::min<uint32_t>(capacity, 4096);
    p->hash_capacity = capacity;
    p->hash_init = false;
  }
  if (!p->d_hash_ident) {
    std::vector<uint64_t> id(ns);
    for (int s = 0; s < ns; s++) {
      const int op = p->spec.slots[s].op;
      id[s] = op == SLOT_MIN_I64 ? 0x7fffffffffffffffull : op == SLOT_MAX_I64 ? 0x8000000000000000ull
            : op == SLOT_MIN_F64 ? 0x7ff8000000000000ull : op == SLOT_MAX_F64 ? 0xfff0000000000000ull : 0ull;
    }
    SD_CUDA(cudaMalloc(&p->d_hash_ident, (size_t)std::max(ns, 1) * 8));
    SD_CUDA(cudaMemcpy(p->d_hash_ident, id.data(), (size_t)ns * 8, cudaMemcpyHostToDevice));
  }
  if (!p->hash_init) {
    int rc = hash_table_init(p->stream, p->hash, capacity, ns, p->d_hash_ident);
    if (rc) return rc;
    p->hash_init = true;
  }
  return 0;
}

int ensure_out(sd_plan* p, int64_t cap_records) {
  const int64_t rec = 8 + 8 * (int64_t)std::max<size_t>(p->spec.proj.size(), 1);
  if (!p->d_out_count) { SD_CUDA(cudaMalloc(&p->d_out_count, 64)); SD_CUDA(cudaMemset(p->d_out_count, 0, 64)); }
  if (cap_records > p->out_cap) {
    if (p->d_out) cudaFree(p->d_out);
    p->d_out = nullptr;
    SD_CUDA(cudaMalloc(&p->d_out, (size_t)(cap_records * rec)));
    p->out_cap = cap_records;
  }
  return 0;
}

// device time of this execution's launches (sum over the per-launch event pairs; the span as a fallback)
static void update_agg_time(sd_plan* p) {
  if (!p->have_timing) return;
  float total = 0;
  bool ok = p->ev_used > 0 && p->ev_used == (size_t)p->metrics[7];
  for (size_t i = 0; ok && i < p->ev_used; i++) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, p->ev_pairs[i].first, p->ev_pairs[i].second) != cudaSuccess) { ok = false; break; }
    total += ms;
  }
  if (!ok) { float ms = 0; if (cudaEventElapsedTime(&ms, p->ev_start, p->ev_stop) == cudaSuccess) total = ms; else return; }
  p->agg_ms = total;
}

// the kernel variant of the plan for this execution / launch
int plan_variant(sd_plan* p, int litnull, int slow, const KernelEntry** out) {
  const int idx = (litnull ? 1 : 0) | (slow ? 2 : 0);
  if (idx == 0) { *out = &p->kernel; return 0; }
  if (!p->variant_state[idx]) {
    CodegenOptions opt;
    opt.lit_nullable = litnull ? 1 : 0;
    opt.slow_paths = slow ? 1 : 0;
    opt.force_hash = p->force_hash ? 1 : 0;
    sd_plan_desc dv = p->spec.desc_view();
    int rc = resolve_kernel(dv, opt, p->device, &p->variant[idx], nullptr);
    if (rc) return rc;
    p->variant_state[idx] = 1;
  }
  *out = &p->variant[idx];
  return 0;
}

int launch_scan(sd_plan* p, const void* d_batches, const int32_t* d_prefix, int nbatches, int total_chunks, int needs_slow,
                const std::vector<const StoredBatch*>* blist = nullptr, int replay_batch_base = -1);

// A dense-table plan (dictionary-string keys) has to give up its table: too many key combinations, or a key column
// arrives as raw variable-width strings (no dictionary ids).  The plan becomes its hash-table variant for good and
// whatever this execution has launched so far is rebuilt (the per-batch tables differ) and launched again.
int switch_to_hash(sd_plan* p) {
  CodegenOptions opt;
  opt.force_hash = 1;
  PlanSpec hspec;
  KernelEntry hk;
  sd_plan_desc dv = p->spec.desc_view();
  int rc = resolve_kernel(dv, opt, p->device, &hk, &hspec);
  if (rc) return rc;
  const std::vector<sd_plan::Launch> earlier(p->launch_log);
  const std::vector<const StoredBatch*> exec(p->exec_batches);
  p->launch_log.clear();
  p->exec_batches.clear();
  p->spec = hspec;
  p->kernel = hk;
  p->force_hash = true;
  for (int v = 0; v < 4; v++) p->variant_state[v] = 0;
  p->active = nullptr;
  p->last_kernel = nullptr;
  p->last_smem = (size_t)-1;
  p->result_init = false;
  p->hash_init = false;
  p->cache.valid = false;
  p->chunk_rows = CHUNK_ROWS;
  SD_CUDA(cudaMemsetAsync(p->d_counters, 0, 64, p->stream));
  for (auto& l : earlier) {
    std::vector<const StoredBatch*> list(exec.begin() + l.batch_base, exec.begin() + l.batch_base + l.nbatches);
    BuiltScan bs;
    rc = build_scan(p, list, p->scratch, p->stream, &bs);
    if (rc) return rc;
    rc = launch_scan(p, bs.d_batches, bs.d_prefix, bs.nbatches, bs.total_chunks, bs.needs_slow, &list);
    if (rc) return rc;
  }
  return 0;
}

int launch_scan(sd_plan* p, const void* d_batches, const int32_t* d_prefix, int nbatches, int total_chunks, int needs_slow,
                const std::vector<const StoredBatch*>* blist, int replay_batch_base) {
  const bool replay = replay_batch_base >= 0;
  if (!replay && p->spec.mode == MODE_GROUPS) {   // dense table still possible with the dictionaries seen so far?
    int64_t ng = 1;
    for (size_t k = 0; k < p->spec.keys.size(); k++) ng *= std::max<int64_t>(1, (int64_t)p->key_vals[k].size());
    if (ng > (1 << 16)) {
      if (!blist) return set_error(SD_ERR_STATE, "dense group table overflow without a batch list");
      int rc = switch_to_hash(p);
      if (rc) return rc;
      BuiltScan bs;
      rc = build_scan(p, *blist, p->scratch, p->stream, &bs);
      if (rc) return rc;
      return launch_scan(p, bs.d_batches, bs.d_prefix, bs.nbatches, bs.total_chunks, bs.needs_slow, blist);
    }
  }
  int batch_base = replay ? replay_batch_base : (int)p->exec_batches.size();
  if (!replay && p->spec.mode != MODE_NOKEY) {
    p->launch_log.push_back({d_batches, d_prefix, nbatches, total_chunks, batch_base, needs_slow});
    if (blist) p->exec_batches.insert(p->exec_batches.end(), blist->begin(), blist->end());
  }
  p->finished_nrows = -1;
  p->dev_rows_len = -1;
  if (nbatches == 0 || total_chunks == 0) return 0;
  const PlanSpec& sp = p->spec;
  const int ns = (int)sp.slots.size(), nk = (int)sp.keys.size();
  // group radices from the current key dictionaries
  int32_t radix[MAX_KEYS] = {1, 1, 1, 1};
  int ngroups = 1;
  for (int k = 0; k < nk && sp.mode == MODE_GROUPS; k++) {
    radix[k] = std::max<int>(1, (int)p->key_vals[k].size());
    ngroups *= radix[k];
  }
  const size_t ne = (size_t)ngroups * ns;
  // Where the dense group table lives (decided per launch from its size):
  //   registers (kernel variant, <= 8 groups) > per-thread private shared-memory copies > one shared-memory copy
  //   per CTA with atomics > global atomics on the running result.
  // What is left of the SM's shared memory (per target CTA) becomes the ring of the staged fast path.
  const KernelEntry* k = &p->kernel;
  {   // a NULL literal needs the variant whose generated code carries literal null flags; batches with deltas, deletes
      // or encodings that are not directly addressable need the variant that carries the per-row paths
    bool any_null = false;
    for (auto& l : p->lits) any_null = any_null || l.is_null;
    static const bool force_full = getenv("SD_TUNE_FULL_KERNEL") != nullptr;   // measurement aid
    int rc = plan_variant(p, any_null, needs_slow || force_full, &k);
    if (rc) return rc;
  }
  int table_mode = TABLE_PRIVATE;
  int fresh = 0;
  int target_ctas = std::max(1, sp.min_ctas);
  if (sp.mode == MODE_GROUPS && ngroups <= REG_GROUPS_MAX && k == &p->kernel && p->kernel.staged && getenv("SD_TUNE_REG_GROUPS")) {
    // experimental (opt-in): measured slower than the private shared-memory tables, see DESIGN.md
    if (p->kernel_reg_state == 0) {
      CodegenOptions opt;
      opt.reg_groups = REG_GROUPS_MAX;
      sd_plan_desc dv = sp.desc_view();
      p->kernel_reg_state = resolve_kernel(dv, opt, p->device, &p->kernel_reg, nullptr) == 0 ? 1 : -1;
    }
    if (p->kernel_reg_state == 1 && 2 * p->kernel_reg.stage_bytes >= ne * THREADS * 8) {
      k = &p->kernel_reg;
      table_mode = TABLE_REGS;
      target_ctas = 1;
    }
  }
  const size_t tile_smem = k->tile_smem;
  const size_t ring_fixed = 2 * MAX_STAGES * 8;
  const size_t min_ring = k->staged ? ring_fixed + 2 * k->stage_bytes + 128 : 0;
  size_t table_bytes = (size_t)std::max(ns, 1) * (THREADS / 32) * 8;
  if (sp.mode == MODE_GROUPS && table_mode != TABLE_REGS) {
    const size_t priv = ne * THREADS * 8, shared = ne * 8;
    // private copies are worth giving up CTAs per SM for; atomics are the last resort
    while (target_ctas > 1 && tile_smem + priv + min_ring > (size_t)p->smem_optin / target_ctas - 1024) target_ctas--;
    const size_t budget1 = (size_t)p->smem_optin / target_ctas - (target_ctas > 1 ? 1024 : 0);
    if (tile_smem + priv + min_ring <= budget1) { table_mode = TABLE_PRIVATE; table_bytes = priv; }
    else if (shared <= 64 * 1024 && tile_smem + shared + min_ring <= budget1) { table_mode = TABLE_SHARED_ATOMIC; table_bytes = shared; }
    else { table_mode = TABLE_GLOBAL_ATOMIC; table_bytes = 64; }
  }
  const size_t budget = (size_t)p->smem_optin / target_ctas - (target_ctas > 1 ? 1024 : 0);
  size_t ring_off = (tile_smem + table_bytes + 127) & ~size_t(127);
  int nstages = 0;
  size_t smem = ring_off;
  if (k->staged) {
    if (ring_off + min_ring > (size_t)p->smem_optin) return set_error(SD_ERR_UNSUPPORTED, "plan does not fit the shared-memory ring (%zu bytes per stage)", k->stage_bytes);
    const size_t avail = std::max(budget, ring_off + min_ring) - ring_off - ring_fixed;
    nstages = (int)std::min<size_t>(MAX_STAGES, avail / k->stage_bytes);
    if (const char* e = getenv("SD_TUNE_NSTAGES")) { int v = atoi(e); if (v >= 2 && v <= nstages) nstages = v; }
    smem = ring_off + ring_fixed + (size_t)nstages * k->stage_bytes;
  }
  if (p->active != k) { p->active = k; p->kernel_name = k->origin + ":" + k->name + (table_mode == TABLE_REGS ? "+regtable" : ""); }
  if (sp.mode == MODE_HASH) {
    int rc = hash_ensure(p, p->hash_capacity ? p->hash_capacity : (1u << 16));
    if (rc) return rc;
  } else if (sp.mode == MODE_PROJECT) {
    int rc = ensure_out(p, p->out_cap ? p->out_cap : (int64_t(1) << 20));
    if (rc) return rc;
  } else if (!p->result_init) {
    if (table_mode == TABLE_GLOBAL_ATOMIC) {   // the kernel adds into the running result: it must hold the identities
      int rc = init_result(p, ngroups);
      if (rc) return rc;
    } else {                                     // the last CTA overwrites it (ScanArgs.fresh): nothing to upload
      int rc = ensure_result(p, ne);
      if (rc) return rc;
      fresh = 1;
      p->result_init = true;
    }
  } else if (ngroups != p->ngroups || memcmp(radix, p->radix, sizeof(radix)) != 0) {
    int rc = remap_result(p, p->radix, p->ngroups, radix, ngroups);
    if (rc) return rc;
  }
  p->ngroups = ngroups;
  memcpy(p->radix, radix, sizeof(radix));

  if (smem != p->last_smem || k != p->last_kernel) {
    int occ = 0;
    int rc = kernel_prepare(*k, smem, &occ);
    if (rc) return rc;
    if (occ < 1) return set_error(SD_ERR_CUDA, "kernel does not fit on an SM (smem %zu)", smem);
    p->max_ctas_per_sm = occ;
    p->last_smem = smem;
    p->last_kernel = k;
  }
  const int occ = p->max_ctas_per_sm;
  const int grid = std::min(total_chunks, p->num_sms * occ);
  if (table_mode != TABLE_GLOBAL_ATOMIC && (size_t)grid * ne > p->partials_cap) {
    if (p->d_partials) cudaFree(p->d_partials);
    p->partials_cap = (size_t)grid * ne * 2;
    SD_CUDA(cudaMalloc(&p->d_partials, p->partials_cap * 8));
  }
  ScanArgs args;
  memset(&args, 0, sizeof(args));
  args.batches = d_batches;
  args.chunk_prefix = d_prefix;
  args.nbatches = nbatches;
  args.total_chunks = total_chunks;
  args.partials = p->d_partials;
  args.result = p->d_result;
  args.ticket = p->d_ticket;
  args.counters = p->d_counters;
  args.ngroups = ngroups;
  args.table_mode = table_mode;
  args.ring_off = (int32_t)ring_off;
  args.nstages = nstages;
  args.hash = p->hash;
  args.out_rows = p->d_out;
  args.out_count = p->d_out_count;
  args.out_cap = p->out_cap;
  args.batch_base = batch_base;
  args.chunk_rows = p->chunk_rows;
  args.fresh = fresh;
  memcpy(args.radix, radix, sizeof(radix));
  if (p->litpool_dirty) {   // STRING literal bytes -> device (once per set of literal values)
    std::vector<uint8_t> pool;
    p->lit_packed.assign(p->lits.size(), 0);
    for (size_t i = 0; i < p->lits.size(); i++) {
      if (p->lits[i].type != SD_STRING) continue;
      p->lit_packed[i] = (int64_t)(((uint64_t)pool.size() << 32) | (uint32_t)p->lits[i].slen);
      pool.insert(pool.end(), p->lits[i].s, p->lits[i].s + p->lits[i].slen);
    }
    if (!pool.empty()) {
      if (pool.size() > p->litpool_cap) {
        if (p->d_litpool) cudaFree(p->d_litpool);
        p->litpool_cap = pool.size() * 2 + 256;
        SD_CUDA(cudaMalloc(&p->d_litpool, p->litpool_cap));
      }
      uint8_t* h = p->pinned.alloc(pool.size());
      if (!h) return SD_ERR_CUDA;
      memcpy(h, pool.data(), pool.size());
      SD_CUDA(cudaMemcpyAsync(p->d_litpool, h, pool.size(), cudaMemcpyHostToDevice, p->stream));
    }
    p->litpool_dirty = false;
  }
  args.lit_pool = p->d_litpool;
  for (size_t i = 0; i < p->lits.size(); i++) {
    args.lits.i[i] = p->lits[i].type == SD_STRING ? p->lit_packed[i] : p->lits[i].i;
    args.lits.d[i] = p->lits[i].d;
    if (p->lits[i].is_null) args.lits.nullmask |= 1ull << i;
  }
  void* kargs[] = {&args};
  if (p->ev_used == p->ev_pairs.size() && p->ev_pairs.size() < 4096) {
    cudaEvent_t a = nullptr, b = nullptr;
    SD_CUDA(cudaEventCreate(&a));
    SD_CUDA(cudaEventCreate(&b));
    p->ev_pairs.emplace_back(a, b);
  }
  const bool own_pair = p->ev_used < p->ev_pairs.size();
  if (own_pair) SD_CUDA(cudaEventRecord(p->ev_pairs[p->ev_used].first, p->stream));
  if (!p->have_timing) SD_CUDA(cudaEventRecord(p->ev_start, p->stream));
  { int rc = kernel_launch(*k, grid, smem, p->stream, kargs); if (rc) return rc; }
  SD_CUDA(cudaEventRecord(p->ev_stop, p->stream));
  if (own_pair) { SD_CUDA(cudaEventRecord(p->ev_pairs[p->ev_used].second, p->stream)); p->ev_used++; }
  p->have_timing = true;
  p->metrics[7]++;
  return 0;
}

int flush_pending(sd_plan* p) {
  if (p->pending.empty()) return 0;
  if (p->priv) {   // queue the expansion of compressed buffers first: it runs while the descriptors are being built
    int rc0 = store_flush_lz4(p->priv);
    if (rc0) return rc0;
  }
  std::vector<const StoredBatch*> list;
  for (auto& x : p->pending) list.push_back(x.sb);
  BuiltScan bs;
  // nothing below blocks the submitting thread: descriptors go out on the copy stream from page-locked staging, and
  // the scan is ordered after the copies and the expansions with events
  int rc = build_scan(p, list, p->scratch, p->priv ? p->priv->copy_stream : p->stream, &bs);
  if (rc) return rc;
  if (bs.needs_hash && p->spec.mode == MODE_GROUPS) {   // a key column came as raw strings: dense ids do not exist for it
    rc = switch_to_hash(p);
    if (rc) return rc;
    bs = BuiltScan();
    rc = build_scan(p, list, p->scratch, p->priv ? p->priv->copy_stream : p->stream, &bs);
    if (rc) return rc;
  }
  if (p->priv) {
    if (!p->priv->copies_done) SD_CUDA(cudaEventCreateWithFlags(&p->priv->copies_done, cudaEventDisableTiming));
    SD_CUDA(cudaEventRecord(p->priv->copies_done, p->priv->copy_stream));
    SD_CUDA(cudaStreamWaitEvent(p->stream, p->priv->copies_done, 0));
    for (int k = 0; k + 1 < p->priv->num_copy_streams; k++) {
      SD_CUDA(cudaEventRecord(p->priv->extra_done[k], p->priv->extra_streams[k]));
      SD_CUDA(cudaStreamWaitEvent(p->stream, p->priv->extra_done[k], 0));
    }
    rc = store_lz4_order(p->priv, p->stream);   // the scan reads what the expansions write
    if (rc) return rc;
  }
  p->metrics[3] += bs.updated_cols;
  p->metrics[4] += bs.deleted_batches;
  p->metrics[9] += bs.algo_bytes;
  rc = launch_scan(p, bs.d_batches, bs.d_prefix, bs.nbatches, bs.total_chunks, bs.needs_slow, &list);
  p->pending.clear();
  p->pending_bytes = 0;
  return rc;
}

bool batch_passes_stats(const sd_plan* p, const StoredBatch& sb) {
  if (p->spec.filter < 0 || sb.stats.empty()) return true;
  StatEval ev{p->spec, p->lits, sb.stats.data(), (int64_t)sb.stats.size(), 1 + 3 * sb.stats_ncols, sb.num_rows};
  Tri r;
  if (!ev.eval(p->spec.filter, &r)) return true;
  return r.isnull || r.v;   // only a definite FALSE skips the batch (:948-957)
}

// test hook (host only, no CUDA call): the batch-skipping decision of a plan for one stats row
int stats_pass_hook(const sd_plan_desc* desc, const sd_literal* lits, int32_t nlits, const void* stats, int64_t stats_len,
                    int32_t stats_ncols, int32_t num_rows, int32_t* pass) {
  if (!desc || !pass || (nlits > 0 && !lits)) return set_error(SD_ERR_INVALID, "sdx_stats_pass: null argument");
  PlanSpec spec;
  std::string err;
  int rc = analyze_plan(desc, spec, err);
  if (rc) return set_error(rc, "sdx_stats_pass: %s", err.c_str());
  if (nlits != (int)spec.literal_types.size()) return set_error(SD_ERR_INVALID, "sdx_stats_pass: plan has %zu literal slots", spec.literal_types.size());
  std::vector<sd_literal> lv(lits, lits + nlits);
  *pass = 1;
  if (spec.filter < 0 || !stats || stats_len <= 0) return 0;
  StatEval ev{spec, lv, reinterpret_cast<const uint8_t*>(stats), stats_len, 1 + 3 * stats_ncols, num_rows};
  Tri r;
  if (ev.eval(spec.filter, &r)) *pass = (r.isnull || r.v) ? 1 : 0;
  return 0;
}

// find the kernel for a codegen variant of the plan: ahead-of-time registry first, else NVRTC
int resolve_kernel(const sd_plan_desc& desc, const CodegenOptions& opt, int device, KernelEntry* out, PlanSpec* spec_out) {
  PlanSpec spec;
  std::string err;
  int rc = analyze_plan(&desc, spec, err, &opt);
  if (rc) return set_error(rc, "%s", err.c_str());
  static std::mutex registry_mutex;   // plans are created concurrently from many task threads
  std::lock_guard<std::mutex> lock(registry_mutex);
  for (auto& k : kernel_registry()) if (k.signature == spec.signature) { *out = k; if (spec_out) *spec_out = spec; return 0; }
  KernelEntry k;
  rc = jit_compile(spec, device, k);
  if (rc) return rc;
  kernel_registry().push_back(k);
  *out = k;
  if (spec_out) *spec_out = spec;
  return 0;
}

// MIN / MAX over STRING: slots hold device addresses of [len][bytes] records -> their bytes, for every group at once
typedef std::unordered_map<uint64_t, std::string> StrMap;
int fetch_agg_strings(sd_plan* p, const uint64_t* slots, size_t ngroups, StrMap& out) {
  const PlanSpec& sp = p->spec;
  const size_t ns = sp.slots.size();
  std::vector<int64_t> ptrs;
  for (auto& m : sp.agg_map) {
    if (m.buf_type != SD_STRING) continue;
    for (size_t g = 0; g < ngroups; g++) { const uint64_t a = slots[g * ns + m.value_slot]; if (a && !out.count(a)) { out.emplace(a, std::string()); ptrs.push_back((int64_t)a); } }
  }
  if (ptrs.empty()) return 0;
  int64_t* d_ptrs = nullptr;
  SD_CUDA(cudaMalloc(&d_ptrs, ptrs.size() * 8));
  std::vector<std::string> strs;
  cudaError_t ce = cudaMemcpyAsync(d_ptrs, ptrs.data(), ptrs.size() * 8, cudaMemcpyHostToDevice, p->stream);
  int rc = ce == cudaSuccess ? fetch_string_records(p->stream, d_ptrs, (int64_t)ptrs.size(), 1, strs) : set_error(SD_ERR_CUDA, "cudaMemcpyAsync: %s", cudaGetErrorString(ce));
  cudaFree(d_ptrs);
  if (rc) return rc;
  for (size_t i = 0; i < ptrs.size(); i++) out[(uint64_t)ptrs[i]] = strs[i];
  return 0;
}

// partial-row fields of one group from its slot values (shared by the dense and the hash paths)
void append_agg_fields(const PlanSpec& sp, const uint64_t* sv, std::vector<HVal>& vals, const StrMap* strs = nullptr) {
  for (auto& m : sp.agg_map) {
    HVal v;
    const uint64_t raw = sv[m.value_slot];
    const int64_t cnt = m.count_slot >= 0 ? (int64_t)sv[m.count_slot] : 1;
    if (m.buf_type == SD_STRING) {
      if ((m.buf_nullable && cnt == 0) || raw == 0 || !strs) v.isnull = true;
      else { auto it = strs->find(raw); if (it == strs->end()) v.isnull = true; else v.s = it->second; }
      vals.push_back(v);
      continue;
    }
    if (m.value_slot2 >= 0) {   // DECIMAL SUM / AVG: high and low halves summed separately (sd_codegen.cpp build_slots)
      v.w = (i128)(int64_t)sv[m.value_slot] * ((i128)1 << 32) + (i128)(int64_t)sv[m.value_slot2];
      v.i = (int64_t)v.w;
      if (m.fn == SD_AGG_SUM) { if (m.buf_nullable && cnt == 0) v.isnull = true; vals.push_back(v); }
      else { vals.push_back(v); HVal c; c.i = cnt; vals.push_back(c); }
      continue;
    }
    switch (m.fn) {
      case SD_AGG_COUNT_STAR: case SD_AGG_COUNT: v.i = (int64_t)raw; vals.push_back(v); break;
      case SD_AGG_SUM:
        if (m.buf_nullable && cnt == 0) v.isnull = true;
        else if (m.buf_type == SD_DOUBLE) memcpy(&v.d, &raw, 8); else v.i = (int64_t)raw;
        vals.push_back(v); break;
      case SD_AGG_AVG: {
        memcpy(&v.d, &raw, 8); vals.push_back(v);
        HVal c; c.i = cnt; vals.push_back(c); break;
      }
      default:
        if (m.buf_nullable && cnt == 0) v.isnull = true;
        else if (type_is_fp(m.buf_type)) memcpy(&v.d, &raw, 8); else { v.i = (int64_t)raw; v.w = v.i; }
        vals.push_back(v); break;
    }
  }
}

// MODE_HASH: grow + replay on overflow, compact the occupied entries, emit partial rows
int finish_hash(sd_plan* p) {
  const PlanSpec& sp = p->spec;
  const int ns = (int)sp.slots.size(), nk = (int)sp.keys.size();
  if (!p->hash_capacity) { int rc = hash_ensure(p, 1u << 16); if (rc) return rc; }
  uint32_t flags[16];
  for (;;) {
    SD_CUDA(cudaMemcpyAsync(flags, p->hash.overflow, 64, cudaMemcpyDeviceToHost, p->stream));
    SD_CUDA(cudaStreamSynchronize(p->stream));
    const uint32_t overflow = flags[0], count = flags[8];
    if (!overflow && (uint64_t)count * 2 <= p->hash_capacity) break;
    // too full (or an insert gave up): grow and replay every launch of this execution over the same bytes
    if (p->hash_capacity >= (1u << 28)) return set_error(SD_ERR_UNSUPPORTED, "group-by hash table would exceed 2^28 entries");
    uint32_t ncap = p->hash_capacity * 8;
    while ((uint64_t)count * 4 > ncap && ncap < (1u << 28)) ncap *= 2;
    int rc = hash_ensure(p, ncap);
    if (rc) return rc;
    SD_CUDA(cudaMemsetAsync(p->d_counters, 0, 64, p->stream));
    for (auto& l : p->launch_log) { rc = launch_scan(p, l.d_batches, l.d_prefix, l.nbatches, l.total_chunks, l.needs_slow, nullptr, l.batch_base); if (rc) return rc; }
  }
  const uint32_t count = flags[8];
  // compact -> host
  int64_t* d_keys = nullptr; uint32_t* d_knull = nullptr; uint64_t* d_vals = nullptr; uint32_t* d_cursor = nullptr;
  const size_t n = std::max<uint32_t>(count, 1);
  SD_CUDA(cudaMalloc(&d_keys, n * std::max(nk, 1) * 8));
  SD_CUDA(cudaMalloc(&d_knull, n * 4));
  SD_CUDA(cudaMalloc(&d_vals, n * ns * 8));
  SD_CUDA(cudaMalloc(&d_cursor, 64));
  int rc = hash_table_compact(p->stream, p->hash, p->hash_capacity, nk, ns, d_keys, d_knull, d_vals, d_cursor);
  if (rc) return rc;
  std::vector<int64_t> hk((size_t)count * nk);
  std::vector<uint32_t> hn(count);
  std::vector<uint64_t> hv((size_t)count * ns);
  unsigned long long counters[2] = {0, 0};
  if (count) {
    SD_CUDA(cudaMemcpyAsync(hk.data(), d_keys, hk.size() * 8, cudaMemcpyDeviceToHost, p->stream));
    SD_CUDA(cudaMemcpyAsync(hn.data(), d_knull, hn.size() * 4, cudaMemcpyDeviceToHost, p->stream));
    SD_CUDA(cudaMemcpyAsync(hv.data(), d_vals, hv.size() * 8, cudaMemcpyDeviceToHost, p->stream));
  }
  SD_CUDA(cudaMemcpyAsync(counters, p->d_counters, 16, cudaMemcpyDeviceToHost, p->stream));
  SD_CUDA(cudaStreamSynchronize(p->stream));
  if (getenv("SD_DEBUG_VERIFY")) {   // diagnostic builds (SD_JIT_DEFINES=-DSD_EXP_VERIFY=1): staged tile vs global memory
    unsigned long long dbg[8] = {0};
    SD_CUDA(cudaMemcpy(dbg, p->d_counters, 64, cudaMemcpyDeviceToHost));
    uint32_t hist[64] = {0};
    SD_CUDA(cudaMemcpy(hist, p->hash.overflow, 256, cudaMemcpyDeviceToHost));
    SD_CUDA(cudaMemset(p->hash.overflow + 16, 0, 192));
    fprintf(stderr, "[verify] by consumer warp:");
    for (int i = 0; i < 8; i++) fprintf(stderr, " %u", hist[16 + i]);
    fprintf(stderr, "   by eighth of the tile (128 rows each):");
    for (int i = 0; i < 8; i++) fprintf(stderr, " %u", hist[24 + i]);
    fprintf(stderr, "   by column:");
    for (int i = 0; i < 4; i++) fprintf(stderr, " %u", hist[32 + i]);
    fprintf(stderr, "   tiles with any mismatch (warp-level): %u of %u warp-tiles\n", hist[40], hist[41]);
    fprintf(stderr, "[verify] of the mismatches: %llu = the stage's PREVIOUS occupant (read before the copy landed), %llu = its NEXT occupant (overwritten before the read)\n", dbg[2], dbg[3]);
    fprintf(stderr, "[verify] mismatching values %llu; first: column %llu stage %llu row %llu staged %016llx true %016llx\n", dbg[4],
            (dbg[5] >> 56) - (dbg[5] ? 1 : 0), (dbg[5] >> 48) & 0xff, dbg[5] & 0xffffffffffffull, dbg[6], dbg[7]);
  }
  // STRING keys are held by reference (address of the [len][bytes] record in a resident buffer): fetch their bytes
  std::vector<std::vector<std::string>> key_strings((size_t)nk);
  for (int k = 0; k < nk && count; k++) {
    if (sp.exprs[sp.keys[k]].type != SD_STRING) continue;
    rc = fetch_string_records(p->stream, d_keys + k, (int64_t)count, nk, key_strings[(size_t)k]);
    if (rc) { cudaFree(d_keys); cudaFree(d_knull); cudaFree(d_vals); cudaFree(d_cursor); return rc; }
  }
  cudaFree(d_keys); cudaFree(d_knull); cudaFree(d_vals); cudaFree(d_cursor);
  update_agg_time(p);
  p->metrics[6] = (int64_t)(p->agg_ms * 1e6);
  p->metrics[8] = (int64_t)counters[0];
  p->metrics[11] = (int64_t)counters[0];
  const std::vector<int> types = partial_field_types(sp);
  StrMap agg_strs;
  rc = fetch_agg_strings(p, hv.data(), count, agg_strs);
  if (rc) return rc;
  std::vector<uint8_t>& out = p->finished_rows;
  out.clear();
  for (uint32_t g = 0; g < count; g++) {
    std::vector<HVal> vals;
    for (int k = 0; k < nk; k++) {
      HVal v;
      const int64_t code = hk[(size_t)g * nk + k];
      if ((hn[g] >> k) & 1u) v.isnull = true;
      else if (types[k] == SD_STRING) v.s = key_strings[(size_t)k][g];
      else if (type_is_fp(types[k])) memcpy(&v.d, &code, 8);
      else { v.i = code; v.w = code; }
      vals.push_back(v);
    }
    append_agg_fields(sp, &hv[(size_t)g * ns], vals, &agg_strs);
    emit_unsafe_row(out, types, vals);
  }
  p->finished_nrows = count;
  return 0;
}

// MODE_PROJECT: grow + replay when the record buffer was too small, then records -> UnsafeRows
// The projected rows built by the GPU (sd_rows.cu).  *done = false: this execution needs the host writer below (more than 32
// fields, or a projected STRING column whose dictionary has entries that exist only on the host: values brought by an update
// delta).
static int project_rows_on_device(sd_plan* p, unsigned long long count, bool* done) {
  *done = false;
  const bool off = getenv("SD_TUNE_HOST_ROWS") != nullptr;   // (tests compare the two writers byte for byte)
  const PlanSpec& sp = p->spec;
  const int np = (int)sp.proj.size();
  if (off || np > ROW_MAX_FIELDS || count >= (1ull << 31) - 2) return 0;
  uint8_t kinds[ROW_MAX_FIELDS];
  std::vector<int> str_cols;
  for (int j = 0; j < np; j++) {
    const sd_expr& e = sp.exprs[sp.proj[j]];
    switch (e.type) {
      case SD_BOOLEAN: kinds[j] = ROW_KIND_BOOL; break;
      case SD_BYTE: kinds[j] = ROW_KIND_1; break;
      case SD_SHORT: kinds[j] = ROW_KIND_2; break;
      case SD_INT: case SD_DATE: kinds[j] = ROW_KIND_4; break;
      case SD_FLOAT: kinds[j] = ROW_KIND_FLOAT; break;
      case SD_STRING: kinds[j] = ROW_KIND_STRING; str_cols.push_back(e.a); break;
      default: kinds[j] = ROW_KIND_8; break;
    }
  }
  RowWriterBuffers& b = p->roww;
  const int ns = (int)str_cols.size(), nb = (int)p->exec_batches.size();
  if (ns > 0) {
    std::vector<const void*> key;
    key.reserve((size_t)nb + 1);
    key.push_back(reinterpret_cast<const void*>((uintptr_t)ns));
    for (const StoredBatch* sb : p->exec_batches) key.push_back(reinterpret_cast<const void*>((uintptr_t)sb->uid));
    if (key != b.src_key || !b.d_src) {
      b.src_key.clear();
      std::vector<RowStrSrc> src((size_t)nb * ns);
      std::vector<int32_t> recoff;
      std::vector<size_t> first((size_t)nb * ns, SIZE_MAX);
      for (int bi = 0; bi < nb; bi++) {
        const StoredBatch& sb = *p->exec_batches[bi];
        for (int k = 0; k < ns; k++) {
          const StoredCol& sc = sb.cols[sb.positional ? str_cols[k] : sp.cols[str_cols[k]].table_ordinal];
          RowStrSrc& r = src[(size_t)bi * ns + k];
          if (sc.raw_str) { r = RowStrSrc{sc.dev.dict, nullptr, -1, 0}; continue; }
          if (sc.dict_rec_off.size() != sc.dict_strings.size() || !sc.dev_base) return 0;   // host-only dictionary entries
          first[(size_t)bi * ns + k] = recoff.size();
          for (int64_t o : sc.dict_rec_off) {
            if (o < 0 || o > INT32_MAX) return 0;
            recoff.push_back((int32_t)o);
          }
          r = RowStrSrc{sc.dev_base, nullptr, sc.dev.dict_n, (int32_t)sc.dict_strings.size()};
        }
      }
      if (recoff.size() * 4 > b.recoff_cap) {
        if (b.d_recoff) cudaFree(b.d_recoff);
        b.d_recoff = nullptr; b.recoff_cap = 0;
        SD_CUDA(cudaMalloc(reinterpret_cast<void**>(&b.d_recoff), recoff.size() * 4 + 4096));
        b.recoff_cap = recoff.size() * 4 + 4096;
      }
      if (src.size() * sizeof(RowStrSrc) > b.src_cap) {
        if (b.d_src) cudaFree(b.d_src);
        b.d_src = nullptr; b.src_cap = 0;
        SD_CUDA(cudaMalloc(reinterpret_cast<void**>(&b.d_src), src.size() * sizeof(RowStrSrc) + 4096));
        b.src_cap = src.size() * sizeof(RowStrSrc) + 4096;
      }
      for (size_t i = 0; i < src.size(); i++) if (first[i] != SIZE_MAX) src[i].rec_off = b.d_recoff + first[i];
      // (pageable sources: both copies are staged by the driver before the calls return)
      if (!recoff.empty()) SD_CUDA(cudaMemcpyAsync(b.d_recoff, recoff.data(), recoff.size() * 4, cudaMemcpyHostToDevice, p->stream));
      SD_CUDA(cudaMemcpyAsync(b.d_src, src.data(), src.size() * sizeof(RowStrSrc), cudaMemcpyHostToDevice, p->stream));
      SD_CUDA(cudaStreamSynchronize(p->stream));
      b.src_key.swap(key);
    }
  }
  int64_t total = 0;
  int rc = device_write_rows(p->stream, reinterpret_cast<const uint64_t*>(p->d_out), (int64_t)count, np, kinds, nb, b, &total);
  if (rc) return rc;
  unsigned long long counters[2] = {0, 0};
  SD_CUDA(cudaMemcpyAsync(counters, p->d_counters, 16, cudaMemcpyDeviceToHost, p->stream));
  SD_CUDA(cudaStreamSynchronize(p->stream));
  update_agg_time(p);
  p->metrics[6] = (int64_t)(p->agg_ms * 1e6);
  p->metrics[8] = (int64_t)counters[0];
  p->metrics[11] = (int64_t)counters[0];
  p->finished_rows.clear();
  p->dev_rows_len = total;
  p->finished_nrows = (int64_t)count;
  *done = true;
  return 0;
}

// rows of this execution on the host (callers that post-process them: the exchange, the partial merge)
static int rows_to_host(sd_plan* p) {
  if (p->dev_rows_len < 0) return 0;
  p->finished_rows.resize((size_t)p->dev_rows_len);
  if (p->dev_rows_len) SD_CUDA(cudaMemcpyAsync(p->finished_rows.data(), p->roww.d_rows, (size_t)p->dev_rows_len, cudaMemcpyDeviceToHost, p->stream));
  SD_CUDA(cudaStreamSynchronize(p->stream));
  p->dev_rows_len = -1;
  return 0;
}

int finish_project(sd_plan* p) {
  const PlanSpec& sp = p->spec;
  const int np = (int)sp.proj.size();
  const int64_t rec = 8 + 8 * (int64_t)np;
  int rc = ensure_out(p, p->out_cap ? p->out_cap : 1024);
  if (rc) return rc;
  unsigned long long count = 0;
  for (;;) {
    SD_CUDA(cudaMemcpyAsync(&count, p->d_out_count, 8, cudaMemcpyDeviceToHost, p->stream));
    SD_CUDA(cudaStreamSynchronize(p->stream));
    if ((int64_t)count <= p->out_cap) break;
    rc = ensure_out(p, (int64_t)count + (int64_t)count / 8 + 1024);
    if (rc) return rc;
    SD_CUDA(cudaMemsetAsync(p->d_out_count, 0, 8, p->stream));
    SD_CUDA(cudaMemsetAsync(p->d_counters, 0, 64, p->stream));
    for (auto& l : p->launch_log) { rc = launch_scan(p, l.d_batches, l.d_prefix, l.nbatches, l.total_chunks, l.needs_slow, nullptr, l.batch_base); if (rc) return rc; }
  }
  {
    bool done = false;
    rc = project_rows_on_device(p, count, &done);
    if (rc || done) return rc;
  }
  static const bool dbg = getenv("SD_DEBUG_TIMING") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) { if (dbg) fprintf(stderr, "[finish_project] %s at %.2f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count()); };
  // (page-locked staging for the records: a pageable destination makes the driver bounce the copy through its own buffers)
  const size_t rec_bytes = (size_t)count * (size_t)rec;
  if (rec_bytes > p->h_recs_cap) {
    if (p->h_recs) cudaFreeHost(p->h_recs);
    p->h_recs = nullptr; p->h_recs_cap = 0;
    SD_CUDA(cudaMallocHost(&p->h_recs, rec_bytes + rec_bytes / 4 + 4096));
    p->h_recs_cap = rec_bytes + rec_bytes / 4 + 4096;
  }
  const uint64_t* recs = p->h_recs;
  unsigned long long counters[2] = {0, 0};
  if (count) SD_CUDA(cudaMemcpyAsync(p->h_recs, p->d_out, rec_bytes, cudaMemcpyDeviceToHost, p->stream));
  SD_CUDA(cudaMemcpyAsync(counters, p->d_counters, 16, cudaMemcpyDeviceToHost, p->stream));
  SD_CUDA(cudaStreamSynchronize(p->stream));
  update_agg_time(p);
  p->metrics[6] = (int64_t)(p->agg_ms * 1e6);
  p->metrics[8] = (int64_t)counters[0];
  p->metrics[11] = (int64_t)counters[0];
  std::vector<int> types;
  std::vector<int> str_col(np, -1);
  for (int j = 0; j < np; j++) {
    const sd_expr& e = sp.exprs[sp.proj[j]];
    types.push_back(field_type(e.type, e.type == SD_DECIMAL ? decimal_ps(sp, sp.proj[j]) : 0));
    if (e.type == SD_STRING) str_col[j] = e.a;
  }
  lap("records on the host");
  // strings of raw (variable-width) batches are projected by reference: record = position in the batch's body
  std::vector<int64_t> raw_ptrs;
  for (unsigned long long i = 0; i < count && !str_col.empty(); i++) {
    const uint64_t* r = &recs[(size_t)i * (size_t)(rec / 8)];
    const uint32_t bidx = (uint32_t)(r[0] & 0xffffffffu), pnull = (uint32_t)(r[0] >> 32);
    if (bidx >= p->exec_batches.size()) return set_error(SD_ERR_CUDA, "corrupt projection record (batch %u)", bidx);
    const StoredBatch& sb = *p->exec_batches[bidx];
    for (int j = 0; j < np; j++) {
      if (str_col[j] < 0 || ((pnull >> j) & 1u)) continue;
      const StoredCol& sc = sb.cols[sb.positional ? str_col[j] : sp.cols[str_col[j]].table_ordinal];
      if (sc.raw_str) raw_ptrs.push_back((int64_t)(uintptr_t)(sc.dev.dict + (uint32_t)r[1 + j]));
    }
  }
  std::vector<std::string> raw_strings;
  if (!raw_ptrs.empty()) {
    int64_t* d_ptrs = nullptr;
    SD_CUDA(cudaMalloc(&d_ptrs, raw_ptrs.size() * 8));
    cudaError_t ce = cudaMemcpyAsync(d_ptrs, raw_ptrs.data(), raw_ptrs.size() * 8, cudaMemcpyHostToDevice, p->stream);
    rc = ce == cudaSuccess ? fetch_string_records(p->stream, d_ptrs, (int64_t)raw_ptrs.size(), 1, raw_strings) : set_error(SD_ERR_CUDA, "cudaMemcpyAsync: %s", cudaGetErrorString(ce));
    cudaFree(d_ptrs);
    if (rc) return rc;
  }
  // records -> UnsafeRows, written in place (no per-row value objects: C4 emits ~1 M rows per execution)
  size_t next_raw = 0;
  std::vector<uint8_t>& out = p->finished_rows;
  out.clear();
  const int64_t bits = ((np + 63) / 64) * 8, fixed = bits + 8 * (int64_t)np;
  // pass 1: sizes (strings are the only variable part)
  size_t total = 0;
  {
    size_t nr = 0;
    for (unsigned long long i = 0; i < count; i++) {
      const uint64_t* r = &recs[(size_t)i * (size_t)(rec / 8)];
      const uint32_t bidx = (uint32_t)(r[0] & 0xffffffffu), pnull = (uint32_t)(r[0] >> 32);
      const StoredBatch& sb = *p->exec_batches[bidx];
      int64_t var = 0;
      for (int j = 0; j < np; j++) {
        if (str_col[j] < 0 || ((pnull >> j) & 1u)) continue;
        const StoredCol& sc = sb.cols[sb.positional ? str_col[j] : sp.cols[str_col[j]].table_ordinal];
        if (sc.raw_str) { var += ((int64_t)raw_strings[nr++].size() + 7) & ~int64_t(7); continue; }
        const int64_t code = (int64_t)r[1 + j];
        if (code == sc.dev.dict_n && code >= 0) continue;   // NULL code
        if (code < 0 || code >= (int64_t)sc.dict_strings.size()) return set_error(SD_ERR_CUDA, "dictionary code %lld out of range", (long long)code);
        var += ((int64_t)sc.dict_strings[(size_t)code].size() + 7) & ~int64_t(7);
      }
      total += (size_t)(8 + fixed + var);
    }
  }
  lap("sizes");
  out.assign(total, 0);
  lap("zero fill");
  uint8_t* w = out.data();
  for (unsigned long long i = 0; i < count; i++) {
    const uint64_t* r = &recs[(size_t)i * (size_t)(rec / 8)];
    const uint32_t bidx = (uint32_t)(r[0] & 0xffffffffu), pnull = (uint32_t)(r[0] >> 32);
    const StoredBatch& sb = *p->exec_batches[bidx];
    uint8_t* row = w + 8;
    int64_t voff = fixed;
    for (int j = 0; j < np; j++) {
      uint8_t* slot = row + bits + 8 * (int64_t)j;
      if ((pnull >> j) & 1u) { row[j >> 3] |= (uint8_t)(1u << (j & 7)); continue; }
      const uint64_t raw = r[1 + j];
      const int t = ft_base(types[j]);
      if (t == SD_STRING) {
        const StoredCol& sc = sb.cols[sb.positional ? str_col[j] : sp.cols[str_col[j]].table_ordinal];
        const std::string* sv;
        if (sc.raw_str) sv = &raw_strings[next_raw++];
        else {
          const int64_t code = (int64_t)raw;
          if (code == sc.dev.dict_n) { row[j >> 3] |= (uint8_t)(1u << (j & 7)); continue; }
          sv = &sc.dict_strings[(size_t)code];
        }
        const int64_t ol = (voff << 32) | (int64_t)sv->size();
        memcpy(slot, &ol, 8);
        memcpy(row + voff, sv->data(), sv->size());
        voff += ((int64_t)sv->size() + 7) & ~int64_t(7);
        continue;
      }
      switch (t) {
        case SD_BOOLEAN: slot[0] = raw != 0; break;
        case SD_BYTE: memcpy(slot, &raw, 1); break;
        case SD_SHORT: memcpy(slot, &raw, 2); break;
        case SD_INT: case SD_DATE: memcpy(slot, &raw, 4); break;
        case SD_FLOAT: { double d; memcpy(&d, &raw, 8); const float f = (float)d; memcpy(slot, &f, 4); break; }
        default: memcpy(slot, &raw, 8); break;   // LONG, TIMESTAMP, DOUBLE (its bits), DECIMAL(p <= 18)
      }
    }
    const int64_t sz = voff;
    memcpy(w, &sz, 8);
    w += 8 + sz;
  }
  lap("rows written");
  p->finished_nrows = (int64_t)count;
  return 0;
}

int ensure_private_store(sd_plan* p) {
  if (p->priv) return 0;
  // the private store's "table" is exactly the plan's scan columns
  std::vector<sd_column> schema(p->spec.cols);
  return sd_store_create(p->device, (int32_t)schema.size(), schema.data(), &p->priv);
}

}  // namespace

extern "C" {

int sd_init(int device) {
  int n = 0;
  SD_CUDA(cudaGetDeviceCount(&n));
  if (device < 0 || device >= n) return set_error(SD_ERR_INVALID, "sd_init: device %d of %d", device, n);
  SD_CUDA(cudaSetDevice(device));
  t_device = device;
  return 0;
}
int sd_device_count(int* out) { SD_CUDA(cudaGetDeviceCount(out)); return 0; }
const char* sd_version(void) { return "snappydata_b200 0.1.0 (sm_100a)"; }

int sdx_stats_pass(const sd_plan_desc* desc, const sd_literal* lits, int32_t nlits, const void* stats, int64_t stats_len,
                   int32_t stats_ncols, int32_t num_rows, int32_t* pass) {
  return stats_pass_hook(desc, lits, nlits, stats, stats_len, stats_ncols, num_rows, pass);
}

int sd_plan_create(const sd_plan_desc* desc, sd_plan** out) {
  if (!out) return set_error(SD_ERR_INVALID, "sd_plan_create: null out");
  std::unique_ptr<sd_plan, void (*)(sd_plan*)> p(new sd_plan(), sd_plan_destroy);   // a failure below releases what was created
  std::string err;
  int rc = analyze_plan(desc, p->spec, err);
  if (rc) return set_error(rc, "sd_plan_create: %s", err.c_str());
  p->device = t_device;
  SD_CUDA(cudaSetDevice(p->device));
  // kernel: ahead-of-time compiled plan, else NVRTC
  {
    CodegenOptions opt;
    sd_plan_desc dv = p->spec.desc_view();
    rc = resolve_kernel(dv, opt, p->device, &p->kernel, nullptr);
    if (rc) return rc;
    p->active = &p->kernel;
  }
  p->kernel_name = p->kernel.origin + ":" + p->kernel.name;
  cudaDeviceProp prop;
  SD_CUDA(cudaGetDeviceProperties(&prop, p->device));
  p->num_sms = prop.multiProcessorCount;
  p->smem_optin = (int)prop.sharedMemPerBlockOptin;
  SD_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  p->own_stream = true;
  SD_CUDA(cudaEventCreate(&p->ev_start));
  SD_CUDA(cudaEventCreate(&p->ev_stop));
  SD_CUDA(cudaMalloc(&p->d_ticket, 64));
  SD_CUDA(cudaMemset(p->d_ticket, 0, 64));
  rc = ensure_result(p.get(), 64);   // d_state: counters + room for a small group table
  if (rc) return rc;
  p->scratch.device = p->device;
  p->scratch.slab_bytes = size_t(8) << 20;
  p->cache_arena.device = p->device;
  p->cache_arena.slab_bytes = size_t(8) << 20;
  const int nk = (int)p->spec.keys.size();
  p->key_ids.resize(nk);
  p->key_vals.resize(nk);
  p->key_null_id.assign(nk, -1);
  p->lits.resize(p->spec.literal_types.size());
  p->lit_strs.resize(p->spec.literal_types.size());
  for (size_t i = 0; i < p->lits.size(); i++) { memset(&p->lits[i], 0, sizeof(sd_literal)); p->lits[i].type = p->spec.literal_types[i]; }
  if (p->spec.mode == MODE_GROUPS) p->chunk_rows = 2 * CHUNK_ROWS;   // measured: tools/sweep.sh, profiles/r01_tuning.txt
  if (const char* e = getenv("SD_TUNE_CHUNK_ROWS")) { int v = atoi(e); if (v >= 2048 && v % 2048 == 0 && v <= (1 << 20)) p->chunk_rows = v; }
  p->lits_set = p->lits.empty();
  *out = p.release();
  return 0;
}

int sd_plan_set_literals(sd_plan* p, const sd_literal* vals, int32_t n) {
  if (!p) return set_error(SD_ERR_INVALID, "null plan");
  if (n != (int)p->lits.size()) return set_error(SD_ERR_INVALID, "sd_plan_set_literals: plan has %zu literal slots, got %d", p->lits.size(), n);
  if (!p->pending.empty()) return set_error(SD_ERR_STATE, "sd_plan_set_literals: batches already submitted for this execution");
  for (int i = 0; i < n; i++) {
    p->lits[i] = vals[i];
    p->lit_strs[i].assign(vals[i].s ? vals[i].s : "", vals[i].s ? (size_t)std::max(0, vals[i].slen) : 0);
    p->lits[i].s = p->lit_strs[i].data();
    p->lits[i].slen = (int32_t)p->lit_strs[i].size();
  }
  p->lits_set = true;
  p->litpool_dirty = true;
  return 0;
}

int sd_plan_set_option(sd_plan* p, int32_t option, int64_t value) {
  if (!p) return set_error(SD_ERR_INVALID, "null plan");
  if (option == SD_OPT_RETAIN_BUFFERS) {
    int rc = ensure_private_store(p);
    if (rc) return rc;
    p->priv->retain_buffers = value != 0;
    return 0;
  }
  return set_error(SD_ERR_INVALID, "unknown option %d", option);
}

int sd_plan_set_stream(sd_plan* p, void* cuda_stream) {
  if (!p) return set_error(SD_ERR_INVALID, "null plan");
  if (p->own_stream && p->stream) { cudaSetDevice(p->device); cudaStreamDestroy(p->stream); }
  p->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  p->own_stream = false;
  return 0;
}

const char* sd_plan_kernel_name(sd_plan* p) { return p ? p->kernel_name.c_str() : ""; }

int sd_batch_submit(sd_plan* p, const sd_batch* b) {
  if (!p || !b) return set_error(SD_ERR_INVALID, "sd_batch_submit: null argument");
  if (!p->lits_set) return set_error(SD_ERR_STATE, "sd_batch_submit: literals not set");
  if (b->ncols != (int)p->spec.cols.size()) return set_error(SD_ERR_INVALID, "sd_batch_submit: batch has %d columns, plan scans %zu", b->ncols, p->spec.cols.size());
  SD_CUDA(cudaSetDevice(p->device));
  p->metrics[2]++;   // columnBatchesSeen
  // stats-row skipping before any byte moves (ColumnTableScan.scala:532-543)
  if (b->stats_row && b->stats_len > 0 && p->spec.filter >= 0) {
    StatEval ev{p->spec, p->lits, reinterpret_cast<const uint8_t*>(b->stats_row), b->stats_len, 1 + 3 * b->stats_ncols, b->num_rows};
    Tri r;
    if (ev.eval(p->spec.filter, &r) && !r.isnull && !r.v) { p->metrics[5]++; return 0; }
  }
  int rc = ensure_private_store(p);
  if (rc) return rc;
  const int64_t before = p->priv->h2d_bytes;
  // the private store is indexed by scan column: buffers arrive in plan order already
  sd_batch local = *b;
  rc = store_put(p->priv, &local, nullptr);
  if (rc) return rc;
  p->metrics[10] += p->priv->h2d_bytes - before;
  p->priv->batches.back()->positional = true;
  const StoredBatch* sb = p->priv->batches.back().get();
  // scan columns of the private store are positional: re-point table ordinals on the fly in build_scan
  p->pending.push_back({sb});
  p->pending_bytes += p->priv->h2d_bytes - before;
  // (compressed inputs: one expansion launch per flush, as long as its longest buffer; the launches of successive
  // flushes overlap on their own streams, so the same flush size serves both)
  int64_t threshold = int64_t(256) << 20;
  if (const char* e = getenv("SD_TUNE_FLUSH_MB")) { const long v = atol(e); if (v >= 1 && v <= 65536) threshold = int64_t(v) << 20; }
  if (p->pending_bytes >= threshold) return flush_pending(p);
  return 0;
}

int sd_plan_scan_store(sd_plan* p, sd_store* s, const int32_t* bucket_ids, int32_t nbuckets) {
  if (!p || !s) return set_error(SD_ERR_INVALID, "sd_plan_scan_store: null argument");
  if (!p->lits_set) return set_error(SD_ERR_STATE, "sd_plan_scan_store: literals not set");
  if (s->device != p->device) return set_error(SD_ERR_INVALID, "store lives on device %d, plan on %d", s->device, p->device);
  SD_CUDA(cudaSetDevice(p->device));
  int rc = flush_pending(p);
  if (rc) return rc;
  // snapshot under the store's lock: the batches present NOW are what this execution scans (ingest may continue meanwhile;
  // the reference's scan likewise sees the batches of its snapshot, ColumnFormatIterator over the bucket's entries)
  std::vector<const StoredBatch*> snapshot;
  int64_t snap_version;
  {
    std::lock_guard<std::mutex> lock(s->mu);
    rc = store_flush_lz4(s);
    if (rc) return rc;
    rc = store_lz4_check(s);   // resident stores: expansions happen once, before the first scan
    if (rc) return rc;
    snapshot.reserve(s->batches.size());
    for (auto& sbp : s->batches) snapshot.push_back(sbp.get());
    snap_version = s->version;
  }
  for (auto& c : p->spec.cols) {
    if (c.table_ordinal < 0 || c.table_ordinal >= (int)s->schema.size()) return set_error(SD_ERR_INVALID, "plan column ordinal %d outside the store schema", c.table_ordinal);
    if (s->schema[c.table_ordinal].type != c.type || s->schema[c.table_ordinal].nullable != c.nullable)
      return set_error(SD_ERR_INVALID, "plan column %d (type %d nullable %d) does not match the store schema (type %d nullable %d)", c.table_ordinal,
                       c.type, c.nullable, s->schema[c.table_ordinal].type, s->schema[c.table_ordinal].nullable);
  }
  std::vector<int32_t> buckets(bucket_ids, bucket_ids + (bucket_ids ? nbuckets : 0));
  const std::string lk = literal_key(p);
  sd_plan::ScanCache& c = p->cache;
  const bool same_query = c.valid && c.store == s && c.buckets == buckets && c.lit_key == lk;
  // descriptors (+ stats skipping, aux tables) of snapshot[from, end) as one segment in the cache arena
  auto build_segment = [&](size_t from, sd_plan::ScanSegment* seg, int64_t* seen, int64_t* skipped) -> int {
    std::vector<const StoredBatch*> list;
    for (size_t i = from; i < snapshot.size(); i++) {
      const StoredBatch& sb = *snapshot[i];
      if (!buckets.empty() && std::find(buckets.begin(), buckets.end(), sb.bucket_id) == buckets.end()) continue;
      (*seen)++;
      if (!batch_passes_stats(p, sb)) { (*skipped)++; continue; }
      list.push_back(&sb);
    }
    BuiltScan bs;
    int rc2 = build_scan(p, list, p->cache_arena, p->stream, &bs);
    if (rc2) return rc2;
    if (bs.needs_hash && p->spec.mode == MODE_GROUPS) return -1000;   // the caller switches the plan and rebuilds everything
    seg->d_batches = bs.d_batches; seg->d_prefix = bs.d_prefix; seg->nbatches = bs.nbatches; seg->total_chunks = bs.total_chunks;
    seg->needs_slow = bs.needs_slow; seg->rows = bs.rows; seg->algo_bytes = bs.algo_bytes;
    seg->updated_cols = bs.updated_cols; seg->deleted_batches = bs.deleted_batches;
    seg->covered = snapshot.size() - from;
    seg->batches.swap(list);
    return 0;
  };
  bool rebuilt = false;
  if (same_query && c.version != snap_version && snapshot.size() >= c.snap_uids.size() && !getenv("SD_TUNE_NO_INCREMENTAL_SCAN")) {
    // the store changed: still the batches we know, plus new ones at the end?
    bool prefix = true;
    for (size_t i = 0; i < c.snap_uids.size() && prefix; i++) prefix = snapshot[i]->uid == c.snap_uids[i];
    if (prefix) {
      size_t from = c.snap_uids.size();
      // keep the tail short: fold the small segments after the first one into the new segment when there are several
      if (c.segs.size() > 4 && c.consolidations < 64) {
        size_t keep = c.segs[0].covered;
        c.segs.resize(1);
        // (seen / skipped of the folded segments are recounted by build_segment)
        int64_t seen0 = 0, skipped0 = 0;
        for (size_t i = 0; i < keep; i++) {
          const StoredBatch& sb = *snapshot[i];
          if (!buckets.empty() && std::find(buckets.begin(), buckets.end(), sb.bucket_id) == buckets.end()) continue;
          seen0++;
        }
        skipped0 = seen0 - (int64_t)c.segs[0].batches.size();
        c.seen = seen0; c.skipped = skipped0;
        from = keep;
        c.consolidations++;
      }
      if (c.consolidations < 64) {
        sd_plan::ScanSegment seg;
        rc = build_segment(from, &seg, &c.seen, &c.skipped);
        if (rc == 0) {
          if (seg.covered) c.segs.push_back(std::move(seg));
          c.snap_uids.resize(snapshot.size());
          for (size_t i = from; i < snapshot.size(); i++) c.snap_uids[i] = snapshot[i]->uid;
          c.version = snap_version;
          rebuilt = true;
        } else if (rc != -1000) {
          return rc;
        }
      }
    }
  }
  if (!rebuilt && !(same_query && c.version == snap_version)) {
    // (re)build everything: stats skipping + descriptors + tables, kept on the device for repeated executions
    c.valid = false;
    c.segs.clear();
    c.consolidations = 0;
    p->cache_arena.reset();
    sd_plan::ScanSegment seg;
    int64_t seen = 0, skipped = 0;
    rc = build_segment(0, &seg, &seen, &skipped);
    if (rc == -1000) {
      rc = switch_to_hash(p);
      if (rc) return rc;
      p->cache_arena.reset();
      seen = skipped = 0;
      seg = sd_plan::ScanSegment();
      rc = build_segment(0, &seg, &seen, &skipped);
    }
    if (rc) return rc;
    c.store = s; c.version = snap_version; c.buckets = buckets; c.lit_key = lk;
    c.seen = seen; c.skipped = skipped;
    c.segs.push_back(std::move(seg));
    c.snap_uids.resize(snapshot.size());
    for (size_t i = 0; i < snapshot.size(); i++) c.snap_uids[i] = snapshot[i]->uid;
    c.valid = true;
  }
  p->metrics[2] += c.seen;
  p->metrics[5] += c.skipped;
  for (const sd_plan::ScanSegment& seg : c.segs) {
    p->metrics[3] += seg.updated_cols;
    p->metrics[4] += seg.deleted_batches;
    p->metrics[9] += seg.algo_bytes;
  }
  bool launched = false;
  for (const sd_plan::ScanSegment& seg : c.segs) {
    if (seg.nbatches == 0 && launched) continue;
    rc = launch_scan(p, seg.d_batches, seg.d_prefix, seg.nbatches, seg.total_chunks, seg.needs_slow, &seg.batches);
    if (rc) return rc;
    launched = true;
  }
  return 0;
}

// dense / no-key state (the [ngroups][slots] table the kernel leaves) -> partial rows appended to `out`:
// UnsafeRow(group keys ++ aggregate buffers) (SnappyHashAggregateExec.scala:1148-1178).  The key dictionaries are arguments
// because the exchange's dense form carries ANOTHER rank's state and dictionaries (ids are private to a partition).
static int64_t dense_rows_from_state(const PlanSpec& sp, const uint64_t* h, int ngroups, const int32_t* radix, const std::vector<int>& key_null_id,
                                     const std::vector<std::vector<std::string>>& key_vals, const StrMap* agg_strs, std::vector<uint8_t>& out) {
  const int ns = (int)sp.slots.size(), nk = (int)sp.keys.size();
  const std::vector<int> types = partial_field_types(sp);
  int64_t nrows = 0;
  std::vector<HVal> vals;
  for (int g = 0; g < ngroups; g++) {
    const uint64_t* sv = &h[(size_t)g * ns];
    if (nk > 0 && sv[sp.rows_slot] == 0) continue;   // group never seen
    vals.clear();
    int rem = g, idx[MAX_KEYS];
    for (int k = nk - 1; k >= 0; k--) { idx[k] = rem % radix[k]; rem /= radix[k]; }
    for (int k = 0; k < nk; k++) {
      HVal v;
      if (idx[k] == key_null_id[k]) v.isnull = true; else v.s = key_vals[k][idx[k]];
      vals.push_back(v);
    }
    append_agg_fields(sp, sv, vals, agg_strs);
    emit_unsafe_row(out, types, vals);
    nrows++;
  }
  return nrows;
}

// partial rows of this execution's dense / no-key result -> p->finished_rows
static int finish_dense(sd_plan* p) {
  const PlanSpec& sp = p->spec;
  const int ns = (int)sp.slots.size();
  int rc = 0;
  if (!p->result_init) { rc = init_result(p, 1); if (rc) return rc; p->ngroups = 1; }
  const size_t ne = (size_t)p->ngroups * ns;
  rc = ensure_result(p, ne);
  if (rc) return rc;
  SD_CUDA(cudaMemcpyAsync(p->h_pinned, p->d_state, (STATE_HDR + ne) * 8, cudaMemcpyDeviceToHost, p->stream));   // counters + table in one copy
  SD_CUDA(cudaStreamSynchronize(p->stream));
  const uint64_t* h = p->h_pinned + STATE_HDR;
  const unsigned long long counters[2] = {p->h_pinned[0], p->h_pinned[1]};
  update_agg_time(p);
  p->metrics[6] = (int64_t)(p->agg_ms * 1e6);
  p->metrics[8] = (int64_t)counters[0];
  p->metrics[11] = (int64_t)counters[0];
  StrMap agg_strs;
  {
    std::vector<uint64_t> copy(h, h + ne);   // (the pinned mirror is reused by the fetch's own read-backs)
    rc = fetch_agg_strings(p, copy.data(), (size_t)p->ngroups, agg_strs);
    if (rc) return rc;
  }
  std::vector<uint8_t>& out = p->finished_rows;
  out.clear();
  p->finished_nrows = dense_rows_from_state(sp, h, p->ngroups, p->radix, p->key_null_id, p->key_vals, &agg_strs, out);
  return 0;
}

// run what is pending and materialise this partition's partial rows in p->finished_rows (once per execution)
static int launch_what_is_pending(sd_plan* p) {
  SD_CUDA(cudaSetDevice(p->device));
  int rc = flush_pending(p);
  if (rc) return rc;
  if (p->priv) { rc = store_lz4_check(p->priv); if (rc) return rc; }   // a corrupt compressed buffer fails the execution
  return 0;
}
static int collect_partial_rows(sd_plan* p) {
  int rc = launch_what_is_pending(p);
  if (rc) return rc;
  if (p->finished_nrows >= 0) return 0;
  const int mode = p->spec.mode;
  rc = mode == MODE_HASH ? finish_hash(p) : mode == MODE_PROJECT ? finish_project(p) : finish_dense(p);
  if (rc) return rc;
  p->metrics[0] = p->finished_nrows;
  return 0;
}

int sd_plan_finish(sd_plan* p, void* out_rows, int64_t cap, int64_t* out_len, int64_t* out_nrows) {
  if (!p || !out_len) return set_error(SD_ERR_INVALID, "sd_plan_finish: null argument");
  int rc = collect_partial_rows(p);
  if (rc) return rc;
  if (p->dev_rows_len >= 0) {   // projected rows written by the GPU: one copy, straight into the caller's buffer
    *out_len = p->dev_rows_len;
    if (out_nrows) *out_nrows = p->finished_nrows;
    if (p->dev_rows_len > cap) return set_error(SD_ERR_OVERFLOW, "sd_plan_finish: output needs %lld bytes", (long long)p->dev_rows_len);
    if (p->dev_rows_len) SD_CUDA(cudaMemcpyAsync(out_rows, p->roww.d_rows, (size_t)p->dev_rows_len, cudaMemcpyDeviceToHost, p->stream));
    SD_CUDA(cudaStreamSynchronize(p->stream));
    return 0;
  }
  *out_len = (int64_t)p->finished_rows.size();
  if (out_nrows) *out_nrows = p->finished_nrows;
  if ((int64_t)p->finished_rows.size() > cap) return set_error(SD_ERR_OVERFLOW, "sd_plan_finish: output needs %zu bytes", p->finished_rows.size());
  if (!p->finished_rows.empty()) memcpy(out_rows, p->finished_rows.data(), p->finished_rows.size());
  return 0;
}

int sd_plan_reset(sd_plan* p) {
  if (!p) return set_error(SD_ERR_INVALID, "null plan");
  SD_CUDA(cudaSetDevice(p->device));
  SD_CUDA(cudaStreamSynchronize(p->stream));
  p->pending.clear();
  p->pending_bytes = 0;
  p->result_init = false;
  p->hash_init = false;
  p->launch_log.clear();
  p->exec_batches.clear();
  p->finished_nrows = -1;
  p->dev_rows_len = -1;
  if (p->d_out_count) SD_CUDA(cudaMemsetAsync(p->d_out_count, 0, 8, p->stream));
  p->have_timing = false;
  p->ev_used = 0;
  p->agg_ms = 0;
  SD_CUDA(cudaMemsetAsync(p->d_counters, 0, 64, p->stream));
  memset(p->metrics, 0, sizeof(p->metrics));
  p->scratch.reset();
  p->pinned.reset();
  if (p->priv) {
    cudaStreamSynchronize(p->priv->copy_stream);
    for (int k = 0; k + 1 < p->priv->num_copy_streams; k++) cudaStreamSynchronize(p->priv->extra_streams[k]);
    p->priv->batches.clear(); p->priv->arena.reset(); p->priv->version++; p->priv->h2d_bytes = 0;
    p->priv->pending_lz4.clear();
    store_lz4_check(p->priv);   // waits for queued expansions (their result is being discarded) and releases the staging
    p->priv->lz4_stage.reset();
  }
  // key dictionaries persist across executions of a cached plan only if the scan cache refers to them
  if (!p->cache.valid) {
    for (auto& m : p->key_ids) m.clear();
    for (auto& v : p->key_vals) v.clear();
    std::fill(p->key_null_id.begin(), p->key_null_id.end(), -1);
  }
  return 0;
}

int sd_plan_metrics(sd_plan* p, int64_t out[SD_NUM_METRICS]) {
  if (!p) return set_error(SD_ERR_INVALID, "null plan");
  memcpy(out, p->metrics, sizeof(p->metrics));
  return 0;
}

void sd_plan_destroy(sd_plan* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  if (p->own_stream && p->stream) cudaStreamDestroy(p->stream);
  if (p->ev_start) cudaEventDestroy(p->ev_start);
  if (p->ev_stop) cudaEventDestroy(p->ev_stop);
  for (auto& e : p->ev_pairs) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  if (p->d_state) cudaFree(p->d_state);
  if (p->h_pinned) cudaFreeHost(p->h_pinned);
  if (p->d_partials) cudaFree(p->d_partials);
  hash_free(p);
  if (p->d_out) cudaFree(p->d_out);
  if (p->h_recs) cudaFreeHost(p->h_recs);
  p->roww.release();
  if (p->d_out_count) cudaFree(p->d_out_count);
  if (p->d_hash_ident) cudaFree(p->d_hash_ident);
  if (p->d_ticket) cudaFree(p->d_ticket);
  if (p->d_litpool) cudaFree(p->d_litpool);
  if (p->priv) sd_store_destroy(p->priv);
  delete p;
}

// ---- dense partial table export/import for an on-device exchange (NCCL all-reduce) --------------
int sd_plan_partials_layout(sd_plan* p, int32_t* ngroups, int32_t* nslots, int32_t* slot_is_f64) {
  if (!p) return set_error(SD_ERR_INVALID, "null plan");
  const int ns = (int)p->spec.slots.size();
  if (ngroups) *ngroups = p->ngroups;
  if (nslots) *nslots = ns;
  if (slot_is_f64) for (int s = 0; s < ns; s++) { const int op = p->spec.slots[s].op; slot_is_f64[s] = op == SLOT_ADD_F64 ? 1 : (op == SLOT_ADD_I64 ? 0 : -1); }
  return 0;
}
int sd_plan_export_partials(sd_plan* p, void* dev_out, int64_t cap_bytes) {
  if (!p || !dev_out) return set_error(SD_ERR_INVALID, "null argument");
  if (p->spec.mode == MODE_HASH || p->spec.mode == MODE_PROJECT) return set_error(SD_ERR_UNSUPPORTED, "dense partials exist only for no-key / dictionary-keyed plans");
  SD_CUDA(cudaSetDevice(p->device));
  int rc = flush_pending(p);
  if (rc) return rc;
  if (!p->result_init) { rc = init_result(p, 1); if (rc) return rc; p->ngroups = 1; }
  const size_t bytes = (size_t)p->ngroups * p->spec.slots.size() * 8;
  if ((int64_t)bytes > cap_bytes) return set_error(SD_ERR_OVERFLOW, "export needs %zu bytes", bytes);
  SD_CUDA(cudaMemcpyAsync(dev_out, p->d_result, bytes, cudaMemcpyDeviceToDevice, p->stream));
  return 0;
}
int sd_plan_import_partials(sd_plan* p, const void* dev_in, int64_t bytes) {
  if (!p || !dev_in) return set_error(SD_ERR_INVALID, "null argument");
  if (p->spec.mode == MODE_HASH || p->spec.mode == MODE_PROJECT) return set_error(SD_ERR_UNSUPPORTED, "dense partials exist only for no-key / dictionary-keyed plans");
  p->finished_nrows = -1;
  p->dev_rows_len = -1;
  SD_CUDA(cudaSetDevice(p->device));
  const size_t want = (size_t)p->ngroups * p->spec.slots.size() * 8;
  if ((size_t)bytes != want) return set_error(SD_ERR_INVALID, "import expects %zu bytes", want);
  SD_CUDA(cudaMemcpyAsync(p->d_result, dev_in, want, cudaMemcpyDeviceToDevice, p->stream));
  return 0;
}

// ---- row-buffer rows: the hybrid scan's first element (ColumnTableScan.scala:572-588) ----------------
// nrows UnsafeRows of the plan's scan columns -> one Uncompressed/Dictionary pseudo-batch.
int sd_rows_submit(sd_plan* p, const void* rows, int64_t len, int32_t nrows) {
  if (!p || (!rows && nrows > 0)) return set_error(SD_ERR_INVALID, "sd_rows_submit: null argument");
  if (nrows <= 0) return 0;
  const int nc = (int)p->spec.cols.size();
  std::vector<std::vector<uint8_t>> bufs(nc);
  std::vector<std::vector<HVal>> colv(nc, std::vector<HVal>((size_t)nrows));
  const uint8_t* r = reinterpret_cast<const uint8_t*>(rows);
  int64_t pos = 0;
  for (int i = 0; i < nrows; i++) {
    if (pos + 8 > len) return set_error(SD_ERR_INVALID, "sd_rows_submit: truncated row stream");
    const int64_t sz = rd_i64(r + pos);
    if (sz < 0 || pos + 8 + sz > len) return set_error(SD_ERR_INVALID, "sd_rows_submit: bad row size");
    for (int c = 0; c < nc; c++)
      if (!unsafe_field(r + pos + 8, sz, nc, c, p->spec.cols[c].type, &colv[c][i])) return set_error(SD_ERR_INVALID, "sd_rows_submit: malformed UnsafeRow");
    pos += 8 + sz;
  }
  for (int c = 0; c < nc; c++) {
    const int t = p->spec.cols[c].type;
    std::vector<uint64_t> nw(((size_t)nrows + 63) / 64, 0);
    bool any_null = false;
    for (int i = 0; i < nrows; i++) if (colv[c][i].isnull) { nw[i >> 6] |= 1ull << (i & 63); any_null = true; }
    if (any_null && !p->spec.cols[c].nullable) return set_error(SD_ERR_INVALID, "NULL in NOT NULL column %d of the row buffer", c);
    while (!nw.empty() && nw.back() == 0) nw.pop_back();
    std::vector<uint8_t>& b = bufs[c];
    auto put32 = [&](int32_t v) { b.insert(b.end(), reinterpret_cast<uint8_t*>(&v), reinterpret_cast<uint8_t*>(&v) + 4); };
    if (t == SD_STRING) {   // first-seen dictionary
      std::unordered_map<std::string, int> ids;
      std::vector<std::string> vals;
      std::vector<int32_t> idx;
      for (int i = 0; i < nrows; i++) {
        if (colv[c][i].isnull) continue;
        auto it = ids.find(colv[c][i].s);
        if (it == ids.end()) { it = ids.emplace(colv[c][i].s, (int)vals.size()).first; vals.push_back(colv[c][i].s); }
        idx.push_back(it->second);
      }
      const bool big = vals.size() > 32767;
      put32(big ? ENC_BIG_DICTIONARY : ENC_DICTIONARY);
      put32((int32_t)nw.size() * 8);
      b.insert(b.end(), reinterpret_cast<uint8_t*>(nw.data()), reinterpret_cast<uint8_t*>(nw.data()) + nw.size() * 8);
      put32((int32_t)vals.size());
      for (auto& s : vals) { put32((int32_t)s.size()); b.insert(b.end(), s.begin(), s.end()); }
      for (int32_t x : idx) { if (big) put32(x); else { int16_t s16 = (int16_t)x; b.insert(b.end(), reinterpret_cast<uint8_t*>(&s16), reinterpret_cast<uint8_t*>(&s16) + 2); } }
    } else {
      put32(ENC_UNCOMPRESSED);
      put32((int32_t)nw.size() * 8);
      b.insert(b.end(), reinterpret_cast<uint8_t*>(nw.data()), reinterpret_cast<uint8_t*>(nw.data()) + nw.size() * 8);
      for (int i = 0; i < nrows; i++) {
        const HVal& v = colv[c][i];
        if (v.isnull) continue;
        switch (t) {
          case SD_BOOLEAN: case SD_BYTE: b.push_back((uint8_t)v.i); break;
          case SD_SHORT: { int16_t x = (int16_t)v.i; b.insert(b.end(), reinterpret_cast<uint8_t*>(&x), reinterpret_cast<uint8_t*>(&x) + 2); break; }
          case SD_INT: case SD_DATE: put32((int32_t)v.i); break;
          case SD_FLOAT: { float x = (float)v.d; b.insert(b.end(), reinterpret_cast<uint8_t*>(&x), reinterpret_cast<uint8_t*>(&x) + 4); break; }
          case SD_DOUBLE: { double x = v.d; b.insert(b.end(), reinterpret_cast<uint8_t*>(&x), reinterpret_cast<uint8_t*>(&x) + 8); break; }
          default: { int64_t x = v.i; b.insert(b.end(), reinterpret_cast<uint8_t*>(&x), reinterpret_cast<uint8_t*>(&x) + 8); break; }
        }
      }
    }
  }
  std::vector<const void*> ptrs(nc);
  std::vector<int64_t> lens(nc);
  for (int c = 0; c < nc; c++) { ptrs[c] = bufs[c].data(); lens[c] = (int64_t)bufs[c].size(); }
  sd_batch b;
  memset(&b, 0, sizeof(b));
  b.num_rows = nrows; b.ncols = nc; b.col_bufs = ptrs.data(); b.col_lens = lens.data(); b.bucket_id = -1; b.batch_id = -1;
  const int64_t seen_before = p->metrics[2];
  int rc = sd_batch_submit(p, &b);
  p->metrics[2] = seen_before;       // the row buffer is not a column batch
  if (!rc) p->metrics[1] += nrows;   // numRowsBuffer
  return rc;
}

// ---- merge of partial rows (host; payload is a handful of rows) ------------------------------------------
// evaluate = true : SnappyHashAggregateExec(Final) / CollectAggregateExec: merged buffers -> results (avg = sum / count)
// evaluate = false: the merged PARTIAL rows (keys ++ buffers), what a combiner in front of the final stage emits
static int merge_rows_impl(const PlanSpec& sp, const void* partial_rows, int64_t len, bool evaluate, std::vector<uint8_t>& out, int64_t* out_nrows);

static int merge_to_caller(const PlanSpec& sp, const void* partial_rows, int64_t len, bool evaluate, void* out_rows, int64_t cap,
                           int64_t* out_len, int64_t* out_nrows) {
  if (!out_len) return set_error(SD_ERR_INVALID, "merge: null out_len");
  std::vector<uint8_t> out;
  int rc = merge_rows_impl(sp, partial_rows, len, evaluate, out, out_nrows);
  if (rc) return rc;
  *out_len = (int64_t)out.size();
  if ((int64_t)out.size() > cap) return set_error(SD_ERR_OVERFLOW, "merge: output needs %zu bytes", out.size());
  if (!out.empty()) memcpy(out_rows, out.data(), out.size());
  return 0;
}

int sd_final_merge(const sd_plan_desc* desc, const void* partial_rows, int64_t len, void* out_rows, int64_t cap,
                   int64_t* out_len, int64_t* out_nrows) {
  PlanSpec sp;
  std::string err;
  int rc = analyze_plan(desc, sp, err);
  if (rc) return set_error(rc, "sd_final_merge: %s", err.c_str());
  return merge_to_caller(sp, partial_rows, len, true, out_rows, cap, out_len, out_nrows);
}

/* same merge, reusing the analysis of an existing plan handle (no per-call plan analysis) */
int sd_plan_final_merge(sd_plan* p, const void* partial_rows, int64_t len, void* out_rows, int64_t cap,
                        int64_t* out_len, int64_t* out_nrows) {
  if (!p) return set_error(SD_ERR_INVALID, "null plan");
  return merge_to_caller(p->spec, partial_rows, len, true, out_rows, cap, out_len, out_nrows);
}

/* partial rows of several partitions -> one merged set of PARTIAL rows (same schema) */
int sd_partial_merge(const sd_plan_desc* desc, const void* partial_rows, int64_t len, void* out_rows, int64_t cap,
                     int64_t* out_len, int64_t* out_nrows) {
  PlanSpec sp;
  std::string err;
  int rc = analyze_plan(desc, sp, err);
  if (rc) return set_error(rc, "sd_partial_merge: %s", err.c_str());
  return merge_to_caller(sp, partial_rows, len, false, out_rows, cap, out_len, out_nrows);
}
int sd_plan_partial_merge(sd_plan* p, const void* partial_rows, int64_t len, void* out_rows, int64_t cap,
                          int64_t* out_len, int64_t* out_nrows) {
  if (!p) return set_error(SD_ERR_INVALID, "null plan");
  return merge_to_caller(p->spec, partial_rows, len, false, out_rows, cap, out_len, out_nrows);
}

static int merge_rows_impl(const PlanSpec& sp, const void* partial_rows, int64_t len, bool evaluate, std::vector<uint8_t>& out, int64_t* out_nrows) {
  out.clear();
  if (sp.mode == MODE_PROJECT) {   // projected rows of several partitions: concatenation
    const uint8_t* r = reinterpret_cast<const uint8_t*>(partial_rows);
    int64_t pos = 0, n = 0;
    while (pos + 8 <= len) {
      const int64_t sz = rd_i64(r + pos);
      if (sz < 0 || pos + 8 + sz > len) return set_error(SD_ERR_INVALID, "merge: bad row size");
      pos += 8 + sz; n++;
    }
    out.assign(r, r + pos);
    if (out_nrows) *out_nrows = n;
    return 0;
  }
  const int nk = (int)sp.keys.size();
  const std::vector<int> types = partial_field_types(sp);
  const int n = (int)types.size();
  struct Group { std::vector<HVal> keys; std::vector<HVal> bufs; };
  std::vector<Group> groups;   // insertion order
  std::map<std::string, size_t> index;
  const uint8_t* r = reinterpret_cast<const uint8_t*>(partial_rows);
  int64_t pos = 0;
  std::vector<HVal> f((size_t)n);
  std::string key;
  while (pos + 8 <= len) {
    const int64_t sz = rd_i64(r + pos);
    if (sz < 0 || pos + 8 + sz > len) return set_error(SD_ERR_INVALID, "sd_final_merge: bad row size");
    for (int i = 0; i < n; i++) if (!unsafe_field(r + pos + 8, sz, n, i, types[i], &f[i])) return set_error(SD_ERR_INVALID, "sd_final_merge: malformed partial row");
    key.clear();
    for (int k = 0; k < nk; k++) {
      key.push_back(f[k].isnull ? 'N' : 'V');
      if (!f[k].isnull) {
        if (types[k] == SD_STRING) { uint32_t l = (uint32_t)f[k].s.size(); key.append(reinterpret_cast<char*>(&l), 4); key.append(f[k].s); }
        else if (type_is_fp(types[k])) { double d = f[k].d; if (d == 0.0) d = 0.0; if (std::isnan(d)) d = NAN; key.append(reinterpret_cast<char*>(&d), 8); }
        else key.append(reinterpret_cast<const char*>(&f[k].i), 8);
      }
    }
    auto it = nk > 0 ? index.find(key) : (groups.empty() ? index.end() : index.begin());
    Group* g;
    if (it == index.end()) {
      index.emplace(key, groups.size());
      groups.push_back(Group());
      g = &groups.back();
      g->keys.assign(f.begin(), f.begin() + nk);
      g->bufs.assign(f.begin() + nk, f.end());
    } else {
      g = &groups[it->second];
      int k = 0;
      for (auto& m : sp.agg_map) {   // mergeExpressions of each function
        HVal& b = g->bufs[k];
        const HVal& in = f[nk + k];
        switch (m.fn) {
          case SD_AGG_COUNT_STAR: case SD_AGG_COUNT: b.i += in.i; k++; break;
          case SD_AGG_SUM:
            if (!in.isnull) {
              if (m.buf_type == SD_DOUBLE) b.d = (b.isnull ? 0.0 : b.d) + in.d;
              else if (m.buf_type == SD_DECIMAL) { b.w = (b.isnull ? (i128)0 : b.w) + in.w; b.i = (int64_t)b.w; }
              else b.i = (int64_t)((uint64_t)(b.isnull ? 0 : b.i) + (uint64_t)in.i);
              b.isnull = false;
            }
            k++; break;
          case SD_AGG_AVG:
            if (m.buf_type == SD_DECIMAL) { b.w += in.w; b.i = (int64_t)b.w; } else b.d += in.d;
            g->bufs[k + 1].i += f[nk + k + 1].i; k += 2; break;
          default:
            if (!in.isnull) {
              if (b.isnull) b = in;
              else { const int c = cmp_hval(in, b, m.buf_type); if ((m.fn == SD_AGG_MIN && c < 0) || (m.fn == SD_AGG_MAX && c > 0)) b = in; }
            }
            k++; break;
        }
      }
    }
    pos += 8 + sz;
  }
  if (nk == 0 && groups.empty()) {   // no-key aggregate over no partitions: one row of empty buffers
    Group g;
    for (auto& m : sp.agg_map) {
      HVal v;
      if (m.fn == SD_AGG_SUM || m.fn == SD_AGG_MIN || m.fn == SD_AGG_MAX) v.isnull = true;
      g.bufs.push_back(v);
      if (m.fn == SD_AGG_AVG) g.bufs.push_back(HVal());
    }
    groups.push_back(g);
  }
  const std::vector<int> otypes = evaluate ? final_field_types(sp) : types;
  std::vector<HVal> vals;
  for (auto& g : groups) {
    vals.assign(g.keys.begin(), g.keys.end());
    if (!evaluate) vals.insert(vals.end(), g.bufs.begin(), g.bufs.end());
    else {
      int k = 0;
      for (auto& m : sp.agg_map) {
        if (m.fn == SD_AGG_AVG) {   // Average.evaluateExpression: sum / count, NULL when count == 0
          HVal v;
          const int64_t cnt = g.bufs[k + 1].i;
          if (cnt == 0) v.isnull = true;
          else if (m.buf_type == SD_DECIMAL) {
            // Cast(Cast(sum, DECIMAL(p+14,s+4)) / Cast(count, ...), DECIMAL(p+4,s+4)): the quotient at scale s+4, HALF_UP
            const int ft = otypes[vals.size()];
            const i128 num = g.bufs[k].w * pow10_128(ft_scale(ft) - (m.in_ps & 0xff)), den = cnt;
            i128 q = num / den, rem = num % den;
            if (rem < 0) rem = -rem;
            if (2 * rem >= den) q += (num < 0 ? -1 : 1);
            if (q >= pow10_128(ft_precision(ft)) || q <= -pow10_128(ft_precision(ft))) v.isnull = true;   // does not fit: NULL
            else { v.w = q; v.i = (int64_t)q; }
          } else v.d = g.bufs[k].d / (double)cnt;
          vals.push_back(v);
          k += 2;
        } else {
          HVal v = g.bufs[k];
          if (m.fn == SD_AGG_SUM && m.buf_type == SD_DECIMAL && !v.isnull) {   // a sum that needs more than p+10 digits is NULL (changePrecision fails)
            const i128 lim = pow10_128(m.buf_ps >> 8);
            if (v.w >= lim || v.w <= -lim) v.isnull = true;
          }
          vals.push_back(v);
          k++;
        }
      }
    }
    emit_unsafe_row(out, otypes, vals);
  }
  if (out_nrows) *out_nrows = (int64_t)groups.size();
  return 0;
}

// ---- the one exchange of the path: every partition's partial rows -> all ranks, merged ----------
// Partial rows travel BY VALUE (keys as bytes, like the reference's shuffle of UnsafeRows between the partial and the
// final HashAggregate, SnappyStrategies.scala:566-604): dictionary ids are private to a partition.  One ncclAllGather
// of a fixed-capacity blob per rank [magic:4][flags:4][len:8][rows]; a rank whose rows do not fit says so in its header,
// every rank sees every header, so all of them grow the capacity and repeat in lock step (the capacity sticks).
struct sd_comm {
  void* nccl = nullptr;
  int rank = 0, world = 1, device = 0;
  size_t cap = 2048;               // bytes per rank, header included
  uint8_t *d_send = nullptr, *d_recv = nullptr, *h_send = nullptr, *h_recv = nullptr;
  size_t alloc_cap = 0;
  int64_t exchanges = 0, regrows = 0;
};
static int comm_buffers(sd_comm* c) {
  if (c->alloc_cap >= c->cap) return 0;
  if (c->d_send) { cudaFree(c->d_send); cudaFree(c->d_recv); cudaFreeHost(c->h_send); cudaFreeHost(c->h_recv); c->d_send = c->d_recv = c->h_send = c->h_recv = nullptr; }
  SD_CUDA(cudaMalloc(&c->d_send, c->cap));
  SD_CUDA(cudaMalloc(&c->d_recv, c->cap * (size_t)c->world));
  SD_CUDA(cudaMallocHost(&c->h_send, c->cap));
  SD_CUDA(cudaMallocHost(&c->h_recv, c->cap * (size_t)c->world));
  c->alloc_cap = c->cap;
  return 0;
}
constexpr uint32_t COMM_MAGIC = 0x53445831u;   // "SDX1"

int sd_comm_unique_id(void* out_id) {
  if (!out_id) return set_error(SD_ERR_INVALID, "sd_comm_unique_id: null argument");
  return comm_unique_id(out_id);
}
int sd_comm_create(const void* id, int32_t rank, int32_t world, int32_t device, sd_comm** out) {
  if (!id || !out || world < 1 || rank < 0 || rank >= world) return set_error(SD_ERR_INVALID, "sd_comm_create: bad arguments");
  SD_CUDA(cudaSetDevice(device));
  std::unique_ptr<sd_comm> c(new sd_comm());
  c->rank = rank; c->world = world; c->device = device;
  int rc = comm_init(id, rank, world, &c->nccl);
  if (rc) return rc;
  *out = c.release();
  return 0;
}
void sd_comm_destroy(sd_comm* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  comm_destroy(c->nccl);
  if (c->d_send) { cudaFree(c->d_send); cudaFree(c->d_recv); cudaFreeHost(c->h_send); cudaFreeHost(c->h_recv); }
  delete c;
}
int sd_comm_info(sd_comm* c, int64_t out[4]) {
  if (!c || !out) return set_error(SD_ERR_INVALID, "null argument");
  out[0] = c->world; out[1] = (int64_t)c->cap; out[2] = c->exchanges; out[3] = c->regrows;
  return 0;
}

// The exchange's blob of one rank: [magic u32][flags u32][len u64][payload].  flags bit 0: the payload did not fit the slot (only
// the header travelled; every rank doubles the slot and repeats).  flags bit 1: DENSE form -- the payload is this rank's key
// dictionaries followed by the raw [counters][ngroups x slots] state exactly as the kernel left it in HBM:
//   [ngroups i32][nk i32][radix i32 x nk][null id i32 x nk] { [n u32] { [len u32][bytes] } x n } x nk  pad to 8  [state u64 x (8 + ngroups*ns)]
// The sender queues header H2D + a device-to-device copy of the state + the all-gather behind its scan kernel and meets the
// result with ONE synchronisation: no read-back of its own state, no row building, no second upload before the collective
// (the by-value form below needs all three).  Receivers turn each rank's state into partial rows with that rank's dictionaries.
// averaged host-side laps of the per-query path (SD_DEBUG_TIMING=1; printed every 64 calls)
struct LapStats {
  static constexpr int N = 8;
  double sum[N] = {0}; int64_t calls = 0; const char* names[N] = {nullptr};
  std::chrono::steady_clock::time_point t0;
  bool on = getenv("SD_DEBUG_TIMING") != nullptr;
  void begin() { if (on) t0 = std::chrono::steady_clock::now(); }
  void lap(int i, const char* name) {
    if (!on) return;
    const auto t = std::chrono::steady_clock::now();
    sum[i] += std::chrono::duration<double, std::micro>(t - t0).count(); names[i] = name; t0 = t;
  }
  void end(const char* what, int rank) {
    if (!on || (++calls % 64) != 0) return;
    fprintf(stderr, "[%s rank %d] avg us over %lld calls:", what, rank, (long long)calls);
    for (int i = 0; i < N; i++) if (names[i]) fprintf(stderr, "  %s %.1f", names[i], sum[i] / (double)calls);
    fprintf(stderr, "\n");
  }
};
static bool dense_exchange_eligible(const sd_plan* p) {
  if (getenv("SD_TUNE_EXCHANGE_ROWS")) return false;
  const PlanSpec& sp = p->spec;
  if (sp.mode != MODE_NOKEY && sp.mode != MODE_GROUPS) return false;
  if (p->finished_nrows >= 0) return false;   // rows already materialised by an earlier call
  for (const auto& sl : sp.slots) if (sl.op == SLOT_MIN_STR || sl.op == SLOT_MAX_STR) return false;   // values live in this rank's HBM
  return true;
}

int sd_plan_exchange(sd_plan* p, sd_comm* c) {
  if (!p || !c) return set_error(SD_ERR_INVALID, "sd_plan_exchange: null argument");
  if (c->device != p->device) return set_error(SD_ERR_INVALID, "communicator lives on device %d, plan on %d", c->device, p->device);
  static thread_local LapStats laps;
  laps.begin();
  int rc = launch_what_is_pending(p);
  if (rc) return rc;
  const PlanSpec& sp = p->spec;
  const int ns = (int)sp.slots.size(), nk = (int)sp.keys.size();
  const bool dense = dense_exchange_eligible(p);
  std::vector<uint8_t> dense_hdr;
  size_t state_bytes = 0;
  if (dense) {
    if (!p->result_init) { rc = init_result(p, 1); if (rc) return rc; p->ngroups = 1; }
    const size_t ne = (size_t)p->ngroups * ns;
    rc = ensure_result(p, ne);
    if (rc) return rc;
    state_bytes = (STATE_HDR + ne) * 8;
    auto put32 = [&](uint32_t v) { const uint8_t* q = reinterpret_cast<const uint8_t*>(&v); dense_hdr.insert(dense_hdr.end(), q, q + 4); };
    put32((uint32_t)p->ngroups); put32((uint32_t)nk);
    for (int k = 0; k < nk; k++) put32((uint32_t)p->radix[k]);
    for (int k = 0; k < nk; k++) put32((uint32_t)p->key_null_id[k]);
    for (int k = 0; k < nk; k++) {
      put32((uint32_t)p->key_vals[k].size());
      for (const std::string& v : p->key_vals[k]) { put32((uint32_t)v.size()); dense_hdr.insert(dense_hdr.end(), v.begin(), v.end()); }
    }
    dense_hdr.resize((dense_hdr.size() + 7) & ~size_t(7), 0);
  } else {
    rc = collect_partial_rows(p);
    if (rc) return rc;
    rc = rows_to_host(p);
    if (rc) return rc;
  }
  const std::vector<uint8_t>& mine = p->finished_rows;
  for (;;) {
    rc = comm_buffers(c);
    if (rc) return rc;
    const size_t room = c->cap - 16, n = dense ? dense_hdr.size() + state_bytes : mine.size();
    const bool fits = n <= room;
    const uint32_t hdr32[2] = {COMM_MAGIC, (fits ? 0u : 1u) | (dense ? 2u : 0u)};
    const uint64_t len64 = n;
    memcpy(c->h_send, hdr32, 8);
    memcpy(c->h_send + 8, &len64, 8);
    if (dense) {
      if (fits) {
        memcpy(c->h_send + 16, dense_hdr.data(), dense_hdr.size());
        SD_CUDA(cudaMemcpyAsync(c->d_send, c->h_send, 16 + dense_hdr.size(), cudaMemcpyHostToDevice, p->stream));
        SD_CUDA(cudaMemcpyAsync(c->d_send + 16 + dense_hdr.size(), p->d_state, state_bytes, cudaMemcpyDeviceToDevice, p->stream));
      } else {
        SD_CUDA(cudaMemcpyAsync(c->d_send, c->h_send, 16, cudaMemcpyHostToDevice, p->stream));
      }
    } else {
      const size_t sent = std::min(n, room);
      if (sent) memcpy(c->h_send + 16, mine.data(), sent);
      SD_CUDA(cudaMemcpyAsync(c->d_send, c->h_send, 16 + sent, cudaMemcpyHostToDevice, p->stream));
    }
    laps.lap(0, "prepare+copies");
    rc = comm_all_gather_bytes(c->nccl, c->d_send, c->d_recv, c->cap, p->stream);
    if (rc) return rc;
    SD_CUDA(cudaMemcpyAsync(c->h_recv, c->d_recv, c->cap * (size_t)c->world, cudaMemcpyDeviceToHost, p->stream));
    laps.lap(1, "enqueue-allgather");
    SD_CUDA(cudaStreamSynchronize(p->stream));
    laps.lap(2, "wait(kernel+allgather+d2h)");
    c->exchanges++;
    size_t maxlen = 0;
    for (int r = 0; r < c->world; r++) {
      const uint8_t* h = c->h_recv + (size_t)r * c->cap;
      uint32_t magic; uint64_t l;
      memcpy(&magic, h, 4); memcpy(&l, h + 8, 8);
      if (magic != COMM_MAGIC) return set_error(SD_ERR_CUDA, "sd_plan_exchange: bad header from rank %d", r);
      maxlen = std::max<size_t>(maxlen, (size_t)l);
    }
    if (maxlen <= room) break;
    size_t ncap = c->cap;
    while (ncap - 16 < maxlen) ncap *= 2;
    c->cap = ncap;   // every rank computes the same value from the same headers
    c->regrows++;
  }
  std::vector<uint8_t> all;
  for (int r = 0; r < c->world; r++) {
    const uint8_t* h = c->h_recv + (size_t)r * c->cap;
    uint32_t flags; uint64_t l;
    memcpy(&flags, h + 4, 4); memcpy(&l, h + 8, 8);
    const uint8_t* q = h + 16;
    const uint8_t* const end = q + l;
    if (!(flags & 2u)) { all.insert(all.end(), q, end); continue; }
    // dense form: that rank's dictionaries, then its state
    auto get32 = [&](uint32_t* v) { if (q + 4 > end) return false; memcpy(v, q, 4); q += 4; return true; };
    auto bad = [&]() { return set_error(SD_ERR_CUDA, "sd_plan_exchange: malformed dense blob from rank %d", r); };
    uint32_t ng = 0, rnk = 0;
    if (!get32(&ng) || !get32(&rnk) || (int)rnk != nk || ng == 0) return bad();
    int32_t radix[MAX_KEYS] = {1, 1, 1, 1};
    std::vector<int> null_id((size_t)nk, -1);
    std::vector<std::vector<std::string>> vals((size_t)nk);
    uint64_t prod = 1;
    for (int k = 0; k < nk; k++) { uint32_t v; if (!get32(&v) || v == 0) return bad(); radix[k] = (int32_t)v; prod *= v; }
    if (prod != ng) return bad();
    for (int k = 0; k < nk; k++) { uint32_t v; if (!get32(&v)) return bad(); null_id[(size_t)k] = (int32_t)v; }
    for (int k = 0; k < nk; k++) {
      uint32_t nv; if (!get32(&nv)) return bad();
      for (uint32_t i = 0; i < nv; i++) {
        uint32_t sl; if (!get32(&sl) || q + sl > end) return bad();
        vals[(size_t)k].emplace_back(reinterpret_cast<const char*>(q), sl); q += sl;
      }
      // (a radix may exceed the dictionary by the NULL id; an id beyond both never has rows)
      if ((size_t)radix[k] > vals[(size_t)k].size()) vals[(size_t)k].resize((size_t)radix[k]);
    }
    q = h + 16 + (((size_t)(q - (h + 16)) + 7) & ~size_t(7));
    if (q + (STATE_HDR + (size_t)ng * ns) * 8 != end) return bad();
    std::vector<uint64_t> st((STATE_HDR + (size_t)ng * ns));
    memcpy(st.data(), q, st.size() * 8);
    const int64_t nr = dense_rows_from_state(sp, st.data() + STATE_HDR, (int)ng, radix, null_id, vals, nullptr, all);
    if (r == c->rank) {   // this rank's own execution metrics come back with its blob
      update_agg_time(p);
      p->metrics[6] = (int64_t)(p->agg_ms * 1e6);
      p->metrics[8] = (int64_t)st[0];
      p->metrics[11] = (int64_t)st[0];
      p->metrics[0] = nr;
    }
  }
  laps.lap(3, "decode-blobs");
  std::vector<uint8_t> merged;
  int64_t nrows = 0;
  rc = merge_rows_impl(p->spec, all.data(), (int64_t)all.size(), false, merged, &nrows);
  if (rc) return rc;
  p->finished_rows.swap(merged);
  p->finished_nrows = nrows;
  p->dev_rows_len = -1;
  laps.lap(4, "merge");
  laps.end("sd_plan_exchange", c->rank);
  return 0;
}

// one execution of a cached plan over a resident store, in one call: reset -> literals -> scan -> [exchange] ->
// partial rows (merged over all ranks when `comm` is given).  What a re-executed cached plan does per query in the
// reference (SnappySession plan cache: new literal values, same generated code).
int sd_plan_execute_store(sd_plan* p, sd_store* s, const int32_t* bucket_ids, int32_t nbuckets, const sd_literal* lits, int32_t nlits,
                          sd_comm* comm, void* out_rows, int64_t cap, int64_t* out_len, int64_t* out_nrows) {
  static thread_local LapStats laps;
  laps.begin();
  int rc = sd_plan_reset(p);
  if (rc) return rc;
  laps.lap(0, "reset");
  if (lits || nlits) { rc = sd_plan_set_literals(p, lits, nlits); if (rc) return rc; }
  rc = sd_plan_scan_store(p, s, bucket_ids, nbuckets);
  if (rc) return rc;
  laps.lap(1, "literals+scan-enqueue");
  if (comm) { rc = sd_plan_exchange(p, comm); if (rc) return rc; }
  laps.lap(2, "exchange");
  rc = sd_plan_finish(p, out_rows, cap, out_len, out_nrows);
  laps.lap(3, "finish");
  laps.end("sd_plan_execute_store", comm ? comm->rank : 0);
  return rc;
}

}  // extern "C"
