/*
 * scan_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the reference's column-table scan -> filter -> partial hash-aggregate
 * path, row at a time, exactly in the order the reference's generated (WholeStageCodegen) loop
 * runs it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the timed CPU baseline.
 *
 * The reference cannot be compiled here (Scala on a Spark fork whose sources are absent; no JVM),
 * so this file follows the reference text; each function cites the file:line it restates
 * (paths relative to /root/reference;
 *   enc  = encoders/src/main/scala/org/apache/spark/sql/execution/columnar/encoding,
 *   core = core/src/main/scala/org/apache/spark/sql).
 * Semantics that live in the absent Spark fork (snappy-spark 2.1.1.9: FilterExec, Sum/Average/
 * Count/Min/Max, UnsafeRow, NaN-safe compare) are restated from upstream Apache Spark 2.1.1
 * behaviour.
 *
 * PARITY PINNING: tests/test_oracle_golden.py checks this oracle against the reference's own
 * known-answer fixture for the path -- TPC-H lineitem.tbl (30,201 rows) -> Q1 / Q6 ==
 * tests/common/src/main/resources/TPCH/RESULT/Snappy_1.out / Snappy_6.out -- and against the
 * BitSet known answers of cluster/src/test/.../store/BitSetTest.scala.  RunLength (typeId 1) has
 * no encoder, test or fixture in the reference: that decoder is "parity unpinned".
 *
 * Two layers:
 *   1. a generic plan interpreter with the same C ABI shape as include/snappy_gpu.h
 *      (oracle_plan_create / oracle_batch_submit / oracle_plan_finish / oracle_final_merge) -- the
 *      parity checker for arbitrary plans;
 *   2. hand-written restatements of the loops the reference's code generator would emit for the
 *      benchmark queries (oracle_q6_batches, oracle_q1_batches, oracle_c1_batches), compiled -O2:
 *      the "reference-algorithm CPU restatement" timed as cpu_baseline.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/snappy_gpu.h"

#define MAXCOLS 256
#define MAXERR 512

static __thread char g_err[MAXERR];
const char* oracle_last_error(void) { return g_err; }
static int fail(int code, const char* msg) {
  snprintf(g_err, MAXERR, "%s", msg);
  return code;
}

/* ------------------------------------------------------------------------------------------ */
/* little-endian loads (enc/ColumnEncoding.scala:913-959)                                      */
static inline int32_t ld_i32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }
static inline int16_t ld_i16(const uint8_t* p) { int16_t v; memcpy(&v, p, 2); return v; }
static inline int64_t ld_i64(const uint8_t* p) { int64_t v; memcpy(&v, p, 8); return v; }
static inline uint64_t ld_u64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline double ld_f64(const uint8_t* p) { double v; memcpy(&v, p, 8); return v; }
static inline float ld_f32(const uint8_t* p) { float v; memcpy(&v, p, 4); return v; }

/* ------------------------------------------------------------------------------------------ */
/* BitSet primitives (enc/BitSet.scala:54-62 isSet, :98-130 nextSetBit, :132-157 cardinality)   */
int oracle_bitset_is_set(const uint64_t* words, int position, int num_words) {
  int w = position >> 6;
  return w < num_words && ((words[w] >> (position & 63)) & 1);
}
int oracle_bitset_next_set_bit(const uint64_t* words, int from, int num_words) {
  int w = from >> 6;
  if (w >= num_words) return INT32_MAX;
  uint64_t cur = words[w] & (~0ULL << (from & 63));
  for (;;) {
    if (cur) return (w << 6) + __builtin_ctzll(cur);
    if (++w >= num_words) return INT32_MAX;
    cur = words[w];
  }
}
/* number of set bits in [0, position) */
int oracle_bitset_cardinality(const uint64_t* words, int position, int num_words) {
  int full = position >> 6, n = 0;
  if (full > num_words) full = num_words;
  for (int i = 0; i < full; i++) n += __builtin_popcountll(words[i]);
  if (full < num_words && (position & 63)) n += __builtin_popcountll(words[full] & ((1ULL << (position & 63)) - 1));
  return n;
}

/* ------------------------------------------------------------------------------------------ */
/* values                                                                                      */
typedef struct val {
  int isnull;
  int64_t i;          /* integral classes (BOOLEAN..LONG, DATE, TIMESTAMP, DECIMAL)  */
  double d;           /* FLOAT (float-rounded) and DOUBLE                            */
  const uint8_t* s;   /* STRING                                                      */
  int32_t slen;
  __int128 w;         /* DECIMAL aggregate buffers / results wider than 18 digits (java.math.BigDecimal unscaled value
                         in the reference: Decimal.scala falls back to BigDecimal beyond MAX_LONG_DIGITS)            */
} val;
/* field type codes for rows: sd_type in the low byte, DECIMAL (precision << 8 | scale) above it */
#define FT(t, ps) ((t) == SD_DECIMAL ? ((t) | ((ps) << 8)) : (t))
#define FT_BASE(ft) ((ft) & 0xff)
#define FT_PREC(ft) (((ft) >> 16) & 0xff)
static __int128 pow10_w(int k) { __int128 r = 1; while (k-- > 0) r *= 10; return r; }

static int is_integral(int t) {
  return t == SD_BOOLEAN || t == SD_BYTE || t == SD_SHORT || t == SD_INT || t == SD_LONG || t == SD_DATE ||
         t == SD_TIMESTAMP || t == SD_DECIMAL;
}
static int is_fp(int t) { return t == SD_FLOAT || t == SD_DOUBLE; }

/* ------------------------------------------------------------------------------------------ */
/* column decoder state (one per projected column per batch), restating the decoder classes     */
typedef struct coldec {
  int type_id;            /* 0 Uncompressed 1 RunLength 2 Dictionary 3 BigDictionary 4 BooleanBitSet */
  int sql_type;
  int nullable;
  const uint64_t* nullw;  /* null words (may be unaligned in memory -> copied)  */
  uint64_t* nullw_own;
  int nwords;
  const uint8_t* body;    /* first encoded value / first index                  */
  const uint8_t* end;
  /* dictionary (enc/DictionaryEncoding.scala:85-137) */
  int dict_n;
  const uint8_t** dict_s; int32_t* dict_slen;
  const uint8_t* dict_fixed;   /* int32 or int64 entries */
  /* run length (enc/RunLengthEncoding.scala:84-173) */
  int run_end;            /* runLengthEndPosition */
  const uint8_t* run_cur; /* currentCursor        */
  val run_val;
  /* incremental null bookkeeping (ColumnTableScan.scala:794-815) */
  int nulls_before;       /* numNullsVar: nulls in [0, next_ordinal) */
  int next_ordinal;
} coldec;

static void coldec_free(coldec* d) {
  free(d->nullw_own); free(d->dict_s); free(d->dict_slen);
  memset(d, 0, sizeof(*d));
}

/* enc/ColumnEncoding.scala:797-832 getColumnDecoder + :1042-1099 initializeNulls
 * + per-encoding initializeCursor.  delta_skip: bytes of [numBaseRows][numDeltas][positions]+pad
 * that initDelta consumes between the null words and the encoded data of a delta buffer. */
static int coldec_init(coldec* d, const uint8_t* buf, int64_t len, int sql_type, int nullable, int64_t delta_skip) {
  memset(d, 0, sizeof(*d));
  if (!buf || len < 8) return fail(SD_ERR_INVALID, "column buffer shorter than its header");
  d->type_id = ld_i32(buf);
  if (d->type_id < 0) return fail(SD_ERR_UNSUPPORTED, "oracle: compressed buffer (decompress first)");
  if (d->type_id > 4) return fail(SD_ERR_INVALID, "unknown encoding typeId");
  int null_bytes = ld_i32(buf + 4);
  if (null_bytes < 0 || (null_bytes & 7) || 8 + (int64_t)null_bytes > len) return fail(SD_ERR_INVALID, "bad null bitset size");
  if (!nullable && null_bytes != 0) return fail(SD_ERR_INVALID, "Nulls bitset found in NOT NULL column");
  d->sql_type = sql_type; d->nullable = nullable;
  d->nwords = null_bytes >> 3;
  if (d->nwords) {
    d->nullw_own = (uint64_t*)malloc(null_bytes);
    memcpy(d->nullw_own, buf + 8, null_bytes);
    d->nullw = d->nullw_own;
  }
  const uint8_t* p = buf + 8 + null_bytes + delta_skip;
  d->end = buf + len;
  if (d->type_id == 2 || d->type_id == 3) {
    if (p + 4 > d->end) return fail(SD_ERR_INVALID, "truncated dictionary");
    int n = ld_i32(p); p += 4;
    if (n < 0) return fail(SD_ERR_INVALID, "negative dictionary size");
    d->dict_n = n;
    if (sql_type == SD_STRING) {
      d->dict_s = (const uint8_t**)malloc(sizeof(void*) * (n + 1));
      d->dict_slen = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
      for (int k = 0; k < n; k++) {
        if (p + 4 > d->end) return fail(SD_ERR_INVALID, "truncated string dictionary");
        int l = ld_i32(p);
        if (l < 0 || p + 4 + l > d->end) return fail(SD_ERR_INVALID, "bad string dictionary entry");
        d->dict_s[k] = p + 4; d->dict_slen[k] = l; p += 4 + l;
      }
    } else if (sql_type == SD_INT || sql_type == SD_DATE) {
      d->dict_fixed = p; p += 4 * (int64_t)n;
    } else if (sql_type == SD_LONG || sql_type == SD_TIMESTAMP) {
      d->dict_fixed = p; p += 8 * (int64_t)n;
    } else return fail(SD_ERR_UNSUPPORTED, "DictionaryDecoder not supported for this type");
  } else if (d->type_id == 1) {
    d->run_end = -1; d->run_cur = p;
    if (sql_type == SD_BYTE || sql_type == SD_BOOLEAN)
      return fail(SD_ERR_UNSUPPORTED, "RunLength BYTE/BOOLEAN: reference decoder inconsistent (enc/RunLengthEncoding.scala:99-110)");
  } else if (d->type_id == 4) {
    if (sql_type != SD_BOOLEAN) return fail(SD_ERR_INVALID, "BooleanBitSet on a non-boolean column");
  }
  d->body = p;
  return 0;
}

/* is row `ordinal` NULL, and which stored value is it (k = ordinal - nulls before it).  Restates
 * genIfNonNullCode + NullableDecoder.numNulls (ColumnTableScan.scala:794-815,
 * enc/ColumnEncoding.scala:1103-1142): incremental while ordinals advance by one, recount when
 * rows were skipped. */
static inline int coldec_is_null(coldec* d, int ordinal, int* non_null_pos) {
  if (d->nwords == 0) { *non_null_pos = ordinal; return 0; }
  if (ordinal != d->next_ordinal) {
    if (ordinal == d->next_ordinal - 1) {           /* same ordinal read twice (SNAP-2118) */
      int isn = oracle_bitset_is_set(d->nullw, ordinal, d->nwords);
      *non_null_pos = ordinal - (d->nulls_before - (isn ? 1 : 0));
      return isn;
    }
    d->nulls_before = oracle_bitset_cardinality(d->nullw, ordinal, d->nwords);
  }
  int isn = oracle_bitset_is_set(d->nullw, ordinal, d->nwords);
  *non_null_pos = ordinal - d->nulls_before;
  d->nulls_before += isn;
  d->next_ordinal = ordinal + 1;
  return isn;
}

/* dictionary index of the k-th stored value (readDictionaryIndex, enc/DictionaryEncoding.scala:126-127,156-157) */
static inline int coldec_dict_index(const coldec* d, int k) {
  return d->type_id == 2 ? (int)ld_i16(d->body + 2 * (int64_t)k) : ld_i32(d->body + 4 * (int64_t)k);
}

/* read the k-th stored (non-null) value */
static int coldec_read(coldec* d, int k, val* out) {
  out->isnull = 0;
  const int t = d->sql_type;
  switch (d->type_id) {
    case 0: /* enc/Uncompressed.scala:74-98 */
      switch (t) {
        case SD_BOOLEAN: out->i = d->body[k] == 1; return 0;
        case SD_BYTE: out->i = (int8_t)d->body[k]; return 0;
        case SD_SHORT: out->i = ld_i16(d->body + 2 * (int64_t)k); return 0;
        case SD_INT: case SD_DATE: out->i = ld_i32(d->body + 4 * (int64_t)k); return 0;
        case SD_LONG: case SD_TIMESTAMP: case SD_DECIMAL: out->i = ld_i64(d->body + 8 * (int64_t)k); return 0;
        case SD_FLOAT: out->d = ld_f32(d->body + 4 * (int64_t)k); return 0;
        case SD_DOUBLE: out->d = ld_f64(d->body + 8 * (int64_t)k); return 0;
        case SD_STRING: {
          /* sequential [len][bytes] cursor with rewind (enc/Uncompressed.scala:100-146); the oracle
           * keeps (run_cur, run_end) = (cursor, index of the value the cursor points at) */
          if (d->run_cur == NULL || k < d->run_end) { d->run_cur = d->body; d->run_end = 0; }
          while (d->run_end < k) { d->run_cur += 4 + ld_i32(d->run_cur); d->run_end++; }
          out->slen = ld_i32(d->run_cur); out->s = d->run_cur + 4; return 0;
        }
      }
      break;
    case 2: case 3: { /* enc/DictionaryEncoding.scala:118-137,148-166 */
      int idx = coldec_dict_index(d, k);
      if (idx < 0 || idx >= d->dict_n) return fail(SD_ERR_INVALID, "dictionary index out of range");
      if (t == SD_STRING) { out->s = d->dict_s[idx]; out->slen = d->dict_slen[idx]; }
      else if (t == SD_INT || t == SD_DATE) out->i = ld_i32(d->dict_fixed + 4 * (int64_t)idx);
      else out->i = ld_i64(d->dict_fixed + 8 * (int64_t)idx);
      return 0;
    }
    case 4: /* enc/BooleanBitSetEncoding.scala:57-59 */
      out->i = (ld_u64(d->body + 8 * (int64_t)(k >> 6)) >> (k & 63)) & 1; return 0;
    case 1: { /* enc/RunLengthEncoding.scala:112-172 */
      while (d->run_end < k) {
        const uint8_t* c = d->run_cur;
        if (c >= d->end) return fail(SD_ERR_INVALID, "RunLengthEncoding: reading next run after data end");
        switch (t) {
          case SD_SHORT: d->run_val.i = ld_i16(c); d->run_end += ld_i32(c + 2); d->run_cur = c + 6; break;
          case SD_INT: case SD_DATE: d->run_val.i = ld_i32(c); d->run_end += ld_i32(c + 4); d->run_cur = c + 8; break;
          case SD_LONG: case SD_TIMESTAMP: d->run_val.i = ld_i64(c); d->run_end += ld_i32(c + 8); d->run_cur = c + 12; break;
          case SD_STRING: {
            int l = ld_i32(c); d->run_val.s = c + 4; d->run_val.slen = l;
            d->run_end += ld_i32(c + 4 + l); d->run_cur = c + 8 + l; break;
          }
          default: return fail(SD_ERR_UNSUPPORTED, "RunLengthDecoder not supported for this type");
        }
      }
      *out = d->run_val; out->isnull = 0; return 0;
    }
  }
  return fail(SD_ERR_UNSUPPORTED, "decoder/type combination not supported");
}

/* ------------------------------------------------------------------------------------------ */
/* update deltas: ColumnDeltaDecoder (enc/ColumnDeltaDecoder.scala:32-83) and the 2-way merge
 * UpdatedColumnDecoderBase (enc/UpdatedColumnDecoder.scala:70-128)                             */
typedef struct deltadec {
  int present;
  coldec real;              /* realDecoder over the delta's encoded values         */
  const uint8_t* pos_cur;   /* positionCursor                                      */
  const uint8_t* pos_end;   /* positionEndCursor                                   */
  int decoder_position;     /* relative entry being read                           */
  int non_null_position;    /* relative index among non-null entries               */
  int not_null;
} deltadec;

static int deltadec_init(deltadec* d, const uint8_t* buf, int64_t len, int sql_type, int nullable) {
  memset(d, 0, sizeof(*d));
  if (!buf) return 0;
  if (len < 16) return fail(SD_ERR_INVALID, "delta buffer too short");
  int null_bytes = ld_i32(buf + 4);
  const uint8_t* c = buf + 8 + null_bytes;      /* cursor handed to initDelta */
  int npos = ld_i32(c + 4);
  const uint8_t* pend = c + 8 + 4 * (int64_t)npos;
  int64_t data_off = (((pend - buf) + 7) >> 3) << 3;   /* round to nearest word */
  /* a delta's null bits index relative entries; the decoder class is chosen by field.nullable */
  int rc = coldec_init(&d->real, buf, len, sql_type, nullable || null_bytes != 0, data_off - (8 + null_bytes));
  if (rc) return rc;
  d->present = 1;
  d->pos_cur = c + 8; d->pos_end = pend;
  d->decoder_position = -1; d->non_null_position = -1;
  return 0;
}
static inline int deltadec_read_updated_position(const deltadec* d) {
  return (d->present && d->pos_cur < d->pos_end) ? ld_i32(d->pos_cur) : INT32_MAX;
}
static inline void deltadec_move(deltadec* d) {
  d->pos_cur += 4; d->decoder_position++;
  d->not_null = !oracle_bitset_is_set(d->real.nullw, d->decoder_position, d->real.nwords);
  if (d->not_null) d->non_null_position++;
}

typedef struct updec {
  int present;
  deltadec d1, d2;
  int d1_pos, d2_pos;
  deltadec* current;
  int next_updated;
} updec;

static int updec_move_to_next(updec* u) {            /* moveToNextUpdatedPosition :85-113 */
  int next = INT32_MAX, first = 0;
  if (u->d1_pos != INT32_MAX) { next = u->d1_pos; first = 1; }
  if (u->d2.present) {
    if (u->d2_pos <= next) {
      if (u->d2_pos < next) { next = u->d2_pos; u->current = &u->d2; first = 0; }
      deltadec_move(&u->d2);                          /* skip on equality in any case */
      u->d2_pos = deltadec_read_updated_position(&u->d2);
    }
  }
  if (first) { u->current = &u->d1; deltadec_move(&u->d1); u->d1_pos = deltadec_read_updated_position(&u->d1); }
  return next;
}
static void updec_start(updec* u) {
  u->d1_pos = deltadec_read_updated_position(&u->d1);
  u->d2_pos = deltadec_read_updated_position(&u->d2);
  u->current = NULL;
  u->next_updated = updec_move_to_next(u);
}
static inline int updec_unchanged(updec* u, int ordinal) {   /* unchanged :124-128 + skipUntil */
  if (u->next_updated > ordinal) return 1;
  if (u->next_updated == ordinal) return 0;
  for (;;) {
    u->next_updated = updec_move_to_next(u);
    if (u->next_updated > ordinal) return 1;
    if (u->next_updated == ordinal) return 0;
  }
}

/* delete mask (enc/ColumnDeleteDecoder.scala:25-55): sequential cursor over ascending positions */
typedef struct deldec { const uint8_t* cur; const uint8_t* end; int next; } deldec;
static void deldec_init(deldec* d, const uint8_t* buf, int64_t len) {
  if (!buf) { d->cur = d->end = NULL; d->next = INT32_MAX; return; }
  d->cur = buf + 12; d->end = buf + len;
  if (d->cur < d->end) { d->next = ld_i32(d->cur); d->cur += 4; } else d->next = INT32_MAX;
}
static inline int deldec_deleted(deldec* d, int ordinal) {
  if (d->next != ordinal) return 0;
  if (d->cur < d->end) { d->next = ld_i32(d->cur); d->cur += 4; } else d->next = INT32_MAX;
  return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* expression evaluation (Spark 2.1.1 semantics)                     */
static int cmp_bytes(const uint8_t* a, int la, const uint8_t* b, int lb) {
  int n = la < lb ? la : lb;
  int c = n ? memcmp(a, b, n) : 0;
  return c ? c : (la - lb);
}
/* Utils.nanSafeCompareDoubles: NaN == NaN, NaN greater than anything, -0.0 == 0.0 */
static int cmp_f64(double x, double y) {
  int xn = isnan(x), yn = isnan(y);
  if (xn || yn) return xn && yn ? 0 : (xn ? 1 : -1);
  return x < y ? -1 : (x > y ? 1 : 0);
}
static int cmp_val(const val* a, const val* b, int t) {
  if (t == SD_STRING) return cmp_bytes(a->s, a->slen, b->s, b->slen);
  if (is_fp(t)) return cmp_f64(a->d, b->d);
  return a->i < b->i ? -1 : (a->i > b->i ? 1 : 0);
}
static int64_t wrap_int(int64_t v, int t) {
  switch (t) {
    case SD_BYTE: return (int8_t)v;
    case SD_SHORT: return (int16_t)v;
    case SD_INT: case SD_DATE: return (int32_t)v;
    default: return v;
  }
}
/* Java (long)/(int) casts of a double: NaN -> 0, saturating */
static int64_t f64_to_i64(double d) {
  if (isnan(d)) return 0;
  if (d >= 9223372036854775807.0) return INT64_MAX;
  if (d <= -9223372036854775808.0) return INT64_MIN;
  return (int64_t)d;
}
static int64_t f64_to_i32(double d) {
  if (isnan(d)) return 0;
  if (d >= 2147483647.0) return INT32_MAX;
  if (d <= -2147483648.0) return INT32_MIN;
  return (int32_t)d;
}

typedef struct oracle_plan oracle_plan;
struct rowctx { const val* cols; const oracle_plan* plan; };

struct oracle_plan {
  sd_plan_desc desc;
  sd_column* cols; sd_expr* exprs; int32_t* keys; sd_agg* aggs; int32_t* proj; int32_t* lit_types;
  sd_literal* lits; char** lit_strs;
  /* aggregation state */
  int nbuf;                 /* number of buffer fields per group                   */
  int* buf_type;            /* sd_type per buffer field                            */
  int* buf_ps;              /* DECIMAL buffer fields: (precision << 8) | scale     */
  int* buf_nullable;
  int* expr_nullable;       /* static nullability per expression node              */
  /* groups in insertion order (SHAMap appends value bytes back to back: SHAMap.scala:21-41) */
  int ngroups, cap_groups;
  val** gkeys;              /* [group][nkeys] (string keys own their bytes)        */
  val** gbufs;              /* [group][nbuf]                                       */
  uint64_t* ghash;
  int* htab; int hcap;      /* open addressing, quadratic probe, load 0.75 (ByteBufferHashMap.scala:136-187) */
  /* projection output */
  uint8_t* out; int64_t out_len, out_cap; int64_t out_rows;
  int64_t metrics[SD_NUM_METRICS];
};

static val eval(const oracle_plan* p, int node, const val* cols);
/* (precision << 8) | scale of a DECIMAL-typed node: columns carry it, LIT / CAST nodes in sd_expr.c */
static int dec_ps(const oracle_plan* p, int node) {
  const sd_expr* e = &p->exprs[node];
  if (e->op == SD_OP_COL) return (p->cols[e->a].precision << 8) | p->cols[e->a].scale;
  if (e->op == SD_OP_NEG) return dec_ps(p, e->a);
  return e->c;
}

static val eval_node(const oracle_plan* p, const sd_expr* e, const val* cols) {
  val r; memset(&r, 0, sizeof(r));
  switch (e->op) {
    case SD_OP_COL: { val v = cols[e->a]; if (e->type == SD_DECIMAL) v.w = v.i; return v; }   /* decoders fill .i; DECIMAL consumers read .w */
    case SD_OP_LIT: {
      const sd_literal* l = &p->lits[e->a];
      r.isnull = l->is_null; r.i = l->i; r.d = l->d; r.s = (const uint8_t*)l->s; r.slen = l->slen;
      if (e->type == SD_FLOAT) r.d = (float)l->d;
      return r;
    }
    case SD_OP_ADD: case SD_OP_SUB: case SD_OP_MUL: case SD_OP_DIV: {
      val a = eval(p, e->a, cols), b = eval(p, e->b, cols);
      if (a.isnull || b.isnull) { r.isnull = 1; return r; }
      if (e->type == SD_DOUBLE || e->type == SD_FLOAT) {
        double x = a.d, y = b.d, z;
        if (e->op == SD_OP_DIV) { if (y == 0.0) { r.isnull = 1; return r; } z = x / y; }   /* Divide: x / 0 => NULL */
        else z = e->op == SD_OP_ADD ? x + y : (e->op == SD_OP_SUB ? x - y : x * y);
        r.d = e->type == SD_FLOAT ? (double)(float)z : z;
        if (e->type == SD_FLOAT) {   /* float ops are performed in float */
          float xf = (float)x, yf = (float)y, zf;
          zf = e->op == SD_OP_ADD ? xf + yf : (e->op == SD_OP_SUB ? xf - yf : (e->op == SD_OP_MUL ? xf * yf : xf / yf));
          r.d = zf;
        }
        return r;
      }
      if (e->op == SD_OP_DIV) { r.isnull = 1; return r; }   /* integral Divide is cast to double upstream */
      uint64_t x = (uint64_t)a.i, y = (uint64_t)b.i;
      uint64_t z = e->op == SD_OP_ADD ? x + y : (e->op == SD_OP_SUB ? x - y : x * y);
      r.i = wrap_int((int64_t)z, e->type);
      return r;
    }
    case SD_OP_NEG: {
      val a = eval(p, e->a, cols);
      if (a.isnull) return a;
      if (is_fp(e->type)) r.d = -a.d; else r.i = wrap_int((int64_t)(0 - (uint64_t)a.i), e->type);
      return r;
    }
    case SD_OP_CAST: {
      val a = eval(p, e->a, cols);
      if (a.isnull) { r.isnull = 1; return r; }
      int from = p->exprs[e->a].type, to = e->type;
      /* Spark 2.1.1 Cast (the pairs include/snappy_gpu.h lists; the rest is refused at plan creation) */
      if (from == SD_DECIMAL && is_fp(to)) {                       /* Decimal.toDouble */
        double x = (double)a.i / pow(10.0, dec_ps(p, e->a) & 0xff);
        r.d = to == SD_FLOAT ? (double)(float)x : x;
      } else if (to == SD_DECIMAL && from != SD_DECIMAL) {         /* Decimal(long).changePrecision(p, s) or NULL */
        int ps = dec_ps(p, (int)(e - p->exprs));
        __int128 v = (__int128)a.i * pow10_w(ps & 0xff), lim = pow10_w(ps >> 8);
        if (v >= lim || v <= -lim) r.isnull = 1; else r.i = (int64_t)v;
      } else if (to == SD_DECIMAL) {                                /* DECIMAL -> DECIMAL with a scale that does not shrink */
        int ps0 = dec_ps(p, e->a), ps1 = dec_ps(p, (int)(e - p->exprs));
        __int128 v = (__int128)a.i * pow10_w((ps1 & 0xff) - (ps0 & 0xff)), lim = pow10_w(ps1 >> 8);
        if (v >= lim || v <= -lim) r.isnull = 1; else r.i = (int64_t)v;
      } else if (to == SD_BOOLEAN) r.i = is_fp(from) ? (a.d != 0.0) : (a.i != 0);   /* castToBoolean: _ != 0 */
      else if (is_integral(from) && is_integral(to)) r.i = wrap_int(a.i, to);
      else if (is_integral(from) && to == SD_DOUBLE) r.d = (double)a.i;
      else if (is_integral(from) && to == SD_FLOAT) r.d = (float)a.i;
      else if (is_fp(from) && to == SD_DOUBLE) r.d = a.d;
      else if (is_fp(from) && to == SD_FLOAT) r.d = (float)a.d;
      else if (is_fp(from) && to == SD_LONG) r.i = f64_to_i64(a.d);
      else if (is_fp(from) && is_integral(to)) r.i = wrap_int(f64_to_i32(a.d), to);
      else r = a;
      return r;
    }
    case SD_OP_EQ: case SD_OP_NE: case SD_OP_LT: case SD_OP_LE: case SD_OP_GT: case SD_OP_GE: {
      val a = eval(p, e->a, cols), b = eval(p, e->b, cols);
      if (a.isnull || b.isnull) { r.isnull = 1; return r; }
      int c = cmp_val(&a, &b, p->exprs[e->a].type);
      switch (e->op) {
        case SD_OP_EQ: r.i = c == 0; break; case SD_OP_NE: r.i = c != 0; break;
        case SD_OP_LT: r.i = c < 0; break;  case SD_OP_LE: r.i = c <= 0; break;
        case SD_OP_GT: r.i = c > 0; break;  default: r.i = c >= 0; break;
      }
      return r;
    }
    case SD_OP_AND: {   /* Kleene AND */
      val a = eval(p, e->a, cols);
      if (!a.isnull && !a.i) { r.i = 0; return r; }
      val b = eval(p, e->b, cols);
      if (!b.isnull && !b.i) { r.i = 0; return r; }
      if (a.isnull || b.isnull) { r.isnull = 1; return r; }
      r.i = 1; return r;
    }
    case SD_OP_OR: {
      val a = eval(p, e->a, cols);
      if (!a.isnull && a.i) { r.i = 1; return r; }
      val b = eval(p, e->b, cols);
      if (!b.isnull && b.i) { r.i = 1; return r; }
      if (a.isnull || b.isnull) { r.isnull = 1; return r; }
      r.i = 0; return r;
    }
    case SD_OP_NOT: { val a = eval(p, e->a, cols); if (a.isnull) return a; r.i = !a.i; return r; }
    case SD_OP_ISNULL: { val a = eval(p, e->a, cols); r.i = a.isnull; return r; }
    case SD_OP_ISNOTNULL: { val a = eval(p, e->a, cols); r.i = !a.isnull; return r; }
    case SD_OP_IN: {
      val a = eval(p, e->a, cols);
      if (a.isnull) { r.isnull = 1; return r; }
      int has_null = 0, t = p->exprs[e->a].type;
      for (int k = 0; k < e->c; k++) {
        const sd_literal* l = &p->lits[e->b + k];
        if (l->is_null) { has_null = 1; continue; }
        val b; memset(&b, 0, sizeof(b)); b.i = l->i; b.d = t == SD_FLOAT ? (double)(float)l->d : l->d;
        b.s = (const uint8_t*)l->s; b.slen = l->slen;
        if (cmp_val(&a, &b, t) == 0) { r.i = 1; return r; }
      }
      if (has_null) r.isnull = 1; else r.i = 0;
      return r;
    }
    case SD_OP_STARTSWITH: {
      val a = eval(p, e->a, cols), b = eval(p, e->b, cols);
      if (a.isnull || b.isnull) { r.isnull = 1; return r; }
      r.i = a.slen >= b.slen && (b.slen == 0 || memcmp(a.s, b.s, b.slen) == 0);
      return r;
    }
  }
  r.isnull = 1;
  return r;
}
static val eval(const oracle_plan* p, int node, const val* cols) { return eval_node(p, &p->exprs[node], cols); }

/* static nullability (Catalyst Expression.nullable) */
static void compute_nullability(oracle_plan* p) {
  for (int i = 0; i < p->desc.nexprs; i++) {
    const sd_expr* e = &p->exprs[i];
    int n = 0;
    switch (e->op) {
      case SD_OP_COL: n = p->cols[e->a].nullable; break;
      case SD_OP_LIT: n = 0; break;
      case SD_OP_DIV: n = 1; break;
      case SD_OP_ISNULL: case SD_OP_ISNOTNULL: n = 0; break;
      case SD_OP_CAST: n = p->expr_nullable[e->a] || e->type == SD_DECIMAL; break;   /* Cast.forceNullable(_, DecimalType) */
      case SD_OP_NEG: case SD_OP_NOT: n = p->expr_nullable[e->a]; break;
      case SD_OP_IN: n = p->expr_nullable[e->a]; break;
      default: n = p->expr_nullable[e->a] || p->expr_nullable[e->b]; break;
    }
    p->expr_nullable[i] = n;
  }
}

/* buffer schema (SnappyHashAggregateExec.scala:174-210) */
static int sum_buffer_type(int t) { return is_fp(t) ? SD_DOUBLE : (t == SD_DECIMAL ? SD_DECIMAL : SD_LONG); }
static int imin(int a, int b) { return a < b ? a : b; }

/* casts the path executes (mirrors the product's plan validation) */
static int check_casts(const oracle_plan* p) {
  for (int i = 0; i < p->desc.nexprs; i++) {
    const sd_expr* e = &p->exprs[i];
    if (e->op != SD_OP_CAST) continue;
    int from = p->exprs[e->a].type, to = e->type;
    int tf = from == SD_DATE || from == SD_TIMESTAMP, tt = to == SD_DATE || to == SD_TIMESTAMP;
    if (from == SD_STRING || to == SD_STRING) return fail(SD_ERR_UNSUPPORTED, "casts involving STRING");
    if ((tf || tt) && from != to) return fail(SD_ERR_UNSUPPORTED, "casts involving DATE / TIMESTAMP");
    if (from == SD_DECIMAL && !(is_fp(to) || to == SD_DECIMAL)) return fail(SD_ERR_UNSUPPORTED, "this cast from DECIMAL");
    if (to == SD_DECIMAL && !(from == SD_BYTE || from == SD_SHORT || from == SD_INT || from == SD_LONG || from == SD_DECIMAL))
      return fail(SD_ERR_UNSUPPORTED, "this cast to DECIMAL");
    if (from == SD_DECIMAL && to == SD_DECIMAL && (dec_ps(p, i) & 0xff) < (dec_ps(p, e->a) & 0xff))
      return fail(SD_ERR_UNSUPPORTED, "DECIMAL cast that reduces the scale");
  }
  return 0;
}

static int build_buffer_schema(oracle_plan* p) {
  int n = 0;
  for (int i = 0; i < p->desc.naggs; i++) n += p->aggs[i].fn == SD_AGG_AVG ? 2 : 1;
  p->nbuf = n;
  p->buf_type = (int*)calloc(n + 1, sizeof(int));
  p->buf_ps = (int*)calloc(n + 1, sizeof(int));
  p->buf_nullable = (int*)calloc(n + 1, sizeof(int));
  { int rc0 = check_casts(p); if (rc0) return rc0; }
  int k = 0, keyed = p->desc.nkeys > 0;
  for (int i = 0; i < p->desc.naggs; i++) {
    const sd_agg* a = &p->aggs[i];
    int ct = a->expr >= 0 ? p->exprs[a->expr].type : SD_LONG;
    int cn = a->expr >= 0 ? p->expr_nullable[a->expr] : 0;
    if (a->expr >= 0 && (ct == SD_STRING) && a->fn != SD_AGG_COUNT && a->fn != SD_AGG_MIN && a->fn != SD_AGG_MAX)
      return fail(SD_ERR_UNSUPPORTED, "aggregates over STRING: only MIN / MAX / COUNT");
    if (a->expr >= 0 && ct == SD_STRING && (a->fn == SD_AGG_MIN || a->fn == SD_AGG_MAX) && p->exprs[a->expr].op != SD_OP_COL)
      return fail(SD_ERR_UNSUPPORTED, "aggregates over STRING: only MIN / MAX / COUNT of a STRING column");
    int cps = ct == SD_DECIMAL ? dec_ps(p, a->expr) : 0;
    /* Sum / Average over DECIMAL(p,s): sumDataType = DecimalType.bounded(p + 10, s) (Spark 2.1.1 Sum.scala / Average.scala) */
    int sum_ps = (imin(38, (cps >> 8) + 10) << 8) | (cps & 0xff);
    switch (a->fn) {
      case SD_AGG_COUNT_STAR: case SD_AGG_COUNT: p->buf_type[k] = SD_LONG; p->buf_nullable[k++] = 0; break;
      /* grouped: non-nullable when the child is (bufferAttributesForGroup :174-204);
       * no keys: plain Spark buffers, nullable (doProduceWithoutKeys :337-346) */
      case SD_AGG_SUM: p->buf_type[k] = sum_buffer_type(ct); p->buf_ps[k] = sum_ps; p->buf_nullable[k++] = keyed ? cn : 1; break;
      case SD_AGG_AVG:
        p->buf_type[k] = ct == SD_DECIMAL ? SD_DECIMAL : SD_DOUBLE; p->buf_ps[k] = sum_ps; p->buf_nullable[k++] = 0;   /* Average.sum starts at 0 */
        p->buf_type[k] = SD_LONG; p->buf_nullable[k++] = 0; break;
      case SD_AGG_MIN: case SD_AGG_MAX: p->buf_type[k] = ct; p->buf_ps[k] = cps; p->buf_nullable[k++] = keyed ? cn : 1; break;
      default: return fail(SD_ERR_INVALID, "unknown aggregate function");
    }
  }
  return 0;
}

static void init_buffers(const oracle_plan* p, val* b) {
  int k = 0, keyed = p->desc.nkeys > 0;
  for (int i = 0; i < p->desc.naggs; i++) {
    const sd_agg* a = &p->aggs[i];
    int cn = a->expr >= 0 ? p->expr_nullable[a->expr] : 0;
    switch (a->fn) {
      case SD_AGG_COUNT_STAR: case SD_AGG_COUNT: memset(&b[k], 0, sizeof(val)); k++; break;
      case SD_AGG_SUM: memset(&b[k], 0, sizeof(val)); b[k].isnull = !(keyed && !cn); k++; break;
      case SD_AGG_AVG: memset(&b[k], 0, sizeof(val)); memset(&b[k + 1], 0, sizeof(val)); k += 2; break;
      default: memset(&b[k], 0, sizeof(val)); b[k].isnull = 1; k++; break;
    }
  }
}

static double to_f64(const val* v, int t) { return is_fp(t) ? v->d : (double)v->i; }

/* updateExpressions of Sum / Average / Count / Min / Max */
static void update_buffers(const oracle_plan* p, val* b, const val* cols) {
  int k = 0;
  for (int i = 0; i < p->desc.naggs; i++) {
    const sd_agg* a = &p->aggs[i];
    val v; int t = SD_LONG;
    if (a->expr >= 0) { v = eval(p, a->expr, cols); t = p->exprs[a->expr].type; }
    else { memset(&v, 0, sizeof(v)); }
    switch (a->fn) {
      case SD_AGG_COUNT_STAR: b[k++].i += 1; break;
      case SD_AGG_COUNT: if (!v.isnull) b[k].i += 1; k++; break;
      case SD_AGG_SUM:
        if (!v.isnull) {
          if (is_fp(t)) b[k].d = (b[k].isnull ? 0.0 : b[k].d) + v.d;
          else if (t == SD_DECIMAL) { b[k].w = (b[k].isnull ? (__int128)0 : b[k].w) + v.i; b[k].i = (int64_t)b[k].w; }   /* Decimal + Decimal, exact */
          else b[k].i = (int64_t)((uint64_t)(b[k].isnull ? 0 : b[k].i) + (uint64_t)v.i);
          b[k].isnull = 0;
        }
        k++; break;
      case SD_AGG_AVG:
        if (!v.isnull) {
          if (t == SD_DECIMAL) { b[k].w += v.i; b[k].i = (int64_t)b[k].w; } else b[k].d += to_f64(&v, t);
          b[k + 1].i += 1;
        }
        k += 2; break;
      case SD_AGG_MIN: case SD_AGG_MAX:
        if (!v.isnull) {
          v.w = v.i;
          int take = b[k].isnull;
          if (!take) {
            int c = cmp_val(&v, &b[k], t);     /* UTF8String.compareTo for strings: unsigned bytes, then length */
            take = (a->fn == SD_AGG_MIN && c < 0) || (a->fn == SD_AGG_MAX && c > 0);
          }
          if (take) {
            if (t == SD_STRING) {               /* the buffer owns its bytes (the batch's memory goes away) */
              uint8_t* own = (uint8_t*)malloc(v.slen > 0 ? v.slen : 1);
              memcpy(own, v.s, v.slen);
              if (!b[k].isnull) free((void*)b[k].s);
              v.s = own;
            }
            b[k] = v;
          }
        }
        k++; break;
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* group lookup: hash of the key tuple, open addressing with quadratic probing                 */
static uint64_t hash_mix(uint64_t h, uint64_t v) {
  h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
  h *= 0xff51afd7ed558ccdULL; h ^= h >> 33;
  return h;
}
static uint64_t hash_keys(const oracle_plan* p, const val* k) {
  uint64_t h = 42;
  for (int i = 0; i < p->desc.nkeys; i++) {
    int t = p->exprs[p->keys[i]].type;
    if (k[i].isnull) { h = hash_mix(h, 0x5bd1e995); continue; }
    if (t == SD_STRING) { for (int j = 0; j < k[i].slen; j++) h = hash_mix(h, k[i].s[j]); h = hash_mix(h, (uint64_t)k[i].slen); }
    else if (is_fp(t)) { double d = k[i].d; if (d == 0.0) d = 0.0; if (isnan(d)) d = NAN; uint64_t u; memcpy(&u, &d, 8); h = hash_mix(h, isnan(d) ? 0x7ff8000000000000ULL : u); }
    else h = hash_mix(h, (uint64_t)k[i].i);
  }
  return h;
}
static int keys_equal(const oracle_plan* p, const val* a, const val* b) {
  for (int i = 0; i < p->desc.nkeys; i++) {
    if (a[i].isnull != b[i].isnull) return 0;
    if (a[i].isnull) continue;
    if (cmp_val(&a[i], &b[i], p->exprs[p->keys[i]].type) != 0) return 0;
  }
  return 1;
}
static void rehash(oracle_plan* p) {
  int ncap = p->hcap ? p->hcap * 2 : 8192;      /* sql.initialCapacityOfSHABBMap = 8192 (Literals.scala:312-315) */
  int* t = (int*)malloc(sizeof(int) * ncap);
  for (int i = 0; i < ncap; i++) t[i] = -1;
  for (int g = 0; g < p->ngroups; g++) {
    uint64_t pos = p->ghash[g] & (ncap - 1);
    for (int d = 1; t[pos] >= 0; d++) pos = (pos + d) & (ncap - 1);
    t[pos] = g;
  }
  free(p->htab); p->htab = t; p->hcap = ncap;
}
static val* find_or_insert_group(oracle_plan* p, const val* k) {
  if (p->hcap == 0 || (p->ngroups + 1) * 4 > p->hcap * 3) rehash(p);
  uint64_t h = hash_keys(p, k);
  uint64_t pos = h & (p->hcap - 1);
  for (int d = 1; p->htab[pos] >= 0; d++) {
    int g = p->htab[pos];
    if (p->ghash[g] == h && keys_equal(p, p->gkeys[g], k)) return p->gbufs[g];
    pos = (pos + d) & (p->hcap - 1);
  }
  if (p->ngroups == p->cap_groups) {
    p->cap_groups = p->cap_groups ? p->cap_groups * 2 : 64;
    p->gkeys = (val**)realloc(p->gkeys, sizeof(val*) * p->cap_groups);
    p->gbufs = (val**)realloc(p->gbufs, sizeof(val*) * p->cap_groups);
    p->ghash = (uint64_t*)realloc(p->ghash, sizeof(uint64_t) * p->cap_groups);
  }
  int g = p->ngroups++;
  p->gkeys[g] = (val*)malloc(sizeof(val) * (p->desc.nkeys + 1));
  for (int i = 0; i < p->desc.nkeys; i++) {
    p->gkeys[g][i] = k[i];
    if (!k[i].isnull && p->exprs[p->keys[i]].type == SD_STRING) {
      uint8_t* c = (uint8_t*)malloc(k[i].slen + 1); memcpy(c, k[i].s, k[i].slen); p->gkeys[g][i].s = c;
    }
  }
  p->gbufs[g] = (val*)malloc(sizeof(val) * (p->nbuf + 1));
  init_buffers(p, p->gbufs[g]);
  p->ghash[g] = h;
  p->htab[pos] = g;
  return p->gbufs[g];
}

/* ------------------------------------------------------------------------------------------ */
/* UnsafeRow writer                                                    */
static void out_reserve(oracle_plan* p, int64_t more) {
  if (p->out_len + more > p->out_cap) {
    p->out_cap = (p->out_len + more) * 2 + 1024;
    p->out = (uint8_t*)realloc(p->out, p->out_cap);
  }
}
static void emit_row(oracle_plan* p, int n, const int* types, const val* vals) {
  int64_t bits = ((n + 63) / 64) * 8, fixed = bits + 8 * (int64_t)n, var = 0;
  for (int i = 0; i < n; i++) {
    if (types[i] == SD_STRING && !vals[i].isnull) var += (vals[i].slen + 7) & ~7;
    if (FT_BASE(types[i]) == SD_DECIMAL && FT_PREC(types[i]) > 18) var += 16;   /* UnsafeRowWriter.write(Decimal): 16 bytes always reserved */
  }
  int64_t sz = fixed + var;
  out_reserve(p, 8 + sz);
  uint8_t* r = p->out + p->out_len;
  memcpy(r, &sz, 8); r += 8;
  memset(r, 0, sz);
  int64_t voff = fixed;
  for (int i = 0; i < n; i++) {
    uint8_t* slot = r + bits + 8 * (int64_t)i;
    if (FT_BASE(types[i]) == SD_DECIMAL && FT_PREC(types[i]) > 18) {
      /* precision > MAX_LONG_DIGITS: BigInteger.toByteArray() (minimal big-endian two's complement) in a 16-byte region */
      if (vals[i].isnull) { r[i >> 3] |= (uint8_t)(1u << (i & 7)); int64_t ol = voff << 32; memcpy(slot, &ol, 8); voff += 16; continue; }
      uint8_t be[16]; __int128 v = vals[i].w;
      for (int k = 15; k >= 0; k--) { be[k] = (uint8_t)(v & 0xff); v >>= 8; }
      int first = 0;
      while (first < 15 && ((be[first] == 0 && !(be[first + 1] & 0x80)) || (be[first] == 0xff && (be[first + 1] & 0x80)))) first++;
      memcpy(r + voff, be + first, 16 - first);
      int64_t ol = (voff << 32) | (16 - first); memcpy(slot, &ol, 8); voff += 16;
      continue;
    }
    if (vals[i].isnull) { r[i >> 3] |= (uint8_t)(1u << (i & 7)); continue; }
    if (FT_BASE(types[i]) == SD_DECIMAL) { int64_t v = (int64_t)vals[i].w; memcpy(slot, &v, 8); continue; }
    switch (types[i]) {
      case SD_STRING: { int64_t ol = (voff << 32) | (uint32_t)vals[i].slen; memcpy(slot, &ol, 8);
                        memcpy(r + voff, vals[i].s, vals[i].slen); voff += (vals[i].slen + 7) & ~7; break; }
      case SD_BOOLEAN: slot[0] = vals[i].i != 0; break;
      case SD_BYTE: { int8_t v = (int8_t)vals[i].i; memcpy(slot, &v, 1); break; }
      case SD_SHORT: { int16_t v = (int16_t)vals[i].i; memcpy(slot, &v, 2); break; }
      case SD_INT: case SD_DATE: { int32_t v = (int32_t)vals[i].i; memcpy(slot, &v, 4); break; }
      case SD_FLOAT: { float v = (float)vals[i].d; memcpy(slot, &v, 4); break; }
      case SD_DOUBLE: memcpy(slot, &vals[i].d, 8); break;
      default: memcpy(slot, &vals[i].i, 8); break;
    }
  }
  p->out_len += 8 + sz; p->out_rows++;
}

/* ------------------------------------------------------------------------------------------ */
/* stats-row batch skipping (ColumnTableScan.generateStatPredicate, ColumnTableScan.scala:820-963) */
typedef struct { int isnull; int v; } tri;
static int unsafe_field(const uint8_t* row, int64_t len, int nfields, int idx, int type, val* out) {
  int64_t bits = ((nfields + 63) / 64) * 8;
  if (bits + 8 * (int64_t)nfields > len) return -1;
  memset(out, 0, sizeof(*out));
  if (row[idx >> 3] & (1u << (idx & 7))) { out->isnull = 1; return 0; }
  const uint8_t* slot = row + bits + 8 * (int64_t)idx;
  if (FT_BASE(type) == SD_DECIMAL) {   /* UnsafeRow.getDecimal */
    if (FT_PREC(type) <= 18) { out->i = ld_i64(slot); out->w = out->i; return 0; }
    int64_t ol = ld_i64(slot); const uint8_t* b = row + (ol >> 32); int ln = (int)(ol & 0xffffffff);
    __int128 v = (ln > 0 && (b[0] & 0x80)) ? -1 : 0;
    for (int k = 0; k < ln; k++) v = (v << 8) | b[k];
    out->w = v; out->i = (int64_t)v; return 0;
  }
  switch (type) {
    case SD_STRING: { int64_t ol = ld_i64(slot); out->s = row + (ol >> 32); out->slen = (int32_t)(ol & 0xffffffff); break; }
    case SD_BOOLEAN: out->i = slot[0] != 0; break;
    case SD_BYTE: out->i = (int8_t)slot[0]; break;
    case SD_SHORT: out->i = ld_i16(slot); break;
    case SD_INT: case SD_DATE: out->i = ld_i32(slot); break;
    case SD_FLOAT: out->d = ld_f32(slot); break;
    case SD_DOUBLE: out->d = ld_f64(slot); break;
    default: out->i = ld_i64(slot); break;
  }
  return 0;
}
static tri tri_and(tri a, tri b) { tri r = {0, 0}; if ((!a.isnull && !a.v) || (!b.isnull && !b.v)) return r; if (a.isnull || b.isnull) { r.isnull = 1; return r; } r.v = 1; return r; }
static tri tri_or(tri a, tri b) { tri r = {0, 1}; if ((!a.isnull && a.v) || (!b.isnull && b.v)) return r; if (a.isnull || b.isnull) { r.isnull = 1; r.v = 0; return r; } r.v = 0; return r; }
static tri tri_cmp(const val* a, const val* b, int t, int op) {   /* op: 0 '<', 1 '<=' */
  tri r = {0, 0};
  if (a->isnull || b->isnull) { r.isnull = 1; return r; }
  int c = cmp_val(a, b, t);
  r.v = op == 0 ? c < 0 : c <= 0;
  return r;
}
/* returns 1 if a stats predicate could be derived for `node` (buildFilter.isDefinedAt), result in *out */
static int stat_filter(const oracle_plan* p, int node, const uint8_t* stats, int64_t slen, int nfields, int num_rows, tri* out) {
  const sd_expr* e = &p->exprs[node];
  tri l, r;
  switch (e->op) {
    case SD_OP_AND: {
      int dl = stat_filter(p, e->a, stats, slen, nfields, num_rows, &l), dr = stat_filter(p, e->b, stats, slen, nfields, num_rows, &r);
      if (!dl && !dr) return 0;
      *out = dl && dr ? tri_and(l, r) : (dl ? l : r);
      return 1;
    }
    case SD_OP_OR: {
      int dl = stat_filter(p, e->a, stats, slen, nfields, num_rows, &l), dr = stat_filter(p, e->b, stats, slen, nfields, num_rows, &r);
      if (!(dl && dr)) return 0;
      *out = tri_or(l, r); return 1;
    }
    case SD_OP_EQ: case SD_OP_LT: case SD_OP_LE: case SD_OP_GT: case SD_OP_GE: {
      const sd_expr *ea = &p->exprs[e->a], *eb = &p->exprs[e->b];
      int col_left = ea->op == SD_OP_COL && eb->op == SD_OP_LIT, col_right = eb->op == SD_OP_COL && ea->op == SD_OP_LIT;
      if (!col_left && !col_right) return 0;
      const sd_expr* ec = col_left ? ea : eb; const sd_expr* el = col_left ? eb : ea;
      int t = ec->type, ord = p->cols[ec->a].table_ordinal;
      if (1 + 3 * ord + 2 >= nfields) return 0;
      val lo, hi, lit; memset(&lit, 0, sizeof(lit));
      if (unsafe_field(stats, slen, nfields, 1 + 3 * ord, t, &lo) || unsafe_field(stats, slen, nfields, 2 + 3 * ord, t, &hi)) return 0;
      const sd_literal* L = &p->lits[el->a];
      lit.isnull = L->is_null; lit.i = L->i; lit.d = t == SD_FLOAT ? (double)(float)L->d : L->d; lit.s = (const uint8_t*)L->s; lit.slen = L->slen;
      int op = e->op;
      if (col_right) { if (op == SD_OP_LT) op = SD_OP_GT; else if (op == SD_OP_LE) op = SD_OP_GE; else if (op == SD_OP_GT) op = SD_OP_LT; else if (op == SD_OP_GE) op = SD_OP_LE; }
      switch (op) {
        case SD_OP_EQ: *out = tri_and(tri_cmp(&lo, &lit, t, 1), tri_cmp(&lit, &hi, t, 1)); break;
        case SD_OP_LT: *out = tri_cmp(&lo, &lit, t, 0); break;
        case SD_OP_LE: *out = tri_cmp(&lo, &lit, t, 1); break;
        case SD_OP_GT: *out = tri_cmp(&lit, &hi, t, 0); break;
        default: *out = tri_cmp(&lit, &hi, t, 1); break;
      }
      return 1;
    }
    case SD_OP_IN: {
      const sd_expr* ea = &p->exprs[e->a];
      if (ea->op != SD_OP_COL || e->c > 200 || e->c < 1) return 0;
      int t = ea->type, ord = p->cols[ea->a].table_ordinal;
      if (1 + 3 * ord + 2 >= nfields) return 0;
      val lo, hi, mn, mx; int have = 0;
      if (unsafe_field(stats, slen, nfields, 1 + 3 * ord, t, &lo) || unsafe_field(stats, slen, nfields, 2 + 3 * ord, t, &hi)) return 0;
      memset(&mn, 0, sizeof(mn)); memset(&mx, 0, sizeof(mx));
      for (int k = 0; k < e->c; k++) {   /* Greatest / Least skip nulls */
        const sd_literal* L = &p->lits[e->b + k]; if (L->is_null) continue;
        val v; memset(&v, 0, sizeof(v)); v.i = L->i; v.d = t == SD_FLOAT ? (double)(float)L->d : L->d; v.s = (const uint8_t*)L->s; v.slen = L->slen;
        if (!have) { mn = mx = v; have = 1; } else { if (cmp_val(&v, &mn, t) < 0) mn = v; if (cmp_val(&v, &mx, t) > 0) mx = v; }
      }
      if (!have) { mn.isnull = mx.isnull = 1; }
      *out = tri_and(tri_cmp(&lo, &mx, t, 1), tri_cmp(&mn, &hi, t, 1));
      return 1;
    }
    case SD_OP_STARTSWITH: {
      const sd_expr *ea = &p->exprs[e->a], *eb = &p->exprs[e->b];
      if (ea->op != SD_OP_COL || eb->op != SD_OP_LIT) return 0;
      int ord = p->cols[ea->a].table_ordinal;
      if (1 + 3 * ord + 2 >= nfields) return 0;
      val lo, hi;
      if (unsafe_field(stats, slen, nfields, 1 + 3 * ord, SD_STRING, &lo) || unsafe_field(stats, slen, nfields, 2 + 3 * ord, SD_STRING, &hi)) return 0;
      const sd_literal* L = &p->lits[eb->a];
      /* StartsWithForStats.doGenCode (ColumnTableScan.scala:1028-1088); never NULL */
      tri r0 = {0, 1};
      if (!L->is_null) {
        int plen = L->slen, last = plen - 1;
        uint8_t* up = (uint8_t*)malloc(plen + 1); memcpy(up, L->s, plen);
        while (last >= 0 && up[last] == 0xff) last--;
        if (last < 0 || lo.isnull) {
          if (!hi.isnull) r0.v = cmp_bytes((const uint8_t*)L->s, plen, hi.s, hi.slen) <= 0;
        } else {
          up[last] = (uint8_t)(up[last] + 1);
          r0.v = (hi.isnull || cmp_bytes((const uint8_t*)L->s, plen, hi.s, hi.slen) <= 0) && cmp_bytes(lo.s, lo.slen, up, plen) < 0;
        }
        free(up);
      }
      *out = r0; return 1;
    }
    case SD_OP_ISNULL: case SD_OP_ISNOTNULL: {
      const sd_expr* ea = &p->exprs[e->a];
      if (ea->op != SD_OP_COL) return 0;
      int ord = p->cols[ea->a].table_ordinal;
      if (1 + 3 * ord + 2 >= nfields) return 0;
      val nc;
      if (unsafe_field(stats, slen, nfields, 3 + 3 * ord, SD_INT, &nc)) return 0;
      tri r0 = {nc.isnull, 0};
      if (!nc.isnull) r0.v = e->op == SD_OP_ISNULL ? nc.i > 0 : num_rows > nc.i;
      *out = r0; return 1;
    }
  }
  return 0;
}
/* 1 => scan the batch, 0 => skip it (only a definite FALSE skips; :948-957) */
static int batch_passes_stats(const oracle_plan* p, const sd_batch* b) {
  if (p->desc.filter < 0 || !b->stats_row || b->stats_len <= 0) return 1;
  int nfields = 1 + 3 * b->stats_ncols;
  /* the reference splits the condition into conjuncts and ANDs the derived filters; AND handling in
   * stat_filter is exactly that ("either side defined") */
  tri r;
  if (!stat_filter(p, p->desc.filter, (const uint8_t*)b->stats_row, b->stats_len, nfields, b->num_rows, &r)) return 1;
  return r.isnull || r.v;
}

/* ------------------------------------------------------------------------------------------ */
/* plan lifecycle                                                                               */
static void* dup_mem(const void* src, size_t n) { void* d = malloc(n ? n : 1); if (n) memcpy(d, src, n); return d; }

int oracle_plan_create(const sd_plan_desc* desc, oracle_plan** out) {
  if (!desc || !out) return fail(SD_ERR_INVALID, "null argument");
  if (desc->ncols < 0 || desc->ncols > MAXCOLS) return fail(SD_ERR_INVALID, "bad column count");
  oracle_plan* p = (oracle_plan*)calloc(1, sizeof(*p));
  p->desc = *desc;
  p->cols = (sd_column*)dup_mem(desc->cols, sizeof(sd_column) * desc->ncols);
  p->exprs = (sd_expr*)dup_mem(desc->exprs, sizeof(sd_expr) * desc->nexprs);
  p->keys = (int32_t*)dup_mem(desc->keys, sizeof(int32_t) * desc->nkeys);
  p->aggs = (sd_agg*)dup_mem(desc->aggs, sizeof(sd_agg) * desc->naggs);
  p->proj = (int32_t*)dup_mem(desc->proj, sizeof(int32_t) * desc->nproj);
  p->lit_types = (int32_t*)dup_mem(desc->literal_types, sizeof(int32_t) * desc->nliterals);
  p->lits = (sd_literal*)calloc(desc->nliterals + 1, sizeof(sd_literal));
  p->lit_strs = (char**)calloc(desc->nliterals + 1, sizeof(char*));
  p->expr_nullable = (int*)calloc(desc->nexprs + 1, sizeof(int));
  for (int i = 0; i < desc->nexprs; i++) {
    const sd_expr* e = &p->exprs[i];
    int unary = e->op == SD_OP_NEG || e->op == SD_OP_CAST || e->op == SD_OP_NOT || e->op == SD_OP_ISNULL || e->op == SD_OP_ISNOTNULL || e->op == SD_OP_IN;
    if (e->op == SD_OP_COL) { if (e->a < 0 || e->a >= desc->ncols) return fail(SD_ERR_INVALID, "column reference out of range"); }
    else if (e->op == SD_OP_LIT) { if (e->a < 0 || e->a >= desc->nliterals) return fail(SD_ERR_INVALID, "literal slot out of range"); }
    else if (e->a < 0 || e->a >= i || (!unary && (e->b < 0 || e->b >= i))) return fail(SD_ERR_INVALID, "expression children must precede parents");
  }
  compute_nullability(p);
  int rc = build_buffer_schema(p);
  if (rc) return rc;
  *out = p;
  return 0;
}

int oracle_plan_set_literals(oracle_plan* p, const sd_literal* vals, int32_t n) {
  if (n != p->desc.nliterals) return fail(SD_ERR_INVALID, "literal count mismatch");
  for (int i = 0; i < n; i++) {
    free(p->lit_strs[i]); p->lit_strs[i] = NULL;
    p->lits[i] = vals[i];
    if (vals[i].s && vals[i].slen >= 0) { p->lit_strs[i] = (char*)dup_mem(vals[i].s, vals[i].slen); p->lits[i].s = p->lit_strs[i]; }
  }
  return 0;
}

static void consume_row(oracle_plan* p, const val* cols) {
  p->metrics[11]++;                              /* ColumnTableScan numOutputRows */
  if (p->desc.filter >= 0) {                     /* FilterExec: only TRUE passes */
    val f = eval(p, p->desc.filter, cols);
    if (f.isnull || !f.i) return;
  }
  if (p->desc.naggs == 0 && p->desc.nkeys == 0) { /* projection */
    int n = p->desc.nproj; int types[MAXCOLS]; val vals[MAXCOLS];
    for (int i = 0; i < n; i++) { types[i] = p->exprs[p->proj[i]].type; vals[i] = eval(p, p->proj[i], cols); }
    emit_row(p, n, types, vals);
    return;
  }
  val* bufs;
  if (p->desc.nkeys == 0) {                      /* doConsumeWithoutKeys :450-491 */
    if (p->ngroups == 0) { val none; memset(&none, 0, sizeof(none)); bufs = find_or_insert_group(p, &none); }
    else bufs = p->gbufs[0];
  } else {                                       /* doConsumeWithKeysForSHAMap :1278-1580 */
    val k[64];
    for (int i = 0; i < p->desc.nkeys; i++) k[i] = eval(p, p->keys[i], cols);
    bufs = find_or_insert_group(p, k);
  }
  update_buffers(p, bufs, cols);
}

/* the batch loop of ColumnTableScan.doProduce (ColumnTableScan.scala:565-599, 641-661) */
int oracle_batch_submit(oracle_plan* p, const sd_batch* b) {
  if (b->ncols != p->desc.ncols) return fail(SD_ERR_INVALID, "batch column count != plan column count");
  p->metrics[2]++;                               /* columnBatchesSeen */
  if (!batch_passes_stats(p, b)) { p->metrics[5]++; return 0; }
  int nc = b->ncols, rc = 0;
  coldec* dec = (coldec*)calloc(nc + 1, sizeof(coldec));
  updec* upd = (updec*)calloc(nc + 1, sizeof(updec));
  val* cols = (val*)calloc(nc + 1, sizeof(val));
  for (int c = 0; c < nc && !rc; c++) {
    rc = coldec_init(&dec[c], (const uint8_t*)b->col_bufs[c], b->col_lens[c], p->cols[c].type, p->cols[c].nullable, 0);
    const uint8_t* d0 = b->delta0 ? (const uint8_t*)b->delta0[c] : NULL;
    const uint8_t* d1 = b->delta1 ? (const uint8_t*)b->delta1[c] : NULL;
    if (!rc && (d0 || d1)) {
      upd[c].present = 1;
      rc = deltadec_init(&upd[c].d1, d0, d0 ? b->delta0_lens[c] : 0, p->cols[c].type, p->cols[c].nullable);
      if (!rc) rc = deltadec_init(&upd[c].d2, d1, d1 ? b->delta1_lens[c] : 0, p->cols[c].type, p->cols[c].nullable);
      if (!rc) { updec_start(&upd[c]); p->metrics[3]++; }
    }
  }
  deldec del; deldec_init(&del, (const uint8_t*)b->delete_buf, b->delete_len);
  if (b->delete_buf) p->metrics[4]++;
  for (int ordinal = 0; ordinal < b->num_rows && !rc; ordinal++) {
    if (deldec_deleted(&del, ordinal)) continue;                       /* :481-482 */
    for (int c = 0; c < nc && !rc; c++) {                              /* genCodeColumnBuffer :757-786 */
      if (!upd[c].present || updec_unchanged(&upd[c], ordinal)) {
        int k;
        if (coldec_is_null(&dec[c], ordinal, &k)) { memset(&cols[c], 0, sizeof(val)); cols[c].isnull = 1; }
        else rc = coldec_read(&dec[c], k, &cols[c]);
      } else {
        deltadec* cur = upd[c].current;
        if (cur->not_null) rc = coldec_read(&cur->real, cur->non_null_position, &cols[c]);
        else { memset(&cols[c], 0, sizeof(val)); cols[c].isnull = 1; }
      }
    }
    if (!rc) { consume_row(p, cols); p->metrics[8]++; }
  }
  for (int c = 0; c < nc; c++) { coldec_free(&dec[c]); coldec_free(&upd[c].d1.real); coldec_free(&upd[c].d2.real); }
  free(dec); free(upd); free(cols);
  return rc;
}

/* row-buffer rows (ColumnTableScan.scala:572-588: the same loop with numBatchRows = 1) */
int oracle_rows_submit(oracle_plan* p, const void* rows, int64_t len, int32_t nrows) {
  const uint8_t* r = (const uint8_t*)rows; int64_t pos = 0; int nc = p->desc.ncols;
  val* cols = (val*)calloc(nc + 1, sizeof(val));
  for (int i = 0; i < nrows; i++) {
    if (pos + 8 > len) { free(cols); return fail(SD_ERR_INVALID, "truncated row stream"); }
    int64_t sz = ld_i64(r + pos);
    for (int c = 0; c < nc; c++) unsafe_field(r + pos + 8, sz, nc, c, p->cols[c].type, &cols[c]);
    consume_row(p, cols); p->metrics[1]++; p->metrics[8]++;
    pos += 8 + sz;
  }
  free(cols);
  return 0;
}

int oracle_plan_finish(oracle_plan* p, void* out_rows, int64_t cap, int64_t* out_len, int64_t* out_nrows) {
  if (p->desc.naggs > 0 || p->desc.nkeys > 0) {
    p->out_len = 0; p->out_rows = 0;
    int nk = p->desc.nkeys, n = nk + p->nbuf;
    int* types = (int*)malloc(sizeof(int) * (n + 1));
    val* vals = (val*)malloc(sizeof(val) * (n + 1));
    for (int i = 0; i < nk; i++) types[i] = FT(p->exprs[p->keys[i]].type, p->exprs[p->keys[i]].type == SD_DECIMAL ? dec_ps(p, p->keys[i]) : 0);
    for (int i = 0; i < p->nbuf; i++) types[nk + i] = FT(p->buf_type[i], p->buf_ps[i]);
    if (nk == 0 && p->ngroups == 0) {            /* no-key aggregate always outputs one row */
      val none; memset(&none, 0, sizeof(none)); find_or_insert_group(p, &none);
    }
    for (int g = 0; g < p->ngroups; g++) {       /* insertion order */
      for (int i = 0; i < nk; i++) { vals[i] = p->gkeys[g][i]; vals[i].w = vals[i].i; }
      for (int i = 0; i < p->nbuf; i++) vals[nk + i] = p->gbufs[g][i];
      emit_row(p, n, types, vals);
    }
    free(types); free(vals);
  }
  p->metrics[0] = p->out_rows;
  *out_len = p->out_len; if (out_nrows) *out_nrows = p->out_rows;
  if (p->out_len > cap) return fail(SD_ERR_OVERFLOW, "output buffer too small");
  if (p->out_len) memcpy(out_rows, p->out, p->out_len);
  return 0;
}

static void free_groups(oracle_plan* p) {
  for (int g = 0; g < p->ngroups; g++) {
    for (int i = 0; i < p->desc.nkeys; i++)
      if (!p->gkeys[g][i].isnull && p->exprs[p->keys[i]].type == SD_STRING) free((void*)p->gkeys[g][i].s);
    free(p->gkeys[g]); free(p->gbufs[g]);
  }
  p->ngroups = 0;
  if (p->htab) for (int i = 0; i < p->hcap; i++) p->htab[i] = -1;
}
int oracle_plan_reset(oracle_plan* p) {
  free_groups(p); p->out_len = 0; p->out_rows = 0; memset(p->metrics, 0, sizeof(p->metrics)); return 0;
}
int oracle_plan_metrics(oracle_plan* p, int64_t out[SD_NUM_METRICS]) { memcpy(out, p->metrics, sizeof(p->metrics)); return 0; }
void oracle_plan_destroy(oracle_plan* p) {
  if (!p) return;
  free_groups(p);
  free(p->gkeys); free(p->gbufs); free(p->ghash); free(p->htab);
  for (int i = 0; i < p->desc.nliterals; i++) free(p->lit_strs[i]);
  free(p->cols); free(p->exprs); free(p->keys); free(p->aggs); free(p->proj); free(p->lit_types);
  free(p->lits); free(p->lit_strs); free(p->buf_type); free(p->buf_ps); free(p->buf_nullable); free(p->expr_nullable); free(p->out);
  free(p);
}

/* ------------------------------------------------------------------------------------------ */
/* final merge: mergeExpressions + evaluateExpression of each function
 * (SnappyHashAggregateExec.scala:1317-1326; CollectAggregateExec.scala:67-121)                */
int oracle_final_merge(const sd_plan_desc* desc, const void* partial_rows, int64_t len,
                       void* out_rows, int64_t cap, int64_t* out_len, int64_t* out_nrows) {
  oracle_plan* p; int rc = oracle_plan_create(desc, &p);
  if (rc) return rc;
  int nk = desc->nkeys, n = nk + p->nbuf;
  int* types = (int*)malloc(sizeof(int) * (n + 1));
  for (int i = 0; i < nk; i++) types[i] = FT(p->exprs[p->keys[i]].type, p->exprs[p->keys[i]].type == SD_DECIMAL ? dec_ps(p, p->keys[i]) : 0);
  for (int i = 0; i < p->nbuf; i++) types[nk + i] = FT(p->buf_type[i], p->buf_ps[i]);
  val* f = (val*)malloc(sizeof(val) * (n + 1));
  const uint8_t* r = (const uint8_t*)partial_rows; int64_t pos = 0;
  while (pos + 8 <= len) {
    int64_t sz = ld_i64(r + pos);
    for (int i = 0; i < n; i++) unsafe_field(r + pos + 8, sz, n, i, types[i], &f[i]);
    val* b = find_or_insert_group(p, f);
    /* merge: a fresh group starts from "empty" buffers, so merging == the update rules on buffers */
    int k = 0;
    for (int i = 0; i < desc->naggs; i++) {
      const val* in = &f[nk + k];
      switch (p->aggs[i].fn) {
        case SD_AGG_COUNT_STAR: case SD_AGG_COUNT: b[k].i += in->i; k++; break;
        case SD_AGG_SUM:
          if (!in->isnull) {
            if (p->buf_type[k] == SD_DOUBLE) b[k].d = (b[k].isnull ? 0.0 : b[k].d) + in->d;
            else if (p->buf_type[k] == SD_DECIMAL) { b[k].w = (b[k].isnull ? (__int128)0 : b[k].w) + in->w; b[k].i = (int64_t)b[k].w; }
            else b[k].i = (int64_t)((uint64_t)(b[k].isnull ? 0 : b[k].i) + (uint64_t)in->i);
            b[k].isnull = 0;
          }
          k++; break;
        case SD_AGG_AVG:
          if (p->buf_type[k] == SD_DECIMAL) { b[k].w += in->w; b[k].i = (int64_t)b[k].w; } else b[k].d += in->d;
          b[k + 1].i += in[1].i; k += 2; break;
        default:
          if (!in->isnull) {
            if (b[k].isnull) b[k] = *in;
            else { int c = cmp_val(in, &b[k], p->buf_type[k]);
                   if ((p->aggs[i].fn == SD_AGG_MIN && c < 0) || (p->aggs[i].fn == SD_AGG_MAX && c > 0)) b[k] = *in; }
          }
          k++; break;
      }
    }
    pos += 8 + sz;
  }
  if (nk == 0 && p->ngroups == 0) { val none; memset(&none, 0, sizeof(none)); find_or_insert_group(p, &none); }
  int nout = nk + desc->naggs;
  int* otypes = (int*)malloc(sizeof(int) * (nout + 1));
  val* ov = (val*)malloc(sizeof(val) * (nout + 1));
  for (int g = 0; g < p->ngroups; g++) {
    for (int i = 0; i < nk; i++) { otypes[i] = types[i]; ov[i] = p->gkeys[g][i]; ov[i].w = ov[i].i; }
    int k = 0;
    for (int i = 0; i < desc->naggs; i++) {
      val* b = p->gbufs[g]; val o; memset(&o, 0, sizeof(o));
      if (p->aggs[i].fn == SD_AGG_AVG && p->buf_type[k] == SD_DECIMAL) {
        /* Average over DECIMAL(p,s): Cast(Cast(sum, (p+14,s+4)) / Cast(count, (p+14,s+4)), resultType = bounded(p+4, s+4)):
         * java.math.BigDecimal division, the quotient rounded HALF_UP to scale s+4; NULL when it does not fit p+4 digits */
        int in_ps = dec_ps(p, p->aggs[i].expr);
        int rp = imin(38, (in_ps >> 8) + 4), rs = imin(38, (in_ps & 0xff) + 4);
        otypes[nk + i] = FT(SD_DECIMAL, (rp << 8) | rs);
        if (b[k + 1].i == 0) o.isnull = 1;
        else {
          __int128 num = b[k].w * pow10_w(rs - (in_ps & 0xff)), den = b[k + 1].i;
          __int128 q = num / den, rem = num % den;
          if (rem < 0) rem = -rem;
          if (rem * 2 >= den) q += num < 0 ? -1 : 1;
          if (q >= pow10_w(rp) || q <= -pow10_w(rp)) o.isnull = 1; else { o.w = q; o.i = (int64_t)q; }
        }
        k += 2;
      } else if (p->aggs[i].fn == SD_AGG_AVG) {   /* Average.evaluateExpression: sum / count, NULL when count = 0 */
        otypes[nk + i] = SD_DOUBLE;
        if (b[k + 1].i == 0) o.isnull = 1; else o.d = b[k].d / (double)b[k + 1].i;
        k += 2;
      } else {
        otypes[nk + i] = FT(p->buf_type[k], p->buf_ps[k]); o = b[k];
        if (p->aggs[i].fn == SD_AGG_SUM && p->buf_type[k] == SD_DECIMAL && !o.isnull) {   /* more than p+10 digits: changePrecision fails -> NULL */
          __int128 lim = pow10_w(p->buf_ps[k] >> 8);
          if (o.w >= lim || o.w <= -lim) o.isnull = 1;
        }
        k++;
      }
      ov[nk + i] = o;
    }
    emit_row(p, nout, otypes, ov);
  }
  *out_len = p->out_len; if (out_nrows) *out_nrows = p->out_rows;
  rc = 0;
  if (p->out_len > cap) rc = fail(SD_ERR_OVERFLOW, "output buffer too small");
  else if (p->out_len) memcpy(out_rows, p->out, p->out_len);
  free(types); free(f); free(otypes); free(ov);
  oracle_plan_destroy(p);
  return rc;
}

/* ========================================================================================== */
/* Layer 2: restatements of the GENERATED loops for the benchmark queries.  What Janino would be
 * handed for these plans: NOT NULL columns, Uncompressed / Dictionary decoders inlined
 * (readInt = load at base + (ordinal << 2), ...), FilterExec conjuncts short-circuiting with the
 * column reads deferred to first use, aggregate buffers in locals (no keys) or a SHAMap probe with
 * the per-batch dictionary-index -> group-slot array (SnappyHashAggregateExec.scala:1340-1369).   */

typedef struct q1_group { int used; uint8_t k0[8]; int l0; uint8_t k1[8]; int l1;
                          double sum_qty, sum_price, sum_disc_price, sum_charge, avg_qty_s, avg_price_s, avg_disc_s;
                          int64_t avg_qty_c, avg_price_c, avg_disc_c, count; } q1_group;
#define Q1_MAXG 64

static inline const uint8_t* body_notnull(const void* buf) { return (const uint8_t*)buf + 8; }

/* Q6: sum(l_extendedprice*l_discount) where shipdate >= d0 and shipdate < d1 and
 * discount between lo and hi and quantity < q
 * (cluster/src/test/scala/io/snappydata/benchmark/TPCH_Queries.scala:600-613).
 * scan columns: 0 l_shipdate DATE, 1 l_discount, 2 l_quantity, 3 l_extendedprice (DOUBLE) */
int oracle_q6_batches(const sd_batch* b, int nb, int32_t d0, int32_t d1, double lo, double hi, double q,
                      double* sum_out, int* sum_isnull, int64_t* matched) {
  double sum = 0.0; int isnull = 1; int64_t m = 0;
  for (int i = 0; i < nb; i++) {
    const uint8_t* ship = body_notnull(b[i].col_bufs[0]);
    const uint8_t* disc = body_notnull(b[i].col_bufs[1]);
    const uint8_t* qty = body_notnull(b[i].col_bufs[2]);
    const uint8_t* price = body_notnull(b[i].col_bufs[3]);
    int n = b[i].num_rows;
    for (int o = 0; o < n; o++) {
      int32_t sd = ld_i32(ship + ((int64_t)o << 2));
      if (!(sd >= d0)) continue;
      if (!(sd < d1)) continue;
      double dv = ld_f64(disc + ((int64_t)o << 3));
      if (!(cmp_f64(dv, lo) >= 0)) continue;
      if (!(cmp_f64(dv, hi) <= 0)) continue;
      double qv = ld_f64(qty + ((int64_t)o << 3));
      if (!(cmp_f64(qv, q) < 0)) continue;
      double pv = ld_f64(price + ((int64_t)o << 3));
      sum = (isnull ? 0.0 : sum) + pv * dv; isnull = 0; m++;
    }
  }
  *sum_out = sum; *sum_isnull = isnull; if (matched) *matched = m;
  return 0;
}

/* C1: select count(*) where c1 > k over one INT NOT NULL column */
int oracle_c1_batches(const sd_batch* b, int nb, int32_t k, int64_t* count) {
  int64_t c = 0;
  for (int i = 0; i < nb; i++) {
    const uint8_t* v = body_notnull(b[i].col_bufs[0]);
    int n = b[i].num_rows;
    for (int o = 0; o < n; o++) if (ld_i32(v + ((int64_t)o << 2)) > k) c++;
  }
  *count = c; return 0;
}

/* Q1 (TPCH_Queries.scala:125-149). scan columns: 0 l_quantity 1 l_extendedprice 2 l_discount 3 l_tax
 * (DOUBLE) 4 l_returnflag 5 l_linestatus (dictionary STRING) 6 l_shipdate (DATE).
 * Groups found through a tiny open-addressing map keyed by the two strings; per batch the
 * (returnflag index, linestatus index) -> group slot array avoids re-hashing, as the reference's
 * dictionary-array shortcut does for its single-string-key case. */
static int q1_find(q1_group* g, const uint8_t* k0, int l0, const uint8_t* k1, int l1) {
  uint64_t h = 42;
  for (int i = 0; i < l0; i++) h = hash_mix(h, k0[i]);
  for (int i = 0; i < l1; i++) h = hash_mix(h, k1[i] + 256);
  int pos = (int)(h & (Q1_MAXG - 1));
  for (int d = 1;; d++) {
    if (!g[pos].used) { g[pos].used = 1; memcpy(g[pos].k0, k0, l0 < 8 ? l0 : 8); g[pos].l0 = l0; memcpy(g[pos].k1, k1, l1 < 8 ? l1 : 8); g[pos].l1 = l1; return pos; }
    if (g[pos].l0 == l0 && g[pos].l1 == l1 && !memcmp(g[pos].k0, k0, l0) && !memcmp(g[pos].k1, k1, l1)) return pos;
    pos = (pos + d) & (Q1_MAXG - 1);
  }
}
int oracle_q1_batches(const sd_batch* b, int nb, int32_t cutoff, q1_group* groups /* [Q1_MAXG], zeroed by caller */) {
  for (int i = 0; i < nb; i++) {
    const uint8_t* qty = body_notnull(b[i].col_bufs[0]);
    const uint8_t* price = body_notnull(b[i].col_bufs[1]);
    const uint8_t* disc = body_notnull(b[i].col_bufs[2]);
    const uint8_t* tax = body_notnull(b[i].col_bufs[3]);
    coldec rf, ls;
    if (coldec_init(&rf, (const uint8_t*)b[i].col_bufs[4], b[i].col_lens[4], SD_STRING, 0, 0)) return -1;
    if (coldec_init(&ls, (const uint8_t*)b[i].col_bufs[5], b[i].col_lens[5], SD_STRING, 0, 0)) return -1;
    if (rf.dict_n > 64 || ls.dict_n > 64 || rf.type_id != 2 || ls.type_id != 2) return fail(SD_ERR_UNSUPPORTED, "q1 restatement expects small int16 dictionaries");
    int slot[64 * 64];
    for (int x = 0; x < rf.dict_n * ls.dict_n; x++) slot[x] = -1;
    const uint8_t* ship = body_notnull(b[i].col_bufs[6]);
    int n = b[i].num_rows;
    for (int o = 0; o < n; o++) {
      int32_t sd = ld_i32(ship + ((int64_t)o << 2));
      if (!(sd <= cutoff)) continue;
      int i0 = ld_i16(rf.body + ((int64_t)o << 1)), i1 = ld_i16(ls.body + ((int64_t)o << 1));
      int s = slot[i0 * ls.dict_n + i1];
      if (s < 0) s = slot[i0 * ls.dict_n + i1] = q1_find(groups, rf.dict_s[i0], rf.dict_slen[i0], ls.dict_s[i1], ls.dict_slen[i1]);
      q1_group* g = &groups[s];
      double qv = ld_f64(qty + ((int64_t)o << 3)), pv = ld_f64(price + ((int64_t)o << 3));
      double dv = ld_f64(disc + ((int64_t)o << 3)), tv = ld_f64(tax + ((int64_t)o << 3));
      double dp = pv * (1.0 - dv);
      g->sum_qty += qv; g->sum_price += pv; g->sum_disc_price += dp; g->sum_charge += dp * (1.0 + tv);
      g->avg_qty_s += qv; g->avg_qty_c++; g->avg_price_s += pv; g->avg_price_c++; g->avg_disc_s += dv; g->avg_disc_c++;
      g->count++;
    }
    coldec_free(&rf); coldec_free(&ls);
  }
  return 0;
}

/* one thread per partition, like Spark local[N] (one task per partition); partial results merged in
 * partition order afterwards.  kind: 1 = C1, 6 = Q6, 11 = Q1                                    */
typedef struct part_task {
  int kind; const sd_batch* b; int nb;
  int32_t i0, i1; double f0, f1, f2;
  double sum; int sum_isnull; int64_t count; q1_group groups[Q1_MAXG]; int rc;
} part_task;
static void* part_main(void* a) {
  part_task* t = (part_task*)a;
  if (t->kind == 6) t->rc = oracle_q6_batches(t->b, t->nb, t->i0, t->i1, t->f0, t->f1, t->f2, &t->sum, &t->sum_isnull, &t->count);
  else if (t->kind == 1) t->rc = oracle_c1_batches(t->b, t->nb, t->i0, &t->count);
  else { memset(t->groups, 0, sizeof(t->groups)); t->rc = oracle_q1_batches(t->b, t->nb, t->i0, t->groups); }
  return NULL;
}
/* batches are dealt to nthreads partitions in contiguous ranges; out: Q6 -> out_f[0] = sum,
 * out_i[0] = isnull, out_i[1] = matched; C1 -> out_i[0] = count; Q1 -> q1_out[Q1_MAXG] merged groups */
int oracle_run_partitions(int kind, const sd_batch* b, int nb, int nthreads, int32_t i0, int32_t i1,
                          double f0, double f1, double f2, double* out_f, int64_t* out_i, q1_group* q1_out) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > nb && nb > 0) nthreads = nb;
  part_task* t = (part_task*)calloc(nthreads, sizeof(part_task));
  pthread_t* th = (pthread_t*)calloc(nthreads, sizeof(pthread_t));
  for (int i = 0; i < nthreads; i++) {
    int lo = (int)((int64_t)nb * i / nthreads), hi = (int)((int64_t)nb * (i + 1) / nthreads);
    t[i].kind = kind; t[i].b = b + lo; t[i].nb = hi - lo; t[i].i0 = i0; t[i].i1 = i1; t[i].f0 = f0; t[i].f1 = f1; t[i].f2 = f2;
    if (nthreads == 1) part_main(&t[i]); else pthread_create(&th[i], NULL, part_main, &t[i]);
  }
  if (nthreads > 1) for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
  int rc = 0;
  if (kind == 6) {
    double s = 0.0; int isn = 1; int64_t m = 0;
    for (int i = 0; i < nthreads; i++) { rc |= t[i].rc; if (!t[i].sum_isnull) { s = (isn ? 0.0 : s) + t[i].sum; isn = 0; } m += t[i].count; }
    out_f[0] = s; out_i[0] = isn; out_i[1] = m;
  } else if (kind == 1) {
    int64_t c = 0; for (int i = 0; i < nthreads; i++) { rc |= t[i].rc; c += t[i].count; } out_i[0] = c;
  } else {
    memset(q1_out, 0, sizeof(q1_group) * Q1_MAXG);
    for (int i = 0; i < nthreads; i++) {
      rc |= t[i].rc;
      for (int g = 0; g < Q1_MAXG; g++) if (t[i].groups[g].used) {
        q1_group* s = &t[i].groups[g];
        q1_group* d = &q1_out[q1_find(q1_out, s->k0, s->l0, s->k1, s->l1)];
        d->sum_qty += s->sum_qty; d->sum_price += s->sum_price; d->sum_disc_price += s->sum_disc_price; d->sum_charge += s->sum_charge;
        d->avg_qty_s += s->avg_qty_s; d->avg_qty_c += s->avg_qty_c; d->avg_price_s += s->avg_price_s; d->avg_price_c += s->avg_price_c;
        d->avg_disc_s += s->avg_disc_s; d->avg_disc_c += s->avg_disc_c; d->count += s->count;
      }
    }
  }
  free(t); free(th);
  return rc;
}
int oracle_q1_group_size(void) { return (int)sizeof(q1_group); }
int oracle_q1_max_groups(void) { return Q1_MAXG; }
