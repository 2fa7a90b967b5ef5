/*
 * snappy_gpu.h -- C ABI of libsnappygpu.so, the B200-native replacement for the inside of
 * SnappyData's partial-aggregation stage:
 *
 *     ColumnTableScan -> [FilterExec / ProjectExec] -> SnappyHashAggregateExec(Partial)
 *
 * These entry points are what a JNI shim on the reference side binds (INTEGRATION.md shows it).
 * Plain pointers and sizes only; no C++ / torch types.  Every function returns 0 on success and a
 * non-zero sd_status otherwise; sd_last_error() then holds a thread-local message.  There is NO
 * CPU fallback: a plan or buffer the GPU path cannot execute is an error (the reference's
 * CodegenSparkFallback, core/.../execution/CodegenSparkFallback.scala:48-134, is deliberately not
 * mirrored; BASELINE.json north_star).
 *
 * Citations are to /root/reference; enc = encoders/src/main/scala/org/apache/spark/sql/execution/
 * columnar/encoding, core = core/src/main/scala/org/apache/spark/sql.
 *
 * Threading: one thread drives one sd_plan from create to destroy; distinct plans may
 * be driven concurrently from distinct threads; an sd_store may be shared by plans (reads only while
 * plans scan it).
 */
#ifndef SNAPPY_GPU_H
#define SNAPPY_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SD_ABI_VERSION 2

typedef enum sd_status {
  SD_OK = 0,
  SD_ERR_INVALID = 1,      /* malformed descriptor / buffer                         */
  SD_ERR_UNSUPPORTED = 2,  /* plan shape or encoding the GPU path does not execute  */
  SD_ERR_CUDA = 3,         /* CUDA runtime / driver / NVRTC failure                 */
  SD_ERR_OVERFLOW = 4,     /* caller's output buffer too small (outLen = needed)    */
  SD_ERR_STATE = 5         /* call sequence violation                               */
} sd_status;

/* SQL types of scan columns / expression nodes (Catalyst DataType of the attribute;
 * core/execution/columnar/ColumnTableScan.scala:684-760 picks the decoder read method from it). */
typedef enum sd_type {
  SD_BOOLEAN = 1, SD_BYTE = 2, SD_SHORT = 3, SD_INT = 4, SD_LONG = 5, SD_FLOAT = 6, SD_DOUBLE = 7,
  SD_DATE = 8,       /* int32 days since epoch       */
  SD_TIMESTAMP = 9,  /* int64 microseconds           */
  SD_STRING = 10,    /* UTF8String                   */
  SD_DECIMAL = 11    /* column values: precision <= 18, int64 unscaled (enc/Uncompressed.scala:95-98); aggregate
                        buffers / results may be wider (SUM: DECIMAL(p+10,s), up to 128-bit unscaled), see sd_agg */
} sd_type;

/* One projected scan column (ColumnTableScan.output attribute). */
typedef struct sd_column {
  int32_t type;           /* sd_type                                                  */
  int32_t nullable;       /* field.nullable: selects Nullable vs NotNull decoder
                             (enc/ColumnEncoding.scala:817-822)                       */
  int32_t table_ordinal;  /* 0-based column of the table (ColumnFormatKey.columnIndex-1) */
  int32_t scale;          /* SD_DECIMAL scale; else 0                                 */
  int32_t precision;      /* SD_DECIMAL precision (1..18 for a scan column); else 0   */
} sd_column;

/* Expression tree, flattened; children always precede parents.  Mirrors the Catalyst trees that
 * FilterExec / ProjectExec / the aggregate functions' children hold.
 * Semantics restated from Spark 2.1.1: SQL three-valued logic, NULL in any
 * operand of arithmetic/comparison => NULL, x / 0 => NULL, integral arithmetic wraps,
 * FLOAT/DOUBLE comparisons use the NaN-safe total order (NaN == NaN, NaN greatest, -0.0 == 0.0),
 * strings compare as unsigned bytes. */
typedef enum sd_op {
  SD_OP_COL = 1,        /* a = index into sd_plan_desc.cols                      */
  SD_OP_LIT = 2,        /* a = literal slot (runtime value, cf. ParamLiteral,
                           core/catalyst/expressions/ParamLiteral.scala:43-110)  */
  SD_OP_ADD = 10, SD_OP_SUB = 11, SD_OP_MUL = 12, SD_OP_DIV = 13, SD_OP_NEG = 14,
  SD_OP_CAST = 15,      /* a -> node type                                        */
  SD_OP_EQ = 20, SD_OP_NE = 21, SD_OP_LT = 22, SD_OP_LE = 23, SD_OP_GT = 24, SD_OP_GE = 25,
  SD_OP_AND = 30, SD_OP_OR = 31, SD_OP_NOT = 32,
  SD_OP_ISNULL = 33, SD_OP_ISNOTNULL = 34,
  SD_OP_IN = 35,        /* a = expr, b = first literal slot, c = number of literals */
  SD_OP_STARTSWITH = 36 /* a = string expr, b = literal node                     */
} sd_op;

typedef struct sd_expr {
  int32_t op;    /* sd_op                     */
  int32_t type;  /* sd_type of the result     */
  int32_t a, b, c;   /* SD_DECIMAL-typed LIT / CAST nodes: c = (precision << 8) | scale of the node's type (a COL node takes
                        them from its column, NEG from its child; DECIMAL literal values are unscaled at that scale) */
} sd_expr;
#define SD_DEC_PS(precision, scale) (((precision) << 8) | (scale))

/* CAST follows Spark 2.1.1 Cast for the pairs the GPU path executes; every other pair is SD_ERR_UNSUPPORTED:
 *   integral <-> integral (wraps), integral/fp -> fp, fp -> integral (Java (int)/(long): NaN -> 0, saturating),
 *   numeric -> BOOLEAN (v != 0), BOOLEAN -> numeric (1 / 0), DECIMAL(p,s) -> DOUBLE/FLOAT (unscaled / 10^s),
 *   integral -> DECIMAL(p,s) (v * 10^s, NULL when it does not fit p digits), DECIMAL(p1,s1) -> DECIMAL(p2,s2 >= s1)
 *   (rescale, NULL when it does not fit).  Casts involving STRING, DATE or TIMESTAMP (other than the identity),
 *   DECIMAL -> integral and down-scaling DECIMAL casts are refused. */

/* Aggregate functions (Spark DeclarativeAggregate) and their partial
 * buffer fields, in the order SnappyHashAggregateExec lays them out
 * (core/execution/aggregate/SnappyHashAggregateExec.scala:174-210,456-471):
 *   COUNT_STAR / COUNT : [count LONG]
 *   SUM                : [sum LONG (integral input) | DOUBLE (float/double input)]   (nullable)
 *   AVG                : [sum DOUBLE, count LONG]
 *   MIN / MAX          : [value of the input type]                                   (nullable)
 * DECIMAL(p,s) input (Spark 2.1.1 Sum / Average): SUM buffer and result DECIMAL(p+10,s); AVG buffers
 * [sum DECIMAL(p+10,s), count LONG], result DECIMAL(p+4,s+4) = sum / count rounded HALF_UP.  In UnsafeRows a DECIMAL
 * of precision <= 18 is its unscaled int64 in the fixed slot; wider ones are (offset << 32 | size) + the
 * BigInteger two's-complement big-endian bytes in a 16-byte reserved region (UnsafeRowWriter.write(Decimal)). */
typedef enum sd_agg_fn {
  SD_AGG_COUNT_STAR = 1, SD_AGG_COUNT = 2, SD_AGG_SUM = 3, SD_AGG_AVG = 4, SD_AGG_MIN = 5, SD_AGG_MAX = 6
} sd_agg_fn;

typedef struct sd_agg {
  int32_t fn;    /* sd_agg_fn                               */
  int32_t expr;  /* input expression node; -1 for COUNT(*)  */
} sd_agg;

/* The fused plan: scan columns -> filter -> (group keys, aggregates) | projection. */
typedef struct sd_plan_desc {
  int32_t abi_version;          /* SD_ABI_VERSION                                        */
  int32_t ncols;   const sd_column* cols;
  int32_t nexprs;  const sd_expr* exprs;
  int32_t filter;               /* root node of the FilterExec condition, or -1          */
  int32_t nkeys;   const int32_t* keys;    /* grouping expressions (node indexes)        */
  int32_t naggs;   const sd_agg* aggs;
  int32_t nproj;   const int32_t* proj;    /* naggs == 0 && nkeys == 0: output columns   */
  int32_t nliterals; const int32_t* literal_types;   /* sd_type per literal slot         */
  int32_t flags;                /* reserved, 0                                           */
} sd_plan_desc;

typedef struct sd_literal {
  int32_t type;      /* sd_type                              */
  int32_t is_null;
  int64_t i;         /* integral / date / timestamp / boolean / decimal-unscaled value */
  double  d;         /* FLOAT / DOUBLE value                 */
  const char* s;     /* STRING bytes (not NUL terminated)    */
  int32_t slen;
  int32_t pad_;
} sd_literal;

/* One column batch as ColumnBatchIterator serves it to the generated loop
 * (core/execution/columnar/ColumnBatchIterator.scala:53-231): per projected column the value
 * buffer (getColumnLob) and up to two update deltas (getUpdatedColumnDecoder, depth 0 and 1), the
 * delete mask (getDeletedColumnDecoder), the stats row (next()), ids.  Arrays are indexed like
 * sd_plan_desc.cols.  Buffers may be heap or direct memory; the library has finished reading (or
 * copied) them when sd_batch_submit returns (ownership rule).  A buffer whose first
 * int32 is negative is a compressed envelope (encoders/.../store/CompressionUtils.scala:53-61): LZ4 (-1)
 * envelopes are accepted -- only the compressed bytes are copied and the block is expanded on the device (run-length
 * and variable-width STRING bodies, whose layout needs a host walk, are expanded on the host instead); Snappy (-2)
 * envelopes, compressed update deltas and compressed delete masks are decompressed on the host, as the reference's own
 * iterator does (ColumnBatchIterator.scala:102-113). */
typedef struct sd_batch {
  int32_t num_rows;
  int32_t ncols;
  const void* const* col_bufs;  const int64_t* col_lens;
  const void* const* delta0;    const int64_t* delta0_lens;   /* may be NULL; entries may be NULL */
  const void* const* delta1;    const int64_t* delta1_lens;
  const void* delete_buf;       int64_t delete_len;           /* may be NULL                      */
  const void* stats_row;        int64_t stats_len;            /* may be NULL (no batch skipping)  */
  int32_t stats_ncols;          /* number of table columns described by the stats row            */
  int32_t bucket_id;
  int64_t batch_id;
} sd_batch;

typedef struct sd_plan sd_plan;
typedef struct sd_store sd_store;

/* ---- lifecycle -------------------------------------------------------------------------------- */
int sd_init(int device);                       /* bind the calling thread's plans to a GPU          */
int sd_device_count(int* out);
const char* sd_last_error(void);               /* thread-local, valid until the next failing call   */
const char* sd_version(void);

/* page-locked host memory for a binding's staging area (the JNI shim copies heap byte[]s into it, so that no Java array
 * is pinned while CUDA work is queued and host->device copies out of it are real asynchronous DMA) */
int sd_host_alloc(int64_t bytes, void** out);
void sd_host_free(void* p);
/* page-lock memory the caller already owns and keeps for a long time (the region's off-heap column buffers): copies out of
 * it become asynchronous DMA at link speed instead of the driver's staged pageable copies (bench.py: e2e_pageable_unretained
 * is ~4x below the pinned legs) */
int sd_host_register(void* p, int64_t bytes);
int sd_host_unregister(void* p);

/* ---- plan (one per Spark task / partition; ColumnTableScan.doProduce + SnappyHashAggregateExec
 *      doProduce/doConsume fused, core/.../ColumnTableScan.scala:186-672,
 *      core/.../aggregate/SnappyHashAggregateExec.scala:240-263) --------------------------------- */
int sd_plan_create(const sd_plan_desc* desc, sd_plan** out);
int sd_plan_set_literals(sd_plan* p, const sd_literal* vals, int32_t n);
/* scan one batch from host buffers (copied to the device inside the call) */
int sd_batch_submit(sd_plan* p, const sd_batch* b);
/* row-buffer rows of the hybrid scan (core/execution/row/RowFormatScanRDD.scala; consumed through the
 * same loop with batch size 1, ColumnTableScan.scala:572-588): nrows UnsafeRows of the plan's
 * scan columns, each prefixed by its int64 size */
int sd_rows_submit(sd_plan* p, const void* rows, int64_t len, int32_t nrows);
/* finish the partition: run what is pending, emit partial-aggregate rows (or projected rows) as
 * repeated [int64 sizeInBytes][UnsafeRow(group keys ++ aggregate buffers)].  On SD_ERR_OVERFLOW
 * *out_len is the size needed and the call may be repeated. */
int sd_plan_finish(sd_plan* p, void* out_rows, int64_t cap, int64_t* out_len, int64_t* out_nrows);
/* make the handle reusable for another execution of the same (cached) plan */
int sd_plan_reset(sd_plan* p);
/* SQLMetrics of the two operators (ColumnTableScan.scala:111-127, SnappyHashAggregateExec.scala:132-137):
 * [0] numOutputRows (aggregate) [1] numRowsBuffer [2] columnBatchesSeen [3] updatedColumnCount
 * [4] deletedBatchCount [5] columnBatchesSkipped [6] aggTime (device ns) [7] kernel launches
 * [8] rows scanned [9] algorithmic bytes scanned [10] host->device bytes [11] scan numOutputRows */
#define SD_NUM_METRICS 12
int sd_plan_metrics(sd_plan* p, int64_t out[SD_NUM_METRICS]);
/* options.  SD_OPT_RETAIN_BUFFERS = 1: the caller keeps every buffer passed to sd_batch_submit alive and
 * unchanged until sd_plan_finish returns (e.g. ref-counted direct ByteBuffers retained by the operator); host->device
 * copies are then queued without a per-batch synchronisation.  Default 0: the reference's ownership rule (buffers
 * may be released when sd_batch_submit returns, ColumnBatchIterator.scala:165-184). */
#define SD_OPT_RETAIN_BUFFERS 1
int sd_plan_set_option(sd_plan* p, int32_t option, int64_t value);
/* run the plan's kernels on a caller-owned CUDA stream (cudaStream_t as void*), e.g. torch's */
int sd_plan_set_stream(sd_plan* p, void* cuda_stream);
/* kernel variant actually selected for the plan ("aot:<signature>" | "jit:<signature>") */
const char* sd_plan_kernel_name(sd_plan* p);
void sd_plan_destroy(sd_plan* p);

/* ---- device-resident column store (residency policy of this engine; the reference keeps batches
 *      in region memory and faults them in per scan, ColumnBatchIterator.scala:179-223) ---------- */
/* schema = the table's columns in table order (ColumnFormatRelation.schema); type and nullability
 * select the decoders exactly as field.dataType / field.nullable do in the reference */
int sd_store_create(int device, int32_t ncols, const sd_column* schema, sd_store** out);
/* upload one batch; b->col_bufs is indexed by TABLE column here (ncols = table width; NULL entries
 * for columns never scanned); delta arrays likewise */
int sd_store_put_batch(sd_store* s, const sd_batch* b);
/* ---- ingest: ColumnBatch creation ON THE DEVICE.  Raw column values of one batch (what the reference's
 *      generated insert loop feeds its ColumnEncoders row by row, core/.../columnar/ColumnInsertExec.scala:326-822) are copied to
 *      the device once and encoded there into the reference's column buffers -- Uncompressed / Dictionary (first-seen
 *      order, int16 -> int32 indexes at 32767 entries) / BooleanBitSet with trimmed null words, the default encoder choice
 *      of enc/ColumnEncoding.scala:837-844 -- plus the stats row (lower / upper bound, null count per column;
 *      ColumnInsertExec.scala:848-921).  The batch is resident and scannable when the call returns; its bytes are the ones
 *      snappydata_b200/column_format.py writes for the same values (sdx_store_get_buffer / sdx_store_get_stats read them
 *      back).  cols[c] describes TABLE column c (values == NULL: not materialised). ---------------------------------- */
typedef struct sd_raw_column {
  const void* values;        /* num_rows values: BOOLEAN / BYTE 1 byte, SHORT 2, INT / DATE / FLOAT 4, LONG / TIMESTAMP / DOUBLE /
                                DECIMAL (unscaled) 8; STRING: int32 offsets[num_rows + 1] into str_bytes                  */
  const uint8_t* str_bytes;  /* STRING: the values' bytes back to back                                                */
  const uint8_t* nulls;      /* optional, 1 byte per row, non-zero = NULL; must be NULL for a NOT NULL column          */
} sd_raw_column;
int sd_store_encode_batch(sd_store* s, int32_t num_rows, const sd_raw_column* cols, int32_t ncols,
                          int32_t bucket_id, int64_t batch_id);
/* ColumnDeltaEncoder.merge (enc/ColumnDeltaEncoder.scala:348-556), host only: the new update delta of a column merged with what
 * the table holds -- another delta (existing_is_delta = 1: union of positions, the new one wins on equal positions, result is a
 * delta) or the full column of num_rows rows (existing_is_delta = 0: the delta folded into the column, result is a column
 * buffer) -- re-encoded with the type's default encoder.  Either input may be a compressed envelope. */
int sd_delta_merge(const sd_column* column, const void* new_delta, int64_t new_len, const void* existing, int64_t existing_len,
                   int32_t existing_is_delta, int32_t num_rows, void* out, int64_t cap, int64_t* out_len);
int sd_store_num_batches(sd_store* s, int64_t* out);
int sd_store_bytes(sd_store* s, int64_t* out);
/* scan every resident batch of the given buckets (NULL/0 = all) with plan p: stats-row skipping on
 * the host, then the fused kernels over the resident bytes; results are collected by sd_plan_finish */
int sd_plan_scan_store(sd_plan* p, sd_store* s, const int32_t* bucket_ids, int32_t nbuckets);
void sd_store_destroy(sd_store* s);

/* ---- final merge (SnappyHashAggregateExec(Final) / CollectAggregateExec.executeCollect,
 *      core/execution/aggregate/CollectAggregateExec.scala:67-121): merges partial rows of all
 *      partitions (sums add, counts add, min/max combine) and evaluates results (avg = sum/count).
 *      Host-side: payload is a handful of rows.  Output rows: keys ++ one result per aggregate. --- */
int sd_final_merge(const sd_plan_desc* desc, const void* partial_rows, int64_t len,
                   void* out_rows, int64_t cap, int64_t* out_len, int64_t* out_nrows);

/* the same merge reusing the analysis held by a plan handle (no per-call plan analysis) */
int sd_plan_final_merge(sd_plan* p, const void* partial_rows, int64_t len,
                        void* out_rows, int64_t cap, int64_t* out_len, int64_t* out_nrows);

/* partial rows of several partitions -> ONE merged set of partial rows (same schema; a combiner in front of the final
 * stage).  sd_plan_exchange uses it after gathering every rank's rows. */
int sd_partial_merge(const sd_plan_desc* desc, const void* partial_rows, int64_t len,
                     void* out_rows, int64_t cap, int64_t* out_len, int64_t* out_nrows);   /* host only */
int sd_plan_partial_merge(sd_plan* p, const void* partial_rows, int64_t len,
                          void* out_rows, int64_t cap, int64_t* out_len, int64_t* out_nrows);

/* ---- the cross-partition exchange: partial -> Exchange -> final as SnappyStrategies plans it
 *      (core/.../SnappyStrategies.scala:566-604).  One partition (= one sd_plan on one GPU) per rank; the exchange is ONE
 *      ncclAllGather over NVLink of every rank's partial rows BY VALUE, merged on every rank.  NCCL is dlopen'ed
 *      (libnccl.so.2, or $SD_NCCL_LIB); the caller only transports the 128-byte unique id from rank 0 to the others
 *      (torch.distributed / Spark broadcast / any RPC). ------------------------------------------------------------ */
typedef struct sd_comm sd_comm;
#define SD_COMM_ID_BYTES 128
int sd_comm_unique_id(void* out_id);                         /* rank 0 */
int sd_comm_create(const void* id, int32_t rank, int32_t world, int32_t device, sd_comm** out);   /* collective */
void sd_comm_destroy(sd_comm* c);
/* [0] world [1] bytes per rank of the gather slot [2] all-gathers issued [3] times the slot had to grow */
int sd_comm_info(sd_comm* c, int64_t out[4]);
/* collective, after this execution's scans: gathers + merges; sd_plan_finish then returns the MERGED partial rows
 * (identical on every rank; any number of groups -- the gather slot grows in lock step on all ranks) */
int sd_plan_exchange(sd_plan* p, sd_comm* c);
/* one execution of a cached plan over a resident store in one call:
 * reset -> set_literals -> scan_store -> [exchange when comm != NULL] -> finish */
int sd_plan_execute_store(sd_plan* p, sd_store* s, const int32_t* bucket_ids, int32_t nbuckets,
                          const sd_literal* lits, int32_t nlits, sd_comm* comm,
                          void* out_rows, int64_t cap, int64_t* out_len, int64_t* out_nrows);

/* ---- export of the dense partial table for an on-device exchange (NCCL all-reduce over NVLink of
 *      per-GPU partials).  Writes nslots int64/double words per group into dev_out
 *      (device pointer) on the plan's stream.  Only for plans without string/hash keys. ---------- */
int sd_plan_partials_layout(sd_plan* p, int32_t* ngroups, int32_t* nslots, int32_t* slot_is_f64);
int sd_plan_export_partials(sd_plan* p, void* dev_out, int64_t cap_bytes);
int sd_plan_import_partials(sd_plan* p, const void* dev_in, int64_t bytes);

/* ---- synthetic lineitem tables generated on the device straight into a store (bench/test
 *      utility; byte-identical to snappydata_b200/lineitem.py; not part of the reference boundary) */
int sdx_store_gen_lineitem(sd_store* s, int64_t first_row, int64_t nrows, int32_t rows_per_batch,
                           int32_t nbuckets, uint64_t seed, int32_t column_mask);
/* copy a resident buffer back to the host (tests: device generator == host generator) */
int sdx_store_get_buffer(sd_store* s, int64_t batch_index, int32_t table_col, void* out, int64_t cap,
                         int64_t* out_len);
/* host LZ4 prefix decoder used to lay out compressed column buffers (test hook) */
int64_t sdx_lz4_decode_prefix(const void* src, int64_t src_len, void* dst, int64_t want);
/* host decompression of a stored envelope [-codecId][uncompressedLen][payload] (LZ4 = 1, Snappy = 2), as the engine
 * applies it to update deltas, delete masks and Snappy column buffers (test hook; no CUDA call) */
int sdx_decompress_envelope(const void* buf, int64_t len, void* out, int64_t cap, int64_t* out_len);
/* stats row (UnsafeRow) of a resident batch */
int sdx_store_get_stats(sd_store* s, int64_t batch_index, void* out, int64_t cap, int64_t* out_len);
int sdx_store_batch_info(sd_store* s, int64_t batch_index, int32_t* num_rows, int32_t* bucket_id,
                         int64_t* batch_id);
/* the batch-skipping decision (ColumnTableScan.scala:820-963) of a plan's filter for one stats row: *pass = 0 when the
 * batch would be skipped.  Host only, no CUDA call (test hook: tests/test_stats_predicate.py compares it with the oracle). */
int sdx_stats_pass(const sd_plan_desc* desc, const sd_literal* lits, int32_t nlits, const void* stats,
                   int64_t stats_len, int32_t stats_ncols, int32_t num_rows, int32_t* pass);
/* expand n raw LZ4 blocks with the engine's device kernel (bench/test hook used by tools/lz4_bench.py): uploads the
 * blocks, places output i at a 16-byte boundary + dst_misalign, runs `reps` launches timed with CUDA events
 * (ms_per_launch = their mean) and copies output i to outs[i] when outs != NULL.  `dense` selects the kernel variant:
 * bit 0 = the denser shape (SD_TUNE_LZ4_DENSE), bit 1 = the window parse (SD_TUNE_LZ4_PARSE).  A corrupt block ->
 * SD_ERR_INVALID. */
int sdx_lz4_expand(int32_t device, const void* const* blocks, const int64_t* block_lens, const int64_t* out_lens,
                   int32_t n, int32_t dst_misalign, int32_t dense, int32_t reps, void* const* outs,
                   double* ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* SNAPPY_GPU_H */
