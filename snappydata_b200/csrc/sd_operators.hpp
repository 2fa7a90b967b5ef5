// sd_operators.hpp -- C++ host-side mirror of the reference's operator surface for the hot path, layered
// on the C ABI (include/snappy_gpu.h).  The reference's host side is Scala/JVM; no JVM toolchain exists in
// this build environment, so the mirror is C++ (the JVM binding a maintainer would use is under jvm/).
// Names, argument meaning and error behaviour follow the reference so that tests read like its own:
//
//   ColumnBatch / ColumnBatchIterator   encoders/.../columnar/ColumnBatch.scala:36-50 (region-less batch),
//                                       core/execution/columnar/ColumnBatchIterator.scala:36-50,122-163
//   ColumnTableScan                     core/execution/columnar/ColumnTableScan.scala:67-87 (output, partition scan)
//   FilterExec condition                StoreDataSourceStrategy.scala:128-130 (unhandledFilters = all)
//   SnappyHashAggregateExec             core/execution/aggregate/SnappyHashAggregateExec.scala:72-80
//   CollectAggregateExec.executeCollect core/execution/aggregate/CollectAggregateExec.scala:67-121
//
// Errors: every failing C-ABI call throws std::runtime_error(sd_last_error()) -- the counterpart of the
// IOException the generated loop throws (ColumnTableScan.scala:662-668); there is no CPU fallback.
#ifndef SD_OPERATORS_HPP
#define SD_OPERATORS_HPP

#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/snappy_gpu.h"

namespace snappy {

inline void check(int rc) {
  if (rc != 0) throw std::runtime_error(std::string("[sd_status ") + std::to_string(rc) + "] " + sd_last_error());
}

struct StructField {
  std::string name;
  sd_type dataType;
  bool nullable;
  int tableOrdinal;   // ColumnFormatKey.columnIndex - 1
};

// One column batch as the scan sees it.
struct ColumnBatch {
  int numRows = 0;
  std::vector<std::vector<uint8_t>> buffers;          // by table column; empty = not present
  std::vector<uint8_t> statsData;
  std::vector<std::vector<uint8_t>> delta0, delta1;   // by table column; empty = none
  std::vector<uint8_t> deleteMask;
  int64_t batchId = 0;
  int bucketId = 0;
};

// Iteration contract of core/execution/columnar/ColumnBatchIterator.scala (region-less mode).
class ColumnBatchIterator {
 public:
  explicit ColumnBatchIterator(const std::vector<ColumnBatch>* batches) : batches_(batches) {}
  bool hasNext() const { return pos_ + 1 < (int)batches_->size(); }
  const std::vector<uint8_t>& next() { return (*batches_)[++pos_].statsData; }   // -> stats row buffer
  const std::vector<uint8_t>& getColumnLob(int tableColumn) const { return cur().buffers[tableColumn]; }
  const std::vector<uint8_t>* getUpdatedColumnBuffer(int tableColumn, int depth) const {
    const auto& d = depth == 0 ? cur().delta0 : cur().delta1;
    return (tableColumn < (int)d.size() && !d[tableColumn].empty()) ? &d[tableColumn] : nullptr;
  }
  const std::vector<uint8_t>* getDeletedColumnBuffer() const { return cur().deleteMask.empty() ? nullptr : &cur().deleteMask; }
  int getDeletedRowCount() const {   // int at offset 8 of the delete buffer (ColumnBatchIterator.scala:151-163)
    if (cur().deleteMask.size() < 12) return 0;
    int32_t n; memcpy(&n, cur().deleteMask.data() + 8, 4); return n;
  }
  int64_t getCurrentBatchId() const { return cur().batchId; }
  int getCurrentBucketId() const { return cur().bucketId; }
  int numRows() const { return cur().numRows; }
  void close() { pos_ = (int)batches_->size(); }

 private:
  const ColumnBatch& cur() const { return (*batches_)[pos_]; }
  const std::vector<ColumnBatch>* batches_;
  int pos_ = -1;
};

// ---- Catalyst-like expression trees (flattened on demand) --------------------------------------------
class PlanBuilder;
struct Expr {
  int node = -1;
  sd_type type = SD_INT;
};

class PlanBuilder {
 public:
  Expr attr(const StructField& f) {
    cols_.push_back(sd_column{f.dataType, f.nullable ? 1 : 0, f.tableOrdinal, 0});
    return add(SD_OP_COL, f.dataType, (int)cols_.size() - 1, 0, 0);
  }
  Expr literal(sd_type t) {   // tokenised constant: value supplied per execution (ParamLiteral)
    lit_types_.push_back(t);
    return add(SD_OP_LIT, t, (int)lit_types_.size() - 1, 0, 0);
  }
  Expr binary(int op, Expr a, Expr b, sd_type t) { return add(op, t, a.node, b.node, 0); }
  Expr unary(int op, Expr a, sd_type t) { return add(op, t, a.node, 0, 0); }
  Expr And(Expr a, Expr b) { return binary(SD_OP_AND, a, b, SD_BOOLEAN); }
  Expr Or(Expr a, Expr b) { return binary(SD_OP_OR, a, b, SD_BOOLEAN); }
  Expr cmp(int op, Expr a, Expr b) { return binary(op, a, b, SD_BOOLEAN); }
  Expr Multiply(Expr a, Expr b) { return binary(SD_OP_MUL, a, b, a.type); }
  Expr Add(Expr a, Expr b) { return binary(SD_OP_ADD, a, b, a.type); }
  Expr Subtract(Expr a, Expr b) { return binary(SD_OP_SUB, a, b, a.type); }

  std::vector<sd_column> cols_;
  std::vector<sd_expr> exprs_;
  std::vector<int32_t> lit_types_;

 private:
  Expr add(int op, sd_type t, int a, int b, int c) {
    exprs_.push_back(sd_expr{op, t, a, b, c});
    Expr e; e.node = (int)exprs_.size() - 1; e.type = t;
    return e;
  }
};

struct AggregateExpression { sd_agg_fn fn; Expr child; };   // child.node == -1 for COUNT(*)

// ColumnTableScan + FilterExec + SnappyHashAggregateExec(Partial), fused.
class SnappyHashAggregateExec {
 public:
  SnappyHashAggregateExec(PlanBuilder& b, Expr* filterCondition, const std::vector<Expr>& groupingExpressions,
                          const std::vector<AggregateExpression>& aggregateExpressions)
      : builder_(b) {
    for (auto& k : groupingExpressions) keys_.push_back(k.node);
    for (auto& a : aggregateExpressions) aggs_.push_back(sd_agg{a.fn, a.fn == SD_AGG_COUNT_STAR ? -1 : a.child.node});
    memset(&desc_, 0, sizeof(desc_));
    desc_.abi_version = SD_ABI_VERSION;
    desc_.ncols = (int)b.cols_.size(); desc_.cols = b.cols_.data();
    desc_.nexprs = (int)b.exprs_.size(); desc_.exprs = b.exprs_.data();
    desc_.filter = filterCondition ? filterCondition->node : -1;
    desc_.nkeys = (int)keys_.size(); desc_.keys = keys_.data();
    desc_.naggs = (int)aggs_.size(); desc_.aggs = aggs_.data();
    desc_.nliterals = (int)b.lit_types_.size(); desc_.literal_types = b.lit_types_.data();
    check(sd_plan_create(&desc_, &plan_));
  }
  ~SnappyHashAggregateExec() { sd_plan_destroy(plan_); }
  SnappyHashAggregateExec(const SnappyHashAggregateExec&) = delete;

  const sd_plan_desc& desc() const { return desc_; }
  std::string nodeName() const { return "SnappyHashAggregate"; }

  // One task: consume a partition's batches, return the partial rows
  // ([int64 size][UnsafeRow(groupKeys ++ aggBuffers)] ...).
  std::vector<uint8_t> executePartition(ColumnBatchIterator& it, const std::vector<sd_literal>& literals) {
    check(sd_plan_reset(plan_));
    check(sd_plan_set_literals(plan_, literals.data(), (int32_t)literals.size()));
    const int nc = desc_.ncols;
    while (it.hasNext()) {
      const std::vector<uint8_t>& stats = it.next();
      std::vector<const void*> bufs(nc), d0(nc), d1(nc);
      std::vector<int64_t> lens(nc), d0l(nc), d1l(nc);
      for (int c = 0; c < nc; c++) {
        const int t = builder_.cols_[c].table_ordinal;
        const std::vector<uint8_t>& lob = it.getColumnLob(t);
        bufs[c] = lob.data(); lens[c] = (int64_t)lob.size();
        const std::vector<uint8_t>* u0 = it.getUpdatedColumnBuffer(t, 0);
        const std::vector<uint8_t>* u1 = it.getUpdatedColumnBuffer(t, 1);
        d0[c] = u0 ? u0->data() : nullptr; d0l[c] = u0 ? (int64_t)u0->size() : 0;
        d1[c] = u1 ? u1->data() : nullptr; d1l[c] = u1 ? (int64_t)u1->size() : 0;
      }
      sd_batch b;
      memset(&b, 0, sizeof(b));
      b.num_rows = it.numRows(); b.ncols = nc; b.col_bufs = bufs.data(); b.col_lens = lens.data();
      b.delta0 = d0.data(); b.delta0_lens = d0l.data(); b.delta1 = d1.data(); b.delta1_lens = d1l.data();
      const std::vector<uint8_t>* del = it.getDeletedColumnBuffer();
      b.delete_buf = del ? del->data() : nullptr; b.delete_len = del ? (int64_t)del->size() : 0;
      b.stats_row = stats.empty() ? nullptr : stats.data(); b.stats_len = (int64_t)stats.size();
      b.stats_ncols = numTableColumns; b.bucket_id = it.getCurrentBucketId(); b.batch_id = it.getCurrentBatchId();
      check(sd_batch_submit(plan_, &b));
    }
    std::vector<uint8_t> out(1 << 14);
    int64_t len = 0, nrows = 0;
    int rc = sd_plan_finish(plan_, out.data(), (int64_t)out.size(), &len, &nrows);
    if (rc == SD_ERR_OVERFLOW) { out.resize((size_t)len); rc = sd_plan_finish(plan_, out.data(), (int64_t)out.size(), &len, &nrows); }
    check(rc);
    out.resize((size_t)len);
    return out;
  }

  // SQLMetrics of the fused operators (ColumnTableScan.scala:111-127, SnappyHashAggregateExec.scala:132-137)
  std::vector<int64_t> metrics() const {
    std::vector<int64_t> m(SD_NUM_METRICS);
    check(sd_plan_metrics(plan_, m.data()));
    return m;
  }
  int numTableColumns = 16;

 private:
  PlanBuilder& builder_;
  std::vector<int32_t> keys_;
  std::vector<sd_agg> aggs_;
  sd_plan_desc desc_;
  sd_plan* plan_ = nullptr;
};

// Driver-side merge of the partial rows of all partitions.
struct CollectAggregateExec {
  static std::vector<uint8_t> executeCollect(const sd_plan_desc& desc, const std::vector<std::vector<uint8_t>>& partitions) {
    std::vector<uint8_t> all;
    for (auto& p : partitions) all.insert(all.end(), p.begin(), p.end());
    std::vector<uint8_t> out(all.size() * 2 + 4096);
    int64_t len = 0, nrows = 0;
    check(sd_final_merge(&desc, all.data(), (int64_t)all.size(), out.data(), (int64_t)out.size(), &len, &nrows));
    out.resize((size_t)len);
    return out;
  }
};

}  // namespace snappy
#endif
