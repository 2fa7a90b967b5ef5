// test_operators.cpp -- drives the C++ operator mirror (sd_operators.hpp -> C ABI -> CUDA) over a
// synthetic lineitem table, partition by partition like Spark tasks, and checks the merged result
// against the CPU oracle (liboracle.so) on the same ColumnBatch bytes.  Reads like the reference's
// TPCHDUnitTest: build the table, run Q1 and Q6, compare.  Run by tests/test_gpu_cpp_operators.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>

#include "../../snappydata_b200/csrc/sd_operators.hpp"

extern "C" {
// oracle (test infrastructure): same ABI shape under the oracle_ prefix
struct oracle_plan;
int oracle_plan_create(const sd_plan_desc*, oracle_plan**);
int oracle_plan_set_literals(oracle_plan*, const sd_literal*, int32_t);
int oracle_batch_submit(oracle_plan*, const sd_batch*);
int oracle_plan_finish(oracle_plan*, void*, int64_t, int64_t*, int64_t*);
void oracle_plan_destroy(oracle_plan*);
int oracle_final_merge(const sd_plan_desc*, const void*, int64_t, void*, int64_t, int64_t*, int64_t*);
const char* oracle_last_error(void);
// bench/test utilities of the product
int sdx_store_gen_lineitem(sd_store*, int64_t, int64_t, int32_t, int32_t, uint64_t, int32_t);
int sdx_store_get_buffer(sd_store*, int64_t, int32_t, void*, int64_t, int64_t*);
int sdx_store_batch_info(sd_store*, int64_t, int32_t*, int32_t*, int64_t*);
}

using namespace snappy;

static sd_literal lit_i(sd_type t, int64_t v) { sd_literal l; memset(&l, 0, sizeof(l)); l.type = t; l.i = v; return l; }
static sd_literal lit_d(double v) { sd_literal l; memset(&l, 0, sizeof(l)); l.type = SD_DOUBLE; l.d = v; return l; }

// rows of a result stream as (key string -> doubles/longs rendered as doubles)
static std::map<std::string, std::vector<double>> parse(const std::vector<uint8_t>& rows, int nkeys, const std::vector<int>& types) {
  std::map<std::string, std::vector<double>> out;
  size_t pos = 0;
  const int n = (int)types.size();
  while (pos + 8 <= rows.size()) {
    int64_t sz; memcpy(&sz, rows.data() + pos, 8);
    const uint8_t* r = rows.data() + pos + 8;
    const int64_t bits = ((n + 63) / 64) * 8;
    std::string key; std::vector<double> vals;
    for (int i = 0; i < n; i++) {
      const uint8_t* slot = r + bits + 8 * i;
      const bool isnull = r[i >> 3] & (1u << (i & 7));
      if (i < nkeys) {
        int64_t ol; memcpy(&ol, slot, 8);
        key += isnull ? std::string("<null>") : std::string((const char*)r + (ol >> 32), (size_t)(ol & 0xffffffff));
        key += "|";
      } else if (isnull) vals.push_back(NAN);
      else if (types[i] == SD_DOUBLE) { double d; memcpy(&d, slot, 8); vals.push_back(d); }
      else { int64_t v; memcpy(&v, slot, 8); vals.push_back((double)v); }
    }
    out[key] = vals;
    pos += 8 + (size_t)sz;
  }
  return out;
}

static int compare(const char* what, const std::vector<uint8_t>& got, const std::vector<uint8_t>& want, int nkeys, const std::vector<int>& types) {
  auto g = parse(got, nkeys, types), w = parse(want, nkeys, types);
  if (g.size() != w.size()) { printf("FAIL %s: %zu groups, oracle %zu\n", what, g.size(), w.size()); return 1; }
  for (auto& kv : w) {
    auto it = g.find(kv.first);
    if (it == g.end()) { printf("FAIL %s: group %s missing\n", what, kv.first.c_str()); return 1; }
    for (size_t i = 0; i < kv.second.size(); i++) {
      const double a = it->second[i], b = kv.second[i];
      const bool is_count = types[nkeys + i] != SD_DOUBLE;
      const bool ok = (std::isnan(a) && std::isnan(b)) || (is_count ? a == b : std::fabs(a - b) <= 1e-6 * std::fmax(std::fabs(a), std::fabs(b)));
      if (!ok) { printf("FAIL %s: group %s field %zu: %.17g vs oracle %.17g\n", what, kv.first.c_str(), i, a, b); return 1; }
    }
  }
  printf("ok   %s: %zu group(s) match the oracle\n", what, g.size());
  return 0;
}

int main(int argc, char** argv) {
  const int64_t rows = argc > 1 ? atoll(argv[1]) : 150001;
  const int per_batch = 20000, nbuckets = 4;
  check(sd_init(0));
  // ---- the table: generated on the device, copied back as host ColumnBatches (what a region holds) ----
  std::vector<sd_column> schema(16);
  const sd_type types16[16] = {SD_LONG, SD_LONG, SD_LONG, SD_INT, SD_DOUBLE, SD_DOUBLE, SD_DOUBLE, SD_DOUBLE, SD_STRING, SD_STRING,
                               SD_DATE, SD_DATE, SD_DATE, SD_STRING, SD_STRING, SD_STRING};
  for (int i = 0; i < 16; i++) schema[i] = sd_column{types16[i], 0, i, 0};
  sd_store* store = nullptr;
  check(sd_store_create(0, 16, schema.data(), &store));
  check(sdx_store_gen_lineitem(store, 0, rows, per_batch, nbuckets, 77, 0x7f0));
  int64_t nb = 0;
  check(sd_store_num_batches(store, &nb));
  std::vector<ColumnBatch> table((size_t)nb);
  for (int64_t b = 0; b < nb; b++) {
    int32_t n, bucket; int64_t id;
    check(sdx_store_batch_info(store, b, &n, &bucket, &id));
    table[b].numRows = n; table[b].bucketId = bucket; table[b].batchId = id;
    table[b].buffers.resize(16);
    for (int c = 4; c <= 10; c++) {
      int64_t len = 0;
      sdx_store_get_buffer(store, b, c, nullptr, 0, &len);
      table[b].buffers[c].resize((size_t)len);
      check(sdx_store_get_buffer(store, b, c, table[b].buffers[c].data(), len, &len));
    }
  }
  sd_store_destroy(store);

  int failures = 0;
  for (int q = 0; q < 2; q++) {
    PlanBuilder b;
    std::vector<sd_literal> lits;
    std::vector<Expr> keys;
    std::vector<AggregateExpression> aggs;
    Expr cond;
    std::vector<int> ptypes, ftypes;
    if (q == 0) {   // TPC-H Q6 (TPCH_Queries.scala:600-613)
      Expr ship = b.attr({"l_shipdate", SD_DATE, false, 10}), disc = b.attr({"l_discount", SD_DOUBLE, false, 6});
      Expr qty = b.attr({"l_quantity", SD_DOUBLE, false, 4}), price = b.attr({"l_extendedprice", SD_DOUBLE, false, 5});
      Expr d0 = b.literal(SD_DATE), d1 = b.literal(SD_DATE), lo = b.literal(SD_DOUBLE), hi = b.literal(SD_DOUBLE), qq = b.literal(SD_DOUBLE);
      cond = b.And(b.And(b.And(b.And(b.cmp(SD_OP_GE, ship, d0), b.cmp(SD_OP_LT, ship, d1)), b.cmp(SD_OP_GE, disc, lo)), b.cmp(SD_OP_LE, disc, hi)),
                   b.cmp(SD_OP_LT, qty, qq));
      aggs.push_back({SD_AGG_SUM, b.Multiply(price, disc)});
      lits = {lit_i(SD_DATE, 8766), lit_i(SD_DATE, 9131), lit_d(0.05), lit_d(0.07), lit_d(24.0)};
      ptypes = {SD_DOUBLE}; ftypes = {SD_DOUBLE};
    } else {        // TPC-H Q1 (TPCH_Queries.scala:125-149)
      Expr qty = b.attr({"l_quantity", SD_DOUBLE, false, 4}), price = b.attr({"l_extendedprice", SD_DOUBLE, false, 5});
      Expr disc = b.attr({"l_discount", SD_DOUBLE, false, 6}), tax = b.attr({"l_tax", SD_DOUBLE, false, 7});
      Expr rf = b.attr({"l_returnflag", SD_STRING, false, 8}), ls = b.attr({"l_linestatus", SD_STRING, false, 9});
      Expr ship = b.attr({"l_shipdate", SD_DATE, false, 10});
      Expr cutoff = b.literal(SD_DATE), one_a = b.literal(SD_DOUBLE), one_b = b.literal(SD_DOUBLE);
      cond = b.cmp(SD_OP_LE, ship, cutoff);
      keys = {rf, ls};
      Expr disc_price = b.Multiply(price, b.Subtract(one_a, disc));
      aggs = {{SD_AGG_SUM, qty}, {SD_AGG_SUM, price}, {SD_AGG_SUM, disc_price}, {SD_AGG_SUM, b.Multiply(disc_price, b.Add(one_b, tax))},
              {SD_AGG_AVG, qty}, {SD_AGG_AVG, price}, {SD_AGG_AVG, disc}, {SD_AGG_COUNT_STAR, Expr()}};
      lits = {lit_i(SD_DATE, 10136), lit_d(1.0), lit_d(1.0)};
      ftypes = {SD_STRING, SD_STRING, SD_DOUBLE, SD_DOUBLE, SD_DOUBLE, SD_DOUBLE, SD_DOUBLE, SD_DOUBLE, SD_DOUBLE, SD_LONG};
    }
    SnappyHashAggregateExec agg(b, &cond, keys, aggs);
    // one task per bucket (partition = bucket, like Spark local[N])
    std::vector<std::vector<uint8_t>> partials;
    for (int bucket = 0; bucket < nbuckets; bucket++) {
      std::vector<ColumnBatch> part;
      for (auto& cb : table) if (cb.bucketId == bucket) part.push_back(cb);
      ColumnBatchIterator it(&part);
      partials.push_back(agg.executePartition(it, lits));
    }
    std::vector<uint8_t> got = CollectAggregateExec::executeCollect(agg.desc(), partials);
    // oracle over the same bytes, one partition
    oracle_plan* op = nullptr;
    if (oracle_plan_create(&agg.desc(), &op) || oracle_plan_set_literals(op, lits.data(), (int32_t)lits.size())) { printf("oracle: %s\n", oracle_last_error()); return 2; }
    for (auto& cb : table) {
      std::vector<const void*> bufs; std::vector<int64_t> lens;
      for (int c = 0; c < agg.desc().ncols; c++) { const auto& v = cb.buffers[b.cols_[c].table_ordinal]; bufs.push_back(v.data()); lens.push_back((int64_t)v.size()); }
      sd_batch sb; memset(&sb, 0, sizeof(sb));
      sb.num_rows = cb.numRows; sb.ncols = agg.desc().ncols; sb.col_bufs = bufs.data(); sb.col_lens = lens.data();
      if (oracle_batch_submit(op, &sb)) { printf("oracle: %s\n", oracle_last_error()); return 2; }
    }
    std::vector<uint8_t> opart(1 << 16), want(1 << 16);
    int64_t len = 0, n = 0;
    oracle_plan_finish(op, opart.data(), (int64_t)opart.size(), &len, &n);
    opart.resize((size_t)len);
    oracle_final_merge(&agg.desc(), opart.data(), len, want.data(), (int64_t)want.size(), &len, &n);
    want.resize((size_t)len);
    oracle_plan_destroy(op);
    failures += compare(q == 0 ? "Q6 via ColumnTableScan->Filter->SnappyHashAggregate (4 partitions)" : "Q1 via ColumnTableScan->Filter->SnappyHashAggregate (4 partitions)",
                        got, want, (int)keys.size(), ftypes);
    auto m = agg.metrics();
    printf("     metrics: columnBatchesSeen=%lld rowsScanned(last partition)=%lld kernelLaunches=%lld\n", (long long)m[2], (long long)m[8], (long long)m[7]);
  }
  return failures ? 1 : 0;
}
