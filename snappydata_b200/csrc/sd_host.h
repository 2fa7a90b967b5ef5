// sd_host.h -- internal host-side declarations shared by the engine translation units.
#ifndef SD_HOST_H
#define SD_HOST_H

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/snappy_gpu.h"
#include "sd_codegen.h"
#include "sd_device.h"

namespace sd {

// ---- errors -------------------------------------------------------------------------------------
int set_error(int code, const char* fmt, ...);
#define SD_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return sd::set_error(SD_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// ---- kernel registry ------------------------------------------------------------------------------
struct KernelEntry {
  std::string signature;
  const void* func;      // &scan_aggregate_kernel<PLAN> (runtime API launch)
  void* drv_func;        // CUfunction for NVRTC-compiled plans
  size_t tile_smem;      // sizeof(TileSmem<PLAN>) rounded up to 16
  int staged = 0;        // PLAN::STAGES > 0: producer warp + shared-memory ring (block = THREADS + 32)
  size_t stage_bytes = 0;
  std::string origin;    // "aot" | "jit"
  std::string name;      // Plan_<fnv1a(signature)>: the generated struct's name
};
std::vector<KernelEntry>& kernel_registry();
std::string plan_struct_name(const std::string& signature);
struct AotRegistrar {
  AotRegistrar(const char* signature, const void* func, size_t tile_smem, int staged, size_t stage_bytes);
};
// set the dynamic shared-memory limit and query CTAs/SM; launch (runtime API for AOT, driver API for JIT)
int kernel_prepare(const KernelEntry& k, size_t smem, int* ctas_per_sm);
inline int kernel_block_threads(const KernelEntry& k) { return THREADS + (k.staged ? 32 : 0); }
int kernel_launch(const KernelEntry& k, int grid, size_t smem, cudaStream_t stream, void** args);
// NVRTC path (sd_jit.cpp): compile `spec.source` against the embedded kernel headers.
int jit_compile(const PlanSpec& spec, int device, KernelEntry& out);

// ---- MODE_HASH group table housekeeping (sd_hash.cu) ------------------------------------------------
int hash_table_init(cudaStream_t stream, const HashTable& t, uint32_t capacity, int nslot, const uint64_t* d_ident);
int hash_table_compact(cudaStream_t stream, const HashTable& t, uint32_t capacity, int nk, int nslot, int64_t* out_keys,
                       uint32_t* out_knull, uint64_t* out_vals, uint32_t* d_cursor);
// strings held by reference -> host: d_recs[i * stride] = device address of a [len:int32][bytes] record (0: none)
int fetch_string_records(cudaStream_t stream, const int64_t* d_recs, int64_t n, int64_t stride, std::vector<std::string>& out);

// ---- sd_rows.cu: projection records -> UnsafeRows on the device -------------------------------------------------------------
constexpr int ROW_MAX_FIELDS = 32;   // one lane per field; the record's null bits are 32 wide
enum { ROW_KIND_8 = 0, ROW_KIND_BOOL = 1, ROW_KIND_1 = 2, ROW_KIND_2 = 3, ROW_KIND_4 = 4, ROW_KIND_FLOAT = 5, ROW_KIND_STRING = 6 };
// where the strings of one projected STRING column of one batch live on the device
struct RowStrSrc {
  const uint8_t* base;        // dictionary: the column's uploaded buffer; raw strings: its [len][bytes] body
  const int32_t* rec_off;     // dictionary: offset of each code's [len][bytes] record from base; nullptr: raw (the value IS the offset)
  int32_t null_code;          // dictionary code that stands for NULL (-1: none)
  int32_t n;                  // dictionary entries
};
struct RowWriterBuffers {
  int64_t* d_offs = nullptr; size_t offs_cap = 0;
  void* d_tmp = nullptr; size_t tmp_cap = 0;
  uint8_t* d_rows = nullptr; size_t rows_cap = 0;
  int32_t* d_err = nullptr;
  RowStrSrc* d_src = nullptr; size_t src_cap = 0;       // [batches of the execution][string fields]
  int32_t* d_recoff = nullptr; size_t recoff_cap = 0;   // the rec_off tables, back to back
  std::vector<const void*> src_key;                       // the batches d_src / d_recoff were built for
  void release();
};
int device_write_rows(cudaStream_t st, const uint64_t* d_recs, int64_t count, int np, const uint8_t* kinds, int nbatches,
                      RowWriterBuffers& b, int64_t* total_out);

// ---- on-device LZ4 (sd_lz4.cu) ---------------------------------------------------------------------------
struct Lz4Job { const uint8_t* src; uint8_t* dst; int64_t src_len; int64_t dst_len; };
int64_t lz4_decode_prefix(const uint8_t* src, int64_t src_len, uint8_t* dst, int64_t want);
// host Snappy (raw format) decoder: returns the uncompressed length or -1 (sd_lz4.cu)
int64_t snappy_decode(const uint8_t* src, int64_t src_len, uint8_t* dst, int64_t dst_cap);
// a stored buffer [-codecId][uncompressedLen][payload] (CompressionUtils.scala:53-61) -> its uncompressed bytes, on the host
int decompress_envelope_host(const uint8_t* buf, int64_t len, std::vector<uint8_t>& out);
int lz4_launch(cudaStream_t stream, const Lz4Job* d_jobs, int njobs, unsigned int* d_error);

// ---- NCCL behind sd_comm (sd_nccl.cpp; dlopen'ed) --------------------------------------------------------
int comm_unique_id(void* out128);
int comm_init(const void* id128, int rank, int world, void** out);
void comm_destroy(void* c);
int comm_all_gather_bytes(void* c, const void* d_send, void* d_recv, size_t bytes_per_rank, cudaStream_t stream);

// ---- device memory arena: bump allocation out of large slabs ---------------------------------------
struct Arena {
  int device = 0;
  size_t slab_bytes = size_t(512) << 20;
  std::vector<std::pair<uint8_t*, size_t>> slabs;
  size_t cur_slab = 0;
  size_t cur_off = 0;
  size_t used = 0;
  // returns p with (p + misalign) % align == 0, or nullptr (error set)
  uint8_t* alloc(size_t n, size_t align = 256, size_t misalign = 0);
  void reset();     // keep slabs, forget allocations
  void release();   // free slabs
  ~Arena() { release(); }
};

// page-locked host staging for small uploads (descriptors, tables) that must not block the submitting thread:
// copies from it are truly asynchronous, and the memory stays valid until reset()
struct PinnedArena {
  size_t slab_bytes = size_t(4) << 20;
  std::vector<std::pair<uint8_t*, size_t>> slabs;
  size_t cur_slab = 0, cur_off = 0;
  uint8_t* alloc(size_t n);   // 16-byte aligned, nullptr on failure (error set)
  void reset() { cur_slab = 0; cur_off = 0; }
  ~PinnedArena();
};

// ---- resident batches ------------------------------------------------------------------------------
struct StoredDelta {
  bool present = false;
  DevDelta dev;                           // device pointers
  int64_t len = 0;
  std::vector<std::string> dict_strings;  // STRING values dictionary
};

struct StoredCol {
  bool present = false;
  std::string unsupported;                // non-empty: reason the GPU path cannot scan this column
  int64_t len = 0;
  int64_t algo_bytes = 0;                 // len - 8 - dictionary bytes
  uint8_t* dev_base = nullptr;            // device copy of the whole buffer
  int64_t body_off = 0;                   // offset of the first value / index
  DevCol dev;                             // device view (delta pointers filled per plan)
  std::vector<std::string> dict_strings;  // STRING dictionary (or distinct RLE run strings)
  std::vector<int64_t> dict_rec_off;      // per dict_strings entry: offset of its [len][bytes] record in the buffer (-1: none)
  std::vector<int64_t> dict_rec_ptr;      // per unified code: device address of the record (0: NULL placeholder) -- string keys by reference
  bool raw_str = false;                   // Uncompressed variable-width STRING body (dev.enc == ENC_STR_RAW)
  bool has_nulls = false;
  StoredDelta delta[2];
  DevDelta* dev_delta[2] = {nullptr, nullptr};   // DevDelta structs resident on the device
  bool fast = false;                      // vector fast path applies
};

inline uint64_t next_batch_uid() { static std::atomic<uint64_t> n{1}; return n.fetch_add(1, std::memory_order_relaxed); }
struct StoredBatch {
  uint64_t uid = next_batch_uid();        // never reused (an address can be): key of per-plan caches built from a batch
  int32_t num_rows = 0;
  int32_t bucket_id = 0;
  int64_t batch_id = 0;
  std::vector<StoredCol> cols;            // by table column
  std::vector<uint8_t> stats;             // stats UnsafeRow (host copy)
  int32_t stats_ncols = 0;
  int32_t* dev_deletes = nullptr;
  int32_t num_deletes = 0;
  bool has_deltas = false;
  bool positional = false;                // cols are indexed by the plan's scan column (private store)
};

}  // namespace sd

struct sd_store {
  // ingest (sd_store_put_batch) may run concurrently with scans: `mu` guards batches / arena / version / the LZ4 queues.  A scan
  // works on the SNAPSHOT of batch pointers it takes under the lock when it starts (batches are never removed or moved
  // while the store lives; a batch becomes visible only after its bytes have reached the device).
  std::mutex mu;
  int device = 0;
  std::vector<sd_column> schema;
  sd::Arena arena;
  cudaStream_t copy_stream = nullptr;
  // extra H2D streams (SD_TUNE_COPY_STREAMS > 1): large buffer copies rotate over them
  cudaStream_t extra_streams[4] = {nullptr, nullptr, nullptr, nullptr};
  int num_copy_streams = 1;
  int next_stream = 0;
  cudaEvent_t extra_done[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<std::unique_ptr<sd::StoredBatch>> batches;
  int64_t version = 0;
  int64_t h2d_bytes = 0;
  bool retain_buffers = false;   // SD_OPT_RETAIN_BUFFERS: no per-put synchronisation of the copy stream
  cudaEvent_t copies_done = nullptr;
  // compressed payloads waiting to be expanded on the device (one launch for many buffers)
  std::vector<sd::Lz4Job> pending_lz4;
  sd::Arena lz4_stage;
  unsigned int* d_lz4_error = nullptr;
  int64_t lz4_buffers = 0, lz4_in_bytes = 0, lz4_out_bytes = 0;
  // expansions are queued on a few streams of their own so that the launches of successive flushes (each one as
  // long as its longest buffer) overlap each other and the copies that follow
  static constexpr int LZ4_STREAMS = 12;   // enough launches in flight to keep ~2800 buffers resident with 256 MB flushes
  cudaStream_t lz4_streams[LZ4_STREAMS] = {};
  cudaEvent_t lz4_done[LZ4_STREAMS] = {};
  bool lz4_used[LZ4_STREAMS] = {};
  cudaEvent_t lz4_copied = nullptr;
  int lz4_next = 0;
  // compressed payloads of one batch that lie (almost) back to back in host memory travel as ONE copy
  const uint8_t* span_h0 = nullptr; uint8_t* span_d0 = nullptr; size_t span_len = 0;
  sd::PinnedArena lz4_jobs_host;
  sd::PinnedArena enc_host;        // sd_encode.cu: prefixes / descriptors on their way to the device
  std::mutex enc_mu;               // one encoder at a time per store; `mu` is taken only to lay the buffers out and to publish
  cudaStream_t enc_stream = nullptr;
  cudaEvent_t enc_event = nullptr;   // page-locked copies of the job lists (a pageable source would stall the caller per flush)
};

namespace sd {
// upload one batch (columns by table ordinal of `schema`) into the store's arena
int store_put(sd_store* s, const sd_batch* b, const int32_t* table_ordinals /* nullptr: identity */);
// queue the expansion of every pending compressed buffer (asynchronous; no-op when nothing is pending)
int store_flush_lz4(sd_store* s);
// make `stream` wait for every expansion queued so far
int store_lz4_order(sd_store* s, cudaStream_t stream);
// block until the queued expansions are done, release their staging memory, report a corrupt payload
int store_lz4_check(sd_store* s);
}

#endif
