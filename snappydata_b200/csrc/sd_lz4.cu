// sd_lz4.cu -- on-device LZ4 block decompression of stored column buffers.
//
// The reference stores a column value as [-codecId:int32][uncompressedLen:int32][LZ4 block] when it is
// >= 2048 bytes and shrinks to <= 75 % (encoders/.../store/CompressionUtils.scala:53-61,102-110) and
// decompresses it on the CPU whenever a scan needs it (ColumnFormatEntry.scala:498-570,
// ColumnBatchIterator.scala:102-113).  Here only the COMPRESSED bytes cross PCIe; the block is expanded in
// HBM by one warp per buffer, many buffers per launch.  The host decodes just the first bytes it needs to
// lay the buffer out (8-byte header, null words, dictionary) with the small prefix decoder below.
//
// LZ4 block format: sequences of [token][literal length ext*][literals][offset:2][match length ext*];
// token = (literal length << 4) | (match length - 4); the last sequence ends after its literals.
#include <cstring>

#include "sd_host.h"

namespace sd {

// ---- host: decode at most `want` leading bytes of an LZ4 block (returns bytes produced, -1 if corrupt) --
int64_t lz4_decode_prefix(const uint8_t* src, int64_t src_len, uint8_t* dst, int64_t want) {
  int64_t s = 0, o = 0;
  while (s < src_len && o < want) {
    const uint8_t token = src[s++];
    int64_t lit = token >> 4;
    if (lit == 15) { uint8_t b; do { if (s >= src_len) return -1; b = src[s++]; lit += b; } while (b == 255); }
    if (s + lit > src_len) return -1;
    const int64_t lcopy = lit < want - o ? lit : want - o;
    memcpy(dst + o, src + s, (size_t)lcopy);
    o += lcopy; s += lit;
    if (o >= want || s >= src_len) break;
    if (s + 2 > src_len) return -1;
    const int64_t off = src[s] | ((int64_t)src[s + 1] << 8);
    s += 2;
    int64_t ml = (token & 15);
    if (ml == 15) { uint8_t b; do { if (s >= src_len) return -1; b = src[s++]; ml += b; } while (b == 255); }
    ml += 4;
    if (off == 0 || off > o) return -1;
    for (int64_t i = 0; i < ml && o < want; i++, o++) dst[o] = dst[o - off];
  }
  return o;
}

// ---- device: one warp per buffer ---------------------------------------------------------------------
// Measured (round 1, profiles/r01_lz4.txt): the decode of one 1.6 MB column buffer is a serial chain of ~350
// cycles per LZ4 sequence (token load, offset load, match source read from L2), i.e. ~70-130 ms per launch
// however many buffers it covers, so for Q1 the compressed path (46 % of the bytes over PCIe) is currently
// SLOWER end to end than sending the buffers uncompressed (0.43 vs 1.09 G rows/s).  A variant that staged the
// input and a 16 KB output window in shared memory was not faster (more instructions per sequence) and was
// dropped.  Making this path pay needs intra-buffer parallelism (speculative sequence boundaries or a
// two-pass parse/copy split): next round.
// Every lane parses the (uniform) token stream redundantly -- the loads are warp-uniform broadcasts -- and
// the copies are split across lanes.  A match may overlap its own output (offset < length): byte i of the
// match is out[o - off + (i % off)], which always lies in the already written region.
__global__ void lz4_decode_kernel(const Lz4Job* jobs, int njobs, unsigned int* error_flag) {
  const int warp = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (warp >= njobs) return;
  const Lz4Job j = jobs[warp];
  const uint8_t* __restrict__ src = j.src;
  uint8_t* dst = j.dst;
  int64_t s = 0, o = 0;
  bool bad = false;
  while (s < j.src_len) {
    const uint32_t token = src[s++];
    int64_t lit = token >> 4;
    if (lit == 15) { uint32_t b; do { if (s >= j.src_len) { bad = true; break; } b = src[s++]; lit += b; } while (b == 255); }
    if (bad || s + lit > j.src_len || o + lit > j.dst_len) { bad = true; break; }
    for (int64_t i = lane; i < lit; i += 32) dst[o + i] = src[s + i];
    o += lit; s += lit;
    if (s >= j.src_len) break;   // last sequence: literals only
    if (s + 2 > j.src_len) { bad = true; break; }
    const int64_t off = (int64_t)src[s] | ((int64_t)src[s + 1] << 8);
    s += 2;
    int64_t ml = token & 15;
    if (ml == 15) { uint32_t b; do { if (s >= j.src_len) { bad = true; break; } b = src[s++]; ml += b; } while (b == 255); }
    ml += 4;
    if (bad || off == 0 || off > o || o + ml > j.dst_len) { bad = true; break; }
    __syncwarp();   // the match source was written by other lanes (this or earlier sequences)
    const uint8_t* m = dst + o - off;
    if (off >= ml) { for (int64_t i = lane; i < ml; i += 32) dst[o + i] = m[i]; }
    else { for (int64_t i = lane; i < ml; i += 32) dst[o + i] = m[i % off]; }
    o += ml;
  }
  if ((bad || o != j.dst_len) && lane == 0) atomicExch(error_flag, 1u);
}

int lz4_launch(cudaStream_t stream, const Lz4Job* d_jobs, int njobs, unsigned int* d_error) {
  if (njobs <= 0) return 0;
  const int warps_per_block = 4;
  const int blocks = (njobs + warps_per_block - 1) / warps_per_block;
  lz4_decode_kernel<<<blocks, warps_per_block * 32, 0, stream>>>(d_jobs, njobs, d_error);
  SD_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sd

// test hook: the host prefix decoder (tests/test_lz4_prefix.py compares it with liblz4)
extern "C" int64_t sdx_lz4_decode_prefix(const void* src, int64_t src_len, void* dst, int64_t want) {
  return sd::lz4_decode_prefix(reinterpret_cast<const uint8_t*>(src), src_len, reinterpret_cast<uint8_t*>(dst), want);
}
