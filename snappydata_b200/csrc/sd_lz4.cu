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

// ---- device: one warp per buffer, shared-memory staged -----------------------------------------------------
// A sequence-by-sequence decoder that touches HBM per sequence is latency bound (every match source read goes
// to L2: ~350 cycles per sequence measured).  Here each warp owns
//   IN  : a 2 KB staging buffer of compressed input, refilled by coalesced loads,
//   WIN : a 16 KB ring holding the most recent output,
// parses the token stream from IN (every lane redundantly: uniform shared-memory broadcasts), copies literals
// IN -> WIN and matches WIN -> WIN (distance <= 8 KB, the common case for column data) or HBM -> WIN (farther
// back: already flushed), and writes WIN out in coalesced 4 KB blocks as they complete.
constexpr int LZ_IN = 2048;           // input staging bytes per warp
constexpr int LZ_WIN = 16384;         // output window per warp (power of two)
constexpr int LZ_BLK = 4096;          // flush granularity
constexpr int LZ_NEAR = 8192;         // matches at most this far back are served from the window
constexpr int LZ_WARPS = 4;           // warps per CTA

struct LzWarp {
  uint8_t in[LZ_IN];
  uint8_t win[LZ_WIN];
};

__global__ void __launch_bounds__(LZ_WARPS * 32) lz4_decode_kernel(const Lz4Job* jobs, int njobs, unsigned int* error_flag) {
  extern __shared__ __align__(16) uint8_t lz_smem[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = blockIdx.x * LZ_WARPS + wib;
  if (warp >= njobs) return;
  LzWarp& sm = reinterpret_cast<LzWarp*>(lz_smem)[wib];
  const Lz4Job j = jobs[warp];
  const uint8_t* __restrict__ src = j.src;
  uint8_t* dst = j.dst;
  // output position p lives at win[(p + wofs) & mask]: then win and dst share their 16-byte phase and a block
  // that is aligned in HBM is aligned in the window as well
  const int wofs = (int)(reinterpret_cast<uintptr_t>(dst) & 15);
  auto W = [&](int64_t p) -> int { return (int)((p + wofs) & (LZ_WIN - 1)); };
  int64_t s = 0;                      // next input byte to parse
  int64_t in_base = 0, in_end = 0;    // IN holds src[in_base, in_end)
  int64_t o = 0, flushed = 0;         // output produced / written to HBM
  int64_t limit = wofs ? 16 - wofs : LZ_BLK;   // flush when o reaches it (first: the unaligned head)
  if (limit > j.dst_len) limit = j.dst_len;
  bool bad = false;

  auto stage = [&](int need) {        // make sure src[s, s + need) is staged (need <= 64)
    if (s + need <= in_end || in_end >= j.src_len) return;
    __syncwarp();
    in_base = s;
    in_end = s + LZ_IN < j.src_len ? s + LZ_IN : j.src_len;
    for (int i = lane; i < (int)(in_end - in_base); i += 32) sm.in[i] = src[in_base + i];
    __syncwarp();
  };
  auto inb = [&](int64_t p) -> uint32_t { return sm.in[p - in_base]; };
  auto flush = [&]() {                // write [flushed, limit) to HBM; called when o == limit
    __syncwarp();
    if (limit - flushed == LZ_BLK && ((reinterpret_cast<uintptr_t>(dst) + flushed) & 15) == 0) {
      const uint4* w = reinterpret_cast<const uint4*>(sm.win);                // 16-byte vectors of the ring (wraps)
      const int v0 = W(flushed) >> 4;
      uint4* d = reinterpret_cast<uint4*>(dst + flushed);
#pragma unroll
      for (int i = 0; i < LZ_BLK / 16 / 32; i++) d[lane + 32 * i] = w[(v0 + lane + 32 * i) & (LZ_WIN / 16 - 1)];
    } else {
      for (int64_t p = flushed + lane; p < limit; p += 32) dst[p] = sm.win[W(p)];
    }
    flushed = limit;
    limit = flushed + LZ_BLK < j.dst_len ? flushed + LZ_BLK : j.dst_len;
    __syncwarp();
  };

  while (s < j.src_len && !bad) {
    stage(64);
    const uint32_t token = inb(s++);
    int64_t lit = token >> 4;
    if (lit == 15) {
      for (;;) { stage(1); if (s >= j.src_len) { bad = true; break; } const uint32_t b = inb(s++); lit += b; if (b != 255) break; }
    }
    if (bad || s + lit > j.src_len || o + lit > j.dst_len) { bad = true; break; }
    // ---- literals: IN -> window (pieces bounded by the staging buffer and the next flush) -----------------
    while (lit > 0) {
      stage(1);
      int64_t n = in_end - s;
      if (n > lit) n = lit;
      if (n > limit - o) n = limit - o;
      for (int i = lane; i < (int)n; i += 32) sm.win[W(o + i)] = sm.in[s - in_base + i];
      s += n; o += n; lit -= n;
      if (o == limit && o < j.dst_len) flush();
    }
    if (s >= j.src_len) break;   // last sequence: literals only
    stage(2);
    if (s + 2 > j.src_len) { bad = true; break; }
    const int64_t off = (int64_t)inb(s) | ((int64_t)inb(s + 1) << 8);
    s += 2;
    int64_t ml = token & 15;
    if (ml == 15) {
      for (;;) { stage(1); if (s >= j.src_len) { bad = true; break; } const uint32_t b = inb(s++); ml += b; if (b != 255) break; }
    }
    ml += 4;
    if (bad || off == 0 || off > o || o + ml > j.dst_len) { bad = true; break; }
    // ---- match: window -> window when near, HBM -> window when it refers to already flushed output ---------
    while (ml > 0) {
      int64_t n = ml;
      if (n > limit - o) n = limit - o;
      __syncwarp();                                   // sources written by other lanes are visible
      if (off <= LZ_NEAR) {
        if (off >= n) {
          for (int i = lane; i < (int)n; i += 32) sm.win[W(o + i)] = sm.win[W(o - off + i)];
        } else {                                      // overlapping: the pattern of `off` bytes repeats
          for (int i = lane; i < (int)n; i += 32) sm.win[W(o + i)] = sm.win[W(o - off + (i % (int)off))];
        }
      } else {                                        // farther back than the window guarantees: flushed bytes from HBM
        for (int i = lane; i < (int)n; i += 32) {
          const int64_t p = o - off + i;
          sm.win[W(o + i)] = p < flushed ? dst[p] : sm.win[W(p)];
        }
      }
      o += n; ml -= n;
      if (o == limit && o < j.dst_len) flush();
    }
  }
  // tail: what is left in the window
  __syncwarp();
  if (!bad && o == j.dst_len) {
    for (int64_t p = flushed + lane; p < o; p += 32) dst[p] = sm.win[W(p)];
  } else if (lane == 0) atomicExch(error_flag, 1u);
}

int lz4_launch(cudaStream_t stream, const Lz4Job* d_jobs, int njobs, unsigned int* d_error) {
  if (njobs <= 0) return 0;
  const int blocks = (njobs + LZ_WARPS - 1) / LZ_WARPS;
  const size_t smem = sizeof(LzWarp) * LZ_WARPS;
  static bool attr_set = false;
  if (!attr_set) {
    SD_CUDA(cudaFuncSetAttribute(lz4_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  lz4_decode_kernel<<<blocks, LZ_WARPS * 32, smem, stream>>>(d_jobs, njobs, d_error);
  SD_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sd

// test hook: the host prefix decoder (tests/test_lz4_prefix.py compares it with liblz4)
extern "C" int64_t sdx_lz4_decode_prefix(const void* src, int64_t src_len, void* dst, int64_t want) {
  return sd::lz4_decode_prefix(reinterpret_cast<const uint8_t*>(src), src_len, reinterpret_cast<uint8_t*>(dst), want);
}
