// sd_hash.cu -- housekeeping kernels of the MODE_HASH group table (init with slot identities, compaction
// of the occupied entries before the read-back).  The find-or-insert itself is in sd_kernels.cuh.
#include "sd_host.h"

namespace sd {

__global__ void hash_init_kernel(uint32_t* state, uint64_t* vals, uint32_t capacity, int nslot, const uint64_t* ident) {
  const size_t n = (size_t)capacity * nslot;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) vals[i] = ident[i % nslot];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < capacity; i += (size_t)gridDim.x * blockDim.x) state[i] = 0u;
}

__global__ void hash_compact_kernel(const uint32_t* state, const int64_t* keys, const uint32_t* knull, const uint64_t* vals,
                                    uint32_t capacity, int nk, int nslot, int64_t* out_keys, uint32_t* out_knull, uint64_t* out_vals,
                                    uint32_t* cursor) {
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < capacity; e += gridDim.x * blockDim.x) {
    if (state[e] != 2u) continue;
    const uint32_t o = atomicAdd(cursor, 1u);
    for (int k = 0; k < nk; k++) out_keys[(size_t)o * nk + k] = keys[(size_t)e * nk + k];
    out_knull[o] = knull[e];
    for (int s = 0; s < nslot; s++) out_vals[(size_t)o * nslot + s] = vals[(size_t)e * nslot + s];
  }
}

int hash_table_init(cudaStream_t stream, const HashTable& t, uint32_t capacity, int nslot, const uint64_t* d_ident) {
  hash_init_kernel<<<296, 256, 0, stream>>>(t.state, t.vals, capacity, nslot, d_ident);
  SD_CUDA(cudaGetLastError());
  SD_CUDA(cudaMemsetAsync(t.overflow, 0, 4, stream));
  SD_CUDA(cudaMemsetAsync(t.count, 0, 4, stream));
  return 0;
}

int hash_table_compact(cudaStream_t stream, const HashTable& t, uint32_t capacity, int nk, int nslot, int64_t* out_keys,
                       uint32_t* out_knull, uint64_t* out_vals, uint32_t* d_cursor) {
  SD_CUDA(cudaMemsetAsync(d_cursor, 0, 4, stream));
  hash_compact_kernel<<<296, 256, 0, stream>>>(t.state, t.keys, t.knull, t.vals, capacity, nk, nslot, out_keys, out_knull, out_vals, d_cursor);
  SD_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace sd
