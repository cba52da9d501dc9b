// Hardware probe (test infrastructure, not on any product path): does a K-major SWIZZLE_128B UMMA
// shared-memory descriptor address an operand that starts at an arbitrary ROW of a TMA-written
// block (start address = block + r * 128 B), and what must the descriptor's base-offset field
// (bits 49..51) hold for it?  The answer decides whether overlapping (Toeplitz) A tiles can be
// read in place from one tall block instead of being re-fetched per K block.
//   out[v][r][128][32] = A[r .. r+128) x B^T  for row offsets r = 0..15 and
//   v = 0: base_offset = 0        v = 1: base_offset = r & 7
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_host.cuh"

namespace nnab {

constexpr int PROBE_ROWS = 144, PROBE_N = 32, PROBE_R = 16;

__global__ void __launch_bounds__(128, 1)
probe_rowoffset_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                       float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_s = base;                           // 144 rows x 128 B = 18 KB
  const uint32_t b_s = base + 20 * 1024;               // 32 rows x 128 B
  const uint32_t bar_full = base + 28 * 1024, bar_mma = bar_full + 8, tmem_slot = bar_full + 16;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 32);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_full, PROBE_ROWS * 128 + PROBE_N * 128);
    tma_load_3d(a_s, &tm_a, bar_full, 0, 0, 0);
    tma_load_3d(b_s, &tm_b, bar_full, 0, 0, 0);
  }
  mbar_wait(bar_full, 0);
  tcgen05_fence_after();
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(PROBE_N >> 3) << 17) |
                         ((uint32_t)(128 >> 4) << 24);
  uint32_t phase = 0;
  for (int v = 0; v < 2; ++v) {
    for (int r = 0; r < PROBE_R; ++r) {
      if (threadIdx.x == 0) {
        for (int k = 0; k < 4; ++k) {
          uint64_t da = make_smem_desc<64>(a_s + (uint32_t)r * 128u + (uint32_t)k * 32u);
          if (v == 1) da |= (uint64_t)(r & 7) << 49;
          const uint64_t db = make_smem_desc<64>(b_s + (uint32_t)k * 32u);
          umma_bf16(tmem_base, da, db, idesc, k > 0 ? 1u : 0u);
        }
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, phase);
      phase ^= 1u;
      tcgen05_fence_after();
      uint32_t d[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), d);
      tmem_ld_wait();
      float* o = out + (((size_t)v * PROBE_R + r) * 128 + warp * 32 + lane) * PROBE_N;
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(d[j]);
      tcgen05_fence_before();
      __syncthreads();
      tcgen05_fence_after();
    }
  }
  if (warp == 1) tmem_dealloc(tmem_base, 32);
}

// a: (144, 64) bf16, b: (32, 64) bf16, out: (2, 16, 128, 32) fp32 -- all device pointers
int tc_probe_rowoffset(const void* a, const void* b, float* out, cudaStream_t stream) {
  CUtensorMap ma, mb;
  int rc = encode_3d(&ma, const_cast<void*>(a), 64, PROBE_ROWS, 1, 128, (uint64_t)PROBE_ROWS * 128, 64,
                     PROBE_ROWS, 64);
  if (rc) return rc;
  rc = encode_3d(&mb, const_cast<void*>(b), 64, PROBE_N, 1, 128, (uint64_t)PROBE_N * 128, 64, PROBE_N, 64);
  if (rc) return rc;
  const int smem = 32 * 1024;
  probe_rowoffset_kernel<<<1, 128, smem, stream>>>(ma, mb, out);
  NNAB_LAUNCH_CHECK();
  return NNAB_OK;
}

}  // namespace nnab

extern "C" __attribute__((visibility("default"))) int nnab_probe_rowoffset(const void* a, const void* b,
                                                                         float* out, void* stream) {
  if (a == nullptr || b == nullptr || out == nullptr) return NNAB_EINVAL;
  return nnab::tc_probe_rowoffset(a, b, out, (cudaStream_t)stream);
}
