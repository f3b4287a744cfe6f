// Micro-probes used to calibrate the kernels' cost models (not part of the product path):
// tcgen05.mma issue-to-retire rate for a given tile shape / operand layout.
// Compiled only into probe builds: python -m diffbir_b200.build --tag=probes -DDBIR_DEBUG_PROBES
// (loaded with DBIR_LIB_TAG=probes); the product libraries carry no calibration code.
#ifdef DBIR_DEBUG_PROBES
#include "common.cuh"
#include "../../include/diffbir_b200.h"

namespace {

// One CTA, one warp issuing `iters` back-to-back M=128 x N x K=16 MMAs (smem operands, zeros),
// one commit at the end; out[0] = cycles from first issue to the commit's arrival.
__global__ void __launch_bounds__(64, 1) mma_rate_kernel(int n, int b_mn, int iters, int a_tmem, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    const uint64_t a_desc = umma_desc_sw128(smem_u32(smem));
    const uint64_t b_desc = umma_desc_sw128(smem_u32(smem + 16384));
    const uint32_t idesc = umma_idesc(n, 0, b_mn);
    long long t0 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        if (a_tmem == 2) {
          // stage the A slab smem -> TMEM (128 rows x 32 bytes = 8 columns, double buffered), then TS MMA
          asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem + 256 + 8 * (i & 1)),
                       "l"(a_desc + 2 * (i & 3))
                       : "memory");
          asm volatile(
              "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem),
              "r"(tmem + 256 + 8 * (i & 1)), "l"(b_desc + 2 * (i & 3)), "r"(idesc), "r"(i > 0 ? 1u : 0u)
              : "memory");
        } else if (a_tmem) {
          asm volatile(
              "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem),
              "r"(tmem + 256), "l"(b_desc + 2 * (i & 3)), "r"(idesc), "r"(i > 0 ? 1u : 0u)
              : "memory");
        } else {
          umma_f16(tmem, a_desc + 2 * (i & 3), b_desc + 2 * (i & 3), idesc, i > 0 ? 1u : 0u);
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (elect_one()) out[0] = clock64() - t0;
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

}  // namespace

extern "C" int dbir_debug_mma_rate(int32_t n, int32_t b_mn_major, int32_t iters, int32_t a_in_tmem, void* out_cycles,
                                   void* stream) {
  const int smem = 16384 + 32768 + 1024;
  static bool configured = false;
  if (!configured) {
    DBIR_CHECK_CUDA(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  mma_rate_kernel<<<1, 64, smem, reinterpret_cast<cudaStream_t>(stream)>>>(n, b_mn_major, iters, a_in_tmem,
                                                                         reinterpret_cast<long long*>(out_cycles));
  DBIR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

#endif  // DBIR_DEBUG_PROBES
