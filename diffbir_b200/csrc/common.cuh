// diffbir_b200 — shared device/host helpers for the sm_100a kernels.
//
// Everything here is hand-written PTX for Blackwell (mbarrier, TMA, tcgen05/TMEM);
// no CUTLASS/CuTe types. Encodings follow the PTX ISA; the bit layouts of the
// UMMA shared-memory descriptor and instruction descriptor are restated in
// DESIGN.md §"tcgen05 descriptors".
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

// ---------------------------------------------------------------------------
// 16-bit tensor-core operand format. fp16 (11-bit significand) by default:
// same tcgen05 kind::f16 throughput as bf16, 8x smaller operand rounding error,
// which is what lets the 50-step trajectory stay >= 50 dB from the fp32
// reference. Build with -DDBIR_OPERAND_BF16 for bf16 operands.
// Accumulation is always fp32 (TMEM); residual streams are fp32 in HBM.
// ---------------------------------------------------------------------------
#ifdef DBIR_OPERAND_BF16
typedef __nv_bfloat16 op_t;
typedef __nv_bfloat162 op2_t;
#define DBIR_OPERAND_KIND 0
#define DBIR_UMMA_FMT 1u
__device__ __forceinline__ op_t f2op(float x) { return __float2bfloat16_rn(x); }
__device__ __forceinline__ float op2f(op_t x) { return __bfloat162float(x); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
#else
typedef __half op_t;
typedef __half2 op2_t;
#define DBIR_OPERAND_KIND 1
#define DBIR_UMMA_FMT 0u
__device__ __forceinline__ op_t f2op(float x) { return __float2half_rn(x); }
__device__ __forceinline__ float op2f(op_t x) { return __half2float(x); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  __half2 v = *reinterpret_cast<__half2*>(&u);
  return __half22float2(v);
}
#endif

// ---------------------------------------------------------------------------
// Host-side error plumbing (C ABI returns int; message via dbir_last_error()).
// ---------------------------------------------------------------------------
void dbir_set_error(const char* fmt, ...);
#define DBIR_CHECK_CUDA(expr)                                                     \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      dbir_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                \
                     cudaGetErrorString(_e));                                     \
      return -1;                                                                  \
    }                                                                             \
  } while (0)
#define DBIR_REQUIRE(cond, ...)                                                   \
  do {                                                                            \
    if (!(cond)) {                                                                \
      dbir_set_error(__VA_ARGS__);                                                \
      return -2;                                                                  \
    }                                                                             \
  } while (0)

// Encodes a tiled TMA descriptor (driver entry point fetched at run time, so the
// library does not link libcuda). rank <= 5; dims/box innermost first; strides in
// bytes for dims 1..rank-1. swizzle: 0 none, 1 128B, 2 64B, 3 32B.
int dbir_make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, int elem_bytes,
                   int swizzle128);
extern "C" int dbir_sm_count(void);

#ifdef __CUDACC__
// ---------------------------------------------------------------------------
// Programmatic dependent launch (PDL): every kernel is launched with the programmatic-stream-serialization
// attribute (on by default, DBIR_PDL=0 turns it off); each one calls pdl_trigger() at entry (lets the NEXT
// kernel's CTAs be scheduled and run their prologue: barrier init, TMEM alloc, descriptor prefetch) and
// pdl_wait() before touching memory written by its predecessors (griddepcontrol.wait returns when all
// prerequisite grids completed and flushed). Measured on B200 inside the captured forward graph (round 2,
// profiles/r02_forward_ab_4.jsonl): 5.80 ms with it, 5.76 ms without -- the ~200 KB shared-memory GEMM CTAs
// cannot co-reside with their successor's CTAs, so there is little to overlap; it is harmless and stays on.
// ---------------------------------------------------------------------------
extern "C" int dbir_pdl_enabled(void);
template <typename... KArgs, typename... Args>
inline cudaError_t dbir_launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                       cudaStream_t st, unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (dbir_pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t dbir_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                               cudaStream_t st, Args&&... args) {
  return dbir_launch_cluster(kernel, grid, block, smem, st, 1u, static_cast<Args&&>(args)...);
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a fully converged warp (deterministic for a given mask). Single-thread work
// (TMA, tcgen05.mma, tcgen05.commit) sits in `if (elect_one())` inside warp-uniform loops: the
// compiler then keeps addresses / descriptors in uniform registers instead of wrapping every
// such instruction in a per-lane serialisation loop, which a plain `if (lane == 0)` region costs.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier --------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
               : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug traps (-> launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz
      printf("dbir: mbarrier timeout block(%d,%d,%d) thread %d\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}

// ---- proxy / tcgen05 fences ------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- TMA -------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

// ---- TMEM ------------------------------------------------------------------
// Whole-warp (.sync.aligned) calls.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 32-bit, 16 / 32 consecutive columns: thread t of the warp reads TMEM lane
// (warp%4)*32 + t. taddr = (lane << 16) | column.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,"
      "%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,"
      "%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
      "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
      "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ---- UMMA ------------------------------------------------------------------
// Shared-memory matrix descriptor, 128B swizzle, rows of 128 bytes, 8-row groups
// 1024 bytes apart (SBO). Works for K-major operands ([rows][64 x 16-bit]) and for
// MN-major operands ([k rows][64 x 16-bit of the MN dimension]) alike; `lbo_bytes`
// only matters for MN-major operands wider than 64 elements.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes = 16) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((1024u >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16, fp32 accumulate, M=128.
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t n, uint32_t a_mn_major,
                                                  uint32_t b_mn_major, uint32_t m = 128u) {
  return (1u << 4) | (DBIR_UMMA_FMT << 7) | (DBIR_UMMA_FMT << 10) | (a_mn_major << 15) |
         (b_mn_major << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives on the mbarrier once all previously issued MMAs of this thread completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
                   "r"(smem_u32(bar))
               : "memory");
}

// ---- CTA pair (cta_group::2): two CTAs of a cluster on the two SMs of a TPC run one M=256 MMA;
// each CTA holds its own 128 accumulator rows in its TMEM and stages its 128 rows of A plus HALF
// of the B tile, so the B operand crosses the L2 fabric once per pair instead of once per CTA.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// TMA loads whose completion is signalled on a barrier of EITHER CTA of the pair (`bar_cluster` is
// a shared::cluster address, normally the leader's barrier)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster,
                                                 int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "{\n.reg .b16 m;\nmov.b16 m, 3;\n"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n}" ::
          "r"(smem_u32(bar))
      : "memory");
}

// ---- misc math -------------------------------------------------------------
// Packed fp32 pair arithmetic (FFMA2 / FADD2 / FMUL2 on sm_100): one issue slot for two lanes of work.
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1,
                                      float c0, float c1) {
  asm("{ .reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n mov.b64 rc, {%6,%7};\n"
      " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0,%1}, rd; }"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{ .reg .b64 ra, rb, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n"
      " add.rn.f32x2 rd, ra, rb;\n mov.b64 {%0,%1}, rd; }"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fmul2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{ .reg .b64 ra, rb, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n"
      " mul.rn.f32x2 rd, ra, rb;\n mov.b64 {%0,%1}, rd; }"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
// GELU(x) = x * Phi(x) with the exact-erf definition (F.gelu default; attention.py:19-45, swinir.py:27).
// erf by Abramowitz-Stegun 7.1.26 (|abs error| <= 1.5e-7, below fp32 erff's own ulp for the 16-bit
// outputs these epilogues feed): branch-free, one MUFU.RCP + one MUFU.EX2 + 9 FMA-pipe ops, where
// libdevice's erff costs ~30 instructions and a divergent branch per element -- the GEGLU / MLP
// epilogues (K = C, so the mainloop is only 5-20 k-blocks long) were bound by it.
__device__ __forceinline__ float gelu_erf_f(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = ex2_approx(-1.4426950408889634f * z * z);
  const float half_erfc = 0.5f * p * e;                 // 0.5 * erfc(|x| / sqrt 2)
  // Phi(x) = 1 - half_erfc for x >= 0, half_erfc for x < 0
  const float phi = x >= 0.f ? 1.0f - half_erfc : half_erfc;
  return x * phi;
}
// Two GELUs per call on packed fp32 pairs (FFMA2 / FMUL2): same A-S 7.1.26 polynomial as gelu_erf_f, half the
// FMA-pipe issue slots -- the GEGLU epilogue (K = C: the mainloop is only 5-20 k-blocks long) is bound by them.
__device__ __forceinline__ void gelu_erf_f2(float x0, float x1, float& y0, float& y1) {
  float z0 = fabsf(x0), z1 = fabsf(x1);
  fmul2(z0, z1, z0, z1, 0.70710678118654752f, 0.70710678118654752f);
  float d0, d1;
  ffma2(d0, d1, z0, z1, 0.3275911f, 0.3275911f, 1.0f, 1.0f);
  float t0, t1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  float p0, p1;
  ffma2(p0, p1, t0, t1, 1.061405429f, 1.061405429f, -1.453152027f, -1.453152027f);
  ffma2(p0, p1, p0, p1, t0, t1, 1.421413741f, 1.421413741f);
  ffma2(p0, p1, p0, p1, t0, t1, -0.284496736f, -0.284496736f);
  ffma2(p0, p1, p0, p1, t0, t1, 0.254829592f, 0.254829592f);
  fmul2(p0, p1, p0, p1, t0, t1);
  float q0, q1;
  fmul2(q0, q1, z0, z1, z0, z1);
  fmul2(q0, q1, q0, q1, -1.4426950408889634f, -1.4426950408889634f);
  const float e0 = ex2_approx(q0), e1 = ex2_approx(q1);
  float h0, h1;
  fmul2(h0, h1, p0, p1, e0, e1);
  fmul2(h0, h1, h0, h1, 0.5f, 0.5f);                    // 0.5 * erfc(|x| / sqrt 2)
  const float phi0 = x0 >= 0.f ? 1.0f - h0 : h0, phi1 = x1 >= 0.f ? 1.0f - h1 : h1;
  fmul2(y0, y1, x0, x1, phi0, phi1);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif  // __CUDACC__
