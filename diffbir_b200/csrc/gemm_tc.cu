// tcgen05 / TMEM / TMA GEMM with an implicit-GEMM 3x3 convolution mode.
//
//   D[M, N] = A[M, K] * B[N, K]^T   (16-bit operands, fp32 accumulate in TMEM)
//
// A is either a plain row-major [M, K] matrix (Linear layers, 1x1 convs on NHWC
// activations) or an NHWC activation tensor read through a 4-D TMA box, one box per
// filter tap, with the hardware's out-of-bounds zero fill providing the conv padding
// (replaces nn.Conv2d 3x3/pad 1 at reference unet.py:149-153,173-180, vae.py:77-86,
// swinir.py:472,797-811). B is the packed weight [N, K] (K = taps*C, tap-major).
//
// One CTA computes a 128 x BN tile: warp 0 = TMA producer, warp 1 = MMA issuer,
// warp 2 = TMEM allocator + epilogue, warps 3..5 = epilogue (thread == accumulator row).
//
// Epilogue (per warp, 32 rows x 32 columns at a time): tcgen05.ld -> bias / per-image row
// vector (time embedding) / GELU / LeakyReLU / SiLU / GEGLU gating / alpha -> + fp32 residual
// tile (TMA-loaded, double buffered) -> swizzled shared-memory tile -> TMA bulk store
// (double buffered). Threads never touch global memory for the output: with one epilogue warp
// per scheduler every extra instruction is exposed latency, and the TMA unit also clips
// overhanging rows / columns (M, N tails, conv tiles that overhang the image).
//   out = residual + alpha * act(acc + bias + rowvec),  fp32 or 16-bit, any row stride.
#include "common.cuh"
#include "../../include/diffbir_b200.h"
#include <stdlib.h>
#include <array>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                 // 64 x 16-bit = 128 B = one swizzle atom row
constexpr int A_SUB_BYTES = BM * BK * 2;    // one 64-wide k-block of A
constexpr int EPI_TILE_BYTES = 32 * 32 * 4;   // one warp's 32x32 fp32 staging tile

struct GemmParams {
  int M, N, num_kb;
  // conv geometry (mode 1)
  int mode;                // 0 plain, 1 conv taps over NHWC
  int H, W, NI;            // image size / count
  int bw, bh, bni;         // box extents (bw*bh*bni == 128)
  int tiles_x, tiles_y;    // tiles per image row / column
  int cblocks;             // C / 64
  int kw, pad;             // filter width (taps = kw*kw), padding
  int wbw, wbh;            // per-warp sub-box (32 rows) extents in w / h
  // epilogue
  int out_kind;            // 0 fp32, 1 16-bit operand
  const float* bias;
  const float* rowvec;     // [NI or M/rows_per_vec, N]
  int rows_per_vec;
  int has_residual;
  float alpha;
  int act;                 // 0 none, 1 gelu(erf), 2 leaky relu (slope in act_param), 3 silu
  float act_param;
  int geglu;               // 1: tile holds [BN/2 values | BN/2 gates]; output width N/2
  int bias_per_row;        // bias indexed by output row instead of column
  // grouped launch: G same-shape problems back to back along M; group g's weights / bias start at row g*N
  int group_span;          // m-tiles (plain) or images (conv) per group; 0 = one problem
  // split-K: grid.z CTAs share one output tile; partial tiles go through `ws`, the last CTA to
  // arrive (ticket) sums them in fixed order (deterministic) and runs the epilogue.
  int splits, kb_per_split;
  float* ws;
  unsigned int* tickets;
  long long* dbg;          // optional per-CTA clock64 stamps [ctas][4]: start, setup done, acc ready, end
  // fused GroupNorm statistics of the OUTPUT tensor: per (image, 32-row slot, channel) sum / sum of
  // squares over the warp's rows -> stats[(img * nslots + slot) * N + col][2] (fp32 outputs only)
  float* stats;
  int stats_nslots, stats_rows_per_img, stats_slots_x;
  // L2 prefetch of memory a LATER kernel will stream (the next layers' weights): issued by the
  // epilogue warps while they wait for the mainloop
  const char* pf_ptr;
  long long pf_bytes;
};

__device__ __forceinline__ float apply_act(float v, int act, float prm) {
  if (act == 1) return gelu_erf_f(v);
  if (act == 2) return v > 0.f ? v : v * prm;
  return silu_f(v);
}

__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// Per-warp epilogue state: which 32 rows of the tile, their sub-box origin, staging buffers.
struct EpiCtx {
  const CUtensorMap* tm_out;
  const CUtensorMap* tm_res;
  uint32_t epi_base;       // shared address of this warp's 4 staging tiles: out[0], out[1], res[0], res[1]
  uint32_t res_bar;        // shared address of this warp's two residual-load mbarriers
  int row0;                // plain mode: first global row of this warp
  int x, y, n;             // conv mode: origin of this warp's sub-box
  long long stats_base;    // element offset of this warp's (image, slot) row in the stats buffer, < 0: none
  int grp_off;             // grouped launch: row offset of this CTA's problem in the stacked bias
  __device__ __forceinline__ uint32_t out_buf(int b) const { return epi_base + b * EPI_TILE_BYTES; }
  __device__ __forceinline__ uint32_t res_buf(int b) const { return epi_base + (2 + b) * EPI_TILE_BYTES; }
  __device__ __forceinline__ uint32_t bar(int b) const { return res_bar + b * 8; }
};

__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > 4000000000LL) { printf("dbir: residual mbarrier timeout\n"); __trap(); }
  }
}

__device__ __forceinline__ void epi_issue_residual(const GemmParams& p, EpiCtx& e, int lane, int chunk_idx,
                                                   int gc0) {
  // lane 0 only
  if (lane == 0) {
    const int b = chunk_idx & 1;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(e.bar(b)), "r"(EPI_TILE_BYTES) : "memory");
    if (p.mode == 0) {
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
              "r"(e.res_buf(b)), "l"(reinterpret_cast<uint64_t>(e.tm_res)), "r"(e.bar(b)), "r"(gc0), "r"(e.row0)
          : "memory");
    } else {
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
              "r"(e.res_buf(b)), "l"(reinterpret_cast<uint64_t>(e.tm_res)), "r"(e.bar(b)), "r"(gc0), "r"(e.x),
          "r"(e.y), "r"(e.n)
          : "memory");
    }
  }
}

// Finishes one 32-column chunk held as fp32 in v[] (thread == row): epilogue math, residual,
// staging, TMA store. `chunk_idx` counts chunks processed by this warp (buffer parity),
// `gc0` is the first *output* column, `next_gc0` (< 0: none) the column of the chunk after next
// whose residual tile can start loading into the buffer this chunk frees.
__device__ __forceinline__ void finish_chunk(const GemmParams& p, EpiCtx& e, float* v, int lane, int chunk_idx,
                                             int gc0, int ncols, int next_gc0, long long out_row, int vec_idx,
                                             bool raw_math, bool have_pre = false, float pre = 0.f) {
  const int b = chunk_idx & 1;
  if (raw_math && have_pre) {
    // bias + row vector of this chunk were fetched while the mainloop ran (lane == column)
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, pre, j);
    if (p.act) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act, p.act_param);
    }
    if (p.alpha != 1.0f) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
    }
  } else if (raw_math) {
    if (p.bias && p.bias_per_row) {
      const float bv = out_row >= 0 ? __ldg(p.bias + out_row) : 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += bv;
    } else if (p.bias || p.rowvec) {
      // acc + (bias + rowvec): the same association as the prefetched path above, so a sample's bits do not
      // depend on whether its warp's 32 rows share one row vector (tiny images: several images per warp)
      const bool rv_on = p.rowvec && out_row >= 0;
      const float* rv = p.rowvec + static_cast<long long>(vec_idx) * p.N + gc0;
      if (ncols == 32) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 t = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + e.grp_off + gc0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (rv_on) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(rv + j));
            t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w;
          }
          v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ncols) {
            float t = p.bias ? __ldg(p.bias + e.grp_off + gc0 + j) : 0.f;
            if (rv_on) t += __ldg(rv + j);
            v[j] += t;
          }
      }
    }
    if (p.act) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act, p.act_param);
    }
    if (p.alpha != 1.0f) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
    }
  }
  const int sw = lane & 7;
  if (p.has_residual) {
    mbar_wait_s(e.bar(b), (chunk_idx >> 1) & 1);
    const uint32_t rrow = e.res_buf(b) + lane * 128;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = lds_v4(rrow + ((j ^ sw) << 4));
      v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
    }
    fence_proxy_async_smem();                       // generic reads before the async overwrite
    __syncwarp();                                   // every lane has read the residual tile
    if (next_gc0 >= 0) epi_issue_residual(p, e, lane, chunk_idx + 2, next_gc0);
  }
  // the store that used this staging buffer two chunks ago must have finished reading it
  if (lane == 0 && chunk_idx >= 2) bulk_wait_read<1>();
  __syncwarp();
  if (p.out_kind == 0) {
    const uint32_t orow = e.out_buf(b) + lane * 128;          // 128-byte rows, SWIZZLE_128B
#pragma unroll
    for (int j = 0; j < 8; ++j)
      sts_v4(orow + ((j ^ sw) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
             __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
  } else {
    const uint32_t orow = e.out_buf(b) + lane * 64;           // 64-byte rows, SWIZZLE_64B
    const int sw2 = (lane >> 1) & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts_v4(orow + ((j ^ sw2) << 4), pack2(v[8 * j], v[8 * j + 1]), pack2(v[8 * j + 2], v[8 * j + 3]),
             pack2(v[8 * j + 4], v[8 * j + 5]), pack2(v[8 * j + 6], v[8 * j + 7]));
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (p.stats && e.stats_base >= 0 && p.out_kind == 0) {
    // column sums over the warp's valid rows, read back from the staged tile (lane == column)
    const uint32_t valid = __ballot_sync(0xffffffffu, out_row >= 0);
    float sa = 0.f, sb = 0.f;
    const uint32_t col_off = static_cast<uint32_t>(lane & 3) * 4;
    const int chunk = lane >> 2;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float x;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(e.out_buf(b) + i * 128 + (((chunk ^ (i & 7))) << 4) + col_off) : "memory");
      if ((valid >> i) & 1u) { sa += x; sb += x * x; }
    }
    if (lane < ncols)
      *reinterpret_cast<float2*>(p.stats + (e.stats_base + gc0 + lane) * 2) = make_float2(sa, sb);
  }
  if (lane == 0) {
    if (p.mode == 0) tma_store_2d(e.tm_out, e.out_buf(b), gc0, e.row0);
    else tma_store_4d(e.tm_out, e.out_buf(b), gc0, e.x, e.y, e.n);
    bulk_commit();
  }
}

// EW = number of epilogue warps: 4 (two CTAs per SM, epilogue of one overlaps the mainloop of the
// other) or 8 (grids that leave <= 1 CTA per SM: two warps per TMEM lane quarter take alternate
// 32-column chunks, halving the exposed epilogue latency).
// PAIR: the two CTAs of a (2,1,1) cluster (adjacent M tiles, same N tile) run M=256 MMAs issued by
// the even CTA (cta_group::2); each CTA stages its own A rows and half of the B tile.
// KSUB: 64-wide k-blocks per pipeline stage. The producer / MMA-issue loops cost ~350-500 cycles per
// iteration in barrier, TMA-issue and commit latency (measured), more than the MMAs of one k-block
// take for BN < 256; two k-blocks per stage amortise that.
template <int BN, int STAGES, int EW, bool PAIR, int KSUB>
__global__ void __launch_bounds__(64 + 32 * EW, EW == 4 ? 2 : 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_res,
               const GemmParams p) {
  constexpr int B_ROWS = PAIR ? BN / 2 : BN;          // B rows staged by this CTA
  constexpr int B_SUB_BYTES = B_ROWS * BK * 2;        // one k-block of B
  constexpr int A_STAGE_BYTES = KSUB * A_SUB_BYTES;
  constexpr int B_STAGE_BYTES = KSUB * B_SUB_BYTES;
  constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static_assert(STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) >= EW * 4 * EPI_TILE_BYTES, "epilogue staging must fit in the pipeline stages");
  constexpr int CSTEP = 32 * (EW / 4);        // column stride between the chunks of one warp
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint64_t* res_bars = tmem_full + 1;                 // [EW warps][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bars + 2 * EW);
  uint32_t* split_flag = tmem_slot + 1;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const int m_tile = blockIdx.x;
  const int n_tile = blockIdx.y;
  const int tile_lin = blockIdx.y * gridDim.x + blockIdx.x;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
  long long t_start = 0, t_setup = 0, t_acc = 0;
  if (p.dbg) t_start = clock64();

  // tile origin
  int n0 = 0, y0 = 0, x0 = 0;
  if (p.mode == 1) {
    int t = m_tile;
    int tx = t % p.tiles_x; t /= p.tiles_x;
    int ty = t % p.tiles_y; t /= p.tiles_y;
    x0 = tx * p.bw; y0 = ty * p.bh; n0 = t * p.bni;
  }

  // grouped launch: this CTA's problem -> row offset into the stacked weights / bias
  int grp_off = 0;
  if (p.group_span > 0) grp_off = ((p.mode == 1 ? n0 : m_tile) / p.group_span) * p.N;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_out);
    if (p.has_residual) tma_prefetch_desc(&tma_res);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    for (int i = 0; i < 2 * EW; ++i) mbar_init(&res_bars[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) tmem_alloc_pair(tmem_slot, TMEM_COLS);
    else tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  // PAIR: the peer's barriers and TMEM must be live before any remote complete_tx / MMA write
  if constexpr (PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();        // the next kernel may start its prologue
  pdl_wait();           // A operand / residual come from predecessor kernels
  if (p.dbg) t_setup = clock64();

  const int n_stage_iters = (kb1 - kb0 + KSUB - 1) / KSUB;
  if (warp == 0) {
    // ---------------- TMA producer: warp-uniform loop, one elected lane issues -------------
    const uint32_t fbar0 = PAIR ? mapa_u32(smem_u32(&full[0]), 0) : smem_u32(&full[0]);
    const int b_row = grp_off + n_tile * BN + static_cast<int>(cta_rank) * B_ROWS;
    // conv mode: (channel block, tap column, tap row) of the next k-block, advanced incrementally
    // (integer divisions in this loop cost more than the TMA issue itself)
    int cb = 0, tx = 0, ty = 0;
    if (p.mode == 1) {
      const int tap0 = kb0 / p.cblocks;
      cb = kb0 - tap0 * p.cblocks;
      ty = tap0 / p.kw;
      tx = tap0 - ty * p.kw;
    }
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < n_stage_iters; ++it) {
      // coordinates of this stage's k-blocks (all lanes, so they stay in uniform registers).
      // k-blocks past this CTA's range are fetched from out-of-range coordinates: the TMA unit
      // zero-fills them without memory traffic and the byte count per stage stays constant.
      int kc[KSUB], ac[KSUB], ax[KSUB], ay[KSUB];
#pragma unroll
      for (int j = 0; j < KSUB; ++j) {
        const int kb = kb0 + it * KSUB + j;
        const bool live = kb < kb1;
        kc[j] = live ? kb * BK : p.num_kb * BK;
        ac[j] = ax[j] = ay[j] = 0;
        if (p.mode == 1) {
          // conv: A is addressed by (channel block, tap offset) instead of k
          ac[j] = live ? cb * BK : p.cblocks * BK;
          ax[j] = x0 + tx - p.pad;
          ay[j] = y0 + ty - p.pad;
          if (++cb == p.cblocks) { cb = 0; if (++tx == p.kw) { tx = 0; ++ty; } }
        }
      }
      mbar_wait(&empty[s], ph ^ 1);
      if (elect_one()) {
        // PAIR: both CTAs' loads complete on the leader's barrier, which expects the bytes of both
        if (cta_rank == 0) mbar_expect_tx(&full[s], (PAIR ? 2 : 1) * (A_STAGE_BYTES + B_STAGE_BYTES));
        const uint32_t fbar = fbar0 + s * 8;
#pragma unroll
        for (int j = 0; j < KSUB; ++j) {
          uint8_t* a_dst = sA + s * A_STAGE_BYTES + j * A_SUB_BYTES;
          uint8_t* b_dst = sB + s * B_STAGE_BYTES + j * B_SUB_BYTES;
          if (p.mode == 0) {
            if constexpr (PAIR) tma_load_2d_pair(a_dst, &tma_a, fbar, kc[j], m_tile * BM);
            else tma_load_2d(a_dst, &tma_a, &full[s], kc[j], m_tile * BM);
          } else {
            if constexpr (PAIR) tma_load_4d_pair(a_dst, &tma_a, fbar, ac[j], ax[j], ay[j], n0);
            else tma_load_4d(a_dst, &tma_a, &full[s], ac[j], ax[j], ay[j], n0);
          }
          if constexpr (PAIR) tma_load_2d_pair(b_dst, &tma_b, fbar, kc[j], b_row);
          else tma_load_2d(b_dst, &tma_b, &full[s], kc[j], b_row);
        }
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (leader CTA of a pair only): warp-uniform loop -------------
    if (cta_rank == 0) {
      constexpr uint32_t idesc = umma_idesc(BN, 0, 0, PAIR ? 256u : 128u);
      const uint64_t a_desc0 = umma_desc_sw128(smem_u32(sA));
      const uint64_t b_desc0 = umma_desc_sw128(smem_u32(sB));
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < n_stage_iters; ++it) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          // descriptors advance in units of 16 bytes (start-address field, low bits)
          const uint64_t a_desc = a_desc0 + static_cast<uint64_t>((s * A_STAGE_BYTES) >> 4);
          const uint64_t b_desc = b_desc0 + static_cast<uint64_t>((s * B_STAGE_BYTES) >> 4);
#pragma unroll
          for (int j = 0; j < KSUB; ++j) {
            if (kb0 + it * KSUB + j < kb1) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                const uint64_t da = a_desc + static_cast<uint64_t>((j * A_SUB_BYTES) >> 4) + 2 * k;
                const uint64_t db = b_desc + static_cast<uint64_t>((j * B_SUB_BYTES) >> 4) + 2 * k;
                const uint32_t acc = (it > 0 || j > 0 || k > 0) ? 1u : 0u;
                if constexpr (PAIR) umma_f16_pair(tmem_base, da, db, idesc, acc);
                else umma_f16(tmem_base, da, db, idesc, acc);
              }
            }
          }
          // frees the smem stage (in both CTAs of a pair) once these MMAs retire
          if constexpr (PAIR) umma_commit_pair(&empty[s]);
          else umma_commit(&empty[s]);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      // accumulator complete
      if (elect_one()) {
        if constexpr (PAIR) umma_commit_pair(tmem_full);
        else umma_commit(tmem_full);
      }
    }
  } else {
    // ---------------- epilogue: warps 2.., TMEM lane quarter = warp % 4 -------------
    const int q = warp & 3;
    const int ew = warp - 2;                // epilogue warp index
    const int c_first = (ew >> 2) * 32;     // first chunk of this warp (EW == 8: warps 6..9 start at column 32)
    const int r = q * 32 + lane;            // accumulator row within the tile
    long long out_row = -1;                 // global output row (bias_per_row / rowvec index), -1 = masked
    int vec_idx = 0;
    EpiCtx e;
    e.tm_out = &tma_out; e.tm_res = &tma_res;
    e.grp_off = grp_off;
    e.res_bar = smem_u32(res_bars + ew * 2);
    e.row0 = m_tile * BM + q * 32;
    e.x = e.y = e.n = 0;
    if (p.mode == 0) {
      const long long gr = static_cast<long long>(m_tile) * BM + r;
      if (gr < p.M) { out_row = gr; vec_idx = static_cast<int>(gr / p.rows_per_vec); }
    } else {
      const int bx = r % p.bw;
      const int by = (r / p.bw) % p.bh;
      const int bn = r / (p.bw * p.bh);
      const int x = x0 + bx, y = y0 + by, n = n0 + bn;
      if (x < p.W && y < p.H && n < p.NI) {
        out_row = (static_cast<long long>(n) * p.H + y) * p.W + x;
        vec_idx = n;
      }
      const int r0 = q * 32;                // this warp's sub-box origin
      e.x = x0 + r0 % p.bw;
      e.y = y0 + (r0 / p.bw) % p.bh;
      e.n = n0 + r0 / (p.bw * p.bh);
    }
    e.stats_base = -1;
    if (p.stats) {
      if (p.mode == 0) {
        const int img = e.row0 / p.stats_rows_per_img, slot = (e.row0 % p.stats_rows_per_img) / 32;
        if (e.row0 < p.M) e.stats_base = (static_cast<long long>(img) * p.stats_nslots + slot) * p.N;
      } else if (e.x < p.W && e.y < p.H && e.n < p.NI) {
        const int slot = (e.y / p.wbh) * p.stats_slots_x + e.x / p.wbw;
        e.stats_base = (static_cast<long long>(e.n) * p.stats_nslots + slot) * p.N;
      }
    }
    if (p.pf_bytes > 0) {
      // one 128-byte line per prefetch; lines are spread over every epilogue thread of the grid
      const long long nthr = static_cast<long long>(gridDim.x) * gridDim.y * gridDim.z * (32 * EW);
      const long long me = ((static_cast<long long>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * (32 * EW) +
                           (threadIdx.x - 64);
      for (long long off = me * 128; off < p.pf_bytes; off += nthr * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.pf_ptr + off));
    }
    // Per-column addends (bias + the row vector of this warp's image) of every chunk this warp will
    // finish, fetched now so that their L2 latency hides behind the mainloop: lane == column, one
    // register per chunk. Needs one row vector for the whole warp (images of >= 32 rows).
    constexpr int MAX_PRE = BN / CSTEP + 1;
    float pre[MAX_PRE];
    bool have_pre = false;
    if (!p.geglu && !p.bias_per_row && (p.bias || p.rowvec)) {
      const bool any_row = __any_sync(0xffffffffu, out_row >= 0);
      // masked lanes of lane 0 would make v0 meaningless: take the vector of the first valid row
      const unsigned vm = __ballot_sync(0xffffffffu, out_row >= 0);
      const int vsel = any_row ? __shfl_sync(0xffffffffu, vec_idx, __ffs(vm) - 1) : 0;
      const bool uniform = !p.rowvec || !any_row || __all_sync(0xffffffffu, out_row < 0 || vec_idx == vsel);
      if (uniform) {
        have_pre = true;
#pragma unroll
        for (int i = 0; i < MAX_PRE; ++i) {
          const int col = n_tile * BN + c_first + i * CSTEP + lane;
          float x = 0.f;
          if (c_first + i * CSTEP < BN && col < p.N) {
            if (p.bias) x = __ldg(p.bias + grp_off + col);
            if (p.rowvec && any_row) x += __ldg(p.rowvec + static_cast<long long>(vsel) * p.N + col);
          }
          pre[i] = x;
        }
      }
    }
    // GEGLU: value / gate biases of every chunk this warp finishes (lane == column), fetched while the
    // mainloop runs and broadcast with shuffles (the per-element __ldg pair this replaces sat on the
    // epilogue's critical path)
    constexpr int MAX_G = (BN / 2) / CSTEP + 1;
    float pa[MAX_G], pg[MAX_G];
    if (p.geglu) {
#pragma unroll
      for (int i = 0; i < MAX_G; ++i) {
        const int c = c_first + i * CSTEP + lane;
        const bool live = p.bias && c_first + i * CSTEP < BN / 2 && n_tile * (BN / 2) + c < p.N / 2;
        pa[i] = live ? __ldg(p.bias + grp_off + n_tile * BN + c) : 0.f;
        pg[i] = live ? __ldg(p.bias + grp_off + n_tile * BN + BN / 2 + c) : 0.f;
      }
    }
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    if (p.dbg) t_acc = clock64();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // all MMAs have retired -> the pipeline stages are free: 4 staging tiles (4 KB) per warp
    e.epi_base = smem_u32(sA) + ew * (4 * EPI_TILE_BYTES);

    constexpr int OUT_COLS = BN;            // accumulator columns that produce output (GEGLU: BN/2)
    const bool finalize = p.splits == 1;
    if (!p.geglu) {
      const int col_base = n_tile * BN;
      // prefetch the first two residual tiles
      if (finalize && p.has_residual) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int cc = c_first + c * CSTEP;
          if (cc < OUT_COLS && col_base + cc < p.N) epi_issue_residual(p, e, lane, c, col_base + cc);
        }
      }
      int ci = 0;
#pragma unroll 1
      for (int c0 = c_first; c0 < BN; c0 += CSTEP) {
        uint32_t acc[32];
        __syncwarp();
        tmem_ld32(taddr + c0, acc);
        tmem_ld_wait();
        const int gc0 = col_base + c0;
        if (p.splits > 1) {
          // raw partial -> workspace [tile][split][col][row]: coalesced across the warp's rows
          float* wcol = p.ws + ((static_cast<long long>(tile_lin) * p.splits + blockIdx.z) * BN + c0) * BM + r;
#pragma unroll
          for (int j = 0; j < 32; ++j) __stcg(wcol + j * BM, __uint_as_float(acc[j]));
          continue;
        }
        if (gc0 >= p.N) continue;
        const int ncols = min(32, p.N - gc0);
        const int ngc = gc0 + 2 * CSTEP;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
        finish_chunk(p, e, v, lane, ci, gc0, ncols, (c0 + 2 * CSTEP < BN && ngc < p.N) ? ngc : -1, out_row, vec_idx, true,
                     have_pre, pre[(c0 - c_first) / CSTEP]);
        ++ci;
      }
      if (p.splits > 1) {
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
        if (threadIdx.x == 64) {
          const unsigned int t = atomicAdd(&p.tickets[tile_lin], 1u);
          const bool last = (t == static_cast<unsigned int>(p.splits) - 1);
          if (last) p.tickets[tile_lin] = 0;        // self-reset for the next launch
          *split_flag = last ? 1u : 0u;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
        if (*split_flag) {
          __threadfence();
          if (p.has_residual) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int cc = c_first + c * CSTEP;
              if (cc < BN && col_base + cc < p.N) epi_issue_residual(p, e, lane, c, col_base + cc);
            }
          }
          int cj = 0;
#pragma unroll 1
          for (int c0 = c_first; c0 < BN; c0 += CSTEP) {
            const int gc0 = col_base + c0;
            if (gc0 >= p.N) continue;
            const int ncols = min(32, p.N - gc0);
            const int ngc = gc0 + 2 * CSTEP;
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
            for (int z = 0; z < p.splits; ++z) {     // fixed summation order
              const float* wcol = p.ws + ((static_cast<long long>(tile_lin) * p.splits + z) * BN + c0) * BM + r;
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += __ldcg(wcol + j * BM);
            }
            finish_chunk(p, e, v, lane, cj, gc0, ncols, (c0 + 2 * CSTEP < BN && ngc < p.N) ? ngc : -1, out_row, vec_idx, true,
                         have_pre, pre[(c0 - c_first) / CSTEP]);
            ++cj;
          }
        }
      }
    } else {
      // GEGLU: tile columns [0, BN/2) are values, [BN/2, BN) the matching gates
      // (weights are packed that way); output column = n_tile*BN/2 + c.
      constexpr int HB = BN / 2;
      const int n_half = p.N / 2;
      if constexpr (HB % 32 == 0) {
        // Value / gate biases of this warp's chunks go through its (unused: GEGLU has no residual) residual
        // staging tile and come back as broadcast LDS.128: 16 loads per chunk instead of 64 shuffles, and the
        // gating math runs on packed fp32 pairs -- the epilogue's FMA-pipe slots bound this kernel (K = C).
        const uint32_t bb = e.res_buf(0);
#pragma unroll
        for (int i = 0; i < MAX_G; ++i) {
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(bb + (i * 64 + lane) * 4), "f"(pa[i]) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(bb + (i * 64 + 32 + lane) * 4), "f"(pg[i]) : "memory");
        }
        __syncwarp();
        int ci = 0;
#pragma unroll
        for (int i = 0; i < MAX_G; ++i) {
          const int c0 = c_first + i * CSTEP;
          if (c0 >= HB) break;
          uint32_t av[32], ag[32];
          __syncwarp();
          tmem_ld32(taddr + c0, av);
          tmem_ld32(taddr + HB + c0, ag);
          tmem_ld_wait();
          const int gc0 = n_tile * HB + c0;   // output column
          if (gc0 >= n_half) continue;
          const int ncols = min(32, n_half - gc0);
          float v[32];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const float4 ba = lds_v4(bb + i * 256 + jj * 16), bg = lds_v4(bb + i * 256 + 128 + jj * 16);
            float a0, a1, a2, a3, g0, g1, g2, g3, y0, y1, y2, y3;
            fadd2(a0, a1, __uint_as_float(av[4 * jj]), __uint_as_float(av[4 * jj + 1]), ba.x, ba.y);
            fadd2(a2, a3, __uint_as_float(av[4 * jj + 2]), __uint_as_float(av[4 * jj + 3]), ba.z, ba.w);
            fadd2(g0, g1, __uint_as_float(ag[4 * jj]), __uint_as_float(ag[4 * jj + 1]), bg.x, bg.y);
            fadd2(g2, g3, __uint_as_float(ag[4 * jj + 2]), __uint_as_float(ag[4 * jj + 3]), bg.z, bg.w);
            gelu_erf_f2(g0, g1, y0, y1);
            gelu_erf_f2(g2, g3, y2, y3);
            fmul2(v[4 * jj], v[4 * jj + 1], a0, a1, y0, y1);
            fmul2(v[4 * jj + 2], v[4 * jj + 3], a2, a3, y2, y3);
          }
          finish_chunk(p, e, v, lane, ci, gc0, ncols, -1, out_row, vec_idx, false);
          ++ci;
        }
      }
    }
    if (lane == 0) bulk_wait_read<0>();     // staging tiles must outlive the bulk stores' reads
    __syncwarp();
    tc_fence_before();
    if (p.dbg && threadIdx.x == 64) {
      long long* d = p.dbg + 8LL * ((static_cast<long long>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x);
      d[0] = t_start; d[1] = t_setup; d[2] = t_acc; d[3] = clock64();
    }
  }
  if constexpr (PAIR) cluster_sync_all();     // neither CTA may retire while the other can still signal it
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BN, int STAGES, int EW, bool PAIR = false, int KSUB = 1>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tr,
           const GemmParams& p, dim3 grid, cudaStream_t st) {
  constexpr int smem = STAGES * KSUB * (A_SUB_BYTES + (PAIR ? BN / 2 : BN) * BK * 2) + 1024 + 512;
  static_assert(smem <= 227 * 1024, "pipeline does not fit in shared memory");
  static bool configured = false;
  if (!configured) {
    DBIR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, EW, PAIR, KSUB>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  DBIR_CHECK_CUDA(dbir_launch_cluster(gemm_tc_kernel<BN, STAGES, EW, PAIR, KSUB>, grid, dim3(64 + 32 * EW), smem, st,
                                      PAIR ? 2u : 1u, ta, tb, to, tr, p));
  return 0;
}

// Tile width, CTA pairing and split-K factor from a cost model in SM cycles, calibrated on B200
// with the per-CTA clock64 stamps (tools/gpu_pair_timeline.py, profiles/r01_probes.txt):
//   * one 64-wide k-block costs max(MMA time, loop overhead): the MMAs take ~2*BN cycles
//     (128 x BN x 64 at 128 x 256 x 16 per 128 cycles), the producer / issue loops ~520 cycles per
//     stage iteration (barrier, TMA-issue and commit latencies) -> ~290 per k-block with two
//     k-blocks per stage (KSUB = 2);
//   * two CTAs on one SM share its tensor pipe but hide each other's overhead and epilogue;
//   * epilogue ~1500 + 28*BN cycles with 8 warps, ~500 + 55*BN with 4; split-K adds a workspace
//     round trip; CTAs are spread over `sms` SMs, a partial last wave costs a full one.
struct TilePlan { int bn, splits, kb_per_split, pair; };

__host__ int plan_ksub(int bn, int pair, bool wide) {
  if (pair) return (bn == 256 && !wide) ? 1 : 2;
  if (bn == 256) return 1;
  return (wide || bn <= 64) ? 2 : 1;
}

struct PlanQuery {
  int m_tiles, N, num_kb, geglu, forced_bn, split_req, pair_req;
  long long ws_floats_avail;
};

// Calls f(plan, modelled cost) for every plan the kernel family can run for this problem.
template <typename F>
void for_each_plan(const PlanQuery& q, F f) {
  const int cands[5] = {256, 160, 128, 64, 32};
  const int sms = dbir_sm_count();
  // CTA pairs (cta_group::2): DBIR_GEMM_PAIR = 0 never, 1 whenever the shape allows, unset: free choice
  static const int pair_env = [] { const char* e = getenv("DBIR_GEMM_PAIR"); return e ? atoi(e) : -1; }();
  const int pair_mode = q.pair_req == 1 ? 1 : q.pair_req == 2 ? 0 : pair_env;
  const bool can_pair = pair_mode != 0 && !(q.m_tiles & 1);
  // experiment switch (off: wide tiles only when they divide N): SS-mode MMAs run at 75-80 % of peak
  // only for N = 256, so a ragged last tile can still pay for N = 640 / 1920
  static const int ragged_wide = [] { const char* e = getenv("DBIR_GEMM_RAGGED_WIDE"); return e ? atoi(e) : 0; }();
  for (int i = 0; i < 10; ++i) {
    const int bn = cands[i % 5];
    const int pair = 1 - i / 5;
    if (pair && (!can_pair || bn < 64)) continue;
    if (!pair && pair_mode == 1 && can_pair && bn >= 64) continue;     // forced pairing
    if (q.forced_bn > 0 && bn != q.forced_bn) continue;
    if (q.forced_bn <= 0) {
      if (q.geglu && bn < 64) continue;
      if (bn > 64 && q.N % bn != 0 && !(ragged_wide && q.N > bn)) continue;   // ragged N only with the narrow tiles
    }
    const long long tiles = static_cast<long long>(q.m_tiles) * ((q.N + bn - 1) / bn);
    for (int s = 1; s <= 16; ++s) {
      if (s > 1) {
        if (q.split_req == 1 || q.geglu || q.ws_floats_avail <= 0) break;
        if (q.num_kb / s < 6) break;
        if (tiles > 16384 || tiles * s * 128LL * bn > q.ws_floats_avail) break;
        if (tiles * (s - 1) >= 2LL * sms) break;       // already more than two CTAs per SM without it
      }
      if (q.split_req > 1 && s != q.split_req) continue;
      const int kbs = (q.num_kb + s - 1) / s;
      if (static_cast<long long>(kbs) * (s - 1) >= q.num_kb) continue;   // an empty split
      const long long ctas = tiles * s;
      const bool wide = ctas <= sms;                       // one CTA per SM, 8 epilogue warps
      const bool two = !wide && (bn <= 160 || pair);       // 4 epilogue warps, two CTAs per SM
      const int ksub = plan_ksub(bn, pair, wide);
      double ovh = ksub == 2 ? 290.0 : 520.0;
      if (pair) ovh *= 0.92;
      const double mma = 2.0 * bn;
      const long long slots = static_cast<long long>(sms) * (two ? 2 : 1);
      const double waves = static_cast<double>((ctas + slots - 1) / slots);
      double conc = 1.0;                                   // CTAs sharing one tensor pipe
      if (two) conc = ctas >= 2LL * sms ? 2.0 : static_cast<double>(ctas) / sms;
      const double iters = static_cast<double>((kbs + ksub - 1) / ksub) * ksub;
      const double main = iters * (ovh > conc * mma ? ovh : conc * mma);
      double epi = two ? 500.0 + 55.0 * bn : 1500.0 + 28.0 * bn;
      if (s > 1) epi += 30.0 * bn + 6.0 * bn * s;
      if (two) epi *= 0.5;
      f(TilePlan{bn, s, kbs, pair}, waves * (2500.0 + (pair ? 900.0 : 0.0) + main + epi));
    }
  }
}

TilePlan pick_plan(const PlanQuery& q) {
  TilePlan best{64, 1, q.num_kb, 0};
  double best_cost = -1.0;
  for_each_plan(q, [&](const TilePlan& t, double cost) {
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = t; }
  });
  return best;
}

// ---- plan cache + on-device autotuning ---------------------------------------------------------
// The first dbir_gemm call for a problem signature outside stream capture times every plan the model
// rates within 3x of its best (cold L2: the real layers stream their weights from HBM) on the
// caller's own operands, with the outputs redirected to scratch, and caches the winner. Split-K
// factors change the fp32 summation order, so results can differ at rounding level between
// processes; DBIR_GEMM_AUTOTUNE=0 keeps the deterministic model choice.
using PlanKey = std::array<long long, 20>;
std::mutex g_plan_mu;                     // guards the cache
std::mutex g_tune_mu;                     // serialises tuning (shared scratch / events, meaningful timings)
std::map<PlanKey, TilePlan> g_plan_cache;
int g_tuned_problems = 0;

bool autotune_enabled() {
  static const int on = [] { const char* e = getenv("DBIR_GEMM_AUTOTUNE"); return e ? atoi(e) : 1; }();
  return on != 0;
}

PlanKey make_key(const dbir_gemm_args* a) {
  PlanKey k{};
  int i = 0;
  k[i++] = a->M; k[i++] = a->N; k[i++] = a->K; k[i++] = a->a_mode;
  k[i++] = a->img_n; k[i++] = a->img_h; k[i++] = a->img_w; k[i++] = a->img_c; k[i++] = a->ksize;
  k[i++] = a->out_kind; k[i++] = a->geglu; k[i++] = a->force_bn; k[i++] = a->split_k; k[i++] = a->cta_pair;
  k[i++] = a->residual != nullptr; k[i++] = a->gn_partials != nullptr;
  k[i++] = a->splitk_ws ? a->splitk_ws_bytes : 0;
  k[i++] = a->lda; k[i++] = a->ldo; k[i++] = a->act * 8 + (a->groups > 1 ? a->groups : 1);
  return k;
}

int gemm_impl(const dbir_gemm_args* a, cudaStream_t st, const TilePlan* forced);

struct TuneScratch {
  void* out = nullptr; size_t out_cap = 0;
  void* part = nullptr; size_t part_cap = 0;
  void* flush = nullptr;
  static constexpr size_t FLUSH_BYTES = 192u << 20;     // > the 126 MB L2
  cudaEvent_t ev[3][2] = {};
  bool ready = false;
};
TuneScratch g_ts;

bool tune_reserve(void** ptr, size_t* cap, size_t need) {
  if (need <= *cap) return true;
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr; *cap = 0;
  if (cudaMalloc(ptr, need) != cudaSuccess) { cudaGetLastError(); return false; }
  *cap = need;
  return true;
}

// Returns true and the fastest measured plan, or false if tuning could not run (keeps the model's plan).
bool tune_plan(const dbir_gemm_args* a, cudaStream_t st, const PlanQuery& q, TilePlan* out_plan) {
  TuneScratch& ts = g_ts;
  if (!ts.ready) {
    if (cudaMalloc(&ts.flush, TuneScratch::FLUSH_BYTES) != cudaSuccess) { cudaGetLastError(); return false; }
    for (auto& e : ts.ev) for (auto& x : e) if (cudaEventCreate(&x) != cudaSuccess) return false;
    ts.ready = true;
  }
  const int eb = a->out_kind == 0 ? 4 : 2;
  const size_t out_bytes = static_cast<size_t>(a->M) * a->ldo * eb;
  if (!tune_reserve(&ts.out, &ts.out_cap, out_bytes)) return false;
  dbir_gemm_args t = *a;
  t.out = ts.out;
  t.debug_stamps = nullptr;
  t.prefetch_ptr = nullptr; t.prefetch_bytes = 0;
  if (a->gn_partials) {
    const long long imgs = a->a_mode == 1 ? a->img_n : a->M / a->gn_rows_per_img;
    const long long slots = a->a_mode == 1 ? dbir_gemm_gn_slots(a->img_h, a->img_w, 0)
                                           : dbir_gemm_gn_slots(0, 0, a->gn_rows_per_img);
    if (slots <= 0 || !tune_reserve(&ts.part, &ts.part_cap, static_cast<size_t>(imgs) * slots * a->N * 8)) return false;
    t.gn_partials = ts.part;
  }
  std::vector<std::pair<double, TilePlan>> cands;
  double best_model = -1.0;
  for_each_plan(q, [&](const TilePlan& p, double c) {
    cands.emplace_back(c, p);
    if (best_model < 0 || c < best_model) best_model = c;
  });
  float best_ms = -1.f;
  for (const auto& c : cands) {
    if (c.first > 3.0 * best_model) continue;
    if (gemm_impl(&t, st, &c.second) != 0) continue;              // warm-up (also validates the plan)
    bool ok = true;
    for (int r = 0; r < 3 && ok; ++r) {
      ok = cudaMemsetAsync(ts.flush, r, TuneScratch::FLUSH_BYTES, st) == cudaSuccess &&
           cudaEventRecord(ts.ev[r][0], st) == cudaSuccess && gemm_impl(&t, st, &c.second) == 0 &&
           cudaEventRecord(ts.ev[r][1], st) == cudaSuccess;
    }
    if (!ok || cudaEventSynchronize(ts.ev[2][1]) != cudaSuccess) { cudaGetLastError(); continue; }
    float ms = -1.f;
    for (int r = 0; r < 3; ++r) {
      float x = 0.f;
      if (cudaEventElapsedTime(&x, ts.ev[r][0], ts.ev[r][1]) == cudaSuccess && (ms < 0 || x < ms)) ms = x;
    }
    if (ms > 0 && (best_ms < 0 || ms < best_ms)) { best_ms = ms; *out_plan = c.second; }
  }
  return best_ms > 0;
}

TilePlan choose_plan(const dbir_gemm_args* a, cudaStream_t st, const PlanQuery& q) {
  if (!autotune_enabled()) return pick_plan(q);
  const PlanKey key = make_key(a);
  {
    std::lock_guard<std::mutex> lk(g_plan_mu);
    auto it = g_plan_cache.find(key);
    if (it != g_plan_cache.end()) return it->second;
  }
  TilePlan plan = pick_plan(q);
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return plan; }
  if (cs != cudaStreamCaptureStatusNone) return plan;       // cannot time inside a capture: model plan, not cached
  std::lock_guard<std::mutex> tune_lk(g_tune_mu);
  {
    std::lock_guard<std::mutex> lk(g_plan_mu);          // another thread may have tuned it meanwhile
    auto it = g_plan_cache.find(key);
    if (it != g_plan_cache.end()) return it->second;
  }
  TilePlan tuned = plan;
  if (tune_plan(a, st, q, &tuned)) plan = tuned;
  std::lock_guard<std::mutex> lk(g_plan_mu);
  g_plan_cache[key] = plan;
  ++g_tuned_problems;
  return plan;
}

}  // namespace

extern "C" int dbir_gemm(const dbir_gemm_args* a, void* stream) {
  return gemm_impl(a, reinterpret_cast<cudaStream_t>(stream), nullptr);
}
/* The analytic model's plan for a tile grid (host only, no launch): out[0..3] = BN, splits, k-blocks per
 * split, CTA pair flag. Lets the CPU test suite check the planner's invariants. */
extern "C" int dbir_gemm_model_plan(int32_t m_tiles, int32_t N, int32_t num_kb, int32_t geglu, int32_t force_bn,
                                    int32_t split_k, int32_t cta_pair, int64_t ws_floats, int32_t* out) {
  DBIR_REQUIRE(m_tiles > 0 && N > 0 && num_kb > 0 && out, "dbir_gemm_model_plan: bad arguments");
  const TilePlan t = pick_plan(PlanQuery{m_tiles, N, num_kb, geglu, force_bn, split_k, cta_pair, ws_floats});
  out[0] = t.bn; out[1] = t.splits; out[2] = t.kb_per_split; out[3] = t.pair;
  return 0;
}
/* Number of problem signatures tuned so far / drop every cached plan (tests, A/B runs). */
extern "C" int32_t dbir_gemm_tuned_problems(void) { return g_tuned_problems; }
extern "C" void dbir_gemm_clear_plans(void) {
  std::lock_guard<std::mutex> lk(g_plan_mu);
  g_plan_cache.clear();
}

namespace {
int gemm_impl(const dbir_gemm_args* a, cudaStream_t st, const TilePlan* forced) {
  DBIR_REQUIRE(a != nullptr, "dbir_gemm: null args");
  DBIR_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "dbir_gemm: bad shape M=%d N=%d K=%d", a->M,
               a->N, a->K);
  DBIR_REQUIRE(a->a && a->b && a->out, "dbir_gemm: null pointer");
  DBIR_REQUIRE(a->K % 8 == 0, "dbir_gemm: K=%d must be a multiple of 8 (16-byte rows)", a->K);
  DBIR_REQUIRE(!a->out2, "dbir_gemm: out2 is not supported by the TMA-store epilogue");
  const int eb = a->out_kind == 0 ? 4 : 2;
  DBIR_REQUIRE((a->ldo * eb) % 16 == 0, "dbir_gemm: output row stride must be a multiple of 16 bytes (ldo=%lld)",
               (long long)a->ldo);
  DBIR_REQUIRE(!a->residual || a->ldr % 4 == 0, "dbir_gemm: residual row stride must be a multiple of 4");

  GemmParams p{};
  p.M = a->M; p.N = a->N;
  p.mode = a->a_mode;
  p.out_kind = a->out_kind;
  p.bias = a->bias; p.rowvec = a->rowvec;
  p.rows_per_vec = a->rows_per_vec > 0 ? a->rows_per_vec : a->M;
  p.has_residual = a->residual != nullptr;
  p.alpha = a->alpha; p.act = a->act; p.act_param = a->act_param; p.geglu = a->geglu;
  p.bias_per_row = a->bias_per_row;
  const int groups = a->groups > 1 ? a->groups : 1;
  if (groups > 1) {
    DBIR_REQUIRE(!a->bias_per_row, "dbir_gemm: groups and bias_per_row are exclusive");
    DBIR_REQUIRE(a->N % 4 == 0, "dbir_gemm: grouped launches need N %% 4 == 0 (stacked bias alignment)");
    if (a->a_mode == 0)
      DBIR_REQUIRE(a->M % groups == 0 && (a->M / groups) % BM == 0,
                   "dbir_gemm: grouped launch needs M/groups (= %d) to be a multiple of %d rows", a->M / groups, BM);
    else
      DBIR_REQUIRE(a->img_n % groups == 0, "dbir_gemm: grouped conv needs img_n %% groups == 0");
  }
  p.dbg = reinterpret_cast<long long*>(a->debug_stamps);
  p.pf_ptr = reinterpret_cast<const char*>(a->prefetch_ptr);
  p.pf_bytes = a->prefetch_ptr ? a->prefetch_bytes : 0;
  if (a->geglu)
    DBIR_REQUIRE(a->force_bn >= 64 && a->force_bn != 160 && a->N % a->force_bn == 0 && !a->residual,
                 "dbir_gemm: GEGLU needs force_bn in {64,128,256} dividing N (weights are packed per tile)");
  const int n_out = a->geglu ? a->N / 2 : a->N;

  CUtensorMap ta, tb, to, tr;
  int m_tiles;
  if (a->a_mode == 0) {
    const long long lda = a->lda > 0 ? a->lda : a->K;
    DBIR_REQUIRE(lda % 8 == 0, "dbir_gemm: lda must be a multiple of 8");
    uint64_t dims[2] = {static_cast<uint64_t>(a->K), static_cast<uint64_t>(a->M)};
    uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t box[2] = {BK, BM};
    if (dbir_make_tmap(&ta, a->a, 2, dims, strides, box, 2, 1)) return -3;
    p.num_kb = (a->K + BK - 1) / BK;
    m_tiles = (a->M + BM - 1) / BM;
    if (groups > 1) p.group_span = (a->M / groups) / BM;
    // output / residual: 32 x 32 element boxes per epilogue warp
    uint64_t odims[2] = {static_cast<uint64_t>(n_out), static_cast<uint64_t>(a->M)};
    uint64_t ostr[1] = {static_cast<uint64_t>(a->ldo) * eb};
    uint32_t obox[2] = {32, 32};
    if (dbir_make_tmap(&to, a->out, 2, odims, ostr, obox, eb, eb == 4 ? 1 : 2)) return -3;
    tr = to;
    if (a->residual) {
      uint64_t rstr[1] = {static_cast<uint64_t>(a->ldr) * 4};
      if (dbir_make_tmap(&tr, a->residual, 2, odims, rstr, obox, 4, 1)) return -3;
    }
  } else {
    const int C = a->img_c, H = a->img_h, W = a->img_w, NI = a->img_n;
    const int taps = a->ksize * a->ksize;
    DBIR_REQUIRE(a->ksize == 3 || a->ksize == 1, "dbir_gemm: ksize must be 1 or 3");
    DBIR_REQUIRE(C % BK == 0, "dbir_gemm: conv channels C=%d must be a multiple of 64", C);
    DBIR_REQUIRE(a->K == taps * C, "dbir_gemm: K=%d != taps*C=%d", a->K, taps * C);
    DBIR_REQUIRE(static_cast<long long>(NI) * H * W == a->M, "dbir_gemm: M != N*H*W");
    int bw = 1;
    while (bw < W && bw < 128) bw <<= 1;
    int bh = 1;
    while (bh < H && bw * bh < 128) bh <<= 1;
    int bni = 128 / (bw * bh);
    p.H = H; p.W = W; p.NI = NI;
    p.bw = bw; p.bh = bh; p.bni = bni;
    p.tiles_x = (W + bw - 1) / bw;
    p.tiles_y = (H + bh - 1) / bh;
    p.cblocks = C / BK;
    p.kw = a->ksize; p.pad = a->ksize / 2;
    uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W),
                        static_cast<uint64_t>(H), static_cast<uint64_t>(NI)};
    // lda > 0: the pixels of `a` are lda elements apart and the conv reads their first C channels (dense
    // concat buffers: RRDBNet's cat((x, x1, ..)) lives in one [pixels, nf + 4 gc] tensor, bsrnet.py:51-56)
    const uint64_t ps = static_cast<uint64_t>(a->lda > 0 ? a->lda : C);
    DBIR_REQUIRE(ps >= static_cast<uint64_t>(C) && ps % 8 == 0, "dbir_gemm: conv pixel stride lda=%lld must be >= C and a multiple of 8",
                 (long long)a->lda);
    uint64_t strides[3] = {ps * 2, static_cast<uint64_t>(W) * ps * 2, static_cast<uint64_t>(H) * W * ps * 2};
    uint32_t box[4] = {BK, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh),
                       static_cast<uint32_t>(bni)};
    if (dbir_make_tmap(&ta, a->a, 4, dims, strides, box, 2, 1)) return -3;
    p.num_kb = taps * p.cblocks;
    m_tiles = p.tiles_x * p.tiles_y * ((NI + bni - 1) / bni);
    if (groups > 1) {
      DBIR_REQUIRE((NI / groups) % bni == 0, "dbir_gemm: grouped conv: a %d-image tile would straddle two groups", bni);
      p.group_span = NI / groups;
    }
    // per-warp sub-box of 32 pixels: (wbw x wbh x wbn)
    const int wbw = bw < 32 ? bw : 32;
    const int wbh = bh < 32 / wbw ? bh : 32 / wbw;
    const int wbn = 32 / (wbw * wbh);
    p.wbw = wbw; p.wbh = wbh;
    uint64_t odims[4] = {static_cast<uint64_t>(n_out), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                         static_cast<uint64_t>(NI)};
    uint64_t ostr[3] = {static_cast<uint64_t>(a->ldo) * eb, static_cast<uint64_t>(W) * a->ldo * eb,
                        static_cast<uint64_t>(H) * W * a->ldo * eb};
    uint32_t obox[4] = {32, static_cast<uint32_t>(wbw), static_cast<uint32_t>(wbh), static_cast<uint32_t>(wbn)};
    if (dbir_make_tmap(&to, a->out, 4, odims, ostr, obox, eb, eb == 4 ? 1 : 2)) return -3;
    tr = to;
    if (a->residual) {
      uint64_t rstr[3] = {static_cast<uint64_t>(a->ldr) * 4, static_cast<uint64_t>(W) * a->ldr * 4,
                          static_cast<uint64_t>(H) * W * a->ldr * 4};
      if (dbir_make_tmap(&tr, a->residual, 4, odims, rstr, obox, 4, 1)) return -3;
    }
  }

  if (a->gn_partials) {
    DBIR_REQUIRE(a->out_kind == 0 && !a->geglu, "dbir_gemm: gn_partials needs an fp32, non-GEGLU output");
    p.stats = reinterpret_cast<float*>(a->gn_partials);
    if (a->a_mode == 0) {
      DBIR_REQUIRE(a->gn_rows_per_img > 0 && a->gn_rows_per_img % 32 == 0 && a->M % a->gn_rows_per_img == 0,
                   "dbir_gemm: gn_partials in matrix mode needs rows-per-image that is a multiple of 32");
      p.stats_rows_per_img = a->gn_rows_per_img;
      p.stats_nslots = a->gn_rows_per_img / 32;
    } else {
      DBIR_REQUIRE(p.bw * p.bh >= 32, "dbir_gemm: gn_partials needs images of at least 32 pixels");
      p.stats_slots_x = (p.W + p.wbw - 1) / p.wbw;
      p.stats_nslots = p.stats_slots_x * ((p.H + p.wbh - 1) / p.wbh);
    }
  }
  constexpr long long TICKET_FLOATS = 16384;
  const long long ws_floats = a->splitk_ws ? a->splitk_ws_bytes / 4 - TICKET_FLOATS : 0;
  // CTA pairs share one weight tile between two adjacent m-tiles: in a grouped launch a pair must not
  // straddle two problems
  int pair_req = a->cta_pair;
  if (groups > 1 && ((m_tiles / groups) & 1)) pair_req = 2;
  const PlanQuery pq{m_tiles, a->N, p.num_kb, a->geglu, a->force_bn, a->split_k, pair_req, ws_floats};
  const TilePlan plan = forced ? *forced : choose_plan(a, st, pq);
  const int bn = plan.bn;
  p.splits = plan.splits;
  p.kb_per_split = plan.kb_per_split;
  if (plan.splits > 1) {
    p.tickets = reinterpret_cast<unsigned int*>(a->splitk_ws);
    p.ws = reinterpret_cast<float*>(a->splitk_ws) + TICKET_FLOATS;
  }
  {
    const long long ldb = a->ldb > 0 ? a->ldb : a->K;
    uint64_t dims[2] = {static_cast<uint64_t>(a->K), static_cast<uint64_t>(a->N) * groups};
    uint64_t strides[1] = {static_cast<uint64_t>(ldb) * 2};
    uint32_t box[2] = {BK, static_cast<uint32_t>(plan.pair ? bn / 2 : bn)};
    if (dbir_make_tmap(&tb, a->b, 2, dims, strides, box, 2, 1)) return -3;
  }
  dim3 grid(m_tiles, (a->N + bn - 1) / bn, plan.splits);
  // <= one CTA per SM anyway -> 8 epilogue warps and a deeper pipeline; else two CTAs per SM
  static const int wide_mode = [] { const char* e = getenv("DBIR_GEMM_WIDE"); return e ? atoi(e) : -1; }();
  const bool wide = wide_mode >= 0 ? wide_mode != 0
                                   : static_cast<long long>(grid.x) * grid.y * grid.z <= dbir_sm_count();
  // <BN, STAGES, EW, PAIR, KSUB>: EW == 4 kernels run two CTAs per SM (<= ~104 KB of stages each)
#define DBIR_GO(BN_, ST_, EW_, PAIR_, KS_) return launch<BN_, ST_, EW_, PAIR_, KS_>(ta, tb, to, tr, p, grid, st)
  if (plan.pair) {
    switch (bn) {
      case 64:  if (wide) DBIR_GO(64, 4, 8, true, 2);  else DBIR_GO(64, 2, 4, true, 2);
      case 128: if (wide) DBIR_GO(128, 4, 8, true, 2); else DBIR_GO(128, 2, 4, true, 2);
      case 160: if (wide) DBIR_GO(160, 4, 8, true, 2); else DBIR_GO(160, 2, 4, true, 2);
      case 256: if (wide) DBIR_GO(256, 3, 8, true, 2); else DBIR_GO(256, 3, 4, true, 1);
      default:
        dbir_set_error("dbir_gemm: unsupported paired tile width %d", bn);
        return -2;
    }
  }
  switch (bn) {
    case 32:  if (wide) DBIR_GO(32, 4, 8, false, 2);  else DBIR_GO(32, 2, 4, false, 2);
    case 64:  if (wide) DBIR_GO(64, 4, 8, false, 2);  else DBIR_GO(64, 2, 4, false, 2);
    case 128: if (wide) DBIR_GO(128, 3, 8, false, 2); else DBIR_GO(128, 3, 4, false, 1);
    case 160: if (wide) DBIR_GO(160, 3, 8, false, 2); else DBIR_GO(160, 3, 4, false, 1);
    case 256: DBIR_GO(256, 4, 8, false, 1);
    default:
      dbir_set_error("dbir_gemm: unsupported tile width %d", bn);
      return -2;
  }
}
}  // namespace

extern "C" int32_t dbir_gemm_gn_slots(int32_t conv_h, int32_t conv_w, int32_t rows_per_img) {
  if (conv_h > 0 && conv_w > 0) {
    int bw = 1;
    while (bw < conv_w && bw < 128) bw <<= 1;
    int bh = 1;
    while (bh < conv_h && bw * bh < 128) bh <<= 1;
    const int wbw = bw < 32 ? bw : 32;
    const int wbh = bh < 32 / wbw ? bh : 32 / wbw;
    if (wbw * wbh < 32) return -1;
    return ((conv_w + wbw - 1) / wbw) * ((conv_h + wbh - 1) / wbh);
  }
  return rows_per_img > 0 && rows_per_img % 32 == 0 ? rows_per_img / 32 : -1;
}
