// Flash attention for head_dim 64 on tcgen05 / TMEM (self- and cross-attention of the
// SpatialTransformer: reference attention.py:189-216 `F.scaled_dot_product_attention`,
// with the head split/merge copies of :196-214 folded into TMA coordinates).
//
//   O[b, i, h*64:(h+1)*64] = softmax_j(Q_h[i] . K_h[j] / 8) V_h[j]
//
// Q/K/V are 16-bit matrices [B, S, ld] (any row stride / column offset, e.g. slices of a fused
// QKV projection). One CTA = one (128-query tile, head, batch):
//   warp 0   TMA producer   (Q once, K/V tiles of 128 keys, 2 stages)
//   warp 1   MMA issuer     S = Q K^T -> TMEM[0:128), PV -> TMEM[128:192)
//   warps 2-5 softmax       thread == query row: the whole 128-wide score row is pulled into
//                            registers with one TMEM wait, online softmax in fp32 with exp2,
//                            P (16-bit) -> swizzled smem; O accumulates in TMEM across key tiles
//                            and is rescaled in place (tcgen05.ld/st) only when a row max moves.
// Two CTAs are resident per SM (112 KB smem, 256 TMEM columns each) so one CTA's softmax
// overlaps the other's MMAs.
#include "common.cuh"
#include "../../include/diffbir_b200.h"

namespace {

constexpr int TQ = 128;      // queries per CTA
constexpr int TK = 128;      // keys per tile
constexpr int DH = 64;
constexpr int KV_STAGES = 2;
constexpr int TILE_BYTES = 128 * DH * 2;   // 16 KB
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t TM_S = 0, TM_O = 128;

struct AttnParams {
  int sq, skv, heads;
  void* out;
  long long ldo;       // elements per output row
  float scale_log2;    // dh^-0.5 * log2(e)
};

struct __align__(8) AttnBarriers {
  uint64_t q_full;
  uint64_t kv_full[KV_STAGES];
  uint64_t kv_empty[KV_STAGES];
  uint64_t s_full;
  uint64_t p_full;
  uint64_t o_full;
  uint32_t tmem_slot;
};

constexpr int ATTN_SMEM = TILE_BYTES * (1 + 2 * KV_STAGES + 2) + 256;

__global__ void __launch_bounds__(192, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                const __grid_constant__ CUtensorMap tma_v, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE_BYTES;
  uint8_t* sV = sK + KV_STAGES * TILE_BYTES;
  uint8_t* sP = sV + KV_STAGES * TILE_BYTES;            // two [128 x 64] swizzled sub-tiles
  AttnBarriers* bar = reinterpret_cast<AttnBarriers*>(sP + 2 * TILE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * TQ;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int n_tiles = (p.skv + TK - 1) / TK;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("dbir attention: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q); tma_prefetch_desc(&tma_k); tma_prefetch_desc(&tma_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&bar->q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(&bar->kv_full[s], 1); mbar_init(&bar->kv_empty[s], 1); }
    mbar_init(&bar->s_full, 1);
    mbar_init(&bar->p_full, 128);
    mbar_init(&bar->o_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&bar->tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bar->tmem_slot;
  pdl_trigger();
  pdl_wait();

  // Warps 0 / 1 run warp-uniform loops and elect one lane per TMA / MMA / commit: inside a plain
  // `if (lane == 0)` region the compiler serialises every such instruction through a per-lane loop.
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(&bar->q_full, TILE_BYTES);
      tma_load_3d(sQ, &tma_q, &bar->q_full, head * DH, q0, b);
    }
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j % KV_STAGES;
      const uint32_t ph = (j / KV_STAGES) & 1;
      mbar_wait(&bar->kv_empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&bar->kv_full[s], 2 * TILE_BYTES);
        tma_load_3d(sK + s * TILE_BYTES, &tma_k, &bar->kv_full[s], head * DH, j * TK, b);
        tma_load_3d(sV + s * TILE_BYTES, &tma_v, &bar->kv_full[s], head * DH, j * TK, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = umma_idesc(TK, 0, 0);   // S[128 x 128]: Q, K both K-major
    constexpr uint32_t idesc_o = umma_idesc(DH, 0, 1);   // O[128 x 64]: P K-major, V MN-major
    const uint64_t q_desc = umma_desc_sw128(smem_u32(sQ));
    const uint64_t p_desc = umma_desc_sw128(smem_u32(sP));
    const uint64_t k_desc0 = umma_desc_sw128(smem_u32(sK));
    const uint64_t v_desc0 = umma_desc_sw128(smem_u32(sV));
    mbar_wait(&bar->q_full, 0);
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j % KV_STAGES;
      const uint32_t ph = (j / KV_STAGES) & 1;
      // descriptor start addresses advance in 16-byte units
      const uint64_t k_desc = k_desc0 + static_cast<uint64_t>((s * TILE_BYTES) >> 4);
      const uint64_t v_desc = v_desc0 + static_cast<uint64_t>((s * TILE_BYTES) >> 4);
      mbar_wait(&bar->kv_full[s], ph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_f16(tmem_base + TM_S, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&bar->s_full);
      }
      __syncwarp();
      // P(j) in smem (written by the softmax warps) x V(j)
      mbar_wait(&bar->p_full, j & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < TK / 16; ++k)
          umma_f16(tmem_base + TM_O, p_desc + static_cast<uint64_t>(((k >> 2) * TILE_BYTES + (k & 3) * 32) >> 4),
                   v_desc + static_cast<uint64_t>((k * 16 * 128) >> 4), idesc_o, (j > 0 || k != 0) ? 1u : 0u);
        umma_commit(&bar->o_full);
        umma_commit(&bar->kv_empty[s]);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;                 // query row in the tile == TMEM lane
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + TM_S;
    const uint32_t t_o = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + TM_O;
    const float sl2 = p.scale_log2;
    float m_run = -INFINITY, l_run = 0.f;
    uint8_t* p_row = sP + r * 128;
    const int sw = r & 7;

    for (int j = 0; j < n_tiles; ++j) {
      const int valid = min(TK, p.skv - j * TK);
      mbar_wait(&bar->s_full, j & 1);
      tc_fence_after();
      // whole score row (128 fp32) into registers with a single wait
      uint32_t s[TK];
      __syncwarp();
#pragma unroll
      for (int c0 = 0; c0 < TK; c0 += 32) tmem_ld32(t_s + c0, s + c0);
      tmem_ld_wait();
      // 8 independent max chains (a single chain would expose 127 dependent FMNMX latencies)
      float mx[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) mx[a] = -INFINITY;
      if (valid == TK) {
#pragma unroll
        for (int i = 0; i < TK; ++i) mx[i & 7] = fmaxf(mx[i & 7], __uint_as_float(s[i]));
      } else {
#pragma unroll
        for (int i = 0; i < TK; ++i)
          if (i < valid) mx[i & 7] = fmaxf(mx[i & 7], __uint_as_float(s[i]));
      }
      const float m_tile = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])),
                                 fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
      const float m_new = fmaxf(m_run, m_tile);
      float alpha;
      alpha = ex2_approx((m_run - m_new) * sl2);
      const float mb = m_new * sl2;
      // PV(j-1) must have retired before P is overwritten and before O is rescaled
      if (j > 0) {
        mbar_wait(&bar->o_full, (j - 1) & 1);
        tc_fence_after();
        // O (fp32, TMEM) *= alpha for rows whose running max moved; warp-uniform decision because
        // tcgen05.ld/st are warp-collective
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
          for (int c0 = 0; c0 < DH; c0 += 16) {
            uint32_t o[16];
            tmem_ld16(t_o + c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              float y0, y1;
              fmul2(y0, y1, __uint_as_float(o[i]), __uint_as_float(o[i + 1]), alpha, alpha);
              o[i] = __float_as_uint(y0); o[i + 1] = __float_as_uint(y1);
            }
            tmem_st16(t_o + c0, o);
          }
          tmem_st_wait();
        }
      }
      float ls[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) ls[a] = 0.f;
      const float nmb = -mb;
      if (valid == TK) {
        // full tile: packed FFMA2 / FADD2, no masking
#pragma unroll
        for (int c0 = 0; c0 < TK; c0 += 8) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            float x0, x1;
            ffma2(x0, x1, __uint_as_float(s[c0 + i]), __uint_as_float(s[c0 + i + 1]), sl2, sl2, nmb, nmb);
            // arguments are <= 0: one MUFU.EX2 (ex2.approx.ftz), no range fix-ups needed
            e[i] = ex2_approx(x0);
            e[i + 1] = ex2_approx(x1);
            fadd2(ls[i], ls[i + 1], ls[i], ls[i + 1], e[i], e[i + 1]);
          }
          uint4 t;
          t.x = pack2(e[0], e[1]); t.y = pack2(e[2], e[3]); t.z = pack2(e[4], e[5]); t.w = pack2(e[6], e[7]);
          // 16-byte chunk (c0 % 64) / 8 of sub-tile c0 / 64, 128-byte swizzle
          *reinterpret_cast<uint4*>(p_row + (c0 >> 6) * TILE_BYTES + (((((c0 & 63) >> 3)) ^ sw) << 4)) = t;
        }
      } else {
#pragma unroll
        for (int c0 = 0; c0 < TK; c0 += 8) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float x = ex2_approx(fmaf(__uint_as_float(s[c0 + i]), sl2, nmb));
            e[i] = c0 + i < valid ? x : 0.f;
            ls[i] += e[i];
          }
          uint4 t;
          t.x = pack2(e[0], e[1]); t.y = pack2(e[2], e[3]); t.z = pack2(e[4], e[5]); t.w = pack2(e[6], e[7]);
          *reinterpret_cast<uint4*>(p_row + (c0 >> 6) * TILE_BYTES + (((((c0 & 63) >> 3)) ^ sw) << 4)) = t;
        }
      }
      const float l_tile = ((ls[0] + ls[1]) + (ls[2] + ls[3])) + ((ls[4] + ls[5]) + (ls[6] + ls[7]));
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&bar->p_full);
    }
    // epilogue: O / l, 64 x 16-bit = 128 bytes per row
    mbar_wait(&bar->o_full, (n_tiles - 1) & 1);
    tc_fence_after();
    const int qi = q0 + r;
    const float inv = 1.0f / l_run;
    op_t* o_ptr = reinterpret_cast<op_t*>(p.out) + (static_cast<long long>(b) * p.sq + qi) * p.ldo + head * DH;
#pragma unroll
    for (int c0 = 0; c0 < DH; c0 += 32) {
      uint32_t o[32];
      __syncwarp();
      tmem_ld32(t_o + c0, o);
      tmem_ld_wait();
      if (qi < p.sq) {
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          uint4 t;
          t.x = pack2(__uint_as_float(o[c]) * inv, __uint_as_float(o[c + 1]) * inv);
          t.y = pack2(__uint_as_float(o[c + 2]) * inv, __uint_as_float(o[c + 3]) * inv);
          t.z = pack2(__uint_as_float(o[c + 4]) * inv, __uint_as_float(o[c + 5]) * inv);
          t.w = pack2(__uint_as_float(o[c + 6]) * inv, __uint_as_float(o[c + 7]) * inv);
          *reinterpret_cast<uint4*>(o_ptr + c0 + c) = t;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

int make_qkv_map(CUtensorMap* m, const void* base, int batch, int s, long long ld, int width) {
  uint64_t dims[3] = {static_cast<uint64_t>(width), static_cast<uint64_t>(s),
                      static_cast<uint64_t>(batch)};
  uint64_t strides[2] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(s) * ld * 2};
  uint32_t box[3] = {DH, 128, 1};
  return dbir_make_tmap(m, base, 3, dims, strides, box, 2, 1);
}

}  // namespace

extern "C" int dbir_attention(const void* q, const void* k, const void* v, void* out,
                              int32_t batch, int32_t heads, int32_t sq, int32_t skv,
                              int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, void* stream) {
  DBIR_REQUIRE(q && k && v && out, "dbir_attention: null pointer");
  DBIR_REQUIRE(batch > 0 && heads > 0 && sq > 0 && skv > 0, "dbir_attention: bad shape");
  DBIR_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0,
               "dbir_attention: row strides must be multiples of 8 elements");
  CUtensorMap mq, mk, mv;
  const int width = heads * DH;
  if (make_qkv_map(&mq, q, batch, sq, ldq, width)) return -3;
  if (make_qkv_map(&mk, k, batch, skv, ldk, width)) return -3;
  if (make_qkv_map(&mv, v, batch, skv, ldv, width)) return -3;
  AttnParams p;
  p.sq = sq; p.skv = skv; p.heads = heads;
  p.out = out; p.ldo = ldo;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  static bool configured = false;
  if (!configured) {
    DBIR_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         ATTN_SMEM));
    configured = true;
  }
  dim3 grid((sq + TQ - 1) / TQ, heads, batch);
  DBIR_CHECK_CUDA(dbir_launch(attn_fwd_kernel, grid, dim3(192), ATTN_SMEM, reinterpret_cast<cudaStream_t>(stream),
                              mq, mk, mv, p));
  return 0;
}
