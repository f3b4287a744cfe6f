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
// warp 2 = TMEM allocator + epilogue, warps 3..5 = epilogue (thread == accumulator
// row). Fused epilogue: bias, per-image row vector (time embedding), GELU /
// LeakyReLU, GEGLU gating, alpha scaling and fp32 residual add
// (out = res + alpha * f(acc + bias + rowvec)), fp32 or 16-bit output.
#include "common.cuh"
#include "../../include/diffbir_b200.h"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                 // 64 x 16-bit = 128 B = one swizzle atom row
constexpr int A_STAGE_BYTES = BM * BK * 2;

struct GemmParams {
  int M, N, num_kb;
  // conv geometry (mode 1)
  int mode;                // 0 plain, 1 conv taps over NHWC
  int H, W, NI;            // image size / count
  int bw, bh, bni;         // box extents (bw*bh*bni == 128)
  int tiles_x, tiles_y;    // tiles per image row / column
  int cblocks;             // C / 64
  int kw, pad;             // filter width (taps = kw*kw), padding
  // epilogue
  void* out;
  long long ldo;
  int out_kind;            // 0 fp32, 1 16-bit operand
  const float* bias;
  const float* rowvec;     // [NI or M/rows_per_vec, N]
  int rows_per_vec;
  const float* residual;
  long long ldr;
  float alpha;
  int act;                 // 0 none, 1 gelu(erf), 2 leaky relu (slope in act_param), 3 silu
  float act_param;
  int geglu;               // 1: tile holds [BN/2 values | BN/2 gates]; output width N/2
  int bias_per_row;        // bias indexed by output row instead of column
  void* out2;              // optional second copy of the result as op16 [rows, ldo2]
  long long ldo2;
  // split-K: grid.z CTAs share one output tile; partial tiles go through `ws`, the last CTA to
  // arrive (ticket) sums them in fixed order (deterministic) and runs the epilogue.
  int splits, kb_per_split;
  float* ws;
  unsigned int* tickets;
};

__device__ __forceinline__ float apply_act(float v, int act, float prm) {
  if (act == 1) return gelu_erf_f(v);
  if (act == 2) return v > 0.f ? v : v * prm;
  if (act == 3) return silu_f(v);
  return v;
}

// Epilogue for 32 consecutive accumulator columns x the warp's 32 rows.
// Math (bias / time-embedding row vector / activation / alpha) runs in the accumulator layout
// (thread == row). The chunk is then transposed through a padded per-warp shared-memory tile so
// that the residual read and every store touch whole contiguous row segments (one 128-byte
// line per warp instruction for fp32, 64 bytes x 2 rows for 16-bit) instead of 32 scattered rows.
__device__ __forceinline__ void store_chunk(const GemmParams& p, float* v, float* T, int lane,
                                            long long out_row, int gc0, int ncols);

__device__ __forceinline__ void finish_chunk(const GemmParams& p, float* v, float* T, int lane,
                                             long long out_row, int vec_idx, int gc0, int ncols) {
  if (out_row >= 0) {
    if (p.bias) {
      if (p.bias_per_row) {
        const float bv = __ldg(p.bias + out_row);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += bv;
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (j < ncols) v[j] += __ldg(p.bias + gc0 + j);
      }
    }
    if (p.rowvec) {
      const float* rv = p.rowvec + static_cast<long long>(vec_idx) * p.N + gc0;
#pragma unroll
      for (int j = 0; j < 32; ++j) if (j < ncols) v[j] += __ldg(rv + j);
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
  store_chunk(p, v, T, lane, out_row, gc0, ncols);
}

__device__ __forceinline__ void store_chunk(const GemmParams& p, float* v, float* T, int lane,
                                            long long out_row, int gc0, int ncols) {
#pragma unroll
  for (int j = 0; j < 32; ++j) T[lane * 33 + j] = v[j];
  __syncwarp();
  if (p.out_kind == 0) {
    float* out = reinterpret_cast<float*>(p.out);
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
      const long long orow = __shfl_sync(0xffffffffu, out_row, i);
      if (orow >= 0 && lane < ncols) {
        float val = T[i * 33 + lane];
        if (p.residual) val += p.residual[orow * p.ldr + gc0 + lane];
        out[orow * p.ldo + gc0 + lane] = val;
        if (p.out2) reinterpret_cast<op_t*>(p.out2)[orow * p.ldo2 + gc0 + lane] = f2op(val);
      }
    }
  } else {
    op_t* out = reinterpret_cast<op_t*>(p.out);
    const int sub = lane >> 4, col = (lane & 15) * 2;
#pragma unroll 4
    for (int i = 0; i < 32; i += 2) {
      const long long orow = __shfl_sync(0xffffffffu, out_row, i + sub);
      if (orow >= 0 && col < ncols) {
        float a = T[(i + sub) * 33 + col], b = T[(i + sub) * 33 + col + 1];
        const bool two = col + 1 < ncols;
        if (p.residual) {
          const float* rs = p.residual + orow * p.ldr + gc0 + col;
          a += rs[0];
          if (two) b += rs[1];
        }
        op_t* o = out + orow * p.ldo + gc0 + col;
        if (two) *reinterpret_cast<uint32_t*>(o) = pack2(a, b);
        else *o = f2op(a);
        if (p.out2) {
          op_t* o2 = reinterpret_cast<op_t*>(p.out2) + orow * p.ldo2 + gc0 + col;
          if (two) *reinterpret_cast<uint32_t*>(o2) = pack2(a, b);
          else *o2 = f2op(a);
        }
      }
    }
  }
  __syncwarp();
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const GemmParams p) {
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  uint32_t* split_flag = tmem_slot + 1;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x;
  const int n_tile = blockIdx.y;
  const int tile_lin = blockIdx.y * gridDim.x + blockIdx.x;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);

  // tile origin
  int n0 = 0, y0 = 0, x0 = 0;
  if (p.mode == 1) {
    int t = m_tile;
    int tx = t % p.tiles_x; t /= p.tiles_x;
    int ty = t % p.tiles_y; t /= p.tiles_y;
    x0 = tx * p.bw; y0 = ty * p.bh; n0 = t * p.bni;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = kb0; kb < kb1; ++kb) {
        const int s = (kb - kb0) % STAGES;
        const uint32_t ph = ((kb - kb0) / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], A_STAGE_BYTES + B_STAGE_BYTES);
        if (p.mode == 0) {
          tma_load_2d(sA + s * A_STAGE_BYTES, &tma_a, &full[s], kb * BK, m_tile * BM);
        } else {
          const int tap = kb / p.cblocks;
          const int cb = kb - tap * p.cblocks;
          const int dy = tap / p.kw - p.pad;
          const int dx = tap % p.kw - p.pad;
          tma_load_4d(sA + s * A_STAGE_BYTES, &tma_a, &full[s], cb * BK, x0 + dx, y0 + dy, n0);
        }
        tma_load_2d(sB + s * B_STAGE_BYTES, &tma_b, &full[s], kb * BK, n_tile * BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(BN, 0, 0);
      for (int kb = kb0; kb < kb1; ++kb) {
        const int s = (kb - kb0) % STAGES;
        const uint32_t ph = ((kb - kb0) / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA + s * A_STAGE_BYTES);
        const uint32_t b_addr = smem_u32(sB + s * B_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          umma_f16(tmem_base, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32),
                   idesc, (kb > kb0 || k != 0) ? 1u : 0u);
        }
        umma_commit(&empty[s]);   // frees the smem stage once these MMAs retire
      }
      umma_commit(tmem_full);     // accumulator complete
    }
  } else {
    // ---------------- epilogue: warps 2..5, TMEM lane quarter = warp % 4 -------------
    const int q = warp & 3;
    const int r = q * 32 + lane;            // accumulator row within the tile
    long long out_row = -1;                 // global output row, -1 = masked
    int vec_idx = 0;
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
    }
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // all MMAs have retired -> the pipeline stages are free: per-warp 32x33 fp32 transpose tiles
    float* T = reinterpret_cast<float*>(sA) + (warp - 2) * (32 * 33 + 31);

    if (!p.geglu) {
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t acc[32];
        __syncwarp();
        tmem_ld32(taddr + c0, acc);
        tmem_ld_wait();
        const int gc0 = n_tile * BN + c0;
        if (p.splits > 1) {
          // raw partial -> workspace [tile][split][col][row]: coalesced across the warp's rows
          float* wcol = p.ws + ((static_cast<long long>(tile_lin) * p.splits + blockIdx.z) * BN + c0) * BM + r;
#pragma unroll
          for (int j = 0; j < 32; ++j) __stcg(wcol + j * BM, __uint_as_float(acc[j]));
          continue;
        }
        if (gc0 >= p.N) continue;
        const int ncols = min(32, p.N - gc0);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
        finish_chunk(p, v, T, lane, out_row, vec_idx, gc0, ncols);
      }
      if (p.splits > 1) {
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) {
          const unsigned int t = atomicAdd(&p.tickets[tile_lin], 1u);
          const bool last = (t == static_cast<unsigned int>(p.splits) - 1);
          if (last) p.tickets[tile_lin] = 0;        // self-reset for the next launch
          *split_flag = last ? 1u : 0u;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (*split_flag) {
          __threadfence();
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            const int gc0 = n_tile * BN + c0;
            if (gc0 >= p.N) continue;
            const int ncols = min(32, p.N - gc0);
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
            for (int z = 0; z < p.splits; ++z) {     // fixed summation order
              const float* wcol = p.ws + ((static_cast<long long>(tile_lin) * p.splits + z) * BN + c0) * BM + r;
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += __ldcg(wcol + j * BM);
            }
            finish_chunk(p, v, T, lane, out_row, vec_idx, gc0, ncols);
          }
        }
      }
    } else {
      // GEGLU: tile columns [0, BN/2) are values, [BN/2, BN) the matching gates
      // (weights are packed that way); output column = n_tile*BN/2 + c.
      constexpr int HB = BN / 2;
      const int n_half = p.N / 2;
      if constexpr (HB % 32 == 0) {
#pragma unroll 1
        for (int c0 = 0; c0 < HB; c0 += 32) {
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
          for (int j = 0; j < 32; ++j) {
            float a = __uint_as_float(av[j]);
            float g = __uint_as_float(ag[j]);
            if (p.bias && j < ncols) {
              a += __ldg(p.bias + n_tile * BN + c0 + j);
              g += __ldg(p.bias + n_tile * BN + HB + c0 + j);
            }
            v[j] = a * gelu_erf_f(g);
          }
          store_chunk(p, v, T, lane, out_row, gc0, ncols);
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

template <int BN, int STAGES>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, dim3 grid,
           cudaStream_t st) {
  constexpr int smem = STAGES * (A_STAGE_BYTES + BN * BK * 2) + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    DBIR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  gemm_tc_kernel<BN, STAGES><<<grid, 192, smem, st>>>(ta, tb, p);
  DBIR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Tile width + split-K factor. Cost model (cycles, per SM): the single-CTA mainloop is L2-feed
// bound at ~3 cycles per byte-row, so a k-block costs ~3.05*(128+BN); epilogue ~20*BN; a split
// adds a workspace round trip. CTAs are spread over `sms` SMs, a partial last wave costs a full one.
struct TilePlan { int bn, splits, kb_per_split; };

TilePlan pick_plan(int m_tiles, int N, int num_kb, int geglu, int forced_bn, int split_req,
                   long long ws_floats_avail) {
  const int cands[5] = {256, 160, 128, 64, 32};
  const int sms = dbir_sm_count();
  TilePlan best{64, 1, num_kb};
  double best_cost = -1.0;
  for (int i = 0; i < 5; ++i) {
    const int bn = cands[i];
    if (forced_bn > 0 && bn != forced_bn) continue;
    if (forced_bn <= 0) {
      if (geglu && bn < 64) continue;
      if (bn > 64 && N % bn != 0) continue;          // ragged N only with the narrow tiles
    }
    const long long tiles = static_cast<long long>(m_tiles) * ((N + bn - 1) / bn);
    for (int s = 1; s <= 16; ++s) {
      if (s > 1) {
        if (split_req == 1 || geglu || ws_floats_avail <= 0) break;
        if (num_kb / s < 6) break;
        if (tiles > 16384 || tiles * s * 128LL * bn > ws_floats_avail) break;
      }
      if (split_req > 1 && s != split_req) continue;
      const int kbs = (num_kb + s - 1) / s;
      if (static_cast<long long>(kbs) * (s - 1) >= num_kb) continue;   // an empty split
      const double waves = static_cast<double>((tiles * s + sms - 1) / sms);
      double cta = 3000.0 + kbs * 3.05 * (128 + bn) + 20.0 * bn;
      if (s > 1) cta += 15.0 * bn + 6.0 * bn * s;
      const double cost = waves * cta;
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = TilePlan{bn, s, kbs}; }
    }
  }
  return best;
}

}  // namespace

extern "C" int dbir_gemm(const dbir_gemm_args* a, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DBIR_REQUIRE(a != nullptr, "dbir_gemm: null args");
  DBIR_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "dbir_gemm: bad shape M=%d N=%d K=%d", a->M,
               a->N, a->K);
  DBIR_REQUIRE(a->a && a->b && a->out, "dbir_gemm: null pointer");
  DBIR_REQUIRE(a->K % 8 == 0, "dbir_gemm: K=%d must be a multiple of 8 (16-byte rows)", a->K);
  DBIR_REQUIRE(a->out_kind == 0 || (a->ldo % 2 == 0), "dbir_gemm: 16-bit output needs an even ldo");
  DBIR_REQUIRE(!a->out2 || (a->ldo2 % 2 == 0), "dbir_gemm: out2 needs an even ldo2");

  GemmParams p{};
  p.M = a->M; p.N = a->N;
  p.mode = a->a_mode;
  p.out = a->out; p.ldo = a->ldo; p.out_kind = a->out_kind;
  p.bias = a->bias; p.rowvec = a->rowvec;
  p.rows_per_vec = a->rows_per_vec > 0 ? a->rows_per_vec : a->M;
  p.residual = a->residual; p.ldr = a->ldr;
  p.alpha = a->alpha; p.act = a->act; p.act_param = a->act_param; p.geglu = a->geglu;
  p.bias_per_row = a->bias_per_row; p.out2 = a->out2; p.ldo2 = a->ldo2;
  if (a->geglu)
    DBIR_REQUIRE(a->force_bn >= 64 && a->force_bn != 160 && a->N % a->force_bn == 0,
                 "dbir_gemm: GEGLU needs force_bn in {64,128,256} dividing N (weights are packed per tile)");

  CUtensorMap ta, tb;
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
    uint64_t strides[3] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(W) * C * 2,
                           static_cast<uint64_t>(H) * W * C * 2};
    uint32_t box[4] = {BK, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh),
                       static_cast<uint32_t>(bni)};
    if (dbir_make_tmap(&ta, a->a, 4, dims, strides, box, 2, 1)) return -3;
    p.num_kb = taps * p.cblocks;
    m_tiles = p.tiles_x * p.tiles_y * ((NI + bni - 1) / bni);
  }

  constexpr long long TICKET_FLOATS = 16384;
  const long long ws_floats = a->splitk_ws ? a->splitk_ws_bytes / 4 - TICKET_FLOATS : 0;
  const TilePlan plan = pick_plan(m_tiles, a->N, p.num_kb, a->geglu, a->force_bn, a->split_k, ws_floats);
  const int bn = plan.bn;
  p.splits = plan.splits;
  p.kb_per_split = plan.kb_per_split;
  if (plan.splits > 1) {
    p.tickets = reinterpret_cast<unsigned int*>(a->splitk_ws);
    p.ws = reinterpret_cast<float*>(a->splitk_ws) + TICKET_FLOATS;
  }
  {
    const long long ldb = a->ldb > 0 ? a->ldb : a->K;
    uint64_t dims[2] = {static_cast<uint64_t>(a->K), static_cast<uint64_t>(a->N)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldb) * 2};
    uint32_t box[2] = {BK, static_cast<uint32_t>(bn)};
    if (dbir_make_tmap(&tb, a->b, 2, dims, strides, box, 2, 1)) return -3;
  }
  dim3 grid(m_tiles, (a->N + bn - 1) / bn, plan.splits);
  // <= one CTA per SM anyway: use the whole 227 KB for a deep TMA pipeline (the mainloop is
  // latency/feed bound); otherwise 3-5 stages so two CTAs share an SM (epilogue overlap).
  const bool deep = static_cast<long long>(grid.x) * grid.y * grid.z <= dbir_sm_count();
  switch (bn) {
    case 32:  return deep ? launch<32, 11>(ta, tb, p, grid, st) : launch<32, 5>(ta, tb, p, grid, st);
    case 64:  return deep ? launch<64, 9>(ta, tb, p, grid, st) : launch<64, 4>(ta, tb, p, grid, st);
    case 128: return deep ? launch<128, 7>(ta, tb, p, grid, st) : launch<128, 3>(ta, tb, p, grid, st);
    case 160: return deep ? launch<160, 6>(ta, tb, p, grid, st) : launch<160, 3>(ta, tb, p, grid, st);
    case 256: return launch<256, 4>(ta, tb, p, grid, st);
    default:
      dbir_set_error("dbir_gemm: unsupported tile width %d", bn);
      return -2;
  }
}
