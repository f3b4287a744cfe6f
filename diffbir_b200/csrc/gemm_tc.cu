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
};

__device__ __forceinline__ float apply_act(float v, int act, float prm) {
  if (act == 1) return gelu_erf_f(v);
  if (act == 2) return v > 0.f ? v : v * prm;
  if (act == 3) return silu_f(v);
  return v;
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

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x;
  const int n_tile = blockIdx.y;

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
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
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
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA + s * A_STAGE_BYTES);
        const uint32_t b_addr = smem_u32(sB + s * B_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          umma_f16(tmem_base, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32),
                   idesc, (kb | k) != 0 ? 1u : 0u);
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

    if (!p.geglu) {
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t acc[32];
        __syncwarp();
        tmem_ld32(taddr + c0, acc);
        tmem_ld_wait();
        const int gc0 = n_tile * BN + c0;
        if (out_row < 0 || gc0 >= p.N) continue;
        const int ncols = min(32, p.N - gc0);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
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
        if (p.residual) {
          const float* rs = p.residual + out_row * p.ldr + gc0;
          if (ncols == 32 && ((reinterpret_cast<uintptr_t>(rs) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 t = *reinterpret_cast<const float4*>(rs + j);
              v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < ncols) v[j] += rs[j];
          }
        }
        if (p.out_kind == 0) {
          float* o = reinterpret_cast<float*>(p.out) + out_row * p.ldo + gc0;
          if (ncols == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < ncols) o[j] = v[j];
          }
        } else {
          op_t* o = reinterpret_cast<op_t*>(p.out) + out_row * p.ldo + gc0;
          if (ncols == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 t;
              t.x = pack2(v[j], v[j + 1]); t.y = pack2(v[j + 2], v[j + 3]);
              t.z = pack2(v[j + 4], v[j + 5]); t.w = pack2(v[j + 6], v[j + 7]);
              *reinterpret_cast<uint4*>(o + j) = t;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < ncols) o[j] = f2op(v[j]);
          }
        }
        if (p.out2) {
          op_t* o = reinterpret_cast<op_t*>(p.out2) + out_row * p.ldo2 + gc0;
          if (ncols == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 t;
              t.x = pack2(v[j], v[j + 1]); t.y = pack2(v[j + 2], v[j + 3]);
              t.z = pack2(v[j + 4], v[j + 5]); t.w = pack2(v[j + 6], v[j + 7]);
              *reinterpret_cast<uint4*>(o + j) = t;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < ncols) o[j] = f2op(v[j]);
          }
        }
      }
    } else {
      // GEGLU: tile columns [0, BN/2) are values, [BN/2, BN) the matching gates
      // (weights are packed that way); output column = n_tile*BN/2 + c.
      constexpr int HB = BN / 2;
      const int n_half = p.N / 2;
#pragma unroll 1
      for (int c0 = 0; c0 < HB; c0 += 16) {
        uint32_t av[16], ag[16];
        __syncwarp();
        tmem_ld16(taddr + c0, av);
        tmem_ld16(taddr + HB + c0, ag);
        tmem_ld_wait();
        const int gc0 = n_tile * HB + c0;   // output column
        if (out_row < 0 || gc0 >= n_half) continue;
        const int ncols = min(16, n_half - gc0);
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float a = __uint_as_float(av[j]);
          float g = __uint_as_float(ag[j]);
          if (p.bias && j < ncols) {
            a += __ldg(p.bias + n_tile * BN + c0 + j);
            g += __ldg(p.bias + n_tile * BN + HB + c0 + j);
          }
          v[j] = a * gelu_erf_f(g);
        }
        if (p.out_kind == 0) {
          float* o = reinterpret_cast<float*>(p.out) + out_row * p.ldo + gc0;
#pragma unroll
          for (int j = 0; j < 16; ++j) if (j < ncols) o[j] = v[j];
        } else {
          op_t* o = reinterpret_cast<op_t*>(p.out) + out_row * p.ldo + gc0;
          if (ncols == 16 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
              uint4 t;
              t.x = pack2(v[j], v[j + 1]); t.y = pack2(v[j + 2], v[j + 3]);
              t.z = pack2(v[j + 4], v[j + 5]); t.w = pack2(v[j + 6], v[j + 7]);
              *reinterpret_cast<uint4*>(o + j) = t;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (j < ncols) o[j] = f2op(v[j]);
          }
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

int pick_bn(int m_tiles, int N, int geglu, int forced) {
  if (forced > 0) return forced;
  // The single-CTA mainloop is L2-feed bound, so per-CTA time ~ (128 + BN) bytes per k-block;
  // pick the tile width that minimises (CTAs per SM) x (128 + BN), widest on ties.
  const int cands[5] = {256, 160, 128, 64, 32};
  const int sms = dbir_sm_count();
  int best = 64;
  long long best_cost = -1;
  for (int i = 0; i < 5; ++i) {
    const int bn = cands[i];
    if (geglu && bn < 64) continue;
    if (bn > 64 && N % bn != 0) continue;          // ragged N only with the narrow tiles
    const long long ctas = static_cast<long long>(m_tiles) * ((N + bn - 1) / bn);
    const long long cost = ((ctas + sms - 1) / sms) * (128 + bn);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
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
    DBIR_REQUIRE(a->force_bn >= 64 && a->N % a->force_bn == 0,
                 "dbir_gemm: GEGLU needs force_bn (>= 64) dividing N (weights are packed per tile)");

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

  const int bn = pick_bn(m_tiles, a->N, a->geglu, a->force_bn);
  {
    const long long ldb = a->ldb > 0 ? a->ldb : a->K;
    uint64_t dims[2] = {static_cast<uint64_t>(a->K), static_cast<uint64_t>(a->N)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldb) * 2};
    uint32_t box[2] = {BK, static_cast<uint32_t>(bn)};
    if (dbir_make_tmap(&tb, a->b, 2, dims, strides, box, 2, 1)) return -3;
  }
  dim3 grid(m_tiles, (a->N + bn - 1) / bn, 1);
  switch (bn) {
    case 32:  return launch<32, 5>(ta, tb, p, grid, st);
    case 64:  return launch<64, 4>(ta, tb, p, grid, st);
    case 128: return launch<128, 3>(ta, tb, p, grid, st);
    case 160: return launch<160, 3>(ta, tb, p, grid, st);
    case 256: return launch<256, 4>(ta, tb, p, grid, st);
    default:
      dbir_set_error("dbir_gemm: unsupported tile width %d", bn);
      return -2;
  }
}
