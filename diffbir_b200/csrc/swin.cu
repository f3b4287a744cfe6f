// SwinIR window multi-head self-attention (W-MSA / SW-MSA), one CTA per 8x8 window.
//
// Replaces WindowAttention.forward (reference swinir.py:120-151) together with the data
// movement around it in SwinTransformerBlock.forward (swinir.py:255-276): torch.roll,
// window_partition, window_reverse and the reverse roll are folded into the gather/scatter
// addressing; the shift mask (swinir.py:222-243, -100 across region borders) is evaluated
// analytically; the relative-position bias is gathered from the [ (2w-1)^2, heads ] table.
//
// The window's 64 qkv rows are staged in shared memory (16-bit), one warp per head, each lane
// owns two query rows: scores, softmax and PV in fp32 registers (K/V reads are warp-wide
// broadcasts). The result overwrites the row's own q slot in shared memory and is then
// written back coalesced.
#include "common.cuh"
#include "../../include/diffbir_b200.h"

namespace {

constexpr int WS = 8;
constexpr int NTOK = WS * WS;   // 64

template <int HEADS, int DH>
__global__ void __launch_bounds__(HEADS * 32)
swin_window_attn_kernel(const op_t* __restrict__ qkv, long long ldq, int H, int W, int shift,
                        const float* __restrict__ bias_table, op_t* __restrict__ out,
                        long long ldo) {
  constexpr int C = HEADS * DH;
  constexpr int ROW = 3 * C + 4;                 // padded row (16-bit elements)
  extern __shared__ __align__(16) uint8_t smem_raw[];
  op_t* s_qkv = reinterpret_cast<op_t*>(smem_raw);             // [64][ROW]
  float* s_bias = reinterpret_cast<float*>(s_qkv + NTOK * ROW);  // [(2w-1)^2]
  int* s_tok = reinterpret_cast<int*>(s_bias + (2 * WS - 1) * (2 * WS - 1) * HEADS);  // [64] global row
  int* s_reg = s_tok + NTOK;                                    // [64] mask region id

  const int wins_x = W / WS, wins_y = H / WS;
  int wid = blockIdx.x;
  const int wx = wid % wins_x; wid /= wins_x;
  const int wy = wid % wins_y; wid /= wins_y;
  const int b = wid;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  if (tid < NTOK) {
    const int ty = tid / WS, tx = tid % WS;
    const int sy = wy * WS + ty, sx = wx * WS + tx;            // coordinates in the shifted frame
    const int y = (sy + shift) % H, x = (sx + shift) % W;      // original position (roll by -shift)
    s_tok[tid] = (b * H + y) * W + x;
    int rh = 0, rw = 0;
    if (shift > 0) {
      rh = sy < H - WS ? 0 : (sy < H - shift ? 1 : 2);
      rw = sx < W - WS ? 0 : (sx < W - shift ? 1 : 2);
    }
    s_reg[tid] = rh * 3 + rw;
  }
  for (int i = tid; i < (2 * WS - 1) * (2 * WS - 1) * HEADS; i += blockDim.x) s_bias[i] = bias_table[i];
  __syncthreads();
  // stage 64 rows x 3C 16-bit values (4-byte granules; 3C and ldq are even)
  constexpr int WORDS = 3 * C / 2;
  for (int i = tid; i < NTOK * WORDS; i += blockDim.x) {
    const int t = i / WORDS, wd = i % WORDS;
    const uint32_t v = *reinterpret_cast<const uint32_t*>(qkv + static_cast<long long>(s_tok[t]) * ldq + wd * 2);
    *reinterpret_cast<uint32_t*>(s_qkv + t * ROW + wd * 2) = v;
  }
  __syncthreads();

  const int head = warp;
  const float scale = rsqrtf(static_cast<float>(DH));
  for (int rr = 0; rr < 2; ++rr) {
    const int i = lane + rr * 32;                               // query token
    const int iy = i / WS, ix = i % WS;
    float q[DH];
    const op_t* qrow = s_qkv + i * ROW + head * DH;
#pragma unroll
    for (int d = 0; d < DH; d += 2) {
      const float2 t = unpack2(*reinterpret_cast<const uint32_t*>(qrow + d));
      q[d] = t.x * scale; q[d + 1] = t.y * scale;
    }
    float s[NTOK];
    float m = -INFINITY;
    const int reg_i = s_reg[i];
#pragma unroll
    for (int j = 0; j < NTOK; ++j) {
      const op_t* krow = s_qkv + j * ROW + C + head * DH;
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < DH; d += 2) {
        const float2 t = unpack2(*reinterpret_cast<const uint32_t*>(krow + d));
        acc += q[d] * t.x + q[d + 1] * t.y;
      }
      const int jy = j / WS, jx = j % WS;
      acc += s_bias[((iy - jy + WS - 1) * (2 * WS - 1) + (ix - jx + WS - 1)) * HEADS + head];
      if (s_reg[j] != reg_i) acc += -100.0f;
      s[j] = acc;
      m = fmaxf(m, acc);
    }
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < NTOK; ++j) { s[j] = __expf(s[j] - m); l += s[j]; }
    const float inv = 1.0f / l;
    float o[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < NTOK; ++j) {
      const op_t* vrow = s_qkv + j * ROW + 2 * C + head * DH;
      const float pj = s[j];
#pragma unroll
      for (int d = 0; d < DH; d += 2) {
        const float2 t = unpack2(*reinterpret_cast<const uint32_t*>(vrow + d));
        o[d] += pj * t.x; o[d + 1] += pj * t.y;
      }
    }
    // result -> this row's own q slot (no other thread reads it)
    op_t* orow = s_qkv + i * ROW + head * DH;
#pragma unroll
    for (int d = 0; d < DH; d += 2)
      *reinterpret_cast<uint32_t*>(orow + d) = pack2(o[d] * inv, o[d + 1] * inv);
  }
  __syncthreads();
  constexpr int OWORDS = C / 2;
  for (int i = tid; i < NTOK * OWORDS; i += blockDim.x) {
    const int t = i / OWORDS, wd = i % OWORDS;
    *reinterpret_cast<uint32_t*>(out + static_cast<long long>(s_tok[t]) * ldo + wd * 2) =
        *reinterpret_cast<const uint32_t*>(s_qkv + t * ROW + wd * 2);
  }
}

}  // namespace

extern "C" int dbir_swin_window_attention(const void* qkv, int64_t ldq, int32_t batch, int32_t h,
                                          int32_t w, int32_t heads, int32_t head_dim,
                                          int32_t window, int32_t shift, const float* bias_table,
                                          void* out, int64_t ldo, void* stream) {
  DBIR_REQUIRE(qkv && bias_table && out, "dbir_swin_window_attention: null pointer");
  DBIR_REQUIRE(window == 8 && heads == 6 && head_dim == 30,
               "dbir_swin_window_attention: built for window 8, 6 heads x 30 (configs/inference/swinir.yaml)");
  DBIR_REQUIRE(h % 8 == 0 && w % 8 == 0 && ldq % 2 == 0 && ldo % 2 == 0 && shift >= 0 && shift < 8,
               "dbir_swin_window_attention: bad geometry");
  constexpr int C = 180;
  constexpr int ROW = 3 * C + 4;
  const size_t smem = NTOK * ROW * 2 + 225 * 6 * 4 + 2 * NTOK * 4;
  static bool configured = false;
  if (!configured) {
    DBIR_CHECK_CUDA(cudaFuncSetAttribute(swin_window_attn_kernel<6, 30>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int nwin = batch * (h / 8) * (w / 8);
  swin_window_attn_kernel<6, 30><<<nwin, 192, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const op_t*>(qkv), ldq, h, w, shift, bias_table,
      reinterpret_cast<op_t*>(out), ldo);
  DBIR_CHECK_CUDA(cudaGetLastError());
  return 0;
}
