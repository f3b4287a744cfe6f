// HBM-bound helper kernels of the DiffBIR hot path: small-channel convolutions (stems and
// heads), stride-2 im2col, tiny fp32 linears (time embedding), row softmax, layout changes,
// the fused sampler update and the tile blend. fp32 math throughout; 128-bit accesses where
// the layout allows.
#include "common.cuh"
#include "../../include/diffbir_b200.h"

namespace {

// ---------------------------------------------------------------------------------------
// conv3x3, tiny Cin (<= 16), stride 1 pad 1. Input = virtual concat of two NCHW fp32 tensors
// (ControlNet stem: cat(x, hint), controlnet.py:316). Weights fp32 [9*Cin][Cout]
// (k = tap*Cin + c). Output NHWC fp32.
// One CTA = a strip of 32 consecutive pixels of one image row: the (Cin x 3 x 34) input patch is
// staged in shared memory once (scale/shift applied, zero outside the image), then every thread
// owns 4 output channels of 4 pixels of the strip (p, p+8, p+16, p+24), so each 16-byte weight
// load (L1-resident, coalesced across the warp) feeds 16 FMAs. The stems run first in every
// forward of the step loop: the one-thread-per-(pixel, channel quad) version spent 40-75 us there
// on address arithmetic and one weight load per 4 FMAs.
// ---------------------------------------------------------------------------------------
constexpr int SC_STRIP = 32;
__global__ void __launch_bounds__(256)
conv3x3_small_cin_kernel(const float* __restrict__ in1, const float* __restrict__ in2, int c1, int c2,
                         int n, int h, int w, const float* __restrict__ wt,
                         const float* __restrict__ bias, int cout, float* __restrict__ out,
                         float in_scale, float in_shift) {
  __shared__ float s_in[16][3][SC_STRIP + 4];
  pdl_trigger();
  const int cin = c1 + c2;
  const int cv = cout / 4;
  const int strips_x = (w + SC_STRIP - 1) / SC_STRIP;
  const int x0 = (blockIdx.x % strips_x) * SC_STRIP;
  const int y = (blockIdx.x / strips_x) % h;
  const int b = blockIdx.x / (strips_x * h);
  pdl_wait();
  for (int i = threadIdx.x; i < cin * 3 * (SC_STRIP + 2); i += blockDim.x) {
    const int xx = i % (SC_STRIP + 2);
    const int r = (i / (SC_STRIP + 2)) % 3;
    const int c = i / (3 * (SC_STRIP + 2));
    const int gy = y + r - 1, gx = x0 + xx - 1;
    float v = 0.f;
    if (gy >= 0 && gy < h && gx >= 0 && gx < w) {
      const float* src = c < c1 ? in1 + ((static_cast<long long>(b) * c1 + c) * h + gy) * w + gx
                                : in2 + ((static_cast<long long>(b) * c2 + (c - c1)) * h + gy) * w + gx;
      v = __ldg(src) * in_scale + in_shift;
    }
    s_in[c][r][xx] = v;
  }
  __syncthreads();
  const long long row_base = (static_cast<long long>(b) * h + y) * w + x0;
  for (int i = threadIdx.x; i < (SC_STRIP / 4) * cv; i += blockDim.x) {
    const int co = (i % cv) * 4;
    const int pg = i / cv;                       // pixels pg, pg + 8, pg + 16, pg + 24 of the strip
    const float4 bv = *reinterpret_cast<const float4*>(bias + co);
    float4 acc[4] = {bv, bv, bv, bv};
    for (int tap = 0; tap < 9; ++tap) {
      const int ty = tap / 3, tx = tap - 3 * ty;
#pragma unroll 4
      for (int c = 0; c < cin; ++c) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wt + static_cast<long long>(tap * cin + c) * cout + co));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v = s_in[c][ty][pg + 8 * j + tx];
          acc[j].x += v * wv.x; acc[j].y += v * wv.y; acc[j].z += v * wv.z; acc[j].w += v * wv.w;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (x0 + pg + 8 * j < w)
        *reinterpret_cast<float4*>(out + (row_base + pg + 8 * j) * cout + co) = acc[j];
  }
}

// ---------------------------------------------------------------------------------------
// conv3x3, tiny Cout (<= 8), stride 1 pad 1, op16 NHWC input (already normalised / activated),
// fp32 weights [Cout][9*Cin]. One warp per 4 consecutive pixels of a row; lanes split the
// flattened K = (tap, 8-channel vector) space, each weight vector (L1) is applied to all 4 pixels.
// out = (acc + bias) * post_scale + post_shift[c]; layout NCHW (out_nchw) or NHWC.
// ---------------------------------------------------------------------------------------
constexpr int SO_PX = 4;
template <int COUT>
__global__ void __launch_bounds__(256)
conv3x3_small_cout_kernel(const op_t* __restrict__ in, int n, int h, int w, int cin,
                          const float* __restrict__ wt, const float* __restrict__ bias,
                          float post_scale, const float* __restrict__ post_shift,
                          float* __restrict__ out, int out_nchw) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int groups_x = (w + SO_PX - 1) / SO_PX;
  const long long ngroups = static_cast<long long>(n) * h * groups_x;
  const int vpt = cin / 8;                         // 16-byte vectors per tap
  const int kvec = 9 * vpt;
  for (long long g = warp_global; g < ngroups; g += nwarps) {
    const int x0 = static_cast<int>(g % groups_x) * SO_PX;
    const int y = static_cast<int>((g / groups_x) % h);
    const int b = static_cast<int>(g / (static_cast<long long>(groups_x) * h));
    float acc[SO_PX][COUT];
#pragma unroll
    for (int j = 0; j < SO_PX; ++j)
#pragma unroll
      for (int o = 0; o < COUT; ++o) acc[j][o] = 0.f;
    for (int idx = lane; idx < kvec; idx += 32) {
      const int tap = idx / vpt, v = idx - tap * vpt;
      const int ty = tap / 3, tx = tap - 3 * ty;
      const int yy = y + ty - 1;
      if (yy < 0 || yy >= h) continue;
      float4 w0[COUT], w1[COUT];
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        const float* wr = wt + static_cast<long long>(o) * 9 * cin + tap * cin + v * 8;
        w0[o] = __ldg(reinterpret_cast<const float4*>(wr));
        w1[o] = __ldg(reinterpret_cast<const float4*>(wr + 4));
      }
      const op_t* row = in + (static_cast<long long>(b) * h + yy) * w * cin + v * 8;
#pragma unroll
      for (int j = 0; j < SO_PX; ++j) {
        const int xx = x0 + j + tx - 1;
        if (xx < 0 || xx >= w) continue;
        const uint4 raw = *reinterpret_cast<const uint4*>(row + static_cast<long long>(xx) * cin);
        float f[8];
        float2 t;
        t = unpack2(raw.x); f[0] = t.x; f[1] = t.y;
        t = unpack2(raw.y); f[2] = t.x; f[3] = t.y;
        t = unpack2(raw.z); f[4] = t.x; f[5] = t.y;
        t = unpack2(raw.w); f[6] = t.x; f[7] = t.y;
#pragma unroll
        for (int o = 0; o < COUT; ++o)
          acc[j][o] += f[0] * w0[o].x + f[1] * w0[o].y + f[2] * w0[o].z + f[3] * w0[o].w + f[4] * w1[o].x +
                       f[5] * w1[o].y + f[6] * w1[o].z + f[7] * w1[o].w;
      }
    }
#pragma unroll
    for (int j = 0; j < SO_PX; ++j)
#pragma unroll
      for (int o = 0; o < COUT; ++o) acc[j][o] = warp_sum(acc[j][o]);
    if (lane < SO_PX && x0 + lane < w) {
      const int x = x0 + lane;
      const long long pix = (static_cast<long long>(b) * h + y) * w + x;
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < SO_PX; ++j) if (j == lane) s = acc[j][o];
        const float v = (s + bias[o]) * post_scale + (post_shift ? post_shift[o] : 0.f);
        if (out_nchw) out[((static_cast<long long>(b) * COUT + o) * h + y) * w + x] = v;
        else out[pix * COUT + o] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// im2col for 3x3 stride-2 convs: fp32 NHWC [n,h,w,c] -> op16 [n*ho*wo, 9*c] (k = tap*c + ch).
// pad_lo = 1: symmetric pad 1 (Downsample, unet.py:99); pad_lo = 0: pad (0,1,0,1) then
// valid conv (VAE Downsample, vae.py:51-55).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
im2col_s2_kernel(const float* __restrict__ in, int n, int h, int w, int c, int ho, int wo,
                 int pad_lo, op_t* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int cv = c / 4;
  const long long total = static_cast<long long>(n) * ho * wo * 9 * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % cv) * 4;
    long long r = i / cv;
    const int tap = static_cast<int>(r % 9);
    r /= 9;
    const int ox = static_cast<int>(r % wo);
    const int oy = static_cast<int>((r / wo) % ho);
    const int b = static_cast<int>(r / (static_cast<long long>(wo) * ho));
    const int yy = oy * 2 + tap / 3 - pad_lo, xx = ox * 2 + tap % 3 - pad_lo;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (yy >= 0 && yy < h && xx >= 0 && xx < w)
      v = *reinterpret_cast<const float4*>(in + ((static_cast<long long>(b) * h + yy) * w + xx) * c + ch);
    uint2 o;
    o.x = pack2(v.x, v.y); o.y = pack2(v.z, v.w);
    *reinterpret_cast<uint2*>(out + r * 9 * c + static_cast<long long>(tap) * c + ch) = o;
  }
}

// ---------------------------------------------------------------------------------------
// y[m, n] = act_out(bias[n] + sum_k act_in(x[m, k]) * W[n, k]); all fp32; small m.
// One warp per output feature n, 8 rows of x at a time. (time_embed MLP, emb_layers:
// unet.py:166-172,494-498; 1x1 quant convs: vae.py:569-570.)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
linear_f32_kernel(const float* __restrict__ x, long long ldx, int m, int k,
                  const float* __restrict__ wt, const float* __restrict__ bias, int nout,
                  int silu_in, int silu_out, float* __restrict__ y, long long ldy) {
  const int lane = threadIdx.x & 31;
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= nout) return;
  const float* wr = wt + static_cast<long long>(n) * k;
  for (int m0 = 0; m0 < m; m0 += 8) {
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
    for (int kk = lane; kk < k; kk += 32) {
      const float wv = __ldg(wr + kk);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (m0 + r < m) {
          float xv = x[static_cast<long long>(m0 + r) * ldx + kk];
          if (silu_in) xv = silu_f(xv);
          acc[r] += xv * wv;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float s = warp_sum(acc[r]);
      if (lane == 0 && m0 + r < m) {
        float v = s + (bias ? bias[n] : 0.f);
        if (silu_out) v = silu_f(v);
        y[static_cast<long long>(m0 + r) * ldy + n] = v;
      }
    }
  }
}

// Same contract for many rows of a tiny feature count (k, n <= 16: the VAE's 1x1 quant convs over
// every latent pixel, vae.py:569-570): one thread per row, weights broadcast from shared memory.
// (The warp-per-feature kernel above launches n/8 CTAs and walks the rows serially: 0.8 ms for the
// 4096 x 4 -> 4 post_quant_conv of one 512^2 image.)
__global__ void __launch_bounds__(256)
linear_f32_rows_kernel(const float* __restrict__ x, long long ldx, int m, int k,
                       const float* __restrict__ wt, const float* __restrict__ bias, int nout,
                       int silu_in, int silu_out, float* __restrict__ y, long long ldy) {
  __shared__ float s_w[16 * 16], s_b[16];
  for (int i = threadIdx.x; i < nout * k; i += blockDim.x) s_w[i] = wt[i];
  if (threadIdx.x < nout) s_b[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= m) return;
  float xv[16];
#pragma unroll
  for (int kk = 0; kk < 16; ++kk) {
    xv[kk] = kk < k ? x[static_cast<long long>(row) * ldx + kk] : 0.f;
    if (silu_in) xv[kk] = silu_f(xv[kk]);
  }
  for (int o = 0; o < nout; ++o) {
    float acc = 0.f;
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) if (kk < k) acc += xv[kk] * s_w[o * k + kk];
    float v = acc + s_b[o];
    if (silu_out) v = silu_f(v);
    y[static_cast<long long>(row) * ldy + o] = v;
  }
}

// y = alpha * x + z (fp32 rows of width c; z optional) -> fp32 y (optional) and a 16-bit copy with row
// stride ld16 (the first channels of the next dense block's concat buffer). RRDBNet's block tails:
// `out * 0.2 + x` (bsrnet.py:69-70) and the operand cast of every block output.
__global__ void __launch_bounds__(256)
axpby_cast_kernel(const float* __restrict__ x, float alpha, const float* __restrict__ z, long long rows, int c,
                  float* __restrict__ y, op_t* __restrict__ y16, long long ld16) {
  pdl_trigger();
  pdl_wait();
  const int V = c / 4;
  const long long total = rows * V;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / V;
    const int cc = static_cast<int>(i % V) * 4;
    float4 v = *reinterpret_cast<const float4*>(x + r * c + cc);
    v.x = __fmul_rn(v.x, alpha); v.y = __fmul_rn(v.y, alpha); v.z = __fmul_rn(v.z, alpha); v.w = __fmul_rn(v.w, alpha);
    if (z) {
      const float4 t = *reinterpret_cast<const float4*>(z + r * c + cc);
      v.x = __fadd_rn(v.x, t.x); v.y = __fadd_rn(v.y, t.y); v.z = __fadd_rn(v.z, t.z); v.w = __fadd_rn(v.w, t.w);
    }
    if (y) *reinterpret_cast<float4*>(y + r * c + cc) = v;
    if (y16) {
      uint2 o;
      o.x = pack2(v.x, v.y); o.y = pack2(v.z, v.w);
      *reinterpret_cast<uint2*>(y16 + r * ld16 + cc) = o;
    }
  }
}

// timestep_embedding(t, dim): cat(cos(t*f), sin(t*f)), f_i = exp(-ln(1e4) * i / half)
// (util.py:128-148). out [m, dim]
__global__ void timestep_embedding_kernel(const float* __restrict__ t, int m, int dim,
                                          float* __restrict__ out) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m * half) return;
  const int r = i / half, j = i % half;
  // frequency correctly rounded to fp32 (the reference holds fp32 freqs), argument in fp32
  const float f = static_cast<float>(exp(-9.210340371976184 * static_cast<double>(j) / static_cast<double>(half)));
  const float a = __fmul_rn(t[r], f);
  out[static_cast<long long>(r) * dim + j] = cosf(a);
  out[static_cast<long long>(r) * dim + half + j] = sinf(a);
}

// Row softmax fp32 [rows, cols] -> op16 (VAE mid attention, vae.py:232-282). One CTA per row.
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ s, long long lds, int cols, float scale,
                    op_t* __restrict__ out, long long ldo) {
  const float* row = s + static_cast<long long>(blockIdx.x) * lds;
  op_t* orow = out + static_cast<long long>(blockIdx.x) * ldo;
  __shared__ float red[8];
  __shared__ float bcast;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) m = fmaxf(m, row[c]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) { float v = red[0]; for (int i = 1; i < 8; ++i) v = fmaxf(v, red[i]); bcast = v; }
  __syncthreads();
  m = bcast;
  float sum = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) sum += __expf((row[c] - m) * scale);
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) { float v = 0.f; for (int i = 0; i < 8; ++i) v += red[i]; bcast = v; }
  __syncthreads();
  const float inv = 1.0f / bcast;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) orow[c] = f2op(__expf((row[c] - m) * scale) * inv);
}

// op16 NHWC nearest 2x upsample (SwinIR reconstruction tail, swinir.py:879-884).
__global__ void __launch_bounds__(256)
upsample2x_op16_kernel(const op_t* __restrict__ in, int n, int h, int w, int c, op_t* __restrict__ out) {
  const int cv = c / 8;
  const long long total = static_cast<long long>(n) * h * 2 * w * 2 * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % cv) * 8;
    long long p = i / cv;
    const int ox = static_cast<int>(p % (2 * w));
    const int oy = static_cast<int>((p / (2 * w)) % (2 * h));
    const int b = static_cast<int>(p / (4LL * w * h));
    const uint4 v = *reinterpret_cast<const uint4*>(in + ((static_cast<long long>(b) * h + oy / 2) * w + ox / 2) * c + ch);
    *reinterpret_cast<uint4*>(out + p * c + ch) = v;
  }
}

// SwinIR stem: (x - mean) * range, PixelUnshuffle(r) and NCHW -> NHWC op16 in one pass
// (swinir.py:860-861, 700-705). in [n,3,H,W] fp32 -> out [n, H/r, W/r, cpad] with channel
// index c*r*r + dy*r + dx (torch.nn.PixelUnshuffle order), zero padded to cpad.
__global__ void __launch_bounds__(256)
swin_stem_kernel(const float* __restrict__ in, int n, int hh, int ww, int r, float m0, float m1,
                 float m2, float range, int cpad, op_t* __restrict__ out) {
  const int ho = hh / r, wo = ww / r;
  const long long total = static_cast<long long>(n) * ho * wo * cpad;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % cpad);
    const long long p = i / cpad;
    float v = 0.f;
    if (ch < 3 * r * r) {
      const int c = ch / (r * r), dy = (ch / r) % r, dx = ch % r;
      const int ox = static_cast<int>(p % wo), oy = static_cast<int>((p / wo) % ho);
      const int b = static_cast<int>(p / (static_cast<long long>(wo) * ho));
      const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
      v = (in[((static_cast<long long>(b) * 3 + c) * hh + oy * r + dy) * ww + ox * r + dx] - mean) * range;
    }
    out[i] = f2op(v);
  }
}

// NCHW fp32 <-> NHWC fp32 (tiny tensors: latents, images)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, int n, int c, int hw, float* __restrict__ out) {
  const long long total = static_cast<long long>(n) * c * hw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    const long long p = i / c;
    const long long b = p / hw, q = p % hw;
    out[i] = in[(b * c + ch) * hw + q];
  }
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, int n, int c, int hw, float* __restrict__ out) {
  const long long total = static_cast<long long>(n) * c * hw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long q = i % hw;
    const long long t = i / hw;
    const int ch = static_cast<int>(t % c);
    const long long b = t / c;
    out[i] = in[(b * hw + q) * c + ch];
  }
}

// ---------------------------------------------------------------------------------------
// Fused sampler update (one launch per step): classifier-free guidance mix + x0 prediction +
// posterior / DDIM step. eps_c / eps_u, x, noise, x_out are NCHW fp32 [b, 4, h, w].
//   mode 0 (spaced, eps):  x0 = c0*x - c1*e ; mean = c2*x0 + c3*x ; x' = mean + c4*noise
//   mode 1 (spaced, v):    x0 = c0*x - c1*v   (c0 = sqrt_ac, c1 = sqrt(1-ac)); rest alike
//   mode 2 (ddim, eps):    x0 = (x - c1*e)/c0 ; x' = c2*x0 + c3*e + c4*noise
//   mode 3 (ddim, v):      e = c5*v + c1*x  then as mode 2
// (spaced_sampler.py:118-184, ddim_sampler.py:98-146). Coefficients are the fp32 table
// entries of the current step, read from device memory at coef[step_idx*8 ..].
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sampler_step_kernel(const float* __restrict__ eps_c, const float* __restrict__ eps_u, float cfg,
                    const float* __restrict__ x, const float* __restrict__ noise,
                    const float* __restrict__ coef, int mode, long long total,
                    float* __restrict__ x_out) {
  const float c0 = coef[0], c1 = coef[1], c2 = coef[2], c3 = coef[3], c4 = coef[4], c5 = coef[5];
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // explicit _rn intrinsics: no FMA contraction, so the update is bit-identical to the
    // reference's sequence of unfused torch elementwise ops
    float e = eps_c[i];
    if (eps_u) { const float u = eps_u[i]; e = __fadd_rn(u, __fmul_rn(cfg, __fsub_rn(e, u))); }
    const float xv = x[i];
    const float nz = noise ? noise[i] : 0.f;
    float r;
    if (mode == 0 || mode == 1) {
      const float x0 = __fsub_rn(__fmul_rn(c0, xv), __fmul_rn(c1, e));
      const float mean = __fadd_rn(__fmul_rn(c2, x0), __fmul_rn(c3, xv));
      r = __fadd_rn(mean, __fmul_rn(c4, nz));
    } else {
      if (mode == 3) e = __fadd_rn(__fmul_rn(c5, e), __fmul_rn(c1, xv));
      const float x0 = __fdiv_rn(__fsub_rn(xv, __fmul_rn(c1, e)), c0);
      r = __fadd_rn(__fadd_rn(__fmul_rn(c2, x0), __fmul_rn(c3, e)), __fmul_rn(c4, nz));
    }
    x_out[i] = r;
  }
}

// ---------------------------------------------------------------------------------------
// Tiled sampling (mixture of diffusers, utils/common.py:172-232):
//  gather: x_full NCHW [b,c,H,W] -> tiles NCHW [T*b, c, ts, ts] for the T windows in `coords`
//  blend : out = (sum_t tile_t * w) / (sum_t w), accumulated in the reference's row-major tile
//          order per output element (bit-reproducible, same order on every rank).
// coords: int32 [T][2] = (hi, wi).
// ---------------------------------------------------------------------------------------
__global__ void tile_gather_kernel(const float* __restrict__ full, int b, int c, int H, int W,
                                   const int* __restrict__ coords, int T, int ts,
                                   float* __restrict__ tiles) {
  const long long total = static_cast<long long>(T) * b * c * ts * ts;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % ts);
    const int y = static_cast<int>((i / ts) % ts);
    const int ch = static_cast<int>((i / (static_cast<long long>(ts) * ts)) % c);
    const int bi = static_cast<int>((i / (static_cast<long long>(ts) * ts * c)) % b);
    const int t = static_cast<int>(i / (static_cast<long long>(ts) * ts * c * b));
    const int hi = coords[2 * t], wi = coords[2 * t + 1];
    tiles[i] = full[((static_cast<long long>(bi) * c + ch) * H + hi + y) * W + wi + x];
  }
}

__global__ void tile_blend_kernel(const float* __restrict__ tiles, int b, int c, int H, int W,
                                  const int* __restrict__ coords, int T, int ts,
                                  const float* __restrict__ weights, float* __restrict__ out) {
  const long long total = static_cast<long long>(b) * c * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const int y = static_cast<int>((i / W) % H);
    const int ch = static_cast<int>((i / (static_cast<long long>(W) * H)) % c);
    const int bi = static_cast<int>(i / (static_cast<long long>(W) * H * c));
    float acc = 0.f, cnt = 0.f;
    for (int t = 0; t < T; ++t) {
      const int ty = y - coords[2 * t], tx = x - coords[2 * t + 1];
      if (ty < 0 || ty >= ts || tx < 0 || tx >= ts) continue;
      const float wv = weights[ty * ts + tx];
      const float v = tiles[(((static_cast<long long>(t) * b + bi) * c + ch) * ts + ty) * ts + tx];
      acc = __fadd_rn(acc, __fmul_rn(v, wv));   // unfused, reference accumulation order
      cnt = __fadd_rn(cnt, wv);
    }
    out[i] = __fdiv_rn(acc, cnt);
  }
}

inline int grid_for(long long total, int threads = 256) {
  long long g = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(dbir_sm_count()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int dbir_conv3x3_small_cin(const float* in1, const float* in2, int32_t c1, int32_t c2,
                                      int32_t n, int32_t h, int32_t w, const float* weight_kc,
                                      const float* bias, int32_t cout, float in_scale,
                                      float in_shift, float* out_nhwc, void* stream) {
  DBIR_REQUIRE(in1 && weight_kc && bias && out_nhwc, "dbir_conv3x3_small_cin: null pointer");
  DBIR_REQUIRE(c1 + c2 <= 16 && cout % 4 == 0, "dbir_conv3x3_small_cin: Cin<=16, Cout%%4==0");
  const long long strips = static_cast<long long>(n) * h * ((w + SC_STRIP - 1) / SC_STRIP);
  DBIR_REQUIRE(strips < (1LL << 31), "dbir_conv3x3_small_cin: too many strips");
  DBIR_CHECK_CUDA(dbir_launch(conv3x3_small_cin_kernel, dim3(static_cast<unsigned>(strips)), dim3(256), 0, ST(stream), in1, in2, c1,
                              c2, n, h, w, weight_kc, bias, cout, out_nhwc, in_scale, in_shift));
  return 0;
}

extern "C" int dbir_conv3x3_small_cout(const void* in_nhwc, int32_t n, int32_t h, int32_t w,
                                       int32_t cin, const float* weight, const float* bias,
                                       int32_t cout, float post_scale, const float* post_shift,
                                       float* out, int32_t out_nchw, void* stream) {
  DBIR_REQUIRE(in_nhwc && weight && bias && out, "dbir_conv3x3_small_cout: null pointer");
  DBIR_REQUIRE(cin % 8 == 0, "dbir_conv3x3_small_cout: Cin must be a multiple of 8");
  const long long ngroups = static_cast<long long>(n) * h * ((w + SO_PX - 1) / SO_PX);
  const int grid = grid_for(ngroups * 32);
  const op_t* in = reinterpret_cast<const op_t*>(in_nhwc);
#define LAUNCH_SC(C)                                                                              \
  DBIR_CHECK_CUDA(dbir_launch(conv3x3_small_cout_kernel<C>, dim3(grid), dim3(256), 0, ST(stream), in, n, h, w, cin, \
                              weight, bias, post_scale, post_shift, out, out_nchw))
  switch (cout) {
    case 3: LAUNCH_SC(3); break;
    case 4: LAUNCH_SC(4); break;
    case 8: LAUNCH_SC(8); break;
    default: dbir_set_error("dbir_conv3x3_small_cout: Cout must be 3, 4 or 8 (got %d)", cout); return -2;
  }
#undef LAUNCH_SC
  DBIR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dbir_im2col_s2(const float* in_nhwc, int32_t n, int32_t h, int32_t w, int32_t c,
                              int32_t pad_lo, void* out, void* stream) {
  DBIR_REQUIRE(in_nhwc && out && c % 4 == 0, "dbir_im2col_s2: bad args");
  const int ho = pad_lo ? (h + 2 - 3) / 2 + 1 : (h + 1 - 3) / 2 + 1;
  const int wo = pad_lo ? (w + 2 - 3) / 2 + 1 : (w + 1 - 3) / 2 + 1;
  const long long total = static_cast<long long>(n) * ho * wo * 9 * (c / 4);
  DBIR_CHECK_CUDA(dbir_launch(im2col_s2_kernel, dim3(grid_for(total)), dim3(256), 0, ST(stream), in_nhwc, n, h, w, c,
                              ho, wo, pad_lo, reinterpret_cast<op_t*>(out)));
  return 0;
}

extern "C" int dbir_linear_f32(const float* x, int64_t ldx, int32_t m, int32_t k,
                               const float* weight, const float* bias, int32_t n,
                               int32_t silu_in, int32_t silu_out, float* y, int64_t ldy,
                               void* stream) {
  DBIR_REQUIRE(x && weight && y && m > 0 && n > 0 && k > 0, "dbir_linear_f32: bad args");
  if (k <= 16 && n <= 16 && m >= 256) {
    linear_f32_rows_kernel<<<(m + 255) / 256, 256, 0, ST(stream)>>>(x, ldx, m, k, weight, bias, n, silu_in, silu_out, y, ldy);
    DBIR_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const int warps_per_cta = 8;
  linear_f32_kernel<<<(n + warps_per_cta - 1) / warps_per_cta, 256, 0, ST(stream)>>>(
      x, ldx, m, k, weight, bias, n, silu_in, silu_out, y, ldy);
  DBIR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dbir_axpby_cast(const float* x, float alpha, const float* z, int64_t rows, int32_t c, float* y,
                               void* y16, int64_t ld16, void* stream) {
  DBIR_REQUIRE(x && (y || y16) && rows > 0 && c > 0 && c % 4 == 0, "dbir_axpby_cast: bad args");
  DBIR_REQUIRE(!y16 || (ld16 >= c && ld16 % 4 == 0), "dbir_axpby_cast: ld16 must be >= c and a multiple of 4");
  DBIR_CHECK_CUDA(dbir_launch(axpby_cast_kernel, dim3(grid_for(rows * (c / 4))), dim3(256), 0, ST(stream), x, alpha, z,
                              static_cast<long long>(rows), c, y, reinterpret_cast<op_t*>(y16), static_cast<long long>(ld16)));
  return 0;
}

extern "C" int dbir_timestep_embedding(const float* t, int32_t m, int32_t dim, float* out,
                                       void* stream) {
  DBIR_REQUIRE(t && out && dim % 2 == 0, "dbir_timestep_embedding: bad args");
  const int total = m * (dim / 2);
  timestep_embedding_kernel<<<(total + 255) / 256, 256, 0, ST(stream)>>>(t, m, dim, out);
  DBIR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dbir_softmax_rows(const float* s, int64_t lds, int32_t rows, int32_t cols,
                                 float scale, void* out, int64_t ldo, void* stream) {
  DBIR_REQUIRE(s && out && rows > 0 && cols > 0, "dbir_softmax_rows: bad args");
  softmax_rows_kernel<<<rows, 256, 0, ST(stream)>>>(s, lds, cols, scale,
                                                     reinterpret_cast<op_t*>(out), ldo);
  DBIR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dbir_upsample2x_op16(const void* in, int32_t n, int32_t h, int32_t w, int32_t c,
                                    void* out, void* stream) {
  DBIR_REQUIRE(in && out && c % 8 == 0, "dbir_upsample2x_op16: bad args");
  const long long total = static_cast<long long>(n) * h * 2 * w * 2 * (c / 8);
  upsample2x_op16_kernel<<<grid_for(total), 256, 0, ST(stream)>>>(
      reinterpret_cast<const op_t*>(in), n, h, w, c, reinterpret_cast<op_t*>(out));
  DBIR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dbir_swin_stem(const float* in_nchw, int32_t n, int32_t h, int32_t w, int32_t r,
                              const float* mean3, float range, int32_t cpad, void* out,
                              void* stream) {
  DBIR_REQUIRE(in_nchw && out && mean3 && h % r == 0 && w % r == 0 && cpad >= 3 * r * r,
               "dbir_swin_stem: bad args");
  const long long total = static_cast<long long>(n) * (h / r) * (w / r) * cpad;
  swin_stem_kernel<<<grid_for(total), 256, 0, ST(stream)>>>(in_nchw, n, h, w, r, mean3[0], mean3[1],
                                                            mean3[2], range, cpad,
                                                            reinterpret_cast<op_t*>(out));
  DBIR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dbir_nchw_to_nhwc(const float* in, int32_t n, int32_t c, int32_t hw, float* out,
                                 void* stream) {
  const long long total = static_cast<long long>(n) * c * hw;
  nchw_to_nhwc_kernel<<<grid_for(total), 256, 0, ST(stream)>>>(in, n, c, hw, out);
  DBIR_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int dbir_nhwc_to_nchw(const float* in, int32_t n, int32_t c, int32_t hw, float* out,
                                 void* stream) {
  const long long total = static_cast<long long>(n) * c * hw;
  nhwc_to_nchw_kernel<<<grid_for(total), 256, 0, ST(stream)>>>(in, n, c, hw, out);
  DBIR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dbir_sampler_step(const float* eps_cond, const float* eps_uncond, float cfg_scale,
                                 const float* x, const float* noise, const float* coef,
                                 int32_t mode, int64_t numel, float* x_out, void* stream) {
  DBIR_REQUIRE(eps_cond && x && coef && x_out && mode >= 0 && mode <= 3, "dbir_sampler_step: bad args");
  sampler_step_kernel<<<grid_for(numel), 256, 0, ST(stream)>>>(eps_cond, eps_uncond, cfg_scale, x,
                                                               noise, coef, mode, numel, x_out);
  DBIR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dbir_tile_gather(const float* full, int32_t b, int32_t c, int32_t h, int32_t w,
                                const int32_t* coords, int32_t ntiles, int32_t tile,
                                float* tiles, void* stream) {
  DBIR_REQUIRE(full && coords && tiles, "dbir_tile_gather: null pointer");
  const long long total = static_cast<long long>(ntiles) * b * c * tile * tile;
  tile_gather_kernel<<<grid_for(total), 256, 0, ST(stream)>>>(full, b, c, h, w, coords, ntiles, tile, tiles);
  DBIR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int dbir_tile_blend(const float* tiles, int32_t b, int32_t c, int32_t h, int32_t w,
                               const int32_t* coords, int32_t ntiles, int32_t tile,
                               const float* weights, float* out, void* stream) {
  DBIR_REQUIRE(tiles && coords && weights && out, "dbir_tile_blend: null pointer");
  const long long total = static_cast<long long>(b) * c * h * w;
  tile_blend_kernel<<<grid_for(total), 256, 0, ST(stream)>>>(tiles, b, c, h, w, coords, ntiles, tile, weights, out);
  DBIR_CHECK_CUDA(cudaGetLastError());
  return 0;
}
