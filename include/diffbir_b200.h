/* diffbir_b200 — C ABI of libdiffbir_b200.so (hand-written sm_100a kernels).
 *
 * This is the drop-in boundary for the DiffBIR restoration hot path. The reference
 * (XPixelGroup/DiffBIR) has no FFI: every op below replaces a PyTorch library call made
 * from the reference's nn.Module.forward methods; each entry cites the reference
 * file:line it stands in for (paths relative to the reference checkout).
 *
 * Conventions
 *  - All pointers are DEVICE pointers owned by the caller (PyTorch allocations); the
 *    library never allocates or frees device memory and only enqueues work on `stream`
 *    (a cudaStream_t passed as void*), so every call is CUDA-graph capturable.
 *  - Return value: 0 on success, < 0 on error; dbir_last_error() returns a thread-local
 *    message. No C++ exceptions cross the boundary.
 *  - "op16" is the 16-bit tensor-core operand format reported by dbir_operand_kind()
 *    (1 = IEEE fp16, 0 = bf16). Activations are NHWC; residual streams are fp32.
 */
#ifndef DIFFBIR_B200_H
#define DIFFBIR_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- library info ------------------------------------------------------------------ */
const char* dbir_version(void);
const char* dbir_last_error(void);
int dbir_operand_kind(void);          /* 1 = fp16 operands, 0 = bf16 operands */
int dbir_sm_count(void);              /* SMs of the current device (cached) */

/* ---- tcgen05 GEMM / implicit 3x3 convolution --------------------------------------
 * out = residual + alpha * act(A * B^T + bias + rowvec)          (GEGLU: a * gelu(g))
 * Replaces nn.Linear (attention.py:22,35-42,67-73,310,331; swinir.py:21-31,104,106),
 * nn.Conv2d 3x3 / 1x1 (unet.py:67,99,149-153,173-188; controlnet.py:309-312;
 * vae.py:24-27,77-95,241-252,445-447,520-522; swinir.py:472,700-705,790-811) and the
 * elementwise adds around them (unet.py:216-223, attention.py:265-274,353,
 * controlnet.py:36-43, cldm.py:164).
 */
typedef struct dbir_gemm_args {
  const void* a;        /* op16. a_mode 0: [M, lda]; a_mode 1: NHWC [img_n, img_h, img_w, img_c] */
  const void* b;        /* op16 packed weight [N, ldb], K contiguous (conv: k = tap*C + c) */
  void* out;            /* fp32 or op16 [rows, ldo] */
  const float* bias;    /* [N] or NULL */
  const float* rowvec;  /* [M / rows_per_vec, N] added per row group (time embedding) or NULL */
  const float* residual;/* fp32 [rows, ldr] or NULL */
  int64_t lda, ldb, ldo, ldr;   /* element strides; lda/ldb 0 = K */
  int32_t M, N, K;
  int32_t a_mode;       /* 0 plain matrix, 1 implicit conv (ksize x ksize, stride 1, pad ksize/2) */
  int32_t img_n, img_h, img_w, img_c, ksize;
  int32_t rows_per_vec; /* a_mode 0 only; conv mode uses the image index */
  int32_t out_kind;     /* 0 fp32, 1 op16 */
  int32_t act;          /* 0 none, 1 GELU(erf), 2 LeakyReLU(act_param), 3 SiLU */
  int32_t geglu;        /* 1: B rows packed per force_bn-row tile as [values | gates]; out has N/2 columns */
  int32_t force_bn;     /* 0 = auto tile width, else 32/64/128/160/256 (required with geglu) */
  float alpha;          /* scale applied before the residual add (control strength) */
  float act_param;
} dbir_gemm_args;
int dbir_gemm(const dbir_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFBIR_B200_H */
