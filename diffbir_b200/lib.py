"""ctypes binding of libdiffbir_b200.so (the C ABI declared in include/diffbir_b200.h).

The product path fails loudly when the CUDA library is missing: there is no CPU or
PyTorch-eager fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

import os

_HERE = Path(__file__).resolve().parent
# DBIR_OPERANDS=bf16 loads the -DDBIR_OPERAND_BF16 build (bf16 tensor-core operands); default fp16 operands.
_OPERANDS = os.environ.get("DBIR_OPERANDS", "fp16").lower()
if _OPERANDS not in ("fp16", "bf16"):
    raise ValueError(f"DBIR_OPERANDS={_OPERANDS!r}: expected fp16 or bf16")
LIB_PATH = _HERE / ("libdiffbir_b200_bf16.so" if _OPERANDS == "bf16" else "libdiffbir_b200.so")
if os.environ.get("DBIR_LIB_TAG"):      # experiment builds: python -m diffbir_b200.build --tag=<tag> -D...
    LIB_PATH = _HERE / f"libdiffbir_b200_{os.environ['DBIR_LIB_TAG']}.so"
_lib = None
_launches = 0  # kernels launched through this binding (bench.py reports it)


class DbirError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p), ("out", C.c_void_p),
        ("bias", C.c_void_p), ("rowvec", C.c_void_p), ("residual", C.c_void_p),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldo", C.c_int64), ("ldr", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("a_mode", C.c_int32),
        ("img_n", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32), ("img_c", C.c_int32),
        ("ksize", C.c_int32),
        ("rows_per_vec", C.c_int32), ("out_kind", C.c_int32), ("act", C.c_int32),
        ("geglu", C.c_int32), ("force_bn", C.c_int32),
        ("alpha", C.c_float), ("act_param", C.c_float),
        ("bias_per_row", C.c_int32), ("groups", C.c_int32),
        ("out2", C.c_void_p), ("ldo2", C.c_int64),
        ("splitk_ws", C.c_void_p), ("splitk_ws_bytes", C.c_int64),
        ("split_k", C.c_int32), ("cta_pair", C.c_int32),
        ("debug_stamps", C.c_void_p),
        ("gn_partials", C.c_void_p), ("gn_rows_per_img", C.c_int32), ("reserved2", C.c_int32),
        ("prefetch_ptr", C.c_void_p), ("prefetch_bytes", C.c_int64),
    ]


def load():
    """Loads the shared library (building is the job of __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise DbirError(
            f"{LIB_PATH} not found: build it with `python -m diffbir_b200.build{' --bf16' if _OPERANDS == 'bf16' else ''}` "
            "(there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    lib.dbir_last_error.restype = C.c_char_p
    lib.dbir_version.restype = C.c_char_p
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().dbir_last_error().decode(errors="replace")
        raise DbirError(f"{what} failed (rc={rc}): {msg}")


_profile = None   # list of (kind, flops, start_event, end_event) while profiling


def profile_begin():
    """bench.py's kernel census: brackets every GEMM / attention launch with CUDA events."""
    global _profile
    _profile = []


def profile_end():
    global _profile
    torch.cuda.synchronize()
    out = [(k, info, fl, e0.elapsed_time(e1)) for k, info, fl, e0, e1 in _profile]
    _profile = None
    return out


_record = None    # list of (kind, info, flops, thunk) while recording


def record_begin():
    """Records every GEMM / attention launch (arguments included) so bench.py can replay exactly the
    tensor-core kernels of one forward inside a CUDA graph and time them with CUDA events."""
    global _record
    _record = []


def record_end():
    global _record
    out, _record = _record, None
    return out


class _Prof:
    def __init__(self, kind, info, flops):
        self.kind, self.info, self.flops = kind, info, flops

    def __enter__(self):
        if _profile is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *a):
        if _profile is not None:
            self.e1.record()
            _profile.append((self.kind, self.info, self.flops, self.e0, self.e1))


def count_launch(n: int = 1):
    global _launches
    _launches += n


def launches() -> int:
    return _launches


def operand_dtype() -> torch.dtype:
    return torch.float16 if load().dbir_operand_kind() == 1 else torch.bfloat16


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


ACT = {None: 0, "none": 0, "gelu": 1, "lrelu": 2, "silu": 3}


def gemm(a, b, out, *, M, N, K, bias=None, rowvec=None, rows_per_vec=0, residual=None,
         lda=0, ldb=0, ldo=None, ldr=None, conv=None, act=None, act_param=0.0, alpha=1.0,
         geglu=False, force_bn=0, bias_per_row=False, out2=None, ldo2=None, splitk_ws=None,
         split_k=0, cta_pair=0, debug_stamps=None, gn_partials=None, gn_rows_per_img=0, prefetch=None, groups=1):
    """out = residual + alpha * act(A @ B^T + bias + rowvec). conv = (n, h, w, c, ksize)."""
    lib = load()
    g = GemmArgs()
    g.a, g.b, g.out = a.data_ptr(), b.data_ptr(), out.data_ptr()
    g.bias, g.rowvec, g.residual = _ptr(bias), _ptr(rowvec), _ptr(residual)
    n_out = N // 2 if geglu else N
    g.lda, g.ldb = lda, ldb
    g.ldo = n_out if ldo is None else ldo
    g.ldr = n_out if ldr is None else ldr
    g.M, g.N, g.K = M, N, K
    if conv is not None:
        g.a_mode = 1
        g.img_n, g.img_h, g.img_w, g.img_c, g.ksize = conv
    else:
        g.a_mode = 0
    g.rows_per_vec = rows_per_vec
    g.out_kind = 0 if out.dtype == torch.float32 else 1
    g.act = ACT[act]
    g.act_param = act_param
    g.alpha = alpha
    g.geglu = 1 if geglu else 0
    g.force_bn = force_bn
    g.bias_per_row = 1 if bias_per_row else 0
    g.groups = groups
    g.out2 = _ptr(out2)
    g.ldo2 = (n_out if ldo2 is None else ldo2)
    if splitk_ws is not None:
        g.splitk_ws = splitk_ws.data_ptr()
        g.splitk_ws_bytes = splitk_ws.numel() * splitk_ws.element_size()
    g.split_k = split_k
    g.cta_pair = cta_pair
    g.debug_stamps = _ptr(debug_stamps)
    g.gn_partials = _ptr(gn_partials)
    g.gn_rows_per_img = gn_rows_per_img
    if prefetch is not None:
        g.prefetch_ptr, g.prefetch_bytes = prefetch
    with _Prof("conv" if conv is not None else "gemm", (M, N, K), 2.0 * M * N * K):
        check(lib.dbir_gemm(C.byref(g), C.c_void_p(stream_ptr())), "dbir_gemm")
    if _record is not None:
        keep = (a, b, out, bias, rowvec, residual, splitk_ws, gn_partials)     # keep buffers alive
        _record.append(("conv" if conv is not None else "gemm", (M, N, K), 2.0 * M * N * K,
                        lambda g=g, keep=keep: check(lib.dbir_gemm(C.byref(g), C.c_void_p(stream_ptr())), "dbir_gemm")))
    count_launch()


def _sp():
    return C.c_void_p(stream_ptr())


def _fp(t):
    return C.c_void_p(None if t is None else t.data_ptr())


def attention_ws_bytes(batch, heads, sq, skv) -> int:
    """Bytes of (zero-initialised, per-stream) workspace that let dbir_attention_sk balance this
    shape over all SMs; 0 when whole tiles per CTA are used anyway."""
    lib = load()
    lib.dbir_attention_ws_bytes.restype = C.c_int64
    return int(lib.dbir_attention_ws_bytes(batch, heads, sq, skv))


def attention(q, k, v, out, *, batch, heads, sq, skv, ldq, ldk, ldv, ldo, ws=None):
    """Flash attention, head_dim 64; q/k/v/out are op16 tensors (possibly column slices).
    ws: optional zero-initialised uint8/float32 scratch of attention_ws_bytes() bytes."""
    ws_bytes = ws.numel() * ws.element_size() if ws is not None else 0

    def call():
        check(load().dbir_attention_sk(_fp(q), _fp(k), _fp(v), _fp(out), batch, heads, sq, skv,
                                       C.c_int64(ldq), C.c_int64(ldk), C.c_int64(ldv), C.c_int64(ldo),
                                       _fp(ws), C.c_int64(ws_bytes), _sp()), "dbir_attention_sk")
    with _Prof("attention", (batch, heads, sq, skv), 4.0 * batch * heads * sq * skv * 64):
        call()
    if _record is not None:
        _record.append(("attention", (batch, heads, sq, skv), 4.0 * batch * heads * sq * skv * 64, call))
    count_launch()


def gn_workspace_floats(n, hw, c) -> int:
    lib = load()
    lib.dbir_gn_workspace_floats.restype = C.c_int64
    return int(lib.dbir_gn_workspace_floats(n, hw, c))


def gemm_tuned_problems() -> int:
    """Problem signatures dbir_gemm has planned (autotuned) so far."""
    return int(load().dbir_gemm_tuned_problems())


def gemm_clear_plans() -> None:
    load().dbir_gemm_clear_plans()


def gemm_gn_slots(conv_h=0, conv_w=0, rows_per_img=0) -> int:
    return int(load().dbir_gemm_gn_slots(conv_h, conv_w, rows_per_img))


def gn_finalize(p1, slots1, c1, p2, slots2, c2, n, hw, eps, stats):
    check(load().dbir_gn_finalize(_fp(p1), slots1, c1, _fp(p2), slots2, c2, n, hw, C.c_float(eps), _fp(stats),
                                  _sp()), "dbir_gn_finalize")
    count_launch()


def gn_stats(src1, src2, c1, c2, n, hw, eps, stats, workspace):
    check(load().dbir_gn_stats(_fp(src1), _fp(src2), c1, c2, n, hw, C.c_float(eps), _fp(stats),
                               _fp(workspace), _sp()), "dbir_gn_stats")
    count_launch()


def gn_apply(src1, src2, c1, c2, n, h, w, stats, gamma, beta, out, *, norm=True, silu=True,
             upsample=1, out_raw=None, imgs_per_group=0):
    check(load().dbir_gn_apply(_fp(src1), _fp(src2), c1, c2, n, h, w, _fp(stats), _fp(gamma),
                               _fp(beta), 1 if norm else 0, 1 if silu else 0, upsample, _fp(out),
                               _fp(out_raw), imgs_per_group, _sp()), "dbir_gn_apply")
    count_launch()


def layernorm(x, ldx, rows, c, gamma, beta, out, ldo, eps=1e-5, rows_per_group=0):
    check(load().dbir_layernorm(_fp(x), C.c_int64(ldx), rows, c, _fp(gamma), _fp(beta),
                                C.c_float(eps), _fp(out), C.c_int64(ldo),
                                0 if out.dtype == torch.float32 else 1, rows_per_group, _sp()), "dbir_layernorm")
    count_launch()


def window_attention(qkv, ldq, batch, h, w, heads, head_dim, shift, bias_table, mask_value, out, ldo):
    check(load().dbir_window_attention(_fp(qkv), C.c_int64(ldq), batch, h, w, heads, head_dim, 8, shift, _fp(bias_table),
                                       C.c_float(mask_value), _fp(out), C.c_int64(ldo), _sp()), "dbir_window_attention")
    count_launch()


def swin_window_attention(qkv, ldq, batch, h, w, shift, bias_table, out, ldo):
    check(load().dbir_swin_window_attention(_fp(qkv), C.c_int64(ldq), batch, h, w, 6, 30, 8, shift,
                                            _fp(bias_table), _fp(out), C.c_int64(ldo), _sp()),
          "dbir_swin_window_attention")
    count_launch()


def conv3x3_small_cin(in1, in2, c1, c2, n, h, w, weight_kc, bias, cout, out, in_scale=1.0,
                      in_shift=0.0):
    check(load().dbir_conv3x3_small_cin(_fp(in1), _fp(in2), c1, c2, n, h, w, _fp(weight_kc),
                                        _fp(bias), cout, C.c_float(in_scale), C.c_float(in_shift),
                                        _fp(out), _sp()), "dbir_conv3x3_small_cin")
    count_launch()


def conv3x3_small_cout(x, n, h, w, cin, weight, bias, cout, out, *, nchw=True, post_scale=1.0,
                       post_shift=None):
    check(load().dbir_conv3x3_small_cout(_fp(x), n, h, w, cin, _fp(weight), _fp(bias), cout,
                                         C.c_float(post_scale), _fp(post_shift), _fp(out),
                                         1 if nchw else 0, _sp()), "dbir_conv3x3_small_cout")
    count_launch()


def im2col_s2(x, n, h, w, c, pad_lo, out):
    check(load().dbir_im2col_s2(_fp(x), n, h, w, c, pad_lo, _fp(out), _sp()), "dbir_im2col_s2")
    count_launch()


def linear_f32(x, ldx, m, k, weight, bias, n, y, ldy, silu_in=False, silu_out=False):
    check(load().dbir_linear_f32(_fp(x), C.c_int64(ldx), m, k, _fp(weight), _fp(bias), n,
                                 1 if silu_in else 0, 1 if silu_out else 0, _fp(y),
                                 C.c_int64(ldy), _sp()), "dbir_linear_f32")
    count_launch()


def axpby_cast(x, alpha, z, rows, c, y=None, y16=None, ld16=0):
    check(load().dbir_axpby_cast(_fp(x), C.c_float(alpha), _fp(z), C.c_int64(rows), c, _fp(y), _fp(y16),
                                 C.c_int64(ld16), _sp()), "dbir_axpby_cast")
    count_launch()


def timestep_embedding(t, m, dim, out):
    check(load().dbir_timestep_embedding(_fp(t), m, dim, _fp(out), _sp()), "dbir_timestep_embedding")
    count_launch()


def softmax_rows(s, lds, rows, cols, scale, out, ldo):
    check(load().dbir_softmax_rows(_fp(s), C.c_int64(lds), rows, cols, C.c_float(scale), _fp(out),
                                   C.c_int64(ldo), _sp()), "dbir_softmax_rows")
    count_launch()


def upsample2x_op16(x, n, h, w, c, out):
    check(load().dbir_upsample2x_op16(_fp(x), n, h, w, c, _fp(out), _sp()), "dbir_upsample2x_op16")
    count_launch()


def swin_stem(x, n, h, w, r, mean3, rng, cpad, out):
    arr = (C.c_float * 3)(*mean3)
    check(load().dbir_swin_stem(_fp(x), n, h, w, r, arr, C.c_float(rng), cpad, _fp(out), _sp()),
          "dbir_swin_stem")
    count_launch()


def nchw_to_nhwc(x, n, c, hw, out):
    check(load().dbir_nchw_to_nhwc(_fp(x), n, c, hw, _fp(out), _sp()), "dbir_nchw_to_nhwc")
    count_launch()


def nhwc_to_nchw(x, n, c, hw, out):
    check(load().dbir_nhwc_to_nchw(_fp(x), n, c, hw, _fp(out), _sp()), "dbir_nhwc_to_nchw")
    count_launch()


def sampler_step(eps_c, eps_u, cfg, x, noise, coef, mode, numel, x_out):
    check(load().dbir_sampler_step(_fp(eps_c), _fp(eps_u), C.c_float(cfg), _fp(x), _fp(noise),
                                   _fp(coef), mode, C.c_int64(numel), _fp(x_out), _sp()),
          "dbir_sampler_step")
    count_launch()


def tile_gather(full, b, c, h, w, coords, ntiles, tile, tiles):
    check(load().dbir_tile_gather(_fp(full), b, c, h, w, _fp(coords), ntiles, tile, _fp(tiles),
                                  _sp()), "dbir_tile_gather")
    count_launch()


def tile_blend(tiles, b, c, h, w, coords, ntiles, tile, weights, out):
    check(load().dbir_tile_blend(_fp(tiles), b, c, h, w, _fp(coords), ntiles, tile, _fp(weights),
                                 _fp(out), _sp()), "dbir_tile_blend")
    count_launch()


def wavelet_fix(sample, style, out_u8=None, out_f32=None):
    """Fused colour fix (+ uint8 quantisation): sample NCHW view in [-1, 1], style NCHW view in [0, 1],
    both fp32 with unit column stride; exactly one of out_u8 (NHWC uint8) / out_f32 (NCHW fp32)."""
    b, c, h, w = sample.shape
    assert c == 3 and tuple(style.shape) == (b, c, h, w) and sample.stride(3) == 1 and style.stride(3) == 1
    i64 = C.c_int64
    check(load().dbir_wavelet_fix(_fp(sample), i64(sample.stride(0)), i64(sample.stride(1)), i64(sample.stride(2)),
                                  _fp(style), i64(style.stride(0)), i64(style.stride(1)), i64(style.stride(2)),
                                  b, h, w, _fp(out_u8), _fp(out_f32), _sp()), "dbir_wavelet_fix")
    count_launch()
