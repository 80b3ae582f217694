"""ctypes binding of libvitcap_b200.so (the C ABI declared in include/vitcap_b200.h).

PyTorch is used for device memory and streams only: every wrapper passes ``tensor.data_ptr()`` and the
current CUDA stream handle; no torch type crosses the boundary. There is no fallback: if the shared library
cannot be loaded the import of a compute entry point raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# VITCAP_STORE=fp16 (process-wide, read at import): the fast mode stores EVERY 16-bit operand -- weights, q|k|v, attention
# probabilities and outputs, GELU outputs, LayerNorm-fold row copies, K/V cache -- as IEEE halves instead of bfloat16
# (libvitcap_b200_f16.so: the same sources built with -DVC_STORE_F16; DESIGN.md section 4a''). STORE is the torch dtype of those
# tensors; every "bf16" in the names of this module then reads "the 16-bit storage type".
HALF_STORE = os.environ.get("VITCAP_STORE", "bf16").lower() in ("fp16", "f16", "half")
STORE = torch.float16 if HALF_STORE else torch.bfloat16
# VITCAP_LIB: another build of the same ABI (A/B measurements of a kernel change in tools/)
LIB_PATH = os.environ.get("VITCAP_LIB") or os.path.join(_HERE, "lib", "libvitcap_b200_f16.so" if HALF_STORE else "libvitcap_b200.so")

ACT_NONE, ACT_GELU, ACT_TANH = 0, 1, 2

_c = ctypes
_P, _I, _F, _D, _SZ, _U64, _LL = _c.c_void_p, _c.c_int, _c.c_float, _c.c_double, _c.c_size_t, _c.c_uint64, _c.c_longlong

# name -> argtypes; must list every symbol the header declares (checked by tests/test_host_cpu.py::
# test_header_symbols_are_exported_and_bound)
SIGNATURES = {
    "vc_linear": [_I, _P, _I, _P, _I, _P, _P, _I, _I, _I, _P, _I, _I, _I, _I, _P],
    "vc_linear_simt": [_I, _P, _I, _P, _I, _P, _P, _I, _I, _I, _P, _I, _I, _I, _I, _P],
    "vc_linear_tc": [_P, _I, _P, _I, _P, _P, _I, _I, _I, _P, _I, _I, _I, _I, _I, _P],
    "vc_linear_ln_emit": [_P, _I, _P, _I, _P, _P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _P],
    "vc_linear_ln_fold": [_P, _I, _P, _I, _P, _P, _P, _I, _F, _P, _I, _I, _I, _I, _I, _P],
    "vc_linear_ln_emit_postln": [_P, _I, _P, _I, _P, _P, _I, _P, _I, _P, _I, _P, _P, _F, _P, _I, _P, _I, _I, _I, _P],
    "vc_patchify": [_I, _P, _P, _I, _I, _I, _P],
    "vc_patchify_u8": [_I, _P, _P, _I, _I, _I, _I, _P],
    "vc_resize_crop_plan": [_P, _I, _I, _I, _P, _P, _P],
    "vc_resize_crop_u8": [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "vc_assemble_tokens": [_P, _P, _P, _P, _I, _I, _I, _P],
    "vc_layernorm": [_I, _P, _I, _P, _P, _F, _P, _I, _P, _I, _I, _I, _P],
    "vc_split_bf16x3": [_P, _I, _P, _I, _I, _I, _P],
    "vc_linear_x3": [_P, _I, _P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _P],
    "vc_gather_rows": [_I, _P, _SZ, _P, _I, _I, _I, _P],
    "vc_assemble_ctx": [_I, _P, _P, _P, _P, _I, _I, _I, _P],
    "vc_assemble_ctx_pitched": [_I, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "vc_label_rows": [_I, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _F, _P, _P, _I, _I, _I, _I, _P],
    "vc_attention": [_I, _P, _P, _I, _I, _I, _F, _P],
    "vc_attention_labels": [_I, _P, _P, _I, _I, _I, _F, _I, _P, _P],
    "vc_attention_labels_simt": [_I, _P, _P, _I, _I, _I, _F, _I, _P, _P],
    "vc_decode_attention_labels": [_I, _P, _P, _P, _P, _I, _I, _P, _I, _I, _I, _F, _P],
    "vc_decode_attention_skip": [_P, _P, _P, _P, _I, _I, _P, _I, _I, _I, _F, _P, _P, _P],
    "vc_decode_attention_labels_simt": [_I, _P, _P, _P, _P, _I, _I, _P, _I, _I, _I, _F, _P],
    "vc_attention_simt": [_I, _P, _P, _I, _I, _I, _F, _P],
    "vc_cls_attention": [_I, _P, _I, _P, _P, _I, _I, _I, _I, _F, _P],
    "vc_tag_topk": [_P, _I, _I, _I, _I, _F, _P, _P, _P, _P],
    "vc_embed_ln": [_I, _P, _I, _I, _I, _P, _P, _P, _P, _P, _F, _P, _P, _I, _I, _P],
    "vc_decode_attention": [_I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    "vc_decode_attention_simt": [_I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    "vc_token_step": [_P, _I, _I, _I, _I, _F, _U64, _P, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P],
    "vc_greedy_finalize": [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P],
    "vc_graph_if_any_begin": [_P, _I, _I, _P, _P],
    "vc_graph_if_end": [_P],
    "vc_stream_create": [_P],
    "vc_stream_destroy": [_P],
    "vc_beam_row_topk": [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "vc_beam_advance": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _D, _I, _P, _I, _P],
    "vc_beam_finalize": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P],
    "vc_filter_logits": [_P, _I, _I, _I, _F, _I, _F, _I, _P],
    "vc_dec_linear": [_I, _I, _P, _I, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "vc_dec_vocab_argmax": [_I, _P, _I, _P, _I, _P, _P, _I, _I, _I, _I, _P],
    "vc_finish_ln": [_P, _I, _SZ, _I, _P, _I, _P, _I, _P, _P, _F, _P, _I, _P, _I, _I, _I, _I, _P],
    "vc_token_step_partials": [_P, _I, _I, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P],
}

_lib = None

# bench.py instrumentation: when set to a list, every eager tensor-core GEMM launch is bracketed by CUDA events on the
# launching stream and (flops, start_event, end_event) is appended (graph-captured launches are skipped).
GEMM_PROFILE = None


def load_library(path=None):
    """Loads the shared library (no GPU needed) and declares all prototypes."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            "vitcap_b200: %s is missing -- build it with `python -m vitcap_b200.build` "
            "(there is no CPU or PyTorch fallback for the caption path)" % p)
    lib = ctypes.CDLL(p)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _I
    lib.vc_last_error.restype = ctypes.c_char_p
    lib.vc_last_error.argtypes = []
    lib.vc_abi_version.restype = _I
    lib.vc_launch_count.restype = _LL
    lib.vc_reset_launch_count.restype = None
    lib.vc_set_pdl.restype = None
    lib.vc_set_pdl.argtypes = [_I]
    lib.vc_get_pdl.restype = _I
    lib.vc_check_device.restype = _I
    lib.vc_check_device.argtypes = []
    if path is None:
        _lib = lib
    return lib


def _check(rc, name):
    if rc != 0:
        msg = load_library().vc_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (%d): %s" % (name, rc, msg))


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda, "vitcap_b200 kernels need CUDA tensors"
    return t.data_ptr()


def launch_count():
    return int(load_library().vc_launch_count())


def reset_launch_count():
    load_library().vc_reset_launch_count()


_BODY_STREAMS = {}


def body_stream(device):
    """The stream conditional-node bodies are captured on: one per device, created through the library (not a pooled torch
    stream, which may alias the capturing stream or carry a loader's copies) and kept for the life of the process."""
    idx = torch.device(device).index
    idx = torch.cuda.current_device() if idx is None else idx
    if idx not in _BODY_STREAMS:
        with torch.cuda.device(idx):
            h = ctypes.c_void_p()
            _check(load_library().vc_stream_create(ctypes.byref(h)), "vc_stream_create")
        _BODY_STREAMS[idx] = torch.cuda.ExternalStream(h.value, device=torch.device("cuda", idx))
    return _BODY_STREAMS[idx]


class graph_if_any:
    """Context manager: while the current stream is being captured into a CUDA graph, the kernels launched inside the block become
    the body of a conditional IF node that a replay executes only when any(flags != 0) (invert: any(flags == 0)) holds on the
    device at that point (vc_graph_if_any_begin / vc_graph_if_end); the block runs on `body_stream`. Outside a capture (eager
    runs) the block simply executes."""

    def __init__(self, flags, body_stream, invert=False, enabled=True):
        self.flags, self.body, self.invert = flags, body_stream, invert
        self.active = bool(enabled) and body_stream is not None and torch.cuda.is_current_stream_capturing()
        self._ctx = None

    def __enter__(self):
        if self.active:
            assert self.flags.dtype == torch.int32 and self.flags.is_contiguous()
            _check(load_library().vc_graph_if_any_begin(_ptr(self.flags), self.flags.numel(), int(self.invert), _stream(),
                                                        self.body.cuda_stream), "vc_graph_if_any_begin")
            self._ctx = torch.cuda.stream(self.body)
            self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.active:
            self._ctx.__exit__(*exc)
            rc = load_library().vc_graph_if_end(self.body.cuda_stream)
            if exc[0] is None:
                _check(rc, "vc_graph_if_end")
        return False


def set_pdl(mode):
    """Programmatic dependent launch between consecutive kernels: 0 never, 1 (default) launches captured into a CUDA graph
    only, 2 every launch. Graphs captured earlier keep the setting they were captured with."""
    load_library().vc_set_pdl(int(mode))


def get_pdl():
    return int(load_library().vc_get_pdl())


def check_device():
    """Raises unless the current CUDA device is an sm_100 part (the library holds sm_100a code only)."""
    _check(load_library().vc_check_device(), "vc_check_device")


def _is_bf16(t):
    """1 = the 16-bit storage type of this process (ops.STORE), 0 = fp32. The OTHER 16-bit float type is refused: the library
    would otherwise read it as fp32 and run off the end of the buffer."""
    if t.dtype == STORE:
        return 1
    if t.dtype in (torch.float16, torch.bfloat16):
        raise TypeError("this process stores 16-bit operands as %s (VITCAP_STORE); got a %s tensor" % (STORE, t.dtype))
    return 0


def linear(a, w, bias, out, act=ACT_NONE, resid=None, M=None, lda=None, ldo=None, impl="auto", tile_n=0):
    """out[M,N] = act(a[M,K] @ w[N,K]^T + bias) (+ resid). a/w: both bf16 (tensor cores) or both fp32 (exact mode).
    ``a`` and ``out`` may be strided row views (pitch = stride(0))."""
    lib = load_library()
    K = w.shape[1]
    N = w.shape[0]
    M = a.shape[0] if M is None else M
    lda = a.stride(0) if lda is None else lda
    ldo = out.stride(0) if ldo is None else ldo
    assert a.dtype == w.dtype and a.stride(-1) == 1 and w.stride(-1) == 1 and out.stride(-1) == 1
    out_f32 = 1 if out.dtype == torch.float32 else 0
    ldr = resid.stride(0) if resid is not None else 0
    if resid is not None:
        assert resid.dtype == torch.float32 and out_f32
    args = (_ptr(a), lda, _ptr(w), w.stride(0), _ptr(bias), _ptr(out), ldo, out_f32, act, _ptr(resid), ldr, M, N, K)
    if impl == "simt":
        _check(lib.vc_linear_simt(_is_bf16(a), *args, _stream()), "vc_linear_simt")
    elif impl == "tc":
        assert a.dtype == STORE
        _check(lib.vc_linear_tc(*args, tile_n, _stream()), "vc_linear_tc")
    else:
        if a.dtype == torch.float32:
            assert out_f32, "exact mode writes fp32"
        prof = GEMM_PROFILE
        if prof is not None and a.dtype == STORE and not torch.cuda.is_current_stream_capturing():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _check(lib.vc_linear(1, *args, _stream()), "vc_linear")
            e1.record()
            nbytes = 2.0 * M * K + 2.0 * N * K + (4.0 if out_f32 else 2.0) * M * N + (4.0 * M * N if resid is not None else 0.0)
            tag = "linear%s%s N=%d K=%d" % ("+act" if act != ACT_NONE else "", "+residual (fp32 stream)" if resid is not None else "", N, K)
            prof.append((2.0 * M * N * K, e0, e1, nbytes, tag))
        else:
            _check(lib.vc_linear(_is_bf16(a), *args, _stream()), "vc_linear")
    return out


def _profiled(flops, call, nbytes=0.0, tag=""):
    """(flops, start event, end event, algorithmic bytes, shape tag) of an eager tensor-core GEMM launch -> GEMM_PROFILE."""
    prof = GEMM_PROFILE
    if prof is not None and not torch.cuda.is_current_stream_capturing():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        prof.append((flops, e0, e1, nbytes, tag))
    else:
        call()


def linear_ln_emit(a, w, bias, out, resid, xb, stats, M=None, resid_ln=None):
    """out (fp32) = a @ w^T + bias + resid, plus xb = bf16(out) and stats [M, ceil(N/256), 2] = partial (sum, sum of squares) per
    256-column tile of every output row (the producer half of the folded LayerNorm, vc_linear_ln_emit).
    resid_ln = (rstats, rst_tiles, gamma, beta, eps): ``resid`` holds the RAW rows of a previous emit and the residual added is
    their LayerNorm, evaluated on the fly (vc_linear_ln_emit_postln)."""
    lib = load_library()
    N, K = w.shape
    M = a.shape[0] if M is None else M
    assert a.dtype == w.dtype == xb.dtype == STORE and out.dtype == resid.dtype == stats.dtype == torch.float32
    assert stats.is_contiguous() and stats.numel() >= M * ((N + 255) // 256) * 2
    if resid_ln is not None:
        rstats, rst_tiles, g, b, eps = resid_ln
        assert rstats.dtype == g.dtype == b.dtype == torch.float32 and rstats.is_contiguous() and rstats.numel() >= M * rst_tiles * 2
        _profiled(2.0 * M * N * K, lambda: _check(lib.vc_linear_ln_emit_postln(
            _ptr(a), a.stride(0), _ptr(w), w.stride(0), _ptr(bias), _ptr(out), out.stride(0), _ptr(resid), resid.stride(0), _ptr(rstats),
            rst_tiles, _ptr(g), _ptr(b), float(eps), _ptr(xb), xb.stride(0), _ptr(stats), M, N, K, _stream()), "vc_linear_ln_emit_postln"),
            2.0 * M * K + 2.0 * N * K + 10.0 * M * N, "ln_emit (post-LN residual) N=%d K=%d" % (N, K))
        return out
    _profiled(2.0 * M * N * K, lambda: _check(lib.vc_linear_ln_emit(
        _ptr(a), a.stride(0), _ptr(w), w.stride(0), _ptr(bias), _ptr(out), out.stride(0), _ptr(resid), resid.stride(0), _ptr(xb),
        xb.stride(0), _ptr(stats), M, N, K, _stream()), "vc_linear_ln_emit"),
        2.0 * M * K + 2.0 * N * K + 10.0 * M * N, "ln_emit N=%d K=%d" % (N, K))      # fp32 residual in, fp32 row + bf16 copy out
    return out


def linear_ln_fold(xb, wf, bias_f, colsum, stats, st_tiles, eps, out, act=ACT_NONE, M=None, ldo=None):
    """out (bf16) = act(LayerNorm(x) @ w^T + b) from the raw bf16 row copy xb and the producer's statistics (vc_linear_ln_fold):
    wf = bf16(gamma o w), colsum = fp32 row sums of wf, bias_f = b + w beta."""
    lib = load_library()
    N, K = wf.shape
    M = xb.shape[0] if M is None else M
    ldo = out.stride(0) if ldo is None else ldo
    assert xb.dtype == wf.dtype == out.dtype == STORE and bias_f.dtype == colsum.dtype == stats.dtype == torch.float32
    _profiled(2.0 * M * N * K, lambda: _check(lib.vc_linear_ln_fold(
        _ptr(xb), xb.stride(0), _ptr(wf), wf.stride(0), _ptr(bias_f), _ptr(colsum), _ptr(stats), st_tiles, float(eps), _ptr(out), ldo,
        act, M, N, K, _stream()), "vc_linear_ln_fold"),
        2.0 * M * K + 2.0 * N * K + 2.0 * M * N, "ln_fold%s N=%d K=%d" % ("+act" if act != ACT_NONE else "", N, K))
    return out


def patchify(image, out, patch):
    B, _, S, _ = image.shape
    assert image.dtype == torch.float32 and image.is_contiguous()
    _check(load_library().vc_patchify(_is_bf16(out), _ptr(image), _ptr(out), B, S, patch, _stream()), "vc_patchify")
    return out


def patchify_u8(image, out, patch, bgr=True):
    """image uint8 [B,S,S,3] (HWC) -> normalised patch matrix (ToTensor + Normalize(0.5,0.5) fused)."""
    B, S, S2, C = image.shape
    assert image.dtype == torch.uint8 and image.is_contiguous() and C == 3 and S == S2
    _check(load_library().vc_patchify_u8(_is_bf16(out), _ptr(image), _ptr(out), B, S, patch, int(bool(bgr)), _stream()), "vc_patchify_u8")
    return out


def resize_crop_plan(hw, resize_to, crop):
    """Host-side planning for resize_crop_u8. hw: int32 CPU tensor [B,2] (height, width).
    Returns (kmax, max_rows, tmp_off int64 CPU tensor [B+1])."""
    assert hw.dtype == torch.int32 and not hw.is_cuda and hw.is_contiguous() and hw.dim() == 2 and hw.shape[1] == 2
    B = hw.shape[0]
    kmax, max_rows = _c.c_int(0), _c.c_int(0)
    tmp_off = torch.empty(B + 1, dtype=torch.int64)
    _check(load_library().vc_resize_crop_plan(hw.data_ptr(), B, resize_to, crop, _c.addressof(kmax), tmp_off.data_ptr(),
                                              _c.addressof(max_rows)), "vc_resize_crop_plan")
    return kmax.value, max_rows.value, tmp_off


def resize_crop_u8(src, src_off, hw, resize_to, crop, kmax, max_rows, coef, tmp, tmp_off, out):
    """Ragged batch of uint8 HWC images (packed in ``src``) -> uint8 [B,crop,crop,3]: Resize(resize_to, BICUBIC) +
    CenterCrop(crop), bit-identical to torchvision + Pillow. All tensors on the device; see include/vitcap_b200.h."""
    B = hw.shape[0]
    assert src.dtype == torch.uint8 and out.dtype == torch.uint8 and out.is_contiguous() and tuple(out.shape) == (B, crop, crop, 3)
    assert src_off.dtype == torch.int64 and tmp_off.dtype == torch.int64 and hw.dtype == torch.int32 and coef.dtype == torch.int32
    assert coef.numel() >= B * 2 * (kmax + 2) * crop
    _check(load_library().vc_resize_crop_u8(_ptr(src), _ptr(src_off), _ptr(hw), B, resize_to, crop, kmax, max_rows, _ptr(coef),
                                            _ptr(tmp), _ptr(tmp_off), _ptr(out), _stream()), "vc_resize_crop_u8")
    return out


def assemble_tokens(patch_out, cls, pos, x, B, P, H):
    _check(load_library().vc_assemble_tokens(_ptr(patch_out), _ptr(cls), _ptr(pos), _ptr(x), B, P, H, _stream()),
           "vc_assemble_tokens")
    return x


def layernorm(x, gamma, beta, eps, out_t=None, out_f=None, rows=None, x3=False):
    """x3: out_t (bf16, >= 3H columns) receives the split operand [hi | lo | hi] (include/vitcap_b200.h, VC_OPERAND_BF16X3)."""
    rows = x.shape[0] if rows is None else rows
    H = x.shape[-1]
    bf = _is_bf16(out_t) if out_t is not None else 0
    if x3:
        assert bf and out_t.shape[-1] >= 3 * H
        bf = 2
    _check(load_library().vc_layernorm(bf, _ptr(x), x.stride(0), _ptr(gamma), _ptr(beta), float(eps), _ptr(out_t),
                                       out_t.stride(0) if out_t is not None else 0, _ptr(out_f),
                                       out_f.stride(0) if out_f is not None else 0, rows, H, _stream()), "vc_layernorm")


def split_bf16x3(x, out, rows=None):
    """fp32 [rows, K] -> bf16 [rows, 3K] = [hi | lo | hi]."""
    rows = x.shape[0] if rows is None else rows
    K = x.shape[-1]
    assert x.dtype == torch.float32 and out.dtype == torch.bfloat16 and out.shape[-1] >= 3 * K
    _check(load_library().vc_split_bf16x3(_ptr(x), x.stride(0), _ptr(out), out.stride(0), rows, K, _stream()), "vc_split_bf16x3")


def linear_x3(a3, w3, bias, out, resid, M=None):
    """out(fp32) = a w^T + bias + resid on split operands a3 = [hi | lo | hi], w3 = [w_hi | w_hi | w_lo]: same result as
    linear(a3, w3, ...) up to fp32 summation order, each distinct tile loaded once (vc_linear_x3)."""
    M = a3.shape[0] if M is None else M
    N, K3 = w3.shape
    assert a3.dtype == torch.bfloat16 and w3.dtype == torch.bfloat16 and out.dtype == torch.float32 and resid.dtype == torch.float32
    assert a3.shape[-1] == K3 and a3.stride(-1) == 1 and w3.stride(-1) == 1 and out.stride(-1) == 1 and resid.stride(-1) == 1
    _check(load_library().vc_linear_x3(_ptr(a3), a3.stride(0), _ptr(w3), w3.stride(0), _ptr(bias), _ptr(out), out.stride(0),
                                       _ptr(resid), resid.stride(0), M, N, K3, _stream()), "vc_linear_x3")
    return out


def split_weight_bf16x3(w32):
    """Host-side (torch) layout of a Linear weight for the three-product GEMM: [w_hi | w_hi | w_lo] along K."""
    hi = w32.to(torch.bfloat16)
    lo = (w32.float() - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, hi, lo], dim=1).contiguous()


def gather_rows(x, row_stride, out, rows, H):
    _check(load_library().vc_gather_rows(_is_bf16(out), _ptr(x), row_stride, _ptr(out), out.stride(0), rows, H, _stream()),
           "vc_gather_rows")
    return out


def assemble_ctx(cap, tag, ctx_f, ctx_t, B, N, H, rows_per_image=None):
    if rows_per_image is None or rows_per_image == N + 1:
        _check(load_library().vc_assemble_ctx(_is_bf16(ctx_t), _ptr(cap), _ptr(tag), _ptr(ctx_f), _ptr(ctx_t), B, N, H, _stream()),
               "vc_assemble_ctx")
    else:
        _check(load_library().vc_assemble_ctx_pitched(_is_bf16(ctx_t), _ptr(cap), _ptr(tag), _ptr(ctx_f), _ptr(ctx_t), B, N, H,
                                                      rows_per_image, _stream()), "vc_assemble_ctx_pitched")


def label_rows(tag_idx, sep_id, recipe_ln, pos0, word, pos, type0, gamma, beta, eps, ctx_f, ctx_t, B, rows_per_image, row0):
    """Label rows of the context (see vc_label_rows): tag_idx int32 [B,K]; recipe_ln False = raw word embedding,
    True = word + position(pos0 + i) + type 0 -> LayerNorm."""
    K, H = tag_idx.shape[1], word.shape[1]
    assert tag_idx.dtype == torch.int32 and tag_idx.is_contiguous()
    _check(load_library().vc_label_rows(_is_bf16(ctx_t), _ptr(tag_idx), K, int(sep_id), int(bool(recipe_ln)), int(pos0), _ptr(word),
                                        _ptr(pos), _ptr(type0), _ptr(gamma), _ptr(beta), float(eps), _ptr(ctx_f), _ptr(ctx_t), B,
                                        rows_per_image, row0, H, _stream()), "vc_label_rows")


def attention(qkv, out, B, N, heads, scale, impl="auto", n_base=0, n_extra=None):
    """n_extra (int32 [B]) selects the label-region mask: rows < n_base see keys < n_base, later rows n_extra[b] more."""
    lib = load_library()
    if n_extra is not None:
        assert n_extra.dtype == torch.int32 and n_extra.numel() >= B
        fn = lib.vc_attention_labels_simt if impl == "simt" else lib.vc_attention_labels
        _check(fn(_is_bf16(qkv), _ptr(qkv), _ptr(out), B, N, heads, float(scale), int(n_base), _ptr(n_extra), _stream()),
               "vc_attention_labels")
        return out
    if impl == "simt":
        _check(lib.vc_attention_simt(_is_bf16(qkv), _ptr(qkv), _ptr(out), B, N, heads, float(scale), _stream()), "vc_attention_simt")
    else:
        _check(lib.vc_attention(_is_bf16(qkv), _ptr(qkv), _ptr(out), B, N, heads, float(scale), _stream()), "vc_attention")
    return out


def cls_attention(q, qkv, out, B, N, heads, scale):
    """out[b] = softmax(q[b] K_b^T * scale) V_b: one query row per image against the packed qkv buffer."""
    _check(load_library().vc_cls_attention(_is_bf16(qkv), _ptr(q), q.stride(0), _ptr(qkv), _ptr(out), out.stride(0), B, N, heads,
                                           float(scale), _stream()), "vc_cls_attention")
    return out


def tag_topk(logits, V, K, thresh, out_idx, out_prob, out_len, rows=None):
    rows = logits.shape[0] if rows is None else rows
    _check(load_library().vc_tag_topk(_ptr(logits), logits.stride(0), rows, V, K, float(thresh), _ptr(out_idx), _ptr(out_prob),
                                      _ptr(out_len), _stream()), "vc_tag_topk")


def embed_ln(ids, cur_len, mask_id, word, pos, type0, gamma, beta, eps, out_f, out_t, R):
    H = word.shape[1]
    _check(load_library().vc_embed_ln(_is_bf16(out_t), _ptr(ids), ids.shape[1], cur_len, mask_id, _ptr(word), _ptr(pos),
                                      _ptr(type0), _ptr(gamma), _ptr(beta), float(eps), _ptr(out_f), _ptr(out_t), R, H,
                                      _stream()), "vc_embed_ln")


def decode_attention(ctx_qkv, step_qkv, anc, out, B, C, heads, E, cur_len, scale, impl="auto", ctx_vis=None, seq_unfinished=None,
                     img_done=None):
    """ctx_vis (int32 [B]): C context rows are allocated per image, only the first ctx_vis[b] are visible.
    seq_unfinished (int32 [B*E]) / img_done (int32 [B]): bf16 only -- finished sequences / done images are skipped and their
    output rows left untouched (vc_decode_attention_skip)."""
    if (seq_unfinished is not None or img_done is not None) and ctx_qkv.dtype == STORE and impl != "simt":
        for t in (seq_unfinished, img_done, ctx_vis):
            assert t is None or t.dtype == torch.int32
        _check(load_library().vc_decode_attention_skip(_ptr(ctx_qkv), _ptr(step_qkv), _ptr(anc), _ptr(out), B, C, _ptr(ctx_vis), heads, E,
                                                       cur_len, float(scale), _ptr(seq_unfinished), _ptr(img_done), _stream()),
               "vc_decode_attention_skip")
        return
    if ctx_vis is not None:
        assert ctx_vis.dtype == torch.int32 and ctx_vis.numel() >= B
        fn = load_library().vc_decode_attention_labels_simt if impl == "simt" else load_library().vc_decode_attention_labels
        _check(fn(_is_bf16(ctx_qkv), _ptr(ctx_qkv), _ptr(step_qkv), _ptr(anc), _ptr(out), B, C, _ptr(ctx_vis), heads, E, cur_len,
                  float(scale), _stream()), "vc_decode_attention_labels")
        return
    fn = load_library().vc_decode_attention_simt if impl == "simt" else load_library().vc_decode_attention
    _check(fn(_is_bf16(ctx_qkv), _ptr(ctx_qkv), _ptr(step_qkv), _ptr(anc), _ptr(out), B, C, heads, E, cur_len, float(scale),
              _stream()), "vc_decode_attention")


def token_step(logits, V, rows, do_sample, temperature, seed, cur_len, pad_id, eos_ids, ids, unfinished, sum_lp, n_steps,
               seed_dev=None):
    """seed_dev: optional int64 CUDA tensor [1] holding the seed (read at run time; see include/vitcap_b200.h)."""
    assert seed_dev is None or (seed_dev.dtype == torch.int64 and seed_dev.numel() == 1)
    _check(load_library().vc_token_step(_ptr(logits), logits.stride(0), rows, V, int(do_sample), float(temperature),
                                        int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(seed_dev), cur_len, ids.shape[1], pad_id, _ptr(eos_ids),
                                        eos_ids.numel(), _ptr(ids), _ptr(unfinished), _ptr(sum_lp), _ptr(n_steps), _stream()),
           "vc_token_step")


def greedy_finalize(ids, unfinished, sum_lp, n_steps, eos0, R, out_ids, out_lp):
    _check(load_library().vc_greedy_finalize(_ptr(ids), _ptr(unfinished), _ptr(sum_lp), _ptr(n_steps), eos0, ids.shape[1], R,
                                             _ptr(out_ids), _ptr(out_lp), _stream()), "vc_greedy_finalize")


def beam_row_topk(logits, V, rows, K, cand_val, cand_idx, row_max, row_logsum):
    _check(load_library().vc_beam_row_topk(_ptr(logits), logits.stride(0), rows, V, K, _ptr(cand_val), _ptr(cand_idx),
                                           _ptr(row_max), _ptr(row_logsum), _stream()), "vc_beam_row_topk")


def beam_advance(st, cand_val, cand_idx, row_max, row_logsum, B, nb, V, cur_len, keep, length_penalty, pad_id, eos_ids):
    _check(load_library().vc_beam_advance(_ptr(st["ids"]), _ptr(st["beam_scores"]), _ptr(st["done"]), _ptr(st["anc"]),
                                          _ptr(st["hyp_score"]), _ptr(st["hyp_len"]), _ptr(st["hyp_ids"]), _ptr(st["hyp_count"]),
                                          _ptr(st["worst"]), _ptr(cand_val), _ptr(cand_idx), _ptr(row_max), _ptr(row_logsum),
                                          B, nb, V, cur_len, st["ids"].shape[1], keep, float(length_penalty), pad_id,
                                          _ptr(eos_ids), eos_ids.numel(), _stream()), "vc_beam_advance")


def beam_finalize(st, B, keep, pad_id, eos0, out_ids, out_lp):
    _check(load_library().vc_beam_finalize(_ptr(st["hyp_score"]), _ptr(st["hyp_len"]), _ptr(st["hyp_ids"]), _ptr(st["hyp_count"]),
                                           B, keep, st["ids"].shape[1], pad_id, eos0, _ptr(out_ids), _ptr(out_lp), _stream()),
           "vc_beam_finalize")


DEC_PARTIAL, DEC_BF16, DEC_GELU_BF16, DEC_GELU_SPLIT = 0, 1, 2, 3
VOCAB_TILE = 208          # columns per tile of vc_dec_vocab_argmax (two partials per tile)


def vocab_partials(N):
    return 2 * ((N + VOCAB_TILE - 1) // VOCAB_TILE)


DEC_FMT_BF16, DEC_FMT_BF16X3, DEC_FMT_F16 = 0, 1, 2


def _dec_fmt(a, w, x3):
    """Operand format of a decode-step GEMM from the operand dtype: torch.float16 tensors select the IEEE-half product. (In a
    VITCAP_STORE=fp16 process the library's plain format IS the half: format 0.)"""
    assert a.dtype == w.dtype and a.stride(-1) == 1 and w.stride(-1) == 1
    if HALF_STORE:
        assert a.dtype == torch.float16 and not x3, "half storage: one product on halves, no split operands"
        return DEC_FMT_BF16
    if a.dtype == torch.float16:
        assert not x3, "half operands run as one product"
        return DEC_FMT_F16
    assert a.dtype == torch.bfloat16
    return DEC_FMT_BF16X3 if x3 else DEC_FMT_BF16


def dec_linear(mode, a, w, bias, out, M=None, x3=False, splits=1, m_pad=0, lda=None, ldo=None):
    """Decode-step Linear on CTA-pair tiles (include/vitcap_b200.h, vc_dec_linear). a [M, K] / w [N, K] bf16 (K = 3 Kt split
    layouts with x3) or float16 (one product on IEEE halves). mode DEC_PARTIAL: out fp32 [splits, m_pad, N] partial planes, no
    bias -- or, with splits = 1 and a 2-d out, fp32 [M, N] = a w^T + bias (the materialised logits); DEC_BF16: out bf16 [M, N];
    DEC_GELU_BF16: out [M, N] in the operand dtype; DEC_GELU_SPLIT: out bf16 [M, >= 2N] = [hi | lo] of the GELU output."""
    N, K = w.shape
    M = a.shape[0] if M is None else M
    lda = a.stride(0) if lda is None else lda
    fmt = _dec_fmt(a, w, x3)
    assert out.stride(-1) == 1
    if mode == DEC_PARTIAL and out.dim() == 2:
        assert out.dtype == torch.float32 and splits == 1 and out.shape[0] >= M
        ldo = out.stride(0)
        m_pad = -(-M // 128) * 128
    elif mode == DEC_PARTIAL:
        assert out.dtype == torch.float32 and out.dim() == 3 and out.shape[0] >= splits and out.shape[1] == m_pad and out.is_contiguous()
        assert bias is None
        ldo = out.stride(1)
    else:
        assert out.dtype == (torch.float16 if (fmt == DEC_FMT_F16 and mode == DEC_GELU_BF16) else STORE)
        ldo = out.stride(0) if ldo is None else ldo
    _check(load_library().vc_dec_linear(mode, fmt, _ptr(a), lda, _ptr(w), w.stride(0), _ptr(bias), _ptr(out), ldo, M, N, K,
                                        splits, m_pad, _stream()), "vc_dec_linear")
    return out


def dec_vocab_argmax(a, w, bias, part, M=None, x3=False, lda=None):
    """part fp32 [M, n_part, 4], n_part = vocab_partials(N): per-tile (max, arg max, sum of exponentials) of a w^T + bias."""
    N, K = w.shape
    M = a.shape[0] if M is None else M
    lda = a.stride(0) if lda is None else lda
    n_part = vocab_partials(N)
    fmt = _dec_fmt(a, w, x3)
    assert part.dtype == torch.float32 and part.is_contiguous()
    assert part.shape[-1] == 4 and part.shape[-2] == n_part and part.shape[0] >= M
    _check(load_library().vc_dec_vocab_argmax(fmt, _ptr(a), lda, _ptr(w), w.stride(0), _ptr(bias), _ptr(part), n_part, M, N,
                                              K, _stream()), "vc_dec_vocab_argmax")
    return part


def finish_ln(part, splits, bias, gamma, beta, eps, rows, resid=None, gelu=False, out_f=None, out_t=None, split=False):
    """LayerNorm(act(sum of `splits` partial planes + bias) + resid) -> out_f (fp32) and the operand copy out_t: bf16;
    split=True: the [hi | lo] pair; split=3: [hi | lo | hi]; split='f16': IEEE halves (out_t float16);
    split='bf16+f16': columns [0, H) bf16, [H, 2H) the bit patterns of the halves (out_t bf16, pitch >= 2H)."""
    assert part.dtype == torch.float32 and part.dim() == 3 and part.shape[0] >= splits
    H = part.shape[2]
    if out_t is None:
        mode = 0
    elif split == "f16":
        assert out_t.dtype == torch.float16
        mode = 4
    elif split == "bf16+f16":
        assert out_t.dtype == torch.bfloat16
        mode = 5
    else:
        assert out_t.dtype == STORE
        mode = (3 if split == 3 else 2) if split else 1
    _check(load_library().vc_finish_ln(_ptr(part), splits, part.stride(0), part.stride(1), _ptr(bias), int(bool(gelu)), _ptr(resid),
                                       resid.stride(0) if resid is not None else 0, _ptr(gamma), _ptr(beta), float(eps), _ptr(out_f),
                                       out_f.stride(0) if out_f is not None else 0, _ptr(out_t),
                                       out_t.stride(0) if out_t is not None else 0, mode, rows, H, _stream()), "vc_finish_ln")


def token_step_partials(part, rows, cur_len, pad_id, eos_ids, ids, unfinished, sum_lp, n_steps):
    n_part = part.shape[-2]
    _check(load_library().vc_token_step_partials(_ptr(part), n_part, rows, cur_len, ids.shape[1], pad_id, _ptr(eos_ids),
                                                 eos_ids.numel(), _ptr(ids), _ptr(unfinished), _ptr(sum_lp), _ptr(n_steps),
                                                 _stream()), "vc_token_step_partials")


def filter_logits(logits, V, rows, inv_temperature, top_k, top_p, min_tokens_to_keep=1):
    _check(load_library().vc_filter_logits(_ptr(logits), logits.stride(0), rows, V, float(inv_temperature), int(top_k),
                                           float(top_p), int(min_tokens_to_keep), _stream()), "vc_filter_logits")
