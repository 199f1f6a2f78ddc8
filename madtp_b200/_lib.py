"""ctypes binding of libmadtp_b200.so (C ABI declared in include/madtp_b200.h).

The wrappers take torch CUDA tensors, pass raw device pointers + sizes + the current CUDA stream, and raise
RuntimeError on any non-zero status. There is deliberately no CPU or PyTorch fallback: if the shared library is
missing or a tensor is not on a CUDA device the call fails loudly (the reference raises Python exceptions /
asserts on bad input too, e.g. models/utils.py:28, models/med.py:443).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent / "libmadtp_b200.so"
_lib = None

ABI_VERSION = 4                                # include/madtp_b200.h MADTP_B200_ABI_VERSION
GEMM_F16, GEMM_TF32X3, GEMM_SIMT, GEMM_F16X3 = 0, 1, 2, 3
QK_PLANE_SCALE, V_PLANE_SCALE = 8.0, 16.0     # include/madtp_b200.h MADTP_QK_PLANE_SCALE / MADTP_V_PLANE_SCALE
ACT_NONE, ACT_GELU, ACT_RELU, ACT_QUICKGELU, ACT_GELU_FAST = 0, 1, 2, 3, 4
# erf GELU of the FFNs: MADTP_ACT_GELU_FAST (one tanh.approx; error below the fp16 rounding of its own output) unless
# MADTP_EXACT_GELU=1 asks for the 1.5e-7-accurate erfc form
FFN_GELU = ACT_GELU if os.environ.get("MADTP_EXACT_GELU") else ACT_GELU_FAST

_i64, _i32, _f32, _vp = C.c_int64, C.c_int, C.c_float, C.c_void_p

# name -> argtypes; must list every function include/madtp_b200.h declares (tests/test_abi.py checks this).
SIGNATURES = {
    "madtp_abi_version": [],
    "madtp_last_error_string": [],
    "madtp_launch_count": [],
    "madtp_gemm": [_i32, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _i64, _i32, _f32, _i32, _i32,
                   _i32, _vp, _i32, _vp, _i32, _vp],
    "madtp_layernorm": [_vp, _i64, _i32, _i32, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp],
    "madtp_layernorm_pack": [_vp, _i32, _i32, _i32, _vp, _vp, _f32, _vp, _vp, _i32, _i64, _vp, _vp, _vp],
    "madtp_take_token": [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp],
    "madtp_split_tf32": [_vp, _vp, _vp, _i64, _vp],
    "madtp_attn_cross_tc": [_vp, _i64, _vp, _i64, _i32, _vp, _i64, _i32, _vp, _i32, _i32, _i32, _i32, _f32, _vp, _vp, _i64,
                            _i64, _vp, _vp, _vp, _vp, _vp, _vp],
    "madtp_readback_begin": [_vp, _vp, _i64, _i32, _vp],
    "madtp_readback_wait": [_i32],
    "madtp_lm_nll": [_vp, _i64, _i32, _i32, _vp, _f32, _vp, _vp, _vp],
    "madtp_split_f16": [_vp, _vp, _vp, _i64, C.c_float, _vp],
    "madtp_cast_f16": [_vp, _vp, _i64, _vp],
    "madtp_patchify": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "madtp_assemble_tokens": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp],
    "madtp_bert_embed": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "madtp_attn_fwd": [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _f32, _vp, _vp,
                       _i64, _i64, _vp, _vp, _vp, _i32, _vp],
    "madtp_attn_stats": [_vp, _i64, _i64, _vp, _i64, _i64, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _i32,
                         _vp],
    "madtp_attn_small_self": [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _i32, _i32, _i32, _f32, _vp, _vp, _i64,
                              _i64, _vp, _vp, _vp, _i32, _vp, _vp],
    "madtp_token_colstats": [_vp, _i64, _i64, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _i32, _vp],
    "madtp_query_sdft": [_vp, _i64, _i64, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _f32, _vp, _i32, _vp, _i32,
                         _vp],
    "madtp_query_sdft_tc": [_vp, _i64, _i64, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _i32, _vp,
                            _vp],
    "madtp_query_sdft_planes": [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _f32, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _f32,
                                _vp, _i32, _vp, _vp],
    "madtp_dtp_score": [_i32, _i32, _i32, _vp, _i32, _vp, _vp, _i64, _i64, _f32, _vp, _vp, _vp, _vp, _vp, _i32, _vp],
    "madtp_dtp_select": [_i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp],
    "madtp_dtp_gather": [_i32, _i32, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _vp, _vp],
    "madtp_dtp_apply": [_i32, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _f32, _vp, _vp, _i32, _vp, _vp,
                        _i32, _vp, _vp, _vp, _vp],
    "madtp_gather_rows": [_vp, _i64, _vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "madtp_gemm_qkv": [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _f32, _i32, _i32, _i32, _i32, _vp, _vp, _i64, _vp, _vp, _i64,
                       _vp, _vp],
    "madtp_attn_tc_fwd": [_vp, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _f32, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp,
                          _i32, _vp, _vp, _vp],
    "madtp_attn_tc_stats": [_vp, _vp, _i64, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _i32, _vp,
                            _vp],
}


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Load the shared library (building nothing: run `python -m madtp_b200.csrc.build` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -m madtp_b200.csrc.build`). madtp_b200 has no CPU fallback.")
    lib = C.CDLL(os.fspath(_LIB_PATH))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.madtp_last_error_string.restype = C.c_char_p
    lib.madtp_launch_count.restype = C.c_longlong
    if int(lib.madtp_abi_version()) != ABI_VERSION:
        raise RuntimeError(f"{_LIB_PATH} has ABI version {int(lib.madtp_abi_version())}, this binding needs {ABI_VERSION}: "
                           "rebuild with `python -m madtp_b200.csrc.build --force`")
    _lib = lib
    _bind_fastcall(lib)
    return lib


_fast = {}   # entry point name -> (trampoline, function address); empty when the _fastcall extension is unavailable
_float_pos = {}   # entry point name -> positions of `float` parameters (the trampoline picks registers by Python type)


def _bind_fastcall(lib):
    """Route launches through madtp_b200/_fastcall.so (csrc/fastcall.c) when it is built: ~1 us per call instead of
    ~5 us of ctypes argument conversion. Same C ABI, same library; MADTP_NO_FASTCALL=1 keeps ctypes."""
    if os.environ.get("MADTP_NO_FASTCALL") or not (_LIB_PATH.parent / "_fastcall.so").exists():
        return
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("madtp_b200._fastcall", _LIB_PATH.parent / "_fastcall.so")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    except Exception:
        return
    for name, argtypes in SIGNATURES.items():
        _fast[name] = (mod.call, C.cast(getattr(lib, name), C.c_void_p).value)
        _float_pos[name] = tuple(i for i, t in enumerate(argtypes) if t is _f32)


def launch_count() -> int:
    return int(load().madtp_launch_count())


class LaunchTimer:
    """Optional per-entry-point device timing with CUDA events on the launching stream (bench.py's roofline leg).
    `only` restricts the instrumentation to the named entry points so that a timed region is not perturbed."""

    def __init__(self, only=None):
        self.only = set(only) if only else None
        self.records = []          # (name, start_event, end_event, shape arguments)

    def summary(self):
        out = {}
        for name, e0, e1, meta in self.records:
            d = out.setdefault(name, {"launches": 0, "ms": 0.0, "meta": []})
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["meta"].append(meta)
        return out


_timer = None
# positions of the shape arguments recorded with each timed launch
_META_ARGS = {"madtp_gemm": (0, 15, 16, 17), "madtp_gemm_qkv": (8, 9, 11), "madtp_attn_tc_fwd": (6, 7, 8),
              "madtp_attn_tc_stats": (3, 4, 5), "madtp_attn_fwd": (9, 10, 11, 12), "madtp_attn_stats": (6, 7, 8),
              "madtp_layernorm": (2, 3), "madtp_dtp_gather": (0, 1, 2), "madtp_dtp_score": (0, 1, 2),
              "madtp_dtp_select": (0, 1)}


def set_launch_timer(timer):
    global _timer
    _timer = timer


def _call(name, *args):
    """Invoke one C-ABI entry point (optionally bracketed by CUDA events on the current stream)."""
    lib = _lib or load()
    fast = _fast.get(name)
    if fast is not None:
        tramp, addr = fast
        # The trampoline chooses an integer or a float register from each argument's PYTHON type, so the arguments are
        # brought in line with the C prototype here (what ctypes' argtypes would do): float parameters become float,
        # everything else must be an int or None -- a float in an integer slot would shift every later argument.
        fp = _float_pos[name]
        if len(args) != len(SIGNATURES[name]):
            raise TypeError(f"{name}: expected {len(SIGNATURES[name])} arguments, got {len(args)}")
        if fp or any(type(a) is float for a in args):
            args = list(args)
            for i in fp:
                args[i] = float(args[i])
            for i, a in enumerate(args):
                if type(a) is float and i not in fp:
                    raise TypeError(f"{name}: argument {i} is a float but the C prototype takes an integer/pointer")

        def fn(*a):
            return tramp(addr, *a)
    else:
        fn = getattr(lib, name)
    t = _timer
    if t is None:
        return fn(*args)
    key = name if name != "madtp_gemm" else "madtp_gemm:" + ("f16", "tf32x3", "simt", "f16x3")[args[0]]
    if t.only is not None and key not in t.only:
        return fn(*args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st = fn(*args)
    e1.record()
    t.records.append((key, e0, e1, tuple(args[i] for i in _META_ARGS.get(name, ()))))
    return st


def _check(status: int, what: str):
    if status != 0:
        msg = load().madtp_last_error_string().decode(errors="replace")
        raise RuntimeError(f"{what} failed (status {status}): {msg}")


def _ptr(t, dtype=None, name="tensor"):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"madtp_b200: {name} must be a CUDA tensor (no CPU fallback exists)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"madtp_b200: {name} must have dtype {dtype}, got {t.dtype}")
    if _raw_device is not None and t.device.index != _raw_device():
        # every launch goes to the CURRENT device's stream (one process per GPU): a tensor elsewhere would be an
        # illegal address on the device, so fail here instead
        raise RuntimeError(f"madtp_b200: {name} lives on cuda:{t.device.index} but the current device is "
                           f"cuda:{_raw_device()} (use torch.cuda.set_device / torch.cuda.device)")
    return t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream():
    """Raw cudaStream_t of torch's current stream on the current device (the fast C accessors: the Python
    torch.cuda.current_stream() wrapper costs ~15 us per call, more than a small kernel)."""
    if _raw_stream is not None and _raw_device is not None:
        return _raw_stream(_raw_device())
    return torch.cuda.current_stream().cuda_stream


def _rowmajor(t, name):
    if t.dim() != 2 or t.stride(1) != 1:
        raise RuntimeError(f"madtp_b200: {name} must be a 2-D tensor with unit inner stride")
    return t.stride(0)


# ------------------------------------------------------------------------------------------------------------------
# buffers: plain torch allocations, or an Arena of persistent buffers for the device-resident-length (graph) mode
# ------------------------------------------------------------------------------------------------------------------
class Arena:
    """Persistent buffers handed out in CALL ORDER. The launch sequence of a forward with device-resident lengths does
    not depend on the data, so the i-th allocation of every pass is the same buffer: the first pass creates it
    (zero-initialised), later passes -- the CUDA-graph capture among them -- get the same memory back. Three things
    follow: no allocator call is left on the hot path; every kernel argument is stable across passes (what a CUDA
    graph needs); and the region of a capacity-sized buffer BEYOND the dynamic length only ever holds zeros or finite
    values written by an earlier pass through the same role -- which is what makes it safe for the tensor cores to
    read it (P = 0 times a stale V is 0, never NaN)."""

    def __init__(self):
        self.bufs = []
        self.i = 0
        self.frozen = False

    def begin(self):
        self.i = 0

    def take(self, shape, dtype, device):
        shape = tuple(int(x) for x in shape)
        if self.i < len(self.bufs):
            t = self.bufs[self.i]
            if tuple(t.shape) != shape or t.dtype != dtype:
                raise RuntimeError(f"madtp_b200.Arena: allocation {self.i} changed from {tuple(t.shape)} {t.dtype} to "
                                   f"{shape} {dtype}: the launch sequence must not depend on the data")
        else:
            if self.frozen:
                raise RuntimeError("madtp_b200.Arena: new allocation after the arena was frozen")
            t = torch.zeros(shape, dtype=dtype, device=device)
            self.bufs.append(t)
        self.i += 1
        return t

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.bufs)


_arena = None


def set_arena(arena):
    """Route every buffer the wrappers below (and functional.py) allocate through `arena` (None: plain torch)."""
    global _arena
    prev = _arena
    _arena = arena
    return prev


# functional.value_lane_split(): part of every arena / graph key, because the mode changes the launch sequence
_value_split = [os.environ.get("MADTP_VALUE_LANE", "") == "split"]


def mode_key():
    return ("value_split",) if _value_split[0] else ()


class arena_for:
    """`with arena_for(owner, key):` -- run a forward with device-resident lengths inside the persistent Arena that
    `owner` (a module) keeps for `key` (the input shapes). Such a forward NEEDS one: its buffers are capacity-sized and
    the tensor cores read past the dynamic length, so that region must hold zeros or stale finite values, never the
    arbitrary bits of a fresh allocation. Nested scopes reuse the outer arena. Results that must survive the owner's
    next call with the same key have to be cloned by the caller."""

    def __init__(self, owner, key):
        self.owner, self.key, self.own, self.prev = owner, (key, mode_key()), False, None

    def __enter__(self):
        if _arena is not None:
            return _arena
        arenas = self.owner.__dict__.setdefault("_madtp_arenas", {})
        a = arenas.get(self.key)
        if a is None:
            a = arenas[self.key] = Arena()
        a.begin()
        self.prev = set_arena(a)
        self.own = True
        return a

    def __exit__(self, *exc):
        if self.own:
            set_arena(self.prev)
        return False


def release_arenas(module) -> int:
    """Frees the persistent buffers every sub-module of `module` keeps for its device-resident-length forwards."""
    n = 0
    for m in module.modules():
        n += len(m.__dict__.pop("_madtp_arenas", {}))
    return n


def empty(shape, dtype, device):
    if _arena is not None:
        return _arena.take(shape, dtype, device)
    return torch.empty(shape, dtype=dtype, device=device)


def zeros(shape, dtype, device):
    if _arena is not None:
        return _arena.take(shape, dtype, device).zero_()
    return torch.zeros(shape, dtype=dtype, device=device)


def _dyn(t):
    """Device pointer of an int32 device scalar holding a dynamic token count (None: the host value is exact)."""
    if t is None:
        return None
    if t.dtype != torch.int32 or not t.is_cuda or t.numel() < 1:
        raise RuntimeError("madtp_b200: a dynamic length must be an int32 CUDA tensor")
    return t.data_ptr()


# ------------------------------------------------------------------------------------------------------------------
# thin functional wrappers
# ------------------------------------------------------------------------------------------------------------------
def gemm(precision, a, b, out, *, a_lo=None, b_lo=None, bias=None, residual=None, act=ACT_NONE, alpha=1.0,
         m_dev=None, m_mult=1, n_dev=None, n_mult=1):
    """out[M,N] = act(alpha * a[M,K] @ b[N,K]^T + bias) + residual. out may be a column slice of a wider buffer.
    m_dev / n_dev: device-resident extents M = *m_dev * m_mult, N = *n_dev * n_mult (the tensor shapes are capacities)."""
    lda, ldb, ldc = _rowmajor(a, "a"), _rowmajor(b, "b"), _rowmajor(out, "out")
    M, K = a.shape
    N = b.shape[0]
    if b.shape[1] != K or out.shape[0] != M or out.shape[1] != N:
        raise RuntimeError(f"madtp_b200.gemm: shape mismatch a{tuple(a.shape)} b{tuple(b.shape)} out{tuple(out.shape)}")
    op_dtype = torch.float16 if precision in (GEMM_F16, GEMM_F16X3) else torch.float32
    lo_dtype = torch.float16 if precision == GEMM_F16X3 else torch.float32
    if out.dtype not in (torch.float32, torch.float16):
        raise RuntimeError("madtp_b200.gemm: out must be fp32 or fp16")
    ldr = 0
    if residual is not None:
        ldr = _rowmajor(residual, "residual")
        if residual.shape != out.shape:
            raise RuntimeError("madtp_b200.gemm: residual shape mismatch")
    if a_lo is not None and (a_lo.shape != a.shape or a_lo.stride(0) != lda):
        raise RuntimeError("madtp_b200.gemm: a_lo layout must match a")
    if b_lo is not None and (b_lo.shape != b.shape or b_lo.stride(0) != ldb):
        raise RuntimeError("madtp_b200.gemm: b_lo layout must match b")
    st = _call("madtp_gemm", precision, _ptr(a, op_dtype, "a"), _ptr(a_lo, lo_dtype, "a_lo"), lda,
                           _ptr(b, op_dtype, "b"), _ptr(b_lo, lo_dtype, "b_lo"), ldb, _ptr(out, None, "out"),
                           ldc, 1 if out.dtype == torch.float16 else 0, _ptr(bias, torch.float32, "bias"),
                           _ptr(residual, torch.float32, "residual"), ldr, act, float(alpha), M, N, K, _dyn(m_dev),
                           int(m_mult), _dyn(n_dev), int(n_mult), _stream())
    _check(st, "madtp_gemm")
    return out


def layernorm(x, gamma, beta, eps, *, y_f32=None, y_hi=None, y_lo=None, y_f16=None, x_hi=None, x_lo=None, n_dev=None,
              n_mult=1):
    """x: [rows, d] fp32 (row stride free). Every output is optional and contiguous [rows, d]."""
    ldx = _rowmajor(x, "x")
    rows, d = x.shape
    for nm, t, dt in (("y_f32", y_f32, torch.float32), ("y_hi", y_hi, torch.float16), ("y_lo", y_lo, torch.float16),
                      ("y_f16", y_f16, torch.float16), ("x_hi", x_hi, torch.float16), ("x_lo", x_lo, torch.float16)):
        if t is not None and (t.dtype != dt or not t.is_contiguous() or t.numel() != rows * d):
            raise RuntimeError(f"madtp_b200.layernorm: bad output {nm}")
    st = _call("madtp_layernorm", _ptr(x, torch.float32, "x"), ldx, rows, d, _ptr(gamma, torch.float32, "gamma"),
                                _ptr(beta, torch.float32, "beta"), float(eps), _ptr(y_f32), _ptr(y_hi), _ptr(y_lo),
                                _ptr(y_f16), _ptr(x_hi), _ptr(x_lo), _dyn(n_dev), int(n_mult), _stream())
    _check(st, "madtp_layernorm")


def layernorm_pack(x2d, B, N, gamma, beta, eps, y16, per_group, group_stride, *, y_f32=None, p_out=None, n_dev=None):
    """LayerNorm of the packed stream x2d [B * N, d] -> fp16 cross-attention operand layout (madtp_layernorm_pack)."""
    d = x2d.shape[1]
    if not x2d.is_contiguous():
        raise RuntimeError("madtp_b200.layernorm_pack: x must be dense")
    st = _call("madtp_layernorm_pack", _ptr(x2d, torch.float32, "x"), B, N, d, _ptr(gamma, torch.float32, "gamma"),
               _ptr(beta, torch.float32, "beta"), float(eps), _ptr(y_f32, torch.float32, "y_f32"),
               _ptr(y16, torch.float16, "y16"), int(per_group), int(group_stride), _dyn(p_out), _dyn(n_dev), _stream())
    _check(st, "madtp_layernorm_pack")


def take_token(x2d, B, N, token, *, n_dev=None):
    """x2d: packed stream [B * N, d] -> [B, d] = row `token` of every sequence."""
    d = x2d.shape[1]
    out = empty((B, d), torch.float32, x2d.device)
    _check(_call("madtp_take_token", _ptr(x2d, torch.float32, "x"), B, N, int(token), d, _ptr(out), _dyn(n_dev),
                 _stream()), "madtp_take_token")
    return out


def split_tf32(x):
    x = x.contiguous()
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    _check(_call("madtp_split_tf32", _ptr(x, torch.float32, "x"), _ptr(hi), _ptr(lo), x.numel(), _stream()),
           "madtp_split_tf32")
    return hi, lo


def split_f16(x, scale=1.0):
    """fp16 hi/lo planes of scale * x (scale: a power of two)."""
    x = x.contiguous()
    hi = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    lo = torch.empty_like(hi)
    _check(_call("madtp_split_f16", _ptr(x, torch.float32, "x"), _ptr(hi), _ptr(lo), x.numel(), float(scale), _stream()),
           "madtp_split_f16")
    return hi, lo


def cast_f16(x):
    x = x.contiguous()
    y = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    _check(_call("madtp_cast_f16", _ptr(x, torch.float32, "x"), _ptr(y), x.numel(), _stream()), "madtp_cast_f16")
    return y


def patchify(img, P):
    B, Cc, H, W = img.shape
    img = img.contiguous()
    rows = B * (H // P) * (W // P)
    hi = empty((rows, Cc * P * P), torch.float16, img.device)
    lo = empty((rows, Cc * P * P), torch.float16, img.device)
    _check(_call("madtp_patchify", _ptr(img, torch.float32, "img"), _ptr(hi), _ptr(lo), B, Cc, H, W, P, _stream()),
           "madtp_patchify")
    return hi, lo


def assemble_tokens(patches, cls, pos, B, n, d):
    x = empty((B, n + 1, d), torch.float32, patches.device)
    _check(_call("madtp_assemble_tokens", _ptr(patches, torch.float32, "patches"), _ptr(cls, torch.float32, "cls"),
                                        _ptr(pos, torch.float32, "pos"), _ptr(x), B, n, d, _stream()),
           "madtp_assemble_tokens")
    return x


def bert_embed(ids, word, position):
    B, L = ids.shape
    ids = ids.contiguous()
    d = word.shape[1]
    out = empty((B, L, d), torch.float32, word.device)
    _check(_call("madtp_bert_embed", _ptr(ids, torch.int64, "ids"), _ptr(word, torch.float32, "word"),
                                   _ptr(position, torch.float32, "position"), _ptr(out), B, L, d, word.shape[0],
                                   position.shape[0], _stream()), "madtp_bert_embed")
    return out


def _qkv_strides(t, name):
    """t: [B, N, H*64] view (row stride / batch stride arbitrary, unit inner stride)."""
    if t.dim() != 3 or t.stride(2) != 1:
        raise RuntimeError(f"madtp_b200: {name} must be [B, N, H*64] with unit inner stride")
    return t.stride(1), t.stride(0)


def attn_fwd(q, k, v, H, scale, out_f16, *, key_mask=None, stats=None, causal=False):
    """q [B,Nq,H*64], k/v [B,Nk,H*64] fp32 views; out_f16 [B,Nq,H*64] fp16 view. stats = (row_max,row_sum,out_norm)."""
    B, Nq, _ = q.shape
    Nk = k.shape[1]
    ldq, bsq = _qkv_strides(q, "q")
    ldk, bsk = _qkv_strides(k, "k")
    ldv, bsv = _qkv_strides(v, "v")
    ldo, bso = _qkv_strides(out_f16, "out_f16")
    rm = rs = on = None
    if stats is not None:
        rm, rs, on = stats
    if key_mask is not None and (not key_mask.is_contiguous() or key_mask.numel() != B * Nk):
        raise RuntimeError("madtp_b200.attn_fwd: key_mask must be contiguous [B, Nk]")
    st = _call("madtp_attn_fwd", _ptr(q, torch.float32, "q"), ldq, bsq, _ptr(k, torch.float32, "k"), ldk, bsk,
                               _ptr(v, torch.float32, "v"), ldv, bsv, B, H, Nq, Nk, float(scale),
                               _ptr(key_mask, torch.float32, "key_mask"), _ptr(out_f16, torch.float16, "out_f16"), ldo,
                               bso, _ptr(rm), _ptr(rs), _ptr(on), 1 if causal else 0, _stream())
    _check(st, "madtp_attn_fwd")


_rb_host = {}     # device index -> pinned int32 [8]: one slot per outstanding read-back
_rb_next = [0]


def readback_begin(src):
    """Start the asynchronous device->host copy of the int32 scalar `src` behind everything queued so far on the
    current stream; returns a slot for readback_wait. Kernels launched afterwards overlap the host's wait."""
    dev = src.device.index
    host = _rb_host.get(dev)
    if host is None:
        host = _rb_host[dev] = torch.zeros(8, dtype=torch.int32).pin_memory()
    slot = _rb_next[0]
    _rb_next[0] = (slot + 1) & 7
    _check(_call("madtp_readback_begin", _ptr(src, torch.int32, "src"), host.data_ptr() + 4 * slot, 4, slot, _stream()),
           "madtp_readback_begin")
    return dev, slot


def readback_wait(handle) -> int:
    dev, slot = handle
    _check(_call("madtp_readback_wait", slot), "madtp_readback_wait")
    return int(_rb_host[dev][slot])


def lm_nll(logits, labels=None, label_smoothing=0.0):
    """logits [R, V] fp32 (row stride free). Returns (loss [R] or None, lse [R]) -- see madtp_lm_nll."""
    ld = _rowmajor(logits, "logits")
    R, V = logits.shape
    lse = torch.empty(R, dtype=torch.float32, device=logits.device)
    loss = None
    if labels is not None:
        if labels.dtype != torch.int64 or not labels.is_contiguous() or labels.numel() != R:
            raise RuntimeError("madtp_b200.lm_nll: labels must be contiguous int64 [R]")
        loss = torch.empty(R, dtype=torch.float32, device=logits.device)
    _check(_call("madtp_lm_nll", _ptr(logits, torch.float32, "logits"), ld, R, V, _ptr(labels, torch.int64, "labels"),
                 float(label_smoothing), _ptr(loss), _ptr(lse), _stream()), "madtp_lm_nll")
    return loss, lse


def cross_tc_supported(Lq, Nk):
    return Lq <= 128


def attn_cross_tc_ragged(q16, k16, vt16, H, scale, out_f16, k_start, k_len, max_len, *, v_bias=None, key0_bias=None,
                         lq_dev=None):
    """Cross-attention over RAGGED keys: q16 [B, Lq, H*64] fp16 view; k16 [rows, H*64] fp16 and vt16 [H*64, >= rows] fp16
    hold the keys / values of all sequences packed back to back; sequence b owns k_len[b] of them from row k_start[b]
    (int32 device tensors [B], starts multiples of 8); max_len >= max(k_len). key0_bias [B] fp32: extra logit of key 0."""
    B, Lq, C = q16.shape
    if q16.stride(2) != 1 or (B > 1 and q16.stride(0) != Lq * q16.stride(1)):
        raise RuntimeError("madtp_b200.attn_cross_tc_ragged: q must be a [B*Lq, ld] row-major view")
    if k16.dim() != 2 or k16.stride(1) != 1 or vt16.dim() != 2 or vt16.stride(1) != 1 or vt16.shape[1] < k16.shape[0]:
        raise RuntimeError("madtp_b200.attn_cross_tc_ragged: k [rows, C] and vt [C, >= rows] with unit inner stride")
    for t in (k_start, k_len):
        if t.dtype != torch.int32 or t.numel() != B or not t.is_contiguous():
            raise RuntimeError("madtp_b200.attn_cross_tc_ragged: k_start / k_len must be contiguous int32 [B]")
    ldo, bso = _qkv_strides(out_f16, "out_f16")
    rows = k16.shape[0]
    st = _call("madtp_attn_cross_tc", _ptr(q16, torch.float16, "q"), q16.stride(1), _ptr(k16, torch.float16, "k"),
               k16.stride(0), rows, _ptr(vt16, torch.float16, "vt"), vt16.stride(0), rows,
               _ptr(v_bias, torch.float32, "v_bias"), B, H, Lq, int(max_len), float(scale), None,
               _ptr(out_f16, torch.float16, "out_f16"), ldo, bso, _dyn(lq_dev), None, _ptr(k_start, torch.int32),
               _ptr(k_len, torch.int32), _ptr(key0_bias, torch.float32, "key0_bias"), _stream())
    _check(st, "madtp_attn_cross_tc")


def attn_cross_tc(q16, k16, vt16, H, scale, out_f16, *, keys_per_batch=None, v_bias=None, key_mask=None, lq_dev=None,
                  nk_dev=None):
    """Tensor-core cross-attention (value lane). q16 [B,Lq,H*64] fp16 view of a row-major matrix (batch stride =
    Lq * row stride). keys_per_batch = P > 0 (a multiple of 8, default: Nk rounded up): k16 [B,Nk,H*64] fp16 view with
    batch stride P rows, vt16 [H*64, >= B*P] fp16 (V^T, keys of sequence b at columns b*P ..). keys_per_batch = 0:
    every sequence attends to the same keys, k16 [Nk,H*64], vt16 [H*64, >= Nk]. out_f16 [B,Lq,H*64] fp16 view."""
    B, Lq, C = q16.shape
    if q16.stride(2) != 1 or (B > 1 and q16.stride(0) != Lq * q16.stride(1)):
        raise RuntimeError("madtp_b200.attn_cross_tc: q must be a [B*Lq, ld] row-major view")
    if keys_per_batch == 0:
        Nk, ldk, per = k16.shape[0], k16.stride(0), 0
        need = Nk
    else:
        Nk, ldk = k16.shape[1], k16.stride(1)
        per = (Nk + 7) // 8 * 8 if keys_per_batch is None else int(keys_per_batch)
        if k16.stride(2) != 1 or (B > 1 and k16.stride(0) != per * ldk):
            raise RuntimeError("madtp_b200.attn_cross_tc: k must be a [B*keys_per_batch, ld] row-major view")
        need = (B - 1) * per + Nk
    if vt16.dim() != 2 or vt16.stride(1) != 1 or vt16.shape[0] != C or vt16.shape[1] < need:
        raise RuntimeError("madtp_b200.attn_cross_tc: vt must be [H*64, >= B*keys_per_batch] with unit inner stride")
    if key_mask is not None and (not key_mask.is_contiguous() or key_mask.numel() != B * Nk):
        raise RuntimeError("madtp_b200.attn_cross_tc: key_mask must be contiguous [B, Nk]")
    ldo, bso = _qkv_strides(out_f16, "out_f16")
    st = _call("madtp_attn_cross_tc", _ptr(q16, torch.float16, "q"), q16.stride(1), _ptr(k16, torch.float16, "k"), ldk,
               per, _ptr(vt16, torch.float16, "vt"), vt16.stride(0), per, _ptr(v_bias, torch.float32, "v_bias"), B, H, Lq,
               Nk, float(scale), _ptr(key_mask, torch.float32, "key_mask"), _ptr(out_f16, torch.float16, "out_f16"), ldo,
               bso, _dyn(lq_dev), _dyn(nk_dev), None, None, None, _stream())
    _check(st, "madtp_attn_cross_tc")


def attn_stats(q, k, H, scale, stats, col_part, cls_attn, *, key_mask=None, causal=False):
    B, N, _ = q.shape
    ldq, bsq = _qkv_strides(q, "q")
    ldk, bsk = _qkv_strides(k, "k")
    rm, rs, on = stats
    st = _call("madtp_attn_stats", _ptr(q, torch.float32, "q"), ldq, bsq, _ptr(k, torch.float32, "k"), ldk, bsk, B, H, N,
                                 float(scale), _ptr(key_mask, torch.float32, "key_mask"), _ptr(rm), _ptr(rs), _ptr(on),
                                 _ptr(col_part, torch.float32, "col_part"), _ptr(cls_attn, torch.float32, "cls_attn"),
                                 1 if causal else 0, _stream())
    _check(st, "madtp_attn_stats")


def token_colstats(token_att, n, T, divisor, *, n_dev=None, n_sub=0):
    """token_att: [B, rows>=n, ld] fp32 view whose row j is prunable token j. Returns (col_max, col_sum) [B, T].
    n_dev: device-resident tokens per sequence INCLUDING the n_sub leading tokens the view skips (packed sequences)."""
    B = token_att.shape[0]
    ld, bs = token_att.stride(1), token_att.stride(0)
    cm = empty((B, T), torch.float32, token_att.device)
    cs = empty((B, T), torch.float32, token_att.device)
    _check(_call("madtp_token_colstats", _ptr(token_att, torch.float32, "token_att"), ld, bs, B, n, T, float(divisor),
                                       _ptr(cm), _ptr(cs), _dyn(n_dev), int(n_sub), _stream()), "madtp_token_colstats")
    return cm, cs


def query_sdft(token_att, col_max, col_sum, ft, n, T, divisor, sd_ft, accumulate, *, n_dev=None, n_sub=0):
    B = token_att.shape[0]
    d = ft.shape[-1]
    st = _call("madtp_query_sdft", _ptr(token_att, torch.float32, "token_att"), token_att.stride(1), token_att.stride(0),
                                 _ptr(col_max), _ptr(col_sum), _ptr(ft, torch.float32, "ft"), ft.stride(1),
                                 ft.stride(0), B, n, T, d, float(divisor), _ptr(sd_ft, torch.float32, "sd_ft"),
                                 1 if accumulate else 0, _dyn(n_dev), int(n_sub), _stream())
    _check(st, "madtp_query_sdft")


def dtp_score(col_part, cls_attn, token_att, n, T, temperature, *, n_dev=None, parts_tile=0):
    """Returns (score [B,n], threshold [B], count [B] int32, topk [1] int32). n_dev: device-resident N = n + 1
    (packed buffers); parts_tile: query-tile height of the col_part producer (n_parts = ceil(N / parts_tile)), 0 = fixed."""
    B, n_parts, N = col_part.shape
    assert N == n + 1
    dev = col_part.device
    score = empty((B, n), torch.float32, dev)
    thr = empty((B,), torch.float32, dev)
    cnt = empty((B,), torch.int32, dev)
    topk = zeros((1,), torch.int32, dev)
    st = _call("madtp_dtp_score", B, n, T, _ptr(col_part, torch.float32, "col_part"), n_parts,
                                _ptr(cls_attn, torch.float32, "cls_attn"), _ptr(token_att, torch.float32, "token_att"),
                                token_att.stride(1), token_att.stride(0), float(temperature), _ptr(score), _ptr(thr),
                                _ptr(cnt), _ptr(topk), _dyn(n_dev), int(parts_tile), _stream())
    _check(st, "madtp_dtp_score")
    return score, thr, cnt, topk


def dtp_select(score, topk, *, mask_mode=0, mask_in=None, max_keep=0, n_dev=None, n_out=None, k_out=None):
    """n_dev / n_out / k_out: device-resident lengths -- *n_dev = n + 1 in, the next layer's length and the trajectory
    entry out (see madtp_dtp_select)."""
    B, n = score.shape
    dev = score.device
    keep = empty((B, n), torch.uint8, dev)
    dst = empty((B, n), torch.int32, dev)
    tail_w = empty((B, n), torch.float32, dev)
    tail_idx = empty((B, n), torch.int32, dev)
    mask_out = None
    if mask_mode:
        if mask_in is None or not mask_in.is_contiguous() or mask_in.numel() != B * (n + 1):
            raise RuntimeError("madtp_b200.dtp_select: mask_in must be contiguous [B, n+1]")
        mask_out = empty((B, n + 1), torch.float32, dev)
    st = _call("madtp_dtp_select", B, n, _ptr(score, torch.float32, "score"), _ptr(topk, torch.int32, "topk"), _ptr(keep),
                                 _ptr(dst), _ptr(tail_w), _ptr(tail_idx), mask_mode,
                                 _ptr(mask_in, torch.float32, "mask_in"), _ptr(mask_out), int(max_keep), _dyn(n_dev),
                                 _dyn(n_out), _dyn(k_out), _stream())
    _check(st, "madtp_dtp_select")
    return keep, dst, tail_w, tail_idx, mask_out


def dtp_gather(x, topk, dst, tail_w, tail_idx, k, max_keep=0, want_f16=False, n_dev=None):
    """x: [B, n+1, d] fp32 (unit inner stride, dense rows). Returns [B, k+2, d] (and its fp16 copy if want_f16).
    n_dev: device-resident N = n + 1; pass k = n - 1 so that the output has the capacity of the input."""
    B, N, d = x.shape
    if x.stride(2) != 1 or x.stride(1) != d:
        raise RuntimeError("madtp_b200.dtp_gather: x rows must be dense")
    out = empty((B, k + 2, d), torch.float32, x.device)
    out16 = empty((B, k + 2, d), torch.float16, x.device) if want_f16 else None
    st = _call("madtp_dtp_gather", B, N - 1, d, _ptr(x, torch.float32, "x"), x.stride(0), _ptr(topk, torch.int32, "topk"),
                                 _ptr(dst, torch.int32, "dst"), _ptr(tail_w, torch.float32, "tail_w"),
                                 _ptr(tail_idx, torch.int32, "tail_idx"), _ptr(out), out.stride(0), _ptr(out16),
                                 int(max_keep), _dyn(n_dev), _stream())
    _check(st, "madtp_dtp_gather")
    return (out, out16) if want_f16 else out


def dtp_apply(x, score, topk, k, *, mask_mode=0, mask_in=None, max_keep=0, want_f16=False, ln=None, n_dev=None,
              n_out=None, k_out=None):
    """Fused select + gather + merged token (+ LayerNorm): x [B, n+1, d] fp32 dense rows, score [B, n], topk device
    scalar. k: the host's copy of topk (output rows = k + 2), or n - 1 with device-resident lengths (capacity-sized
    output). ln = (gamma, beta, eps): also LayerNorm(out) as fp16. Returns (out, out16 | None, ln16 | None, keep,
    mask_out | None)."""
    B, N, d = x.shape
    n = N - 1
    if x.stride(2) != 1 or x.stride(1) != d:
        raise RuntimeError("madtp_b200.dtp_apply: x rows must be dense")
    dev = x.device
    out = empty((B, k + 2, d), torch.float32, dev)
    out16 = empty((B, k + 2, d), torch.float16, dev) if want_f16 else None
    ln16 = empty((B, k + 2, d), torch.float16, dev) if ln is not None else None
    keep = empty((B, n), torch.uint8, dev)
    mask_out = None
    if mask_mode:
        if mask_in is None or not mask_in.is_contiguous() or mask_in.numel() != B * (n + 1):
            raise RuntimeError("madtp_b200.dtp_apply: mask_in must be contiguous [B, n+1]")
        mask_out = empty((B, n + 1), torch.float32, dev)
    g, bt, eps = ln if ln is not None else (None, None, 0.0)
    st = _call("madtp_dtp_apply", B, n, d, _ptr(score, torch.float32, "score"), _ptr(topk, torch.int32, "topk"),
               _ptr(x, torch.float32, "x"), x.stride(0), _ptr(out), out.stride(0), _ptr(out16),
               _ptr(g, torch.float32, "ln_gamma"), _ptr(bt, torch.float32, "ln_beta"), float(eps), _ptr(ln16), _ptr(keep),
               int(mask_mode), _ptr(mask_in, torch.float32, "mask_in"), _ptr(mask_out), int(max_keep), _dyn(n_dev),
               _dyn(n_out), _dyn(k_out), _stream())
    _check(st, "madtp_dtp_apply")
    return out, out16, ln16, keep, mask_out


def gather_rows(x, idx):
    """x [B,L,d] fp32 (dense rows), idx [B,K] int32 -> [B,K,d]."""
    B, Ltok, d = x.shape
    K = idx.shape[1]
    if x.stride(2) != 1 or x.stride(1) != d:
        raise RuntimeError("madtp_b200.gather_rows: x rows must be dense")
    out = torch.empty(B, K, d, dtype=torch.float32, device=x.device)
    _check(_call("madtp_gather_rows", _ptr(x, torch.float32, "x"), x.stride(0), _ptr(idx, torch.int32, "idx"), _ptr(out),
                                    B, Ltok, K, d, _stream()), "madtp_gather_rows")
    return out



def gemm_qkv(a_hi, a_lo, w_hi, w_lo, bias, n_tok, heads, alpha=1.0, n_dev=None):
    """Fused q|k|v projection for the tensor-core attention: returns fp16 planes (qk_hi, qk_lo [M, 2*heads*64] of
    QK_PLANE_SCALE * value, vt_hi, vt_lo [B*heads*64, n_pad] of V_PLANE_SCALE * value) -- see madtp_gemm_qkv in
    include/madtp_b200.h."""
    M, K = a_hi.shape
    B = M // n_tok
    n_pad = (n_tok + 7) // 8 * 8
    dev = a_hi.device
    qk_hi = empty((M, 2 * heads * 64), torch.float16, dev)
    qk_lo = empty((M, 2 * heads * 64), torch.float16, dev)
    vt_hi = empty((B * heads * 64, n_pad), torch.float16, dev)
    vt_lo = empty((B * heads * 64, n_pad), torch.float16, dev)
    st = _call("madtp_gemm_qkv", _ptr(a_hi, torch.float16, "a_hi"), _ptr(a_lo, torch.float16, "a_lo"),
               _rowmajor(a_hi, "a_hi"), _ptr(w_hi, torch.float16, "w_hi"), _ptr(w_lo, torch.float16, "w_lo"),
               _rowmajor(w_hi, "w_hi"), _ptr(bias, torch.float32, "bias"), float(alpha), M, K, n_tok, heads, _ptr(qk_hi),
               _ptr(qk_lo), qk_hi.stride(0), _ptr(vt_hi), _ptr(vt_lo), n_pad, _dyn(n_dev), _stream())
    _check(st, "madtp_gemm_qkv")
    return qk_hi, qk_lo, vt_hi, vt_lo


def attn_tc_fwd(qk_hi, qk_lo, vt_hi, vt_lo, B, H, N, scale, out_f16, row_lse, out_norm, *, key_mask=None, cls_p=None,
                cls_tile_max=None, n_dev=None, causal=False, out_f32=None):
    """cls_p [B,H,N] / cls_tile_max [B,H,ceil(N/64)] fp32 (both or neither): the CLS query row for attn_tc_stats.
    out_f32: optional fp32 copy of the context, same shape / strides as out_f16."""
    ldo, bso = _qkv_strides(out_f16, "out_f16")
    if out_f32 is not None and _qkv_strides(out_f32, "out_f32") != (ldo, bso):
        raise RuntimeError("madtp_b200.attn_tc_fwd: out_f32 must have the strides of out_f16")
    st = _call("madtp_attn_tc_fwd", _ptr(qk_hi, torch.float16, "qk_hi"), _ptr(qk_lo, torch.float16, "qk_lo"),
               qk_hi.stride(0), _ptr(vt_hi, torch.float16, "vt_hi"), _ptr(vt_lo, torch.float16, "vt_lo"),
               vt_hi.stride(0), B, H, N, float(scale), _ptr(key_mask, torch.float32, "key_mask"),
               _ptr(out_f16, torch.float16, "out_f16"), ldo, bso, _ptr(row_lse, torch.float32, "row_lse"),
               _ptr(out_norm, torch.float32, "out_norm"), _ptr(cls_p, torch.float32, "cls_p"),
               _ptr(cls_tile_max, torch.float32, "cls_tile_max"), 1 if causal else 0, _dyn(n_dev),
               _ptr(out_f32, torch.float32, "out_f32"), _stream())
    _check(st, "madtp_attn_tc_fwd")


def attn_tc_stats(qk_hi, qk_lo, B, H, N, scale, row_lse, out_norm, col_part, cls_attn, cls_p, cls_tile_max, *,
                  key_mask=None, n_dev=None, causal=False):
    st = _call("madtp_attn_tc_stats", _ptr(qk_hi, torch.float16, "qk_hi"), _ptr(qk_lo, torch.float16, "qk_lo"),
               qk_hi.stride(0), B, H, N, float(scale), _ptr(key_mask, torch.float32, "key_mask"), _ptr(row_lse),
               _ptr(out_norm), _ptr(col_part, torch.float32, "col_part"), col_part.shape[1],
               _ptr(cls_attn, torch.float32, "cls_attn"), _ptr(cls_p, torch.float32, "cls_p"),
               _ptr(cls_tile_max, torch.float32, "cls_tile_max"), 1 if causal else 0, _dyn(n_dev), _stream())
    _check(st, "madtp_attn_tc_stats")


def query_sdft_tc(token_att, col_max, col_sum, x2d, row_stride, first_row, n, T, divisor, sd_ft, accumulate, *,
                  n_dev=None):
    """Tensor-core sd_ft: x2d is the dense fp32 matrix [rows, d] of ALL token rows (token j of batch b at row
    b*row_stride + first_row + j)."""
    B = token_att.shape[0]
    d = x2d.shape[1]
    if not x2d.is_contiguous():
        raise RuntimeError("madtp_b200.query_sdft_tc: x2d must be dense")
    st = _call("madtp_query_sdft_tc", _ptr(token_att, torch.float32, "token_att"), token_att.stride(1),
               token_att.stride(0), _ptr(col_max), _ptr(col_sum), _ptr(x2d, torch.float32, "x"), x2d.shape[0],
               int(row_stride), int(first_row), B, n, T, d, float(divisor), _ptr(sd_ft, torch.float32, "sd_ft"),
               1 if accumulate else 0, _dyn(n_dev), _stream())
    _check(st, "madtp_query_sdft_tc")


def query_sdft_planes(token_att, col_max, col_sum, x_hi, x_lo, row_stride, first_row, n, T, divisor, sd_ft, accumulate, *,
                      x_unscale=1.0, n_dev=None):
    """Tensor-core sd_ft from the fp16 hi/lo planes [rows, d] of ALL token rows (x = x_unscale * (x_hi + x_lo); token j
    of batch b at row b*row_stride + first_row + j): MN-major operands, nothing transposed (madtp_query_sdft_planes)."""
    B = token_att.shape[0]
    d = x_hi.shape[1]
    if not (x_hi.is_contiguous() and x_lo.is_contiguous()) or x_hi.shape != x_lo.shape:
        raise RuntimeError("madtp_b200.query_sdft_planes: x_hi / x_lo must be dense [rows, d] planes of the same shape")
    st = _call("madtp_query_sdft_planes", _ptr(token_att, torch.float32, "token_att"), token_att.stride(1),
               token_att.stride(0), _ptr(col_max), _ptr(col_sum), _ptr(x_hi, torch.float16, "x_hi"),
               _ptr(x_lo, torch.float16, "x_lo"), float(x_unscale), x_hi.shape[0], int(row_stride), int(first_row), B, n,
               T, d, float(divisor), _ptr(sd_ft, torch.float32, "sd_ft"), 1 if accumulate else 0, _dyn(n_dev), _stream())
    _check(st, "madtp_query_sdft_planes")


def attn_small_self(q, k, v, H, scale, out_f16, *, key_mask=None, col_sum=None, cls_attn=None, causal=False,
                    l_dev=None):
    """Self-attention of a short sequence (L <= 64) with optional fused pruning statistics (col_sum, cls_attn [B, L])."""
    B, Ltok, _ = q.shape
    ldq, bsq = _qkv_strides(q, "q")
    ldk, bsk = _qkv_strides(k, "k")
    ldv, bsv = _qkv_strides(v, "v")
    ldo, bso = _qkv_strides(out_f16, "out_f16")
    if key_mask is not None and (not key_mask.is_contiguous() or key_mask.numel() != B * Ltok):
        raise RuntimeError("madtp_b200.attn_small_self: key_mask must be contiguous [B, L]")
    # the scratch tensor stays referenced until the launch has been enqueued (stream-ordered reuse after that is safe)
    scratch = None if col_sum is None else empty((B * H * Ltok * (Ltok + 1),), torch.float32, q.device)
    st = _call("madtp_attn_small_self", _ptr(q, torch.float32, "q"), ldq, bsq, _ptr(k, torch.float32, "k"), ldk, bsk,
               _ptr(v, torch.float32, "v"), ldv, bsv, B, H, Ltok, float(scale), _ptr(key_mask, torch.float32, "key_mask"),
               _ptr(out_f16, torch.float16, "out_f16"), ldo, bso, _ptr(col_sum), _ptr(cls_attn), _ptr(scratch),
               1 if causal else 0, _dyn(l_dev), _stream())
    _check(st, "madtp_attn_small_self")
    del scratch
