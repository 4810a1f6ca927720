"""Device arrays and the ``model.np`` array-module the reference driver talks to.

``utils/train.py::train_transformer`` never imports NumPy for the hot loop; it uses
``npb = model.np`` (utils/train.py:40) and plain operators on whatever ``npb.array`` returns.
``B200Module`` (exported as ``b200np``) is that third array module next to numpy and cupy: it
covers exactly the calls the driver and ``utils/function.py::clip_grads`` make, every
one of them a CUDA kernel behind the C ABI (or a stride-only view).  Nothing here computes on the
host; host<->device copies happen only in ``array``/``asnumpy``/``float()``/``int()``.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib

_F32, _I32 = np.dtype(np.float32), np.dtype(np.int32)


# ---------------------------------------------------------------------------------------------
# memory: a tiny size-bucketed pool on top of nlps_malloc (cudaFree would synchronise the device)
# ---------------------------------------------------------------------------------------------
class _Pool:
    def __init__(self):
        self.free = {}
        self.held = 0

    @staticmethod
    def _bucket(nbytes: int) -> int:
        nbytes = max(int(nbytes), 512)
        return 1 << (nbytes - 1).bit_length() if nbytes < (1 << 20) else ((nbytes + (1 << 20) - 1) >> 20) << 20

    def take(self, nbytes: int):
        b = self._bucket(nbytes)
        lst = self.free.get(b)
        if lst:
            return lst.pop(), b
        p = C.c_void_p()
        try:
            _lib.call("nlps_malloc", C.byref(p), C.c_size_t(b))
        except RuntimeError:
            self.trim()
            _lib.call("nlps_malloc", C.byref(p), C.c_size_t(b))
        return p.value, b

    def give(self, ptr: int, bucket: int):
        self.free.setdefault(bucket, []).append(ptr)

    def trim(self):
        for lst in self.free.values():
            for p in lst:
                _lib.load().nlps_free(C.c_void_p(p))
        self.free.clear()


_pool = _Pool()


class _Buffer:
    """Owning handle of one pooled device allocation."""
    __slots__ = ("ptr", "bucket")

    def __init__(self, nbytes: int):
        self.ptr, self.bucket = _pool.take(nbytes)

    def __del__(self):
        try:
            _pool.give(self.ptr, self.bucket)
        except Exception:
            pass


def _c_contig_strides(shape):
    st, acc = [], 1
    for n in reversed(shape):
        st.append(acc)
        acc *= n
    return tuple(reversed(st))


def _i64arr(vals):
    return (C.c_int64 * max(1, len(vals)))(*vals)


class DeviceArray:
    """N-d strided view of device memory (float32 or int32; strides in elements)."""
    __array_priority__ = 1000     # numpy scalars defer to our reflected operators

    def __init__(self, owner, ptr: int, shape, strides, dtype, is_bool=False, bounds=None):
        self._owner = owner                 # keeps the allocation (or model / cache) alive
        self.ptr = int(ptr)
        self.shape = tuple(int(s) for s in shape)
        self.strides = tuple(int(s) for s in strides)
        self.dtype = np.dtype(dtype)
        self.is_bool = is_bool
        self.bounds = bounds                # (min,max) of int ids known on the host, if any

    # ---- construction ---------------------------------------------------------------------
    @staticmethod
    def empty(shape, dtype=np.float32) -> "DeviceArray":
        shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        dtype = np.dtype(dtype)
        if dtype not in (_F32, _I32):
            raise TypeError("device arrays are float32 or int32, not %s" % dtype)
        n = int(np.prod(shape)) if shape else 1
        buf = _Buffer(max(n, 1) * 4)
        return DeviceArray(buf, buf.ptr, shape, _c_contig_strides(shape), dtype)

    @staticmethod
    def from_host(a, dtype=None) -> "DeviceArray":
        a = np.asarray(a)
        if dtype is None:
            if a.dtype.kind == "f":
                dtype = np.float32
            elif a.dtype.kind in "iub":
                dtype = np.int32
            else:
                raise TypeError("cannot move dtype %s to the device" % a.dtype)
        is_bool = a.dtype.kind == "b"
        h = np.ascontiguousarray(a, dtype=dtype)
        out = DeviceArray.empty(h.shape, dtype)
        out.is_bool = is_bool
        if h.size:
            _lib.call("nlps_memcpy_h2d", _lib.vp(out.ptr), h.ctypes.data_as(C.c_void_p), C.c_size_t(h.nbytes),
                      _lib.stream_ptr())
            # pageable source: the copy is staged before the call returns, `h` may die now
            if np.dtype(dtype) == _I32:
                out.bounds = (int(h.min()), int(h.max()))
        return out

    # ---- basic properties -------------------------------------------------------------------
    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape)) if self.shape else 1

    @property
    def T(self):
        return DeviceArray(self._owner, self.ptr, self.shape[::-1], self.strides[::-1], self.dtype, self.is_bool,
                           self.bounds)

    def __len__(self):
        if not self.shape:
            raise TypeError("len() of unsized object")
        return self.shape[0]

    def is_contiguous(self):
        return self.size <= 1 or all(st == ex for n, st, ex in
                                     zip(self.shape, self.strides, _c_contig_strides(self.shape)) if n != 1)

    def contiguous(self) -> "DeviceArray":
        return self if self.is_contiguous() else self.copy()

    def copy(self) -> "DeviceArray":
        out = DeviceArray.empty(self.shape, self.dtype)
        out.is_bool, out.bounds = self.is_bool, self.bounds
        _strided_copy(self, out)
        return out

    def astype(self, dtype, copy=True):
        dtype = np.dtype(dtype)
        if dtype == self.dtype:
            out = self.copy() if copy else self
            out.is_bool = False
            return out
        if self.dtype == _I32 and dtype == _F32:
            src = self.contiguous()
            out = DeviceArray.empty(self.shape, _F32)
            _lib.call("nlps_cast_i32_f32", _lib.stream_ptr(), C.c_int64(src.size), _lib.vp(src.ptr), _lib.vp(out.ptr))
            return out
        if dtype.kind == "f" and self.dtype == _F32:
            return self.copy() if copy else self      # float64 requests stay float32 on the device
        raise TypeError("unsupported device cast %s -> %s" % (self.dtype, dtype))

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = list(int(s) for s in shape)
        if -1 in shape:
            known = int(np.prod([s for s in shape if s != -1])) or 1
            shape[shape.index(-1)] = self.size // known
        if int(np.prod(shape)) != self.size:
            raise ValueError("cannot reshape %s into %s" % (self.shape, tuple(shape)))
        src = self.contiguous()
        return DeviceArray(src._owner, src.ptr, shape, _c_contig_strides(shape), self.dtype, self.is_bool, self.bounds)

    def flatten(self):
        return self.copy().reshape(-1)

    # ---- host transfer -------------------------------------------------------------------------
    def get(self) -> np.ndarray:
        src = self.contiguous()
        out = np.empty(self.shape, dtype=self.dtype)
        if out.size:
            _lib.call("nlps_memcpy_d2h", out.ctypes.data_as(C.c_void_p), _lib.vp(src.ptr), C.c_size_t(out.nbytes),
                      _lib.stream_ptr())
        return out.astype(bool) if self.is_bool else out

    def __array__(self, dtype=None, copy=None):
        a = self.get()
        return a if dtype is None else a.astype(dtype)

    def item(self):
        if self.size != 1:
            raise ValueError("can only convert an array of size 1 to a Python scalar")
        return self.get().reshape(()).item()

    def __float__(self):
        return float(self.item())

    def __int__(self):
        return int(self.item())

    def __bool__(self):
        return bool(self.item())

    def __repr__(self):
        return "DeviceArray(shape=%s, dtype=%s)" % (self.shape, self.dtype)

    def set(self, value):
        """Copy a host array / device array / scalar into this view (broadcasting like numpy)."""
        if isinstance(value, DeviceArray):
            src = value
            if src.shape != self.shape:
                src = DeviceArray.from_host(np.broadcast_to(value.get(), self.shape), self.dtype)
            elif src.dtype != self.dtype:
                src = src.astype(self.dtype)
        else:
            src = DeviceArray.from_host(np.broadcast_to(np.asarray(value, dtype=self.dtype), self.shape), self.dtype)
        _strided_copy(src, self)
        return self

    # ---- indexing ------------------------------------------------------------------------------
    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        adv = [i for i, k in enumerate(key) if isinstance(k, (DeviceArray, np.ndarray, list))]
        if adv:
            return self._fancy(key, adv)
        return self._basic(key)

    def _basic(self, key) -> "DeviceArray":
        if any(k is Ellipsis for k in key):
            i = [j for j, k in enumerate(key) if k is Ellipsis][0]
            n_real = sum(1 for k in key if k is not None and k is not Ellipsis)
            key = key[:i] + (slice(None),) * (self.ndim - n_real) + key[i + 1:]
        ptr_off, shape, strides, dim = 0, [], [], 0
        for k in key:
            if k is None:
                shape.append(1)
                strides.append(0)
                continue
            if dim >= self.ndim:
                raise IndexError("too many indices for array")
            n, st = self.shape[dim], self.strides[dim]
            if isinstance(k, (int, np.integer)):
                k = int(k)
                if k < 0:
                    k += n
                if not 0 <= k < n:
                    raise IndexError("index %d is out of bounds for axis %d with size %d" % (k, dim, n))
                ptr_off += k * st
            elif isinstance(k, slice):
                start, stop, step = k.indices(n)
                if step <= 0:
                    raise IndexError("device views need a positive slice step")
                shape.append(max(0, (stop - start + step - 1) // step))
                strides.append(st * step)
                ptr_off += start * st
            else:
                raise IndexError("unsupported index %r" % (k,))
            dim += 1
        shape += self.shape[dim:]
        strides += self.strides[dim:]
        return DeviceArray(self._owner, self.ptr + 4 * ptr_off, shape, strides, self.dtype, self.is_bool, self.bounds)

    def _fancy(self, key, adv) -> "DeviceArray":
        def as_index(k):
            k = k if isinstance(k, DeviceArray) else DeviceArray.from_host(np.asarray(k), np.int32)
            if k.dtype != _I32 or k.ndim != 1:
                raise IndexError("index arrays must be 1-D int32")
            return k.contiguous()

        if adv == [0] and all(isinstance(k, slice) and k == slice(None) for k in key[1:]):
            idx = as_index(key[0])                              # a[perm]           (utils/train.py:99-100)
            src = self if self._rows_dense() else self.copy()
            row = int(np.prod(src.shape[1:])) if src.ndim > 1 else 1
            out = DeviceArray.empty((idx.shape[0],) + src.shape[1:], self.dtype)
            out.bounds = self.bounds
            _lib.call("nlps_gather_rows", _lib.stream_ptr(), C.c_int(4), C.c_int64(idx.shape[0]), C.c_int64(row),
                      _lib.vp(src.ptr), C.c_int64(src.strides[0] if src.ndim else 1), _lib.vp(idx.ptr),
                      C.c_int64(src.shape[0]), _lib.vp(out.ptr))
            return out
        if adv == [0, 1] and self.ndim >= 2 and all(isinstance(k, slice) and k == slice(None) for k in key[2:]):
            i0, i1 = as_index(key[0]), as_index(key[1])         # H[idx, t, :]      (utils/train.py:113-115)
            if i0.shape != i1.shape:
                raise IndexError("shape mismatch: indexing arrays could not be broadcast together")
            src = self.contiguous()
            n0, n1 = src.shape[0], src.shape[1]
            for ix, n, ax in ((i0, n0, 0), (i1, n1, 1)):        # host-known ranges raise like NumPy; device-built ones wrap/clamp
                if ix.bounds is not None and not (-n <= ix.bounds[0] and ix.bounds[1] < n):
                    bad = ix.bounds[1] if ix.bounds[1] >= n else ix.bounds[0]
                    raise IndexError("index %d is out of bounds for axis %d with size %d" % (bad, ax, n))
            rest = src.shape[2:]
            row = int(np.prod(rest)) if rest else 1
            out = DeviceArray.empty((i0.shape[0],) + rest, self.dtype)
            # each index wraps on its own axis (t = lens - 1 is -1 for an all-PAD row: NumPy picks that row's last position)
            _lib.call("nlps_gather_rows2", _lib.stream_ptr(), C.c_int(4), C.c_int64(i0.shape[0]), C.c_int64(row),
                      _lib.vp(src.ptr), C.c_int64(n0), C.c_int64(n1), _lib.vp(i0.ptr), _lib.vp(i1.ptr), _lib.vp(out.ptr))
            return out
        raise IndexError("unsupported advanced index pattern on a device array: %r" % (key,))

    def _rows_dense(self):
        """rows may be strided, but each row must be one dense block"""
        if self.ndim <= 1:
            return self.ndim == 0 or self.strides[0] == 1 or self.shape[0] <= 1
        inner = DeviceArray(self._owner, self.ptr, self.shape[1:], self.strides[1:], self.dtype)
        return inner.is_contiguous()

    def __setitem__(self, key, value):
        self[key].set(value) if key is not Ellipsis and key != slice(None) else self.set(value)

    # ---- arithmetic ------------------------------------------------------------------------------
    def _binary(self, other, op, reflected=False, out=None):
        a = self.contiguous()
        if isinstance(other, DeviceArray) and other.size == 1 and other.shape != a.shape:
            other = other.item()
        if isinstance(other, np.ndarray):
            other = other.item() if other.size == 1 else DeviceArray.from_host(other, a.dtype)
        if isinstance(other, DeviceArray):
            b = other.contiguous()
            if b.dtype != a.dtype:
                if _F32 in (a.dtype, b.dtype):
                    a, b = a.astype(_F32, copy=False), b.astype(_F32, copy=False)
            if reflected:
                a, b = b, a
            if a.shape == b.shape:
                mode, cols = 1, 1
            elif b.ndim <= a.ndim and b.shape == a.shape[a.ndim - b.ndim:] and a.dtype == _F32:
                mode, cols = 2, b.size                      # trailing-dims broadcast (x + bias)
            else:
                raise ValueError("operands could not be broadcast together: %s %s" % (a.shape, b.shape))
            res = out if out is not None else DeviceArray.empty(a.shape, a.dtype)
            if a.dtype == _F32:
                _lib.call("nlps_binary_f32", _lib.stream_ptr(), C.c_int(op), C.c_int64(a.size), C.c_int64(cols),
                          _lib.vp(a.ptr), _lib.vp(b.ptr), C.c_float(0), C.c_int(mode), _lib.vp(res.ptr))
            else:
                _lib.call("nlps_binary_i32", _lib.stream_ptr(), C.c_int(op), C.c_int64(a.size), _lib.vp(a.ptr),
                          _lib.vp(b.ptr), C.c_int32(0), C.c_int(mode), _lib.vp(res.ptr))
            return res
        # scalar operand
        if reflected:
            if op == 1:      # s - a  ==  (-a) + s
                return (-self)._binary(other, 0)
            if op == 3:
                raise TypeError("scalar / device array is not supported")
        if a.dtype == _I32 and isinstance(other, (float, np.floating)) and float(other) != int(other):
            a = a.astype(_F32)
        res = out if out is not None else DeviceArray.empty(a.shape, a.dtype)
        if a.dtype == _F32:
            _lib.call("nlps_binary_f32", _lib.stream_ptr(), C.c_int(op), C.c_int64(a.size), C.c_int64(1),
                      _lib.vp(a.ptr), None, C.c_float(float(other)), C.c_int(0), _lib.vp(res.ptr))
        else:
            _lib.call("nlps_binary_i32", _lib.stream_ptr(), C.c_int(op), C.c_int64(a.size), _lib.vp(a.ptr), None,
                      C.c_int32(int(other)), C.c_int(0), _lib.vp(res.ptr))
        return res

    def _inplace(self, other, op):
        if self.is_contiguous():
            self._binary(other, op, out=self)
        else:                                   # strided view (e.g. a Wq gradient): go through a dense copy
            tmp = self.copy()
            tmp._binary(other, op, out=tmp)
            _strided_copy(tmp, self)
        return self

    def __add__(self, o): return self._binary(o, 0)
    def __radd__(self, o): return self._binary(o, 0, reflected=True) if isinstance(o, DeviceArray) else self._binary(o, 0)
    def __sub__(self, o): return self._binary(o, 1)
    def __rsub__(self, o): return self._binary(o, 1, reflected=True)
    def __mul__(self, o): return self._binary(o, 2)
    def __rmul__(self, o): return self._binary(o, 2)
    def __truediv__(self, o): return self._binary(o, 3)
    def __iadd__(self, o): return self._inplace(o, 0)
    def __isub__(self, o): return self._inplace(o, 1)
    def __imul__(self, o): return self._inplace(o, 2)
    def __itruediv__(self, o): return self._inplace(o, 3)

    def __neg__(self):
        if self.dtype == _I32:
            return self._binary(-1, 2)
        return _unary(self, 4)

    def __ne__(self, o):
        if self.dtype != _I32 or not isinstance(o, (int, np.integer)):
            raise TypeError("device `!=` supports int32 arrays against an int scalar")
        src = self.contiguous()
        out = DeviceArray.empty(self.shape, _I32)
        out.is_bool = True
        _lib.call("nlps_compare_ne_i32", _lib.stream_ptr(), C.c_int64(src.size), _lib.vp(src.ptr), C.c_int32(int(o)),
                  _lib.vp(out.ptr))
        return out

    __hash__ = None

    def __matmul__(self, other):
        return matmul(self, other)

    def __rmatmul__(self, other):
        return matmul(DeviceArray.from_host(other, np.float32), self)

    # ---- reductions ---------------------------------------------------------------------------------
    def sum(self, axis=None):
        return _reduce(self, 0, axis)

    def max(self, axis=None):
        return _reduce(self, 1, axis)


class _Squared:
    """Lazy result of ``np.square(g)`` so that ``np.sum(np.square(g))`` is ONE sum-of-squares kernel."""

    def __init__(self, base: DeviceArray):
        self.base = base

    def materialize(self) -> DeviceArray:
        return _unary(self.base, 2)


matmul_precision = "tf32"        # arithmetic of `@` when neither operand belongs to a model (see _operand_precision)


def _operand_precision(*arrays):
    """The arithmetic mode of `@`: that of the model an operand is a view of (a parameter such as ``model.W_vocab``, or an
    activation held by one of its caches) -- per model, not process-wide state."""
    for a in arrays:
        owner = getattr(a, "_owner", None)
        for cand in (owner, getattr(owner, "_model", None)):
            prec = getattr(cand, "precision", None)
            if isinstance(prec, str) and prec in _lib.PREC:
                return prec
    return matmul_precision


def _strided_copy(src: DeviceArray, dst: DeviceArray):
    if src.shape != dst.shape:
        raise ValueError("shape mismatch in device copy: %s vs %s" % (src.shape, dst.shape))
    if src.ndim > 4:
        raise ValueError("device views support at most 4 dimensions")
    if src.size == 0:
        return
    _lib.call("nlps_strided_copy", _lib.stream_ptr(), C.c_int(4), C.c_int(src.ndim), _i64arr(src.shape),
              _lib.vp(src.ptr), _i64arr(src.strides), _lib.vp(dst.ptr), _i64arr(dst.strides))


def _unary(a: DeviceArray, op: int) -> DeviceArray:
    src = a.contiguous()
    if src.dtype != _F32:
        src = src.astype(_F32)
    out = DeviceArray.empty(a.shape, _F32)
    _lib.call("nlps_unary_f32", _lib.stream_ptr(), C.c_int(op), C.c_int64(src.size), _lib.vp(src.ptr), _lib.vp(out.ptr))
    return out


def _reduce(a, op: int, axis=None):
    """sum / max / sum-of-squares.  Full reductions return a host NumPy scalar (one D2H sync, like
    ``float(cupy_scalar)``); the only axis reduction the driver uses is the per-row PAD count."""
    if isinstance(a, _Squared):
        if axis is not None:
            return _reduce(a.materialize(), 0, axis)
        a, op = a.base, 2
    if axis is not None:
        ax = axis if axis >= 0 else a.ndim + axis
        if a.ndim == 2 and ax == 1 and a.is_bool and op == 0:
            raise RuntimeError("internal: bool row sums are routed through B200Module.sum")
        raise NotImplementedError("device reductions along an axis are limited to sum(x != pad, axis=1)")
    src = a.contiguous()
    if src.size == 0:
        raise ValueError("zero-size array to reduction operation")
    out = DeviceArray.empty((), src.dtype)
    name = "nlps_reduce_f32" if src.dtype == _F32 else "nlps_reduce_i32"
    _lib.call(name, _lib.stream_ptr(), C.c_int(op), C.c_int64(src.size), _lib.vp(src.ptr), _lib.vp(out.ptr))
    return out.get().reshape(())[()]


def matmul(a: DeviceArray, b: DeviceArray, precision: Optional[str] = None) -> DeviceArray:
    """``a @ b`` for 1-D/2-D float32 operands on the tensor cores (or the FP32 path)."""
    if not isinstance(b, DeviceArray):
        b = DeviceArray.from_host(b, np.float32)
    if a.dtype != _F32 or b.dtype != _F32:
        raise TypeError("device matmul needs float32 operands")
    squeeze_a = a.ndim == 1
    a2 = a[None, :] if squeeze_a else a
    lead = a2.shape[:-1]
    if a2.ndim > 2:
        a2 = a2.reshape(-1, a2.shape[-1])
    if b.ndim != 2 or a2.shape[1] != b.shape[0]:
        raise ValueError("matmul: shapes %s and %s are not aligned" % (a.shape, b.shape))
    if a2.strides[1] != 1 and a2.shape[1] > 1:
        a2 = a2.copy()
    if b.strides[1] != 1 and b.shape[1] > 1:
        b = b.copy()
    M, K, N = a2.shape[0], a2.shape[1], b.shape[1]
    out = DeviceArray.empty((M, N), _F32)
    prec = _lib.PREC[precision or _operand_precision(b, a)]
    lda = a2.strides[0] if M > 1 else max(K, 1)
    ldb = b.strides[0] if K > 1 else max(N, 1)
    if prec != 0 and (a2.ptr % 16 or b.ptr % 16 or lda % 4 or ldb % 4):
        prec = 0                                   # operands TMA cannot address: CUDA-core FP32 kernel
    _lib.call("nlps_gemm", _lib.stream_ptr(), C.c_int(prec), C.c_int(M), C.c_int(N), C.c_int(K),
              _lib.vp(a2.ptr), C.c_int64(lda), C.c_int(0), _lib.vp(b.ptr), C.c_int64(ldb), C.c_int(0),
              _lib.vp(out.ptr), C.c_int64(N), None, C.c_int(0), None, C.c_int64(0), None, C.c_int64(0), C.c_int(0))
    if squeeze_a:
        return out.reshape(N)
    return out.reshape(*lead, N) if len(lead) > 1 else out


# ---------------------------------------------------------------------------------------------
# the array module
# ---------------------------------------------------------------------------------------------
class _Random:
    """``npb.random``: draws come from NumPy's global generator so a seeded run shuffles exactly like
    the reference (utils/train.py:98); only the result lives on the device."""

    @staticmethod
    def permutation(n):
        return DeviceArray.from_host(np.random.permutation(int(n)), np.int32)

    @staticmethod
    def seed(s):
        np.random.seed(s)

    @staticmethod
    def uniform(low, high, size):
        return DeviceArray.from_host(np.random.uniform(low, high, size), np.float32)

    @staticmethod
    def random(size):
        return DeviceArray.from_host(np.random.random(size), np.float32)


class B200Module:
    """The object behind ``model.np``."""
    float32 = np.float32
    int32 = np.int32
    float64 = np.float64        # accepted as a dtype request, stored as float32
    newaxis = None
    ndarray = DeviceArray
    random = _Random()

    @staticmethod
    def array(obj, dtype=None):
        if isinstance(obj, DeviceArray):
            return obj.astype(dtype) if dtype is not None else obj.copy()
        return DeviceArray.from_host(obj, dtype)

    asarray = array

    @staticmethod
    def asnumpy(a):
        return a.get() if isinstance(a, DeviceArray) else np.asarray(a)

    @staticmethod
    def zeros(shape, dtype=np.float32):
        out = DeviceArray.empty(shape, np.int32 if np.dtype(dtype).kind in "iu" else np.float32)
        _lib.call("nlps_memset", _lib.vp(out.ptr), C.c_int(0), C.c_size_t(out.size * 4), _lib.stream_ptr())
        return out

    @staticmethod
    def zeros_like(a):
        return B200Module.zeros(a.shape, a.dtype)

    @staticmethod
    def ones(shape, dtype=np.float32):
        return B200Module.zeros(shape, dtype) + 1

    @staticmethod
    def arange(*args):
        return DeviceArray.from_host(np.arange(*args), np.int32)

    @staticmethod
    def sum(a, axis=None):
        if isinstance(a, DeviceArray) and axis is not None and a.ndim == 2 and a.dtype == _I32 \
                and (axis == 1 or axis == -1):
            # sum(xb != 0, axis=1): per-row count of non-PAD ids          (utils/train.py:60,108)
            if not a.is_bool:
                raise NotImplementedError("axis sums on the device are limited to boolean masks")
            out = DeviceArray.empty((a.shape[0],), _I32)
            src = a if a.strides[1] == 1 else a.copy()
            _lib.call("nlps_count_nonzero_rows", _lib.stream_ptr(), C.c_int64(src.shape[0]), C.c_int64(src.shape[1]),
                      _lib.vp(src.ptr), C.c_int64(src.strides[0] if src.shape[0] > 1 else src.shape[1]),
                      _lib.vp(out.ptr))
            return out
        return _reduce(a, 0, axis)

    @staticmethod
    def max(a, axis=None):
        return _reduce(a, 1, axis)

    @staticmethod
    def square(a):
        return _Squared(a)

    @staticmethod
    def sqrt(a):
        if isinstance(a, DeviceArray):
            return _unary(a, 1)
        return np.sqrt(a)

    @staticmethod
    def exp(a):
        if isinstance(a, DeviceArray):
            return _unary(a, 0)
        return np.exp(a)

    @staticmethod
    def log(a):
        if isinstance(a, DeviceArray):
            return _unary(a, 3)
        return np.log(a)

    @staticmethod
    def maximum(a, b):
        if isinstance(b, DeviceArray) and not isinstance(a, DeviceArray):
            a, b = b, a
        return a._binary(b, 4)

    @staticmethod
    def matmul(a, b):
        return matmul(a, b)

    @staticmethod
    def get_array_module(*_):
        return b200np


b200np = B200Module()
