"""``Transformer``: drop-in for the reference model object (Transformer/transformer.py:8-572).

Same constructor, attributes and methods, so ``utils/train.py::train_transformer`` and the
reference's own ``__main__`` sequence (transformer.py:635-641) run unmodified against it -- but
every piece of arithmetic is a hand-written sm_100a kernel behind ``libnlps_b200.so``.

    model = Transformer(vocab, 256, 8, 1024, 3, 320, embedding_weight=We)   # reference signature
    logits, cache = model.forward(x)                  # (B,S,V) device array, opaque cache
    loss, dlogits = model.calculate_loss_dlogits(logits, y)
    grads = model.backward(cache, dlogits)            # dict of views into the flat grad buffer
    model.step(grads, 1e-4)

Differences that are deliberate and documented: arrays live on the GPU
(``model.np`` is ``b200np``, not numpy/cupy); activations are FP32 where NumPy >= 2 silently
promotes the reference to FP64; dropout masks come from Philox, not the NumPy global RNG.
``train_step`` is the fused fast path (no host synchronisation inside).
"""
from __future__ import annotations

import ctypes as C
import math
import pickle
from typing import Dict, Optional

import numpy as np

from . import _lib
from . import array as _array
from .array import DeviceArray, b200np

import os as _os
_STATIC_INPUTS_OFF = _os.environ.get("NLPS_STATIC_INPUTS", "1") == "0"

PARAM_NAMES = ("We", "Wq", "Wk", "Wv", "Wo", "W1", "W2", "b1", "b2", "W_vocab", "b_vocab",
               "ln1_gamma", "ln1_beta", "ln2_gamma", "ln2_beta")
# key order of the dict returned by the reference's two backward variants
_FULL_ORDER = ("We", "Wq", "Wk", "Wv", "Wo", "W1", "W2", "b1", "b2", "W_vocab", "b_vocab",
               "ln2_gamma", "ln2_beta", "ln1_gamma", "ln1_beta")       # setdefault order, :369-409
_LAST_ORDER = PARAM_NAMES                                              # :211-227


class Cache:
    """Opaque activations of one forward call (the reference's ``cache`` dict is only ever passed
    back to ``backward*``, utils/train.py:112,120)."""

    def __init__(self, model: "Transformer", handle, B: int, S: int, x_dev: DeviceArray):
        self._model = model
        self._h = handle
        self.B, self.S = B, S
        self.x = x_dev

    def release(self):
        if self._h is not None and self._model._h is not None:
            _lib.load().nlps_cache_release(self._model._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def __getitem__(self, key):              # cache['x'], cache['H_final'] like the reference dict
        if key == "x":
            return self.x
        if key == "H_final":
            return self._model._hidden_view(self)
        if key == "layers":                  # cache['layers'][l]['ffn_hidden'] ... (transformer.py:163-171)
            return [_LayerCache(self, l) for l in range(self._model.num_layers)]
        raise KeyError(key)


class _LayerCache:
    """Read-only view of one layer's cached activations under the reference's key names."""
    _KEYS = {"attn_input": (0, "E"), "ffn_input": (1, "E"), "ffn_hidden": (2, "F"), "ctx": (3, "E"), "qkv": (4, "3E")}

    def __init__(self, cache: Cache, layer: int):
        self._cache, self._layer = cache, layer

    def __getitem__(self, key):
        if key not in self._KEYS:
            raise KeyError("%r is not cached by this build (never materialised: attention weights, masks, dropout masks)" % key)
        which, width = self._KEYS[key]
        m, c = self._cache._model, self._cache
        cols = {"E": m.embed_dim, "F": m.ff_dim, "3E": 3 * m.embed_dim}[width]
        p = C.c_void_p()
        _lib.call("nlps_cache_layer_ptr", m._h, c._h, C.c_int(self._layer), C.c_int(which), C.byref(p))
        return DeviceArray(c, p.value, (c.B, c.S, cols), (c.S * cols, cols, 1), np.float32)


class DeviceScalar:
    """A float that still lives on the device; ``float()`` is the synchronisation point
    (utils/train.py:138 ``float(loss)``)."""

    def __init__(self, arr: DeviceArray, model: "Optional[Transformer]" = None):
        self._arr = arr
        self._model = model

    def __float__(self):
        value = float(self._arr.item())
        if self._model is not None:              # the read synchronised: surface what the kernels flagged on the way
            self._model._check_errors()
        return value

    def item(self):
        return float(self)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(float(self), dtype=dtype or np.float64)

    def __repr__(self):
        return "DeviceScalar(%r)" % float(self)

    def __format__(self, spec):
        return format(float(self), spec)

    def _f(self, o):
        return float(o)

    def __add__(self, o): return float(self) + self._f(o)
    __radd__ = __add__
    def __sub__(self, o): return float(self) - self._f(o)
    def __rsub__(self, o): return self._f(o) - float(self)
    def __mul__(self, o): return float(self) * self._f(o)
    __rmul__ = __mul__
    def __truediv__(self, o): return float(self) / self._f(o)
    def __neg__(self): return -float(self)
    def __lt__(self, o): return float(self) < self._f(o)
    def __gt__(self, o): return float(self) > self._f(o)


class HostScalarSlot:
    """A pinned 4-byte host slot + event for reading a device scalar (a step's loss) WITHOUT stalling the GPU:
    ``slot.fetch(loss)`` enqueues the copy behind the step that produces the value and returns at once; ``slot.value()``
    waits for that copy only.  A loop that fetches step i's loss and reads step i-1's keeps the GPU busy back to back --
    what the reference's ``float(loss)`` every iteration (utils/train.py:138) cannot do."""

    def __init__(self, model: "Transformer"):
        self._model = model
        self._host = C.c_void_p()
        _lib.call("nlps_host_alloc_pinned", C.byref(self._host), C.c_size_t(4))
        self._event = C.c_void_p()
        _lib.call("nlps_event_create", C.byref(self._event))
        self._pending = False

    def fetch(self, scalar: "DeviceScalar"):
        st = C.c_void_p()
        _lib.call("nlps_model_stream", self._model._h, C.byref(st))
        _lib.call("nlps_memcpy_d2h_async", self._host, _lib.vp(scalar._arr.ptr), C.c_size_t(4), st)
        _lib.call("nlps_event_record", self._event, st)
        self._keep = scalar                       # the device value must outlive the copy
        self._pending = True

    def value(self) -> float:
        if not self._pending:
            raise RuntimeError("HostScalarSlot.value() before fetch()")
        _lib.call("nlps_event_synchronize", self._event)
        return float(C.cast(self._host, C.POINTER(C.c_float))[0])

    def close(self):
        if self._event:
            _lib.load().nlps_event_destroy(self._event)
            _lib.load().nlps_host_free_pinned(self._host)
            self._event = self._host = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _lecun_host(shape, fan_in):
    """U(-1/sqrt(fan_in), +1/sqrt(fan_in)) from NumPy's GLOBAL generator, drawn in the reference's
    order (transformer.py:34-54,596-598) so ``np.random.seed(k)`` gives the reference's weights."""
    scale = np.sqrt(1 / fan_in)
    return np.random.uniform(-scale, scale, shape).astype(np.float32)


class Transformer:
    def __init__(self, vocab_size, embed_dim, num_heads, ff_dim, num_layers, max_len,
                 embedding_weight=None, use_gpu=False, dropout_p=0.1, use_adam=True, enable_tf32=True,
                 precision: Optional[str] = None, device: int = 0, seed: int = 0x5EED):
        object.__setattr__(self, "_h", None)
        assert embed_dim % num_heads == 0, "embed_dim must be divisible by num_heads"
        _lib.require_device()                      # no CPU path: fail loudly
        self.use_gpu = True                        # the array-module switch has one position here
        self.np = b200np
        self.vocab_size, self.embed_dim, self.num_heads = int(vocab_size), int(embed_dim), int(num_heads)
        self.ff_dim, self.num_layers, self.max_len = int(ff_dim), int(num_layers), int(max_len)
        self.head_dim = self.embed_dim // self.num_heads
        self.use_adam = bool(use_adam)
        self.adam_b1, self.adam_b2, self.adam_eps = 0.9, 0.999, 1e-8
        # enable_tf32 mirrors the reference's cuBLAS TF32 switch (transformer.py:70-80)
        self.precision = precision or ("tf32" if enable_tf32 else "fp32")
        if self.precision not in _lib.PREC:
            raise ValueError("precision must be one of %s" % sorted(_lib.PREC))
        cfg = _lib.Config(self.vocab_size, self.embed_dim, self.num_heads, self.ff_dim, self.num_layers,
                          self.max_len, float(dropout_p), int(self.use_adam), _lib.PREC[self.precision], int(seed))
        h = C.c_void_p()
        _lib.call("nlps_create", C.byref(cfg), C.c_int(device), C.byref(h))
        object.__setattr__(self, "_h", h)
        _lib.call("nlps_set_stream", self._h, _lib.stream_ptr())
        object.__setattr__(self, "_training", True)
        object.__setattr__(self, "_dropout_p", float(dropout_p))
        object.__setattr__(self, "_seed", int(seed))
        self._views: Dict[int, Dict[str, DeviceArray]] = {}
        ld = C.c_int64()
        _lib.call("nlps_logits_ld", self._h, C.byref(ld))
        self._ldv = ld.value
        n = C.c_int64()
        _lib.call("nlps_flat_size", self._h, C.byref(n))
        self.flat_size = n.value

        L, E, F, V = self.num_layers, self.embed_dim, self.ff_dim, self.vocab_size
        init = {}
        init["We"] = (_lecun_host((V, E), E) if embedding_weight is None
                      else np.array(embedding_weight, dtype=np.float32))
        if init["We"].shape != (V, E):
            raise ValueError("embedding_weight must have shape (%d, %d)" % (V, E))
        for name in ("Wq", "Wk", "Wv", "Wo"):
            init[name] = _lecun_host((L, E, E), E)
        init["W1"] = _lecun_host((L, E, F), E)
        init["W2"] = _lecun_host((L, F, E), F)
        init["W_vocab"] = _lecun_host((E, V), E)
        init["ln1_gamma"] = np.ones((L, E), np.float32)
        init["ln2_gamma"] = np.ones((L, E), np.float32)
        for name, arr in init.items():              # biases / betas stay at the zero fill
            self._view(_lib.BUF_PARAM, name).set(arr)

    # ---- views into the flat buffers ------------------------------------------------------------
    def _view(self, buf: int, name: str) -> DeviceArray:
        cache = self._views.setdefault(buf, {})
        if name not in cache:
            v = _lib.View()
            _lib.call("nlps_param_view", self._h, C.c_int(buf), C.c_int(_lib.PARAM_IDS[name]), C.byref(v))
            nd = v.ndim
            cache[name] = DeviceArray(self, v.d_ptr, tuple(v.shape[:nd]), tuple(v.stride[:nd]), np.float32)
        return cache[name]

    def flat(self, which: str = "grad") -> DeviceArray:
        """The whole flat param / grad / Adam buffer as one 1-D device array."""
        buf = {"param": _lib.BUF_PARAM, "grad": _lib.BUF_GRAD, "m": _lib.BUF_M, "v": _lib.BUF_V}[which]
        p = C.c_void_p()
        _lib.call("nlps_flat_ptr", self._h, C.c_int(buf), C.byref(p))
        return DeviceArray(self, p.value, (self.flat_size,), (1,), np.float32)

    def __getattr__(self, name):
        if name in PARAM_NAMES and self.__dict__.get("_h") is not None:
            return self._view(_lib.BUF_PARAM, name)
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in PARAM_NAMES and self.__dict__.get("_h") is not None:
            self._view(_lib.BUF_PARAM, name).set(value)      # model.We = ndarray copies into the buffer
        else:
            object.__setattr__(self, name, value)

    @property
    def m(self):
        return {k: self._view(_lib.BUF_M, k) for k in PARAM_NAMES}

    @property
    def v(self):
        return {k: self._view(_lib.BUF_V, k) for k in PARAM_NAMES}

    @property
    def pe(self):
        p = C.c_void_p()
        _lib.call("nlps_pe_ptr", self._h, C.byref(p))
        return DeviceArray(self, p.value, (self.max_len, self.embed_dim), (self.embed_dim, 1), np.float32)

    @property
    def adam_t(self):
        t = C.c_int64()
        _lib.call("nlps_adam_t", self._h, C.byref(t), C.c_int(0), C.c_int64(0))
        return t.value

    @adam_t.setter
    def adam_t(self, value):
        _lib.call("nlps_adam_t", self._h, None, C.c_int(1), C.c_int64(int(value)))

    @property
    def training(self):
        return self._training

    @training.setter
    def training(self, flag):
        object.__setattr__(self, "_training", bool(flag))
        _lib.call("nlps_set_training", self._h, C.c_int(int(bool(flag))))

    @property
    def dropout_p(self):
        return self._dropout_p

    @dropout_p.setter
    def dropout_p(self, p):
        object.__setattr__(self, "_dropout_p", float(p))
        _lib.call("nlps_set_dropout", self._h, C.c_float(float(p)), C.c_uint64(self._seed))

    def set_precision(self, precision: str):
        _lib.call("nlps_set_precision", self._h, C.c_int(_lib.PREC[precision]))
        object.__setattr__(self, "precision", precision)

    def train(self):                                   # transformer.py:96-97
        self.training = True

    def eval(self):                                    # transformer.py:99-100
        self.training = False

    def synchronize(self):
        _lib.call("nlps_sync", self._h)
        self._check_errors()

    def _check_errors(self):
        """Out-of-range ids cannot raise from inside an asynchronous kernel: the kernels clamp the id and raise a bit in
        the model's error word; it is read (and cleared) at the host's synchronisation points -- ``float(loss)``,
        ``synchronize()``, ``get_params()`` -- and becomes the IndexError NumPy raises in the reference
        (transformer.py:114 ``We[x]``, :538-546 ``log_probs[rows, y]``)."""
        if self.__dict__.get("_h") is None:
            return
        flags = C.c_int(0)
        _lib.call("nlps_check_errors", self._h, C.byref(flags))
        if flags.value & 1:
            raise IndexError("token id out of bounds for axis 0 with size %d (found by the embedding kernel)" % self.vocab_size)
        if flags.value & 2:
            raise IndexError("target id out of bounds for %d classes (found by the cross-entropy kernel)" % self.vocab_size)

    # ---- dropout state (transformer.py:102-108; masks are Philox, keyed by seed / site / forward-call counter) -------
    @property
    def dropout_counter(self) -> int:
        """Forward calls so far = the Philox `step` the NEXT forward draws its masks with."""
        v = C.c_int64()
        _lib.call("nlps_dropout_counter", self._h, C.byref(v), C.c_int(0), C.c_int64(0))
        return v.value

    @dropout_counter.setter
    def dropout_counter(self, value):
        _lib.call("nlps_dropout_counter", self._h, None, C.c_int(1), C.c_int64(int(value)))

    def dropout_mask(self, layer: int, which: int, step: int, shape) -> np.ndarray:
        """Keep-mask (1.0 / 0.0, host array of `shape` = (B, S, E)) that forward call `step` applied -- and backward
        regenerated -- at (layer, which: 0 attention branch, 1 FFN branch).  Debug / parity hook."""
        n = int(np.prod(shape))
        out = DeviceArray.empty((n,), np.float32)
        _lib.call("nlps_dropout_mask", self._h, C.c_int(int(layer)), C.c_int(int(which)), C.c_int64(int(step)),
                  C.c_int64(n), _lib.vp(out.ptr))
        self.synchronize()
        return out.get().reshape(shape)

    def load_params(self, arrays):
        """Copy host arrays (reference shapes) into the parameters; resets Adam like a fresh model."""
        for k in PARAM_NAMES:
            self._view(_lib.BUF_PARAM, k).set(np.asarray(arrays[k], dtype=np.float32))
        for which in ("m", "v"):
            f = self.flat(which)
            _lib.call("nlps_memset", _lib.vp(f.ptr), C.c_int(0), C.c_size_t(f.size * 4), _lib.stream_ptr())
        self.adam_t = 0

    def get_params(self) -> Dict[str, np.ndarray]:
        self.synchronize()
        return {k: self._view(_lib.BUF_PARAM, k).get() for k in PARAM_NAMES}

    # ---- forward (transformer.py:110-199) ----------------------------------------------------------
    def _ids_to_device(self, x) -> DeviceArray:
        xd = x if isinstance(x, DeviceArray) else DeviceArray.from_host(np.asarray(x), np.int32)
        if xd.dtype != np.int32:
            raise TypeError("token ids must be int32")
        if xd.ndim != 2:
            raise ValueError("x must have shape (B, S)")
        if xd.bounds is not None and (xd.bounds[1] >= self.vocab_size or xd.bounds[0] < -self.vocab_size):
            raise IndexError("index %d is out of bounds for axis 0 with size %d" %
                             (xd.bounds[1] if xd.bounds[1] >= self.vocab_size else xd.bounds[0], self.vocab_size))
        if xd.strides[1] != 1 and xd.shape[1] > 1:
            xd = xd.copy()
        return xd

    def forward(self, x, return_hidden_only=False):
        xd = self._ids_to_device(x)
        B, S = xd.shape
        ch = C.c_void_p()
        _lib.call("nlps_cache_create", self._h, C.c_int(B), C.c_int(S), C.byref(ch))
        cache = Cache(self, ch, B, S, xd)
        _lib.call("nlps_forward", self._h, ch, _lib.vp(xd.ptr), C.c_int64(xd.strides[0] if B > 1 else S),
                  C.c_int(1 if return_hidden_only else 0), None, C.c_int64(0))
        if return_hidden_only:
            return self._hidden_view(cache), cache
        p = C.c_void_p()
        _lib.call("nlps_cache_logits_ptr", ch, C.byref(p), None)
        V, ld = self.vocab_size, self._ldv
        return DeviceArray(cache, p.value, (B, S, V), (S * ld, ld, 1), np.float32), cache

    def _hidden_view(self, cache: Cache) -> DeviceArray:
        p = C.c_void_p()
        _lib.call("nlps_cache_hidden_ptr", cache._h, C.byref(p))
        E = self.embed_dim
        return DeviceArray(cache, p.value, (cache.B, cache.S, E), (cache.S * E, E, 1), np.float32)

    # ---- loss (transformer.py:515-549) ---------------------------------------------------------------
    @staticmethod
    def _rows_2d(a: DeviceArray):
        """(rows, cols) view with unit column stride and one row stride, copying only if needed."""
        if a.ndim == 3:
            B, S, V = a.shape
            if a.strides[2] == 1 and S == 1:
                return a.ptr, B, V, (a.strides[0] if B > 1 else V), a
            if a.strides[2] == 1 and (a.strides[0] == S * a.strides[1] or B == 1):
                return a.ptr, B * S, V, a.strides[1], a
            a = a.copy()
            return a.ptr, B * S, V, V, a
        if a.strides[1] != 1 and a.shape[1] > 1:
            a = a.copy()
        return a.ptr, a.shape[0], a.shape[1], (a.strides[0] if a.shape[0] > 1 else a.shape[1]), a

    def calculate_loss_dlogits(self, logits, y, grad_scale: float = 1.0):
        if not isinstance(logits, DeviceArray):
            logits = DeviceArray.from_host(np.asarray(logits), np.float32)
        yd = y if isinstance(y, DeviceArray) else DeviceArray.from_host(np.asarray(y), np.int32)
        if logits.ndim == 2:
            logits = logits[:, None, :]
        if yd.ndim == 1:
            yd = yd[:, None]
        B, S, V = logits.shape
        if V != self.vocab_size:
            raise ValueError("logits have %d classes, the model %d" % (V, self.vocab_size))
        if yd.shape != (B, S):
            raise ValueError("targets %s do not match logits %s" % (yd.shape, logits.shape))
        if yd.bounds is not None and not (-V <= yd.bounds[0] and yd.bounds[1] < V):
            raise IndexError("target id out of bounds for %d classes" % V)
        ptr, rows, _, ld, keep = self._rows_2d(logits)
        yc = yd.contiguous()
        ldd = (V + 3) // 4 * 4
        store = DeviceArray.empty((rows, ldd), np.float32)
        loss = DeviceArray.empty((), np.float32)
        _lib.call("nlps_loss_dlogits", self._h, _lib.vp(ptr), C.c_int64(ld), _lib.vp(yc.ptr), C.c_int64(rows),
                  C.c_float(grad_scale), _lib.vp(loss.ptr), _lib.vp(store.ptr), C.c_int64(ldd))
        dlogits = DeviceArray(store._owner, store.ptr, (B, S, V), (S * ldd, ldd, 1), np.float32)
        del keep
        return DeviceScalar(loss, self), dlogits

    # ---- backward (transformer.py:327-459 and :201-325) ---------------------------------------------------
    def _grad_dict(self, order):
        return {k: self._view(_lib.BUF_GRAD, k) for k in order}

    def backward(self, cache: Cache, dlogits):
        if not isinstance(dlogits, DeviceArray):
            dlogits = DeviceArray.from_host(np.asarray(dlogits), np.float32)
        if dlogits.shape != (cache.B, cache.S, self.vocab_size):
            raise ValueError("dlogits %s does not match the cached forward (%d,%d,%d)" %
                             (dlogits.shape, cache.B, cache.S, self.vocab_size))
        ptr, _, _, ld, keep = self._rows_2d(dlogits)
        if ld % 4 or ptr % 16:                       # tensor-core path wants a TMA-addressable operand
            pad = DeviceArray.empty((cache.B * cache.S, self._ldv), np.float32)
            DeviceArray(pad._owner, pad.ptr, dlogits.shape, (cache.S * self._ldv, self._ldv, 1), np.float32).set(dlogits)
            ptr, ld, keep = pad.ptr, self._ldv, pad
        _lib.call("nlps_backward", self._h, cache._h, _lib.vp(ptr), C.c_int64(ld))
        del keep
        return self._grad_dict(_FULL_ORDER)

    def backward_last_token(self, cache: Cache, t_index, dlogits_last):
        t = t_index if isinstance(t_index, DeviceArray) else DeviceArray.from_host(np.asarray(t_index), np.int32)
        d = dlogits_last if isinstance(dlogits_last, DeviceArray) else \
            DeviceArray.from_host(np.asarray(dlogits_last), np.float32)
        if d.shape != (cache.B, self.vocab_size) or t.shape != (cache.B,):
            raise ValueError("backward_last_token: t_index %s / dlogits_last %s do not match batch %d" %
                             (t.shape, d.shape, cache.B))
        if t.bounds is not None and not (-cache.S <= t.bounds[0] and t.bounds[1] < cache.S):
            raise IndexError("t_index out of bounds for sequence length %d" % cache.S)
        t = t.contiguous()
        ptr, _, _, ld, keep = self._rows_2d(d)
        if ld % 4 or ptr % 16:
            pad = DeviceArray.empty((cache.B, self._ldv), np.float32)
            DeviceArray(pad._owner, pad.ptr, d.shape, (self._ldv, 1), np.float32).set(d)
            ptr, ld, keep = pad.ptr, self._ldv, pad
        _lib.call("nlps_backward_last_token", self._h, cache._h, _lib.vp(t.ptr), _lib.vp(ptr), C.c_int64(ld))
        del keep
        return self._grad_dict(_LAST_ORDER)

    # ---- optimiser (transformer.py:461-513) ------------------------------------------------------------
    def step(self, gradients, learning_rate=0.001):
        own = self._views.get(_lib.BUF_GRAD, {})
        for k in PARAM_NAMES:
            g = gradients[k]
            if g is not own.get(k):                 # foreign arrays (e.g. NumPy): copy into the flat buffer
                self._view(_lib.BUF_GRAD, k).set(g if isinstance(g, DeviceArray) else np.asarray(g, np.float32))
        _lib.call("nlps_step", self._h, C.c_float(float(learning_rate)), C.c_int(0))

    def sgd_step(self, gradients, learning_rate=0.01):
        self.step(gradients, learning_rate)

    def clip_gradients(self, max_norm: float) -> float:
        """Fused equivalent of ``clip_grads(list(grads.values()), max_norm)`` (function.py:27-38);
        returns the pre-clip global norm."""
        out = C.c_float()
        _lib.call("nlps_clip", self._h, C.c_float(float(max_norm)), C.byref(out))
        return out.value

    def train_step(self, x, y, learning_rate=1e-4, clip_grad_norm: Optional[float] = 1.0, grad_scale: float = 1.0,
                   defer_update: bool = False) -> DeviceScalar:
        """forward -> loss -> backward -> clip -> Adam for the full-sequence objective in one call with
        no host synchronisation (transformer.py:635-641 + utils/train.py:122-131)."""
        xd = self._ids_to_device(x)
        yd = y if isinstance(y, DeviceArray) else DeviceArray.from_host(np.asarray(y), np.int32)
        if yd.shape != xd.shape:
            raise ValueError("y %s must match x %s" % (yd.shape, xd.shape))
        if yd.bounds is not None and not (-self.vocab_size <= yd.bounds[0] and yd.bounds[1] < self.vocab_size):
            raise IndexError("target id out of bounds for %d classes" % self.vocab_size)
        yd = yd.contiguous()
        xd, yd = self._static_inputs(xd, yd)
        B, S = xd.shape
        ch = C.c_void_p()
        _lib.call("nlps_cache_create", self._h, C.c_int(B), C.c_int(S), C.byref(ch))
        cache = Cache(self, ch, B, S, xd)
        loss = DeviceArray.empty((), np.float32)
        _lib.call("nlps_train_step", self._h, ch, _lib.vp(xd.ptr), C.c_int64(xd.strides[0] if B > 1 else S),
                  _lib.vp(yd.ptr), C.c_float(float(learning_rate)),
                  C.c_float(float(clip_grad_norm) if clip_grad_norm else 0.0), C.c_float(float(grad_scale)),
                  C.c_int(int(defer_update)), _lib.vp(loss.ptr))
        object.__setattr__(self, "_last_cache", cache)       # keep x/y/activations alive until the next step
        object.__setattr__(self, "_last_y", yd)
        return DeviceScalar(loss, self)

    def _static_inputs(self, xd: DeviceArray, yd: DeviceArray):
        """A captured CUDA graph replays fixed addresses, and ``nlps_train_step`` keys its graphs on the id buffers.  Every
        batch of a shape is therefore copied (two device copies of 4*B*S bytes, ordered before the step) into one pair of
        staging buffers per model, so that fresh batches from a loader replay the same graph instead of running eagerly
        (and instead of a capture per distinct buffer).  NLPS_STATIC_INPUTS=0 passes the caller's buffers through."""
        if _STATIC_INPUTS_OFF:
            return xd, yd
        st = self.__dict__.get("_stage")
        if st is None or st[0].shape != xd.shape:
            st = (DeviceArray.empty(xd.shape, np.int32), DeviceArray.empty(xd.shape, np.int32))
            object.__setattr__(self, "_stage", st)
        sx, sy = st
        if xd.ptr != sx.ptr:
            _array._strided_copy(xd, sx)
        if yd.ptr != sy.ptr:
            _array._strided_copy(yd, sy)
        sx.bounds, sy.bounds = xd.bounds, yd.bounds
        return sx, sy

    def train_step_last_token(self, xb, yb, lens=None, learning_rate=1e-4, clip_grad_norm: Optional[float] = 1.0,
                              grad_scale: float = 1.0, defer_update: bool = False, t_index=None) -> DeviceScalar:
        """One iteration of the reference driver's hot loop (utils/train.py:104-131) as a single call:
        forward on the (already trimmed) batch, last-real-token vocab head, loss, backward_last_token,
        clip, Adam -- with no host synchronisation.  ``lens`` = non-PAD length per row (host or device
        int array); computed on the device when omitted; ``t_index`` (device int32, = lens - 1) wins over both.
        Batches are staged into fixed buffers per shape, so the library replays the step as a CUDA graph."""
        xd = self._ids_to_device(xb)
        yd = yb if isinstance(yb, DeviceArray) else DeviceArray.from_host(np.asarray(yb), np.int32)
        B, S = xd.shape
        if yd.shape != (B,):
            raise ValueError("yb %s must have shape (%d,)" % (yd.shape, B))
        if yd.bounds is not None and not (-self.vocab_size <= yd.bounds[0] and yd.bounds[1] < self.vocab_size):
            raise IndexError("target id out of bounds for %d classes" % self.vocab_size)
        if t_index is not None:
            t = t_index
        elif lens is None:
            t = b200np.sum(xd != 0, axis=1) - 1
        elif isinstance(lens, DeviceArray):
            t = lens - 1
        else:
            t = DeviceArray.from_host(np.asarray(lens, dtype=np.int64) - 1, np.int32)
        if t.shape != (B,):
            raise ValueError("t_index / lens must have shape (%d,)" % B)
        xd, t, yd = self._static_inputs_last(xd, t, yd)
        ch = C.c_void_p()
        _lib.call("nlps_cache_create", self._h, C.c_int(B), C.c_int(S), C.byref(ch))
        cache = Cache(self, ch, B, S, xd)
        loss = DeviceArray.empty((), np.float32)
        _lib.call("nlps_train_step_last_token", self._h, ch, _lib.vp(xd.ptr), C.c_int64(xd.strides[0] if B > 1 else S),
                  _lib.vp(t.ptr), _lib.vp(yd.ptr), C.c_float(float(learning_rate)),
                  C.c_float(float(clip_grad_norm) if clip_grad_norm else 0.0), C.c_float(float(grad_scale)),
                  C.c_int(int(defer_update)), _lib.vp(loss.ptr))
        object.__setattr__(self, "_last_cache", cache)
        object.__setattr__(self, "_last_y", (yd, t))
        return DeviceScalar(loss, self)

    def _static_inputs_last(self, xd: DeviceArray, t: DeviceArray, yd: DeviceArray):
        """Staging buffers per batch shape for the last-token step (see ``_static_inputs``): the driver loop hands over a
        fresh strided slice every iteration; copied into fixed (B,S) / (B,) buffers, every batch of a shape replays one graph."""
        if _STATIC_INPUTS_OFF:
            return (xd if xd.strides[1] == 1 or xd.shape[1] <= 1 else xd.copy()), t.contiguous(), yd.contiguous()
        stages = self.__dict__.get("_stage_last")
        if stages is None:
            stages = {}
            object.__setattr__(self, "_stage_last", stages)
        st = stages.get(xd.shape)
        if st is None:
            B, S = xd.shape
            st = (DeviceArray.empty((B, S), np.int32), DeviceArray.empty((B,), np.int32), DeviceArray.empty((B,), np.int32))
            stages[xd.shape] = st
        sx, st_t, sy = st
        _array._strided_copy(xd, sx)
        _array._strided_copy(t, st_t)
        _array._strided_copy(yd, sy)
        sx.bounds, sy.bounds, st_t.bounds = xd.bounds, yd.bounds, t.bounds
        return sx, st_t, sy

    # ---- evaluation / schedule / sampling of the driver loop without per-batch synchronisation -----------
    def eval_batch_last_token(self, xb, yb, accum: DeviceArray, lens=None):
        """One batch of ``compute_eval_loss`` (utils/train.py:51-75): adds the batch's summed last-token loss and its
        row count to the 2-float device accumulator ``accum``; nothing is read back."""
        xd = self._ids_to_device(xb)
        yd = yb if isinstance(yb, DeviceArray) else DeviceArray.from_host(np.asarray(yb), np.int32)
        B, S = xd.shape
        if lens is None:
            t = b200np.sum(xd != 0, axis=1) - 1
        elif isinstance(lens, DeviceArray):
            t = lens - 1
        else:
            t = DeviceArray.from_host(np.asarray(lens, dtype=np.int64) - 1, np.int32)
        t, yd = t.contiguous(), yd.contiguous()
        ch = C.c_void_p()
        _lib.call("nlps_cache_create", self._h, C.c_int(B), C.c_int(S), C.byref(ch))
        cache = Cache(self, ch, B, S, xd)
        _lib.call("nlps_eval_last_token", self._h, ch, _lib.vp(xd.ptr), C.c_int64(xd.strides[0] if B > 1 else S),
                  _lib.vp(t.ptr), _lib.vp(yd.ptr), _lib.vp(accum.ptr))
        object.__setattr__(self, "_last_cache", cache)
        object.__setattr__(self, "_last_y", (yd, t))

    def eval_finish(self, accum: DeviceArray, sched: DeviceArray):
        """Close an evaluation pass on the device: mean loss, "halve the learning rate if the loss went up"
        (utils/train.py:86-88).  Returns (eval_loss, learning_rate, halved) -- the one read-back of the pass."""
        _lib.call("nlps_eval_finish", self._h, _lib.vp(accum.ptr), _lib.vp(sched.ptr))
        prev, n, lr, halved, loss = (float(v) for v in sched.get())
        self._check_errors()
        return loss, lr, bool(halved)

    def topk_next(self, prompt, k: int = 5):
        """The sampling block of the driver (utils/train.py:143-163) for one prompt: ids of the k most probable next
        tokens, most probable first, and their probabilities -- softmax and selection on the device."""
        seq = DeviceArray.from_host(np.asarray(prompt, dtype=np.int32)[None, :], np.int32)
        was = self.training
        self.training = False
        try:
            H, _ = self.forward(seq, return_hidden_only=True)
        finally:
            self.training = was
        logits = (H[0, -1] @ self.W_vocab + self.b_vocab).contiguous()
        V = self.vocab_size
        work, idx, prob = DeviceArray.empty((V,), np.float32), DeviceArray.empty((k,), np.int32), DeviceArray.empty((k,), np.float32)
        _lib.call("nlps_topk_softmax", _lib.stream_ptr(), _lib.vp(logits.ptr), C.c_int(V), C.c_int(int(k)), _lib.vp(work.ptr),
                  _lib.vp(idx.ptr), _lib.vp(prob.ptr))
        return idx.get(), prob.get()

    def apply_update(self, learning_rate, clip_grad_norm: Optional[float] = 1.0):
        """Second half of a deferred ``train_step`` (after the data-parallel all-reduce)."""
        if clip_grad_norm:
            _lib.call("nlps_clip_compute", self._h, C.c_float(float(clip_grad_norm)))
        _lib.call("nlps_step", self._h, C.c_float(float(learning_rate)), C.c_int(1 if clip_grad_norm else 0))

    def graph_stats(self):
        """(eager, captured, replayed) counts of the fused steps: how often the CUDA-graph path really served them."""
        e, c, r = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.call("nlps_graph_stats", self._h, C.byref(e), C.byref(c), C.byref(r))
        return e.value, c.value, r.value

    def launch_count(self, reset=False) -> int:
        n = C.c_int64()
        _lib.call("nlps_launch_count", self._h, C.byref(n), C.c_int(int(reset)))
        return n.value

    # ---- checkpoint (transformer.py:551-572) -----------------------------------------------------------
    def save(self, filename):
        with open(filename, "wb") as f:
            pickle.dump(self.get_params(), f)

    def load(self, filename):
        with open(filename, "rb") as f:
            self.load_params(pickle.load(f))

    def save_checkpoint(self, filename):
        """Resume checkpoint (the reference only has the parameter pickle of ``save``, transformer.py:551-572):
        the reference-format parameter dict plus Adam's moments and step count (host float32 arrays) and the dropout
        stream's position (Philox seed + forward-call counter), so that a resumed run with dropout_p > 0 draws the masks
        the uninterrupted run would have drawn."""
        state = {"format": "nlpscratch_b200.checkpoint.v2",
                 "dropout_counter": int(self.dropout_counter), "dropout_seed": int(self._seed),
                 "params": self.get_params(),
                 "m": {k: np.ascontiguousarray(a.get()) for k, a in self.m.items()},
                 "v": {k: np.ascontiguousarray(a.get()) for k, a in self.v.items()},
                 "adam_t": int(self.adam_t)}
        with open(filename, "wb") as f:
            pickle.dump(state, f)

    def load_checkpoint(self, filename):
        """Inverse of ``save_checkpoint``; a plain reference-format pickle (``save``) loads as parameters with fresh Adam state."""
        with open(filename, "rb") as f:
            state = pickle.load(f)
        if "params" not in state:
            self.load_params(state)
            return
        self.load_params(state["params"])
        for k in PARAM_NAMES:
            self._view(_lib.BUF_M, k).set(np.asarray(state["m"][k], dtype=np.float32))
            self._view(_lib.BUF_V, k).set(np.asarray(state["v"][k], dtype=np.float32))
        self.adam_t = int(state["adam_t"])
        if "dropout_counter" in state:               # v2: the Philox stream continues where the writer stopped
            object.__setattr__(self, "_seed", int(state["dropout_seed"]))
            _lib.call("nlps_set_dropout", self._h, C.c_float(self._dropout_p), C.c_uint64(self._seed))
            self.dropout_counter = int(state["dropout_counter"])

    def close(self):
        h = self.__dict__.get("_h")
        if h is not None:
            object.__setattr__(self, "_last_cache", None)
            self._views.clear()
            _lib.load().nlps_destroy(h)
            object.__setattr__(self, "_h", None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
