"""Host-side mirror of the reference's acquisition step, over the C ABI.

The reference has no plugin interface: the "acquisition function" is the config string
``alparams["measure"]`` (/root/reference/active_learning.py:240,252,256), the graph nodes
``pseudo_confidence`` / ``pseudo_mean_confidence`` / ``pseudo_label`` / ``pseudo_mask``
(:234-269) and the closure ``rank_confidence()`` (:682-715) with its return tuple
``(low_conf_examples, unlabelled_confidence)``.  This module keeps those names, argument
meanings and error behaviour; the arithmetic runs in libalscore.so on the GPU.

PyTorch is used for device memory and streams only.  NumPy / host tensors are staged to the
GPU by the library; nothing is ever scored on the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional, Sequence, Tuple

import numpy as np

from . import _lib

# alparams["measure"] values (conf/default_params.json:49) + the MC-dropout extension
MEASURES = ("entropy", "margin", "confidence", "variance")


_MEASURE_IDS = {}


def measure_id(name: str) -> int:
    """Measure name -> C enum; unknown names raise exactly like active_learning.py:259-260."""
    if name in _MEASURE_IDS:
        return _MEASURE_IDS[name]
    out = C.c_int(-1)
    rc = _lib.load().als_measure_from_name(str(name).encode(), C.byref(out))
    _lib.check(rc)
    _MEASURE_IDS[name] = out.value
    return out.value


class _Logits:
    """A borrowed view of a logits tensor: pointer + dtype + [T,N,H,W,C] + where it lives."""
    __slots__ = ("ptr", "dtype", "T", "N", "H", "W", "C", "on_host", "keep", "device_index", "ndim")

    def __init__(self, obj, dtype: Optional[str] = None):
        self.keep = obj
        self.device_index = None
        torch = _torch()
        if torch is not None and isinstance(obj, torch.Tensor):
            if not obj.is_contiguous():
                raise ValueError("logits must be dense C-contiguous NHWC (got a strided tensor)")
            if obj.dtype == torch.float32:
                self.dtype = _lib.ALS_F32
            elif obj.dtype == torch.bfloat16:
                self.dtype = _lib.ALS_BF16
            else:
                raise ValueError("logits dtype must be float32 or bfloat16, got %s" % obj.dtype)
            self.on_host = not obj.is_cuda
            if obj.is_cuda:
                self.device_index = obj.device.index
            self.ptr = obj.data_ptr()
            shape = tuple(obj.shape)
        elif isinstance(obj, np.ndarray):
            if not obj.flags["C_CONTIGUOUS"]:
                raise ValueError("logits must be dense C-contiguous NHWC (got a strided array)")
            if obj.dtype == np.float32 and dtype in (None, "float32"):
                self.dtype = _lib.ALS_F32
            elif obj.dtype == np.uint16 and dtype == "bfloat16":
                self.dtype = _lib.ALS_BF16      # bf16 bit patterns
            else:
                raise ValueError("logits dtype must be float32 (or uint16 bit patterns with dtype='bfloat16'), "
                                 "got %s" % obj.dtype)
            self.on_host = True
            self.ptr = obj.ctypes.data
            shape = obj.shape
        else:
            raise TypeError("logits must be a torch.Tensor or numpy.ndarray (use score_dlpack for other producers)")
        self.ndim = len(shape)
        if len(shape) == 4:
            self.T = 1
            self.N, self.H, self.W, self.C = (int(v) for v in shape)
        elif len(shape) == 5:
            self.T, self.N, self.H, self.W, self.C = (int(v) for v in shape)
        else:
            raise ValueError("logits must be [N,H,W,C] or [T,N,H,W,C], got shape %s" % (tuple(shape),))


def _torch():
    try:
        import torch
        return torch
    except Exception:  # pragma: no cover
        return None


class Scorer:
    """One alscore context (one GPU).  Not thread-safe; create one per thread / GPU."""

    def __init__(self, device: Optional[int] = None, use_torch_stream: bool = True):
        lib = _lib.load()
        torch = _torch()
        if device is None:
            device = torch.cuda.current_device() if (torch is not None and torch.cuda.is_available()) else 0
        self.device = int(device)
        self._lib = lib
        self._ctx = C.c_void_p()
        _lib.check(lib.als_ctx_create(self.device, C.byref(self._ctx)))
        self._stream = None                 # None: the context's own (non-blocking) stream
        self._follow_torch = bool(use_torch_stream and torch is not None and torch.cuda.is_available())
        self._sync_stream()

    # -- lifetime ---------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.als_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    @property
    def launch_count(self) -> int:
        return int(self._lib.als_launch_count(self._ctx))

    def enable_timing(self, on: bool = True) -> None:
        """Bracket every scoring launch sequence of the logits path with CUDA events (see last_scoring_ms)."""
        self._check(self._lib.als_ctx_enable_timing(self._ctx, 1 if on else 0))

    def last_scoring_ms(self) -> float:
        """Device time of the last scoring launch sequence (waits for it)."""
        ms = C.c_float(0.0)
        self._check(self._lib.als_last_scoring_ms(self._ctx, C.byref(ms)))
        return float(ms.value)

    def _check(self, rc: int) -> None:
        _lib.check(rc, self._ctx)

    def _sync_stream(self) -> None:
        """Entries without a `stream` argument (pool_*, mc_*, host staging) run on the context's stream: keep that
        equal to torch's CURRENT stream, so they are ordered after the producer of the logits and before any torch
        consumer of their results, also under ``with torch.cuda.stream(side)``.  (The library orders the new stream
        behind the work already queued on the old one.)"""
        if not self._follow_torch:
            return
        cur = int(_torch().cuda.current_stream(self.device).cuda_stream)
        if cur != self._stream:
            self._check(self._lib.als_ctx_set_stream(self._ctx, C.c_void_p(cur)))
            self._stream = cur

    def _stream_arg(self, device=None):
        """`stream` argument of the device entries: torch's current stream on the tensor's device (handle 0 is the
        legacy default stream, which is what torch's default stream is)."""
        torch = _torch()
        return C.c_void_p(int(torch.cuda.current_stream(self.device if device is None else device).cuda_stream))

    # -- graph-level boundary (active_learning.py:234-269) ---------------------------------
    def score(self, logits, measure: str = "entropy", *, dtype: Optional[str] = None, out=None):
        """pseudo_mean_confidence: per-image f64 mean confidence.

        Device tensors are scored in place (zero copy) and the result is a torch.float64 CUDA
        tensor (asynchronous); host arrays are staged and the result is a NumPy f64 array."""
        m = measure_id(measure)
        lg = _Logits(logits, dtype)
        if lg.on_host:
            self._sync_stream()
            scores = np.empty(lg.N, dtype=np.float64)
            self._check(self._lib.als_score_host(self._ctx, lg.ptr, lg.dtype, lg.T, lg.N, lg.H, lg.W, lg.C, m,
                                                 scores.ctypes.data, None, None, None, 0.0))
            return scores
        torch = _torch()
        if out is None:
            out = torch.empty(lg.N, dtype=torch.float64, device=logits.device)
        self._check(self._lib.als_score(self._ctx, lg.ptr, lg.dtype, lg.T, lg.N, lg.H, lg.W, lg.C, m,
                                        out.data_ptr(), None, None, None, 0.0, self._stream_arg(logits.device)))
        return out

    def pseudo_annotation(self, logits, measure: str = "entropy", threshold: float = 0.9, *,
                          dtype: Optional[str] = None, want_label: bool = True, want_mask: bool = True,
                          out: Optional[dict] = None):
        """The whole PseudoAnnotation scope in one pass: returns a dict with
        pseudo_confidence f32[N,H,W], pseudo_mean_confidence f64[N], pseudo_label u8[N,H,W],
        pseudo_mask u8[N,H,W] (conf < threshold ? 0 : 1).  `out` (device logits only) is a dict
        returned by an earlier call with the same shape: its tensors are overwritten in place."""
        m = measure_id(measure)
        lg = _Logits(logits, dtype)
        shp = (lg.N, lg.H, lg.W)
        if lg.on_host:
            self._sync_stream()
            conf = np.empty(shp, np.float32)
            label = np.empty(shp, np.uint8) if want_label else None
            mask = np.empty(shp, np.uint8) if want_mask else None
            scores = np.empty(lg.N, np.float64)
            self._check(self._lib.als_score_host(
                self._ctx, lg.ptr, lg.dtype, lg.T, lg.N, lg.H, lg.W, lg.C, m, scores.ctypes.data, conf.ctypes.data,
                label.ctypes.data if want_label else None, mask.ctypes.data if want_mask else None, float(threshold)))
        else:
            torch = _torch()
            dev = logits.device
            if out is not None:
                conf, scores = out["pseudo_confidence"], out["pseudo_mean_confidence"]
                label, mask = out["pseudo_label"], out["pseudo_mask"]
                if tuple(conf.shape) != shp or (want_label and label is None) or (want_mask and mask is None):
                    raise ValueError("`out` does not match this call (shape %s)" % (shp,))
            else:
                conf = torch.empty(shp, dtype=torch.float32, device=dev)
                label = torch.empty(shp, dtype=torch.uint8, device=dev) if want_label else None
                mask = torch.empty(shp, dtype=torch.uint8, device=dev) if want_mask else None
                scores = torch.empty(lg.N, dtype=torch.float64, device=dev)
            self._check(self._lib.als_score(
                self._ctx, lg.ptr, lg.dtype, lg.T, lg.N, lg.H, lg.W, lg.C, m, scores.data_ptr(), conf.data_ptr(),
                label.data_ptr() if want_label else None, mask.data_ptr() if want_mask else None, float(threshold),
                self._stream_arg(dev)))
        return {"pseudo_confidence": conf, "pseudo_mean_confidence": scores, "pseudo_label": label,
                "pseudo_mask": mask}

    # -- fused classifier head (models/enet/enet_modules.py:1294-1381 + active_learning.py:234-269) ------
    @staticmethod
    def head_supported(num_classes: int, measure: str = "entropy", mc_samples: int = 1) -> bool:
        """True if a fused `Final`-head kernel is built for this class count (and MC sample count)."""
        return bool(_lib.load().als_head_supported(int(num_classes), measure_id(measure), int(mc_samples)))

    def prepare_head(self, kernel) -> None:
        """Upload the `Final` layer's transposed-convolution kernel, float32 [3,3,C,16] (TF filter layout
        [kh, kw, out_channels, in_channels], enet_modules.py:1341).  Needed once per weight update."""
        k = np.ascontiguousarray(np.asarray(kernel, dtype=np.float32))
        if k.ndim != 4 or k.shape[0] != 3 or k.shape[1] != 3 or k.shape[3] != 16:
            raise ValueError("Final kernel must be [3,3,C,16], got %s" % (k.shape,))
        self._check(self._lib.als_head_prepare(self._ctx, k.ctypes.data, int(k.shape[2])))
        self._head_classes = int(k.shape[2])

    def _features(self, features):
        torch = _torch()
        if torch is None or not isinstance(features, torch.Tensor) or not features.is_cuda:
            raise TypeError("features must be a CUDA torch.Tensor [N,h,w,16] or [T,N,h,w,16] (host batches go through "
                            "pool_score_features_batch)")
        if features.dtype != torch.float32 or features.dim() not in (4, 5) or features.shape[-1] != 16 or not features.is_contiguous():
            raise ValueError("features must be a dense float32 [N,h,w,16] / [T,N,h,w,16] tensor, got %s %s"
                             % (features.dtype, tuple(features.shape)))
        shp = tuple(int(v) for v in features.shape[:-1])
        return torch, ((1,) + shp if len(shp) == 3 else shp)

    def score_features(self, features, measure: str = "entropy", *, out=None):
        """pseudo_mean_confidence of the logits `Final` would produce from `features` [N,h,w,16] (or the T
        Monte-Carlo samples [T,N,h,w,16]) -- without materialising them.  Returns a torch.float64 CUDA tensor [N]
        (asynchronous)."""
        m = measure_id(measure)
        torch, (t, n, h, w) = self._features(features)
        if out is None:
            out = torch.empty(n, dtype=torch.float64, device=features.device)
        self._check(self._lib.als_score_features(self._ctx, features.data_ptr(), t, n, h, w, m, out.data_ptr(), None, None,
                                                 None, 0.0, self._stream_arg(features.device)))
        return out

    def pseudo_annotation_features(self, features, measure: str = "entropy", threshold: float = 0.9):
        """The PseudoAnnotation scope from the `Final`-layer input: dict like pseudo_annotation(), maps are [N,2h,2w]."""
        m = measure_id(measure)
        torch, (t, n, h, w) = self._features(features)
        dev = features.device
        shp = (n, 2 * h, 2 * w)
        conf = torch.empty(shp, dtype=torch.float32, device=dev)
        label = torch.empty(shp, dtype=torch.uint8, device=dev)
        mask = torch.empty(shp, dtype=torch.uint8, device=dev)
        scores = torch.empty(n, dtype=torch.float64, device=dev)
        self._check(self._lib.als_score_features(self._ctx, features.data_ptr(), t, n, h, w, m, scores.data_ptr(),
                                                 conf.data_ptr(), label.data_ptr(), mask.data_ptr(), float(threshold),
                                                 self._stream_arg(dev)))
        return {"pseudo_confidence": conf, "pseudo_mean_confidence": scores, "pseudo_label": label, "pseudo_mask": mask}

    def pool_score_features_batch(self, features, batch_indices, measure: str = "entropy") -> None:
        """pool_score_batch from the `Final`-layer input [B,h,w,16] or [T,B,h,w,16] (CUDA tensor, or host tensor /
        array: staged)."""
        m = measure_id(measure)
        torch = _torch()
        if torch is not None and isinstance(features, torch.Tensor):
            if features.dtype != torch.float32 or not features.is_contiguous():
                raise ValueError("features must be dense float32")
            ptr, on_host, shape = features.data_ptr(), not features.is_cuda, tuple(features.shape)
        elif isinstance(features, np.ndarray):
            if features.dtype != np.float32 or not features.flags["C_CONTIGUOUS"]:
                raise ValueError("features must be dense float32")
            ptr, on_host, shape = features.ctypes.data, True, features.shape
        else:
            raise TypeError("features must be a torch.Tensor or numpy.ndarray")
        if len(shape) not in (4, 5) or shape[-1] != 16:
            raise ValueError("features must be [B,h,w,16] or [T,B,h,w,16], got %s" % (tuple(shape),))
        t, b, h, w = ((1,) + tuple(shape[:3])) if len(shape) == 4 else tuple(shape[:4])
        idx = np.ascontiguousarray(np.asarray(batch_indices, dtype=np.int64))
        if idx.shape != (b,):
            raise ValueError("batch_indices must have one entry per image (%d), got shape %s" % (b, idx.shape))
        self._keep = features
        self._sync_stream()
        self._check(self._lib.als_pool_score_features_batch(self._ctx, ptr, 1 if on_host else 0, int(t), int(b), int(h),
                                                            int(w), m, idx.ctypes.data))

    def score_dlpack(self, producer, measure: str = "entropy") -> np.ndarray:
        """Score any ``__dlpack__`` producer (TF >= 2.2, CuPy, JAX, NumPy, torch): device tensors
        zero-copy, host tensors staged.  Returns NumPy f64[N]."""
        m = measure_id(measure)
        self._sync_stream()
        stream_arg = _lib.ALS_STREAM_CTX
        if hasattr(producer, "__dlpack__"):
            # DLPack stream exchange: hand the producer the stream we will score on, it makes the data ready there.
            # (CUDA convention: 1 = legacy default stream, 2 = per-thread default, larger = a cudaStream_t handle.)
            on_cuda = hasattr(producer, "__dlpack_device__") and producer.__dlpack_device__()[0] in (2, 13)
            if on_cuda:
                # without a torch stream to follow, score on the legacy default stream (handle 0) and say so
                h = self._stream if self._stream is not None else 0
                capsule = producer.__dlpack__(stream=h if h > 2 else 1)
                stream_arg = C.c_void_p(h)
            else:
                capsule = producer.__dlpack__()
        else:
            capsule = producer
        api = C.pythonapi
        api.PyCapsule_GetPointer.restype = C.c_void_p
        api.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
        api.PyCapsule_IsValid.argtypes = [C.py_object, C.c_char_p]
        api.PyCapsule_SetName.argtypes = [C.py_object, C.c_char_p]
        if not api.PyCapsule_IsValid(capsule, b"dltensor"):
            raise ValueError("expected a 'dltensor' DLPack capsule (already consumed or versioned capsule?)")
        managed = api.PyCapsule_GetPointer(capsule, b"dltensor")
        # read N for the output size: DLTensor.shape at offset 24, ndim at 16
        ndim = C.c_int32.from_address(managed + 16).value
        shape_ptr = C.c_void_p.from_address(managed + 24).value
        if ndim not in (4, 5):
            raise ValueError("logits must be [N,H,W,C] or [T,N,H,W,C], got ndim=%d" % ndim)
        n = C.c_int64.from_address(shape_ptr + 8 * (ndim - 4)).value
        scores = np.empty(max(n, 0), np.float64)
        try:
            self._check(self._lib.als_score_dlpack(self._ctx, managed, m, scores.ctypes.data, stream_arg))
        finally:
            # consumer protocol: mark the capsule used and run the producer's deleter
            api.PyCapsule_SetName(capsule, _USED_NAME)
            deleter = C.c_void_p.from_address(managed + 56).value
            if deleter:
                C.CFUNCTYPE(None, C.c_void_p)(deleter)(managed)
        return scores

    # -- loop-level boundary (rank_confidence, active_learning.py:682-715) ------------------
    def pool_begin(self, num_examples: int) -> None:
        self._sync_stream()
        self._check(self._lib.als_pool_begin(self._ctx, int(num_examples)))

    def pool_score_batch(self, logits, batch_indices, measure: str = "entropy", *, dtype: Optional[str] = None) -> None:
        m = measure_id(measure)
        lg = _Logits(logits, dtype)
        idx = np.ascontiguousarray(np.asarray(batch_indices, dtype=np.int64))
        if idx.shape != (lg.N,):
            raise ValueError("batch_indices must have one entry per image (%d), got shape %s" % (lg.N, idx.shape))
        self._sync_stream()
        self._check(self._lib.als_pool_score_batch(self._ctx, lg.ptr, 1 if lg.on_host else 0, lg.dtype, lg.T, lg.N,
                                                   lg.H, lg.W, lg.C, m, idx.ctypes.data))

    def pool_scores(self, num_examples: int) -> np.ndarray:
        out = np.empty(int(num_examples), np.float32)
        self._sync_stream()
        self._check(self._lib.als_pool_scores(self._ctx, out.ctypes.data, int(num_examples)))
        return out

    def pool_select(self, unlabelled, selection_size: int) -> Tuple[np.ndarray, np.ndarray]:
        unl = np.ascontiguousarray(np.asarray(unlabelled, dtype=np.int64))
        if unl.ndim != 1:
            raise ValueError("unlabelled must be a 1-D index array")
        k = int(max(0, min(int(selection_size), unl.size)))
        ids = np.empty(k, np.int64)
        conf = np.empty(unl.size, np.float32)
        cnt = C.c_int64(0)
        self._sync_stream()
        self._check(self._lib.als_pool_select(self._ctx, unl.ctypes.data, unl.size, int(selection_size),
                                              ids.ctypes.data, conf.ctypes.data, C.byref(cnt)))
        return ids[:cnt.value], conf

    def rank_pool(self, logits, unlabelled, selection_size: int, measure: str = "entropy", *, example_index=None,
                  num_examples: Optional[int] = None, dtype: Optional[str] = None) -> Tuple[np.ndarray, np.ndarray]:
        """The whole closure (:682-715) in one library call for a pool held in ONE tensor: np.zeros, score + scatter,
        filter, k lowest.  Returns (low_conf_examples, unlabelled_confidence)."""
        m = measure_id(measure)
        lg = _Logits(logits, dtype)
        unl = np.ascontiguousarray(np.asarray(unlabelled, dtype=np.int64))
        if unl.ndim != 1:
            raise ValueError("unlabelled must be a 1-D index array")
        idx = None
        if example_index is not None:
            idx = np.ascontiguousarray(np.asarray(example_index, dtype=np.int64))
            if idx.shape != (lg.N,):
                raise ValueError("example_index must have one entry per image (%d), got shape %s" % (lg.N, idx.shape))
        if num_examples is None:
            num_examples = int(max(idx.max(initial=-1) if idx is not None else lg.N - 1, unl.max(initial=-1))) + 1
        k = int(max(0, min(int(selection_size), unl.size)))
        ids = np.empty(k, np.int64)
        conf = np.empty(unl.size, np.float32)
        cnt = C.c_int64(0)
        self._sync_stream()
        self._keep = logits
        self._check(self._lib.als_rank_pool(self._ctx, lg.ptr, 1 if lg.on_host else 0, lg.dtype, lg.T, lg.N, lg.H, lg.W, lg.C, m,
                                            None if idx is None else idx.ctypes.data, int(num_examples), unl.ctypes.data,
                                            unl.size, int(selection_size), ids.ctypes.data, conf.ctypes.data, C.byref(cnt)))
        return ids[:cnt.value], conf

    # -- multi-GPU: pool sharded by image (csrc/comm.cu) ------------------------------------------------
    def comm_init(self, rank: int, world: int, unique_id: bytes) -> None:
        """Join the NCCL communicator of a sharded pool pass (one process per GPU).  `unique_id` is the 128-byte id
        rank 0 got from ``comm_unique_id()`` and passed around out of band (e.g. a torch.distributed broadcast)."""
        if len(unique_id) != 128:
            raise ValueError("unique_id must be 128 bytes")
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self._lib.als_comm_init_rank(self._ctx, int(rank), int(world), buf))

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _lib.check(_lib.load().als_comm_unique_id(buf))
        return buf.raw

    @staticmethod
    def comm_init_all(scorers: Sequence["Scorer"]) -> None:
        """Single-process form: scorers[i] (one per GPU) becomes rank i of one communicator (ncclCommInitAll)."""
        arr = (C.c_void_p * len(scorers))(*[s._ctx.value for s in scorers])
        _lib.check(_lib.load().als_comm_init_all(arr, len(scorers)), scorers[0]._ctx)

    def pool_select_global(self, unlabelled, selection_size: int, shard: Tuple[int, int], max_shard: int = 0):
        """:705-715 over the pool sharded across the communicator: this rank owns (and has scored) the example ids
        [shard[0], shard[1]).  Collective; returns the same (low_conf_examples, unlabelled_confidence) on every rank.
        On the device: local k candidates -> ONE ncclAllGather -> merge; one device->host copy."""
        unl = np.ascontiguousarray(np.asarray(unlabelled, dtype=np.int64))
        if unl.ndim != 1:
            raise ValueError("unlabelled must be a 1-D index array")
        k = int(max(0, min(int(selection_size), unl.size)))
        ids = np.empty(k, np.int64)
        conf = np.empty(unl.size, np.float32)
        cnt = C.c_int64(0)
        self._sync_stream()
        self._check(self._lib.als_pool_select_global(self._ctx, unl.ctypes.data, unl.size, int(selection_size),
                                                     int(shard[0]), int(shard[1]), int(max_shard), ids.ctypes.data,
                                                     conf.ctypes.data, C.byref(cnt)))
        return ids[:cnt.value], conf

    @staticmethod
    def pool_select_global_all(scorers: Sequence["Scorer"], unlabelled, selection_size: int,
                               shards: Sequence[Tuple[int, int]], max_shard: int = 0):
        """Single-process form of pool_select_global over scorers joined by comm_init_all."""
        unl = np.ascontiguousarray(np.asarray(unlabelled, dtype=np.int64))
        k = int(max(0, min(int(selection_size), unl.size)))
        ids = np.empty(k, np.int64)
        conf = np.empty(unl.size, np.float32)
        cnt = C.c_int64(0)
        n = len(scorers)
        arr = (C.c_void_p * n)(*[s._ctx.value for s in scorers])
        lo = np.ascontiguousarray([int(s[0]) for s in shards], dtype=np.int64)
        hi = np.ascontiguousarray([int(s[1]) for s in shards], dtype=np.int64)
        for s in scorers:
            s._sync_stream()
        _lib.check(_lib.load().als_pool_select_global_all(arr, n, unl.ctypes.data, unl.size, int(selection_size),
                                                          lo.ctypes.data, hi.ctypes.data, int(max_shard), ids.ctypes.data,
                                                          conf.ctypes.data, C.byref(cnt)), scorers[0]._ctx)
        return ids[:cnt.value], conf

    # -- streamed Monte-Carlo accumulation (one dropout pass at a time; Welford state stays in HBM) -------
    def mc_begin(self, shape, dtype: str = "float32", label=None) -> None:
        """Open an accumulation over a batch [N,H,W,C].  `label` (optional CUDA uint8 [N,H,W]) receives the argmax of
        sample 0.  Then mc_add_sample(logits_t) per stochastic forward pass and mc_finish(measure)."""
        n, h, w, c = (int(v) for v in shape)
        self._sync_stream()
        self._mc_shape = (n, h, w, c)
        self._check(self._lib.als_mc_begin(self._ctx, _lib.ALS_F32 if dtype == "float32" else _lib.ALS_BF16, n, h, w, c,
                                           None if label is None else label.data_ptr()))
        self._mc_dtype = _lib.ALS_F32 if dtype == "float32" else _lib.ALS_BF16
        self._mc_label = label

    def mc_add_sample(self, logits, *, dtype: Optional[str] = None) -> None:
        lg = _Logits(logits, dtype)
        if lg.T != 1 or (lg.N, lg.H, lg.W, lg.C) != self._mc_shape or lg.dtype != self._mc_dtype:
            raise ValueError("sample must be [N,H,W,C] = %s of the dtype given to mc_begin" % (self._mc_shape,))
        self._sync_stream()
        self._keep = logits
        self._check(self._lib.als_mc_add_sample(self._ctx, lg.ptr, 1 if lg.on_host else 0))

    def mc_finish(self, measure: str = "variance", *, batch_indices=None, threshold: float = 0.9, want_maps: bool = False):
        """Close the accumulation.  Returns pseudo_mean_confidence (torch f64 CUDA [N]); with ``batch_indices`` the
        scores are also scattered into the pool vector (like pool_score_batch); with ``want_maps`` returns the dict of
        pseudo_annotation() (pseudo_label is the tensor given to mc_begin)."""
        m = measure_id(measure)
        torch = _torch()
        n, h, w, c = self._mc_shape
        dev = torch.device("cuda", self.device)
        scores = torch.empty(n, dtype=torch.float64, device=dev)
        conf = torch.empty((n, h, w), dtype=torch.float32, device=dev) if want_maps else None
        mask = torch.empty((n, h, w), dtype=torch.uint8, device=dev) if want_maps else None
        idx = None
        if batch_indices is not None:
            idx = np.ascontiguousarray(np.asarray(batch_indices, dtype=np.int64))
            if idx.shape != (n,):
                raise ValueError("batch_indices must have one entry per image (%d), got shape %s" % (n, idx.shape))
        self._sync_stream()
        self._check(self._lib.als_mc_finish(self._ctx, m, scores.data_ptr(), None if idx is None else idx.ctypes.data,
                                            None if conf is None else conf.data_ptr(),
                                            None if mask is None else mask.data_ptr(), float(threshold)))
        if want_maps:
            return {"pseudo_confidence": conf, "pseudo_mean_confidence": scores, "pseudo_label": self._mc_label,
                    "pseudo_mask": mask}
        return scores

    # -- device primitives --------------------------------------------------------------------
    def select_smallest(self, keys, ids, k: int):
        """k smallest (key, id) pairs of CUDA tensors keys f32[M], ids i64[M] -> (keys, ids) ascending."""
        torch = _torch()
        if not (keys.is_cuda and ids.is_cuda and keys.dtype == torch.float32 and ids.dtype == torch.int64):
            raise ValueError("keys must be a CUDA float32 tensor and ids a CUDA int64 tensor")
        keys, ids = keys.contiguous(), ids.contiguous()
        M = keys.numel()
        if ids.numel() != M:
            raise ValueError("keys and ids must have the same length")
        kk = max(0, min(int(k), M))
        ok = torch.empty(kk, dtype=torch.float32, device=keys.device)
        oi = torch.empty(kk, dtype=torch.int64, device=keys.device)
        self._check(self._lib.als_select_smallest(self._ctx, keys.data_ptr(), ids.data_ptr(), M, int(k),
                                                  ok.data_ptr(), oi.data_ptr(), self._stream_arg(keys.device)))
        return ok, oi

    def synth_logits(self, T: int, n0: int, n_imgs: int, H: int, W: int, C_: int, *, dtype: str = "float32",
                     seed: int = 20191013, out=None, squeeze_t: bool = True):
        """Device twin of oracle/synth.py: images n0..n0+n_imgs-1 of the synthetic pool."""
        torch = _torch()
        tdt = {"float32": torch.float32, "bfloat16": torch.bfloat16}[dtype]
        if out is None:
            out = torch.empty((T, n_imgs, H, W, C_), dtype=tdt, device=torch.device("cuda", self.device))
        self._check(self._lib.als_synth_logits(self._ctx, out.data_ptr(), _lib.ALS_F32 if dtype == "float32" else _lib.ALS_BF16,
                                               T, n0, n_imgs, H, W, C_, seed, 1 if T > 1 else 0, self._stream_arg(out.device)))
        return out[0] if (T == 1 and squeeze_t and out.dim() == 5) else out

    def flush_l2(self) -> None:
        self._check(self._lib.als_flush_l2(self._ctx, self._stream_arg()))

    def describe_launch(self, dtype: str, T: int, N: int, H: int, W: int, C_: int, measure: str) -> dict:
        name = C.create_string_buffer(128)
        g, b, s, st, tp = (C.c_int() for _ in range(5))
        self._check(self._lib.als_describe_launch(self._ctx, _lib.ALS_F32 if dtype == "float32" else _lib.ALS_BF16, T, N, H, W,
                                                  C_, measure_id(measure), name, C.byref(g), C.byref(b), C.byref(s),
                                                  C.byref(st), C.byref(tp)))
        return {"kernel": name.value.decode(), "grid": g.value, "block": b.value, "smem_bytes": s.value,
                "stages": st.value, "tile_pixels": tp.value}

    def describe_head_launch(self, T: int, measure: str) -> dict:
        """The fused-head kernel score_features would launch (after prepare_head), from the library's own plan."""
        name = C.create_string_buffer(128)
        g, b, s = (C.c_int() for _ in range(3))
        self._check(self._lib.als_describe_head_launch(self._ctx, int(T), measure_id(measure), name, C.byref(g), C.byref(b),
                                                       C.byref(s)))
        return {"kernel": name.value.decode(), "grid": g.value, "block": b.value, "smem_bytes": s.value,
                "stages": None, "tile_pixels": 512}


def head_mma_flops_per_pixel(num_classes: int) -> float:
    """Tensor-core FLOPs the fused head issues per OUTPUT pixel (zero padding of the narrow operands included):
    per tile of 128 input pixels, 3 products x 2 k-steps x (N0 + N1 + N2 + N3) x 128 x 8 MACs (csrc/head.cu)."""
    cb = (int(num_classes) + 3) // 4 * 4
    r16 = lambda v: (v + 15) // 16 * 16
    n1 = n2 = r16(2 * cb)
    n3 = r16(cb)
    n0 = max(r16(4 * cb), n1, cb + n2, cb + n3)
    return 2.0 * 128 * 8 * 6 * (n0 + n1 + n2 + n3) / 512.0


_USED_NAME = b"used_dltensor"   # module-level: PyCapsule_SetName keeps the pointer, not a copy

_default_scorers = {}


def default_scorer(device: Optional[int] = None) -> Scorer:
    torch = _torch()
    if device is None:
        device = torch.cuda.current_device() if (torch is not None and torch.cuda.is_available()) else 0
    sc = _default_scorers.get(device)
    if sc is None:
        sc = _default_scorers[device] = Scorer(device)
    return sc


def _slice_images(logits, sl: slice):
    return logits[sl] if logits.ndim == 4 else logits[:, sl]


def rank_confidence(logits, unlabelled, selection_size: int, measure: str = "entropy", *,
                    batch_size: Optional[int] = None, example_index=None, num_examples: Optional[int] = None,
                    dtype: Optional[str] = None, scorer: Optional[Scorer] = None, head_kernel=None):
    """Drop-in for the ``rank_confidence()`` closure (active_learning.py:682-715).

    With ``head_kernel`` (the `Final` layer's [3,3,C,16] kernel) the first argument is the layer's INPUT
    feature map [N,h,w,16] (or batches of it) and the classifier head runs fused inside the scoring kernel.

    logits            the pool's logits [N,H,W,C] / [T,N,H,W,C] (device or host), or an iterable of
                      ``(batch_logits, batch_indices)`` pairs as ``sess.run`` would hand them out (:697-698)
    unlabelled        index array into the example list (:705)
    selection_size    alparams["selection_size"] (:708); must be > 0 like at the call site (:779)
    Returns ``(low_conf_examples, unlabelled_confidence)`` (:715): the min(len(unlabelled),
    selection_size) lowest-confidence example ids (ascending confidence, ties by id) and the
    float32 confidences of all unlabelled examples (examples never visited keep 0.0, :685).
    """
    sc = scorer or default_scorer()
    measure_id(measure)   # unknown measure fails before any work, like graph construction does
    unlabelled = np.asarray(unlabelled, dtype=np.int64)
    if (head_kernel is None and hasattr(logits, "ndim") and hasattr(logits, "shape")
            and (not batch_size or int(batch_size) >= int(logits.shape[-4]))):
        # the pool is one tensor and goes through in one batch: the whole closure is one library call
        return sc.rank_pool(logits, unlabelled, selection_size, measure, example_index=example_index,
                            num_examples=num_examples, dtype=dtype)
    if hasattr(logits, "ndim") and hasattr(logits, "shape"):
        n = int(logits.shape[-4])
        if example_index is None:
            example_index = np.arange(n, dtype=np.int64)
        example_index = np.asarray(example_index, dtype=np.int64)
        if num_examples is None:
            num_examples = int(max(example_index.max(initial=-1), unlabelled.max(initial=-1))) + 1
        bs = n if not batch_size else int(batch_size)
        batches = ((_slice_images(logits, slice(i, i + bs)), example_index[i:i + bs]) for i in range(0, n, max(bs, 1)))
        if logits.ndim == 5 and bs < n and not getattr(logits, "is_cuda", False):
            # a [T, batch] slice of a host array is strided: make each batch dense before staging
            batches = ((np.ascontiguousarray(b) if isinstance(b, np.ndarray) else b.contiguous(), i) for b, i in batches)
        elif logits.ndim == 5 and bs < n:
            batches = ((b.contiguous(), i) for b, i in batches)
    else:
        batches = logits
        if num_examples is None:
            raise ValueError("num_examples is required when logits is an iterable of batches")
    if head_kernel is not None:
        sc.prepare_head(head_kernel)
    sc.pool_begin(int(num_examples))
    for batch_logits, batch_indices in batches:
        if head_kernel is not None:
            sc.pool_score_features_batch(batch_logits, batch_indices, measure)
        else:
            sc.pool_score_batch(batch_logits, batch_indices, measure, dtype=dtype)
    return sc.pool_select(unlabelled, selection_size)
