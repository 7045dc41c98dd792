"""ctypes binding over include/granne_b200.h."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.environ.get("GRANNE_B200_LIB") or os.path.join(_HERE, "libgranne_b200.so")  # env: tuning variants

ANGULAR, ANGULAR_INT, EMBEDDINGS = 0, 1, 2
QUERY_RAW_F32, QUERY_ELEMENT = 0, 1
_ELEMENT_TYPES = {"angular": ANGULAR, "angular_int": ANGULAR_INT, "embeddings": EMBEDDINGS}
STATS_PER_QUERY = 4
DEFAULT_MAX_SEARCH = 200   # py/src/lib.rs:14
DEFAULT_NUM_ELEMENTS = 10  # py/src/lib.rs:15


class GranneError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("granne_b200 error %d: %s" % (code, message))
        self.code = code


def library_path():
    return _LIB


_lib = None


def load_library():
    """Loads libgranne_b200.so; raises if it has not been built (python -m granne_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise ImportError("granne_b200: %s is missing — build it with `python -m granne_b200.build` "
                          "(there is no CPU fallback)" % _LIB)
    L = C.CDLL(_LIB)
    vp, sz, u64, u32, i32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint32, C.c_int
    sig = {
        "granne_b200_abi_version": (i32, []),
        "granne_b200_last_error": (C.c_char_p, []),
        "granne_b200_open": (i32, [vp, sz, i32, vp, sz, vp, sz, i32, C.POINTER(vp)]),
        "granne_b200_open_files": (i32, [C.c_char_p, i32, C.c_char_p, C.c_char_p, i32, C.POINTER(vp)]),
        "granne_b200_open_device_elements": (i32, [vp, sz, i32, vp, u64, u32, i32, C.POINTER(vp)]),
        "granne_b200_builder_new_device_elements": (i32, [vp, i32, vp, u64, u32, i32, C.POINTER(vp)]),
        "granne_b200_elements_from_raw_device": (i32, [i32, vp, u64, u32, i32, vp, vp]),
        "granne_b200_close": (None, [vp]),
        "granne_b200_len": (u64, [vp]),
        "granne_b200_num_layers": (u64, [vp]),
        "granne_b200_layer_len": (u64, [vp, u64]),
        "granne_b200_get_neighbors": (i32, [vp, u64, u64, vp, sz, C.POINTER(sz)]),
        "granne_b200_num_elements": (u64, [vp]),
        "granne_b200_dim": (u64, [vp]),
        "granne_b200_element_kind": (i32, [vp]),
        "granne_b200_get_element": (i32, [vp, u64, vp]),
        "granne_b200_search_batch": (i32, [vp, vp, sz, i32, u32, u32, vp, vp, vp, vp]),
        "granne_b200_search_batch_device": (i32, [vp, vp, sz, i32, u32, u32, vp, vp, vp, vp, vp]),
        "granne_b200_search_batch_device_gather": (i32, [vp, vp, sz, i32, u32, u32, vp, vp, vp, vp]),
        "granne_b200_stream_status": (i32, [vp]),
        "granne_b200_release_stream": (i32, [vp, vp]),
        "granne_b200_merge_topk_device": (i32, [i32, vp, vp, vp, sz, sz, u32, vp, vp, vp]),
        "granne_b200_inspect_index": (i32, [vp, sz, vp, vp, vp, vp, sz]),
        "granne_b200_decode_layer": (i32, [vp, sz, u64, vp, sz]),
        "granne_b200_compute_distances": (i32, [i32, vp, vp, u64, u32, i32, vp]),
        "granne_b200_write_index": (i32, [vp, vp, sz, C.POINTER(sz)]),
        "granne_b200_reencode_index": (i32, [vp, sz, vp, sz, C.POINTER(sz)]),
        "granne_b200_compute_order": (i32, [vp, vp, u64]),
        "granne_b200_order_from_trails": (i32, [vp, u32, vp, u64, vp]),
        "granne_b200_order_by_keys": (i32, [vp, sz, vp, u64, u32, vp]),
        "granne_b200_embedding_reorder_keys": (i32, [vp, sz, vp, sz, vp]),
        "granne_b200_apply_order": (i32, [vp, sz, i32, vp, sz, vp, u64, vp, sz, C.POINTER(sz), vp, sz, C.POINTER(sz)]),
        "granne_b200_build_config_default": (None, [vp]),
        "granne_b200_builder_new": (i32, [vp, i32, vp, sz, vp, sz, i32, C.POINTER(vp)]),
        "granne_b200_builder_append": (i32, [vp, vp, sz]),
        "granne_b200_builder_build": (i32, [vp, u64]),
        "granne_b200_builder_len": (u64, [vp]),
        "granne_b200_builder_num_layers": (u64, [vp]),
        "granne_b200_builder_layer_len": (u64, [vp, u64]),
        "granne_b200_builder_num_elements": (u64, [vp]),
        "granne_b200_builder_get_neighbors": (i32, [vp, u64, u64, vp, sz, C.POINTER(sz)]),
        "granne_b200_builder_write_index": (i32, [vp, vp, sz, C.POINTER(sz)]),
        "granne_b200_builder_get_index": (i32, [vp, C.POINTER(vp)]),
        "granne_b200_builder_free": (None, [vp]),
        "granne_b200_elements_from_raw": (i32, [i32, vp, u64, u32, i32, vp, sz, C.POINTER(sz)]),
        "granne_b200_launch_count": (u64, [vp]),
        "granne_b200_multi_open": (i32, [i32, vp, sz, i32, vp, vp, vp, vp, sz, vp, sz, C.POINTER(vp)]),
        "granne_b200_multi_close": (None, [vp]),
        "granne_b200_multi_len": (u64, [vp]),
        "granne_b200_multi_num_parts": (sz, [vp]),
        "granne_b200_multi_shard_base": (u64, [vp, sz]),
        "granne_b200_multi_search_batch": (i32, [vp, vp, sz, i32, u32, u32, vp, vp, vp]),
        "granne_b200_device_bytes": (u64, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise GranneError(rc, load_library().granne_b200_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _kind(element_type):
    if isinstance(element_type, str):
        try:
            return _ELEMENT_TYPES[element_type.lower()]
        except KeyError:
            raise ValueError("Invalid element type")  # py/src/lib.rs:208
    return int(element_type)


class Granne:
    """granne.Granne (py/src/lib.rs:149-344) on the GPU.

    Granne(index_path, element_type, elements_path, embeddings_path=None, words_path=None, device=0)
    element_type: "angular" | "angular_int" | "embeddings".  With `words_path` an "embeddings" index also takes string
    queries and reports elements as words (WordEmbeddingsGranne, py/src/variants/index.rs:41-139); the lookup and the
    ordered row sum happen on the host, the search itself on the GPU.
    """

    def __init__(self, index_path, element_type, elements_path, embeddings_path=None, words_path=None, device=0):
        L = load_library()
        h = C.c_void_p()
        kind = _kind(element_type)
        if kind == EMBEDDINGS and embeddings_path is None:
            raise ValueError("embeddings_path required for this element type!")
        _check(L.granne_b200_open_files(os.fsencode(index_path), kind, os.fsencode(elements_path),
                                        os.fsencode(embeddings_path) if embeddings_path else None, device,
                                        C.byref(h)))
        self._h = h
        self.device = device
        self._elements_src = ("path", elements_path)
        self._embeddings_src = ("path", embeddings_path) if embeddings_path else None
        self._words_path = words_path

    @classmethod
    def from_bytes(cls, index_bytes, element_type, elements_bytes, embeddings_bytes=None, device=0):
        """Granne::from_bytes (src/index/mod.rs:108-113) + Vectors::from_bytes / SumEmbeddings::from_bytes."""
        L = load_library()
        self = cls.__new__(cls)
        h = C.c_void_p()
        ib = np.frombuffer(index_bytes, dtype=np.uint8)
        eb = np.frombuffer(elements_bytes, dtype=np.uint8)
        mb = np.frombuffer(embeddings_bytes, dtype=np.uint8) if embeddings_bytes is not None else None
        _check(L.granne_b200_open(_ptr(ib), ib.size, _kind(element_type), _ptr(eb), eb.size,
                                  _ptr(mb) if mb is not None else None, mb.size if mb is not None else 0, device,
                                  C.byref(h)))
        self._h = h
        self.device = device
        self._elements_src = ("bytes", elements_bytes)
        self._embeddings_src = ("bytes", embeddings_bytes) if embeddings_bytes is not None else None
        return self

    @classmethod
    def from_device_elements(cls, index_bytes, element_type, elements, device=None):
        """Granne::from_bytes with a device-resident element container: `elements` is a contiguous CUDA torch tensor
        [n, dim] of ELEMENTS (normalised float32 for "angular", int8 for "angular_int") on the target device — e.g.
        the output of elements_from_raw_device.  Nothing bounces through the host (100M x 128 f32 = 51 GB)."""
        L = load_library()
        kind = _kind(element_type)
        _check_device_rows(kind, elements)
        device = elements.device.index if device is None else device
        if elements.device.index != device:
            raise ValueError("elements tensor lives on cuda:%d, not cuda:%d" % (elements.device.index, device))
        self = cls.__new__(cls)
        h = C.c_void_p()
        ib = np.frombuffer(index_bytes, dtype=np.uint8)
        _check(L.granne_b200_open_device_elements(_ptr(ib), ib.size, kind, C.c_void_p(elements.data_ptr()),
                                                  elements.shape[0], elements.shape[1], device, C.byref(h)))
        self._h = h
        self.device = device
        self._elements_src = None
        self._embeddings_src = None
        return self

    @staticmethod
    def _source_bytes(src):
        if src is None:
            raise GranneError(-1, "this handle does not own a host image of that data")
        kind, value = src
        if kind == "path":
            with open(value, "rb") as f:
                return f.read()
        return bytes(value)

    def index_bytes(self):
        """Index::write_index (src/index/io.rs:11-70) into memory, from the staged graph."""
        return _write_index(load_library().granne_b200_write_index, self._h,
                            [self.layer_len(l) for l in range(self.num_layers())]).tobytes()

    def elements_bytes(self):
        """The elements file image (u64 dim + rows, src/slice_vector/mod.rs:460-466; offsets + 3-byte ids for
        "embeddings")."""
        return self._source_bytes(getattr(self, "_elements_src", None))

    def save_index(self, path):
        """Granne.save_index(path) (py/src/lib.rs:325-329)."""
        with open(path, "wb") as f:
            f.write(self.index_bytes())

    def save_elements(self, path):
        """Granne.save_elements(path) (py/src/lib.rs:339-343)."""
        with open(path, "wb") as f:
            f.write(self.elements_bytes())

    # ---- reorder (src/index/reorder.rs) ----
    def compute_order(self):
        """Granne::compute_order (reorder.rs:126-174) on the GPU: order[i] == j moves element j to position i."""
        order = np.zeros(len(self), dtype=np.uint64)
        _check(load_library().granne_b200_compute_order(self._h, _ptr(order), order.size))
        return order

    def _reopen_reordered(self, order):
        kind = self.element_kind
        index_bytes = self.index_bytes()
        elements_bytes = self._source_bytes(getattr(self, "_elements_src", None))
        new_index, new_elements = apply_order(index_bytes, kind, elements_bytes, order)
        emb = getattr(self, "_embeddings_src", None)
        emb_bytes = self._source_bytes(emb) if emb is not None else None
        fresh = Granne.from_bytes(new_index, kind, new_elements, emb_bytes, device=self.device)
        fresh._words_path = getattr(self, "_words_path", None)
        self.close()
        self._words = None
        self.__dict__.update(fresh.__dict__)
        fresh._h = None
        return [int(x) for x in order]

    def reorder(self, show_progress=False):
        """Granne.reorder(show_progress) (py/src/lib.rs:311-315, reorder.rs:59-82): reorders index and elements in
        place for locality; returns the permutation (new -> old).  "embeddings" indexes are reordered by
        compute_keys_for_reordering like the reference's Python class (py/src/variants/index.rs:70-75)."""
        if self.element_kind == EMBEDDINGS:
            keys = compute_keys_for_reordering(self._source_bytes(self._elements_src),
                                               self._source_bytes(self._embeddings_src))
            return self.reorder_by_keys(keys)
        return self._reopen_reordered(self.compute_order())

    def reorder_by_keys(self, keys, show_progress=False):
        """Granne::reorder_by_keys (reorder.rs:89-124): layer-preserving sort by `keys` (uint64 [n] or [n, kw])."""
        return self._reopen_reordered(order_by_keys(self.index_bytes(), keys))

    def close(self):
        if getattr(self, "_h", None):
            load_library().granne_b200_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- Index trait ----
    def __len__(self):
        return int(load_library().granne_b200_len(self._h))

    def num_layers(self):
        return int(load_library().granne_b200_num_layers(self._h))

    def layer_len(self, layer):
        return int(load_library().granne_b200_layer_len(self._h, layer))

    def get_neighbors(self, idx, layer=None):
        if layer is None:
            layer = self.num_layers() - 1  # py/src/lib.rs:268-276
        out = np.empty(256, dtype=np.uint32)
        n = C.c_size_t()
        _check(load_library().granne_b200_get_neighbors(self._h, idx, layer, _ptr(out), out.size, C.byref(n)))
        return out[:n.value].tolist()

    # ---- ElementContainer ----
    @property
    def dim(self):
        return int(load_library().granne_b200_dim(self._h))

    @property
    def element_kind(self):
        return int(load_library().granne_b200_element_kind(self._h))

    def num_elements(self):
        return int(load_library().granne_b200_num_elements(self._h))

    def get_element(self, idx):
        out = np.empty(self.dim, dtype=np.int8 if self.element_kind == ANGULAR_INT else np.float32)
        _check(load_library().granne_b200_get_element(self._h, idx, _ptr(out)))
        return out

    # ---- search ----
    def search_batch(self, queries, max_search=DEFAULT_MAX_SEARCH, num_elements=DEFAULT_NUM_ELEMENTS,
                     already_element=False, with_stats=False):
        """nq independent Granne::search calls (src/index/mod.rs:140-150) in one launch; host buffers in and out.

        queries: (nq, dim) float32 raw vectors (normalised / quantised by the library like the reference's Python
        binding), or with already_element=True rows that already are elements (normalised f32; int8 for
        angular_int).  Returns (ids uint32 [nq,k] padded 0xFFFFFFFF, dists float32 [nq,k] padded +inf, counts)."""
        q = np.asarray(queries)
        fmt = QUERY_ELEMENT if already_element else QUERY_RAW_F32
        if q.dtype == np.int8:
            if self.element_kind != ANGULAR_INT:
                raise ValueError("int8 queries need an angular_int index")
            fmt = QUERY_ELEMENT
            q = np.ascontiguousarray(q)
        else:
            q = np.ascontiguousarray(q, dtype=np.float32)
            if fmt == QUERY_ELEMENT and self.element_kind == ANGULAR_INT:
                raise ValueError("angular_int elements are int8")
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError("queries must have shape (nq, %d)" % self.dim)
        nq, k = q.shape[0], int(num_elements)
        ids = np.empty((nq, k), dtype=np.uint32)
        dists = np.empty((nq, k), dtype=np.float32)
        counts = np.empty(nq, dtype=np.uint32)
        stats = np.zeros((nq, STATS_PER_QUERY), dtype=np.uint64) if with_stats else None
        _check(load_library().granne_b200_search_batch(self._h, _ptr(q), nq, fmt, int(max_search), k, _ptr(ids),
                                                       _ptr(dists), _ptr(counts),
                                                       _ptr(stats) if with_stats else None))
        if with_stats:
            return ids, dists, counts, stats
        return ids, dists, counts

    def _word_tables(self):
        """(WordDict, embeddings table, per-element id lists) for an "embeddings" index opened with words_path."""
        if getattr(self, "_words", None) is None:
            from . import words as W

            if self.element_kind != EMBEDDINGS or getattr(self, "_words_path", None) is None:
                raise ValueError("string queries need an \"embeddings\" index opened with words_path")
            self._words = (W.WordDict(self._words_path), W.read_dense_f32(self._source_bytes(self._embeddings_src)),
                           W.read_sum_terms(self._source_bytes(self._elements_src)))
        return self._words

    def get_internal_element(self, idx):
        """Granne.get_internal_element(idx) (py/src/lib.rs:255-257): the words of an "embeddings" element
        (py/src/variants/index.rs:133-138); the element itself for the vector types."""
        if self.element_kind != EMBEDDINGS:
            return self.get_element(idx)
        words, _, terms = self._word_tables()
        return words.get_words(terms[idx])

    def search(self, element, max_search=DEFAULT_MAX_SEARCH, num_elements=DEFAULT_NUM_ELEMENTS):
        """Granne.search(element, max_search=200, num_elements=10) -> [(id, distance)] (py/src/lib.rs:227-233); a str
        query is the sum of its words' embeddings (py/src/variants/index.rs:112-125)."""
        if isinstance(element, str):
            from . import words as W

            words, table, _ = self._word_tables()
            element = W.create_embedding(table, words.get_word_ids(element))
        ids, dists, counts = self.search_batch(np.asarray(element)[None, :], max_search, num_elements)
        return [(int(ids[0, i]), float(dists[0, i])) for i in range(int(counts[0]))]

    def search_batch_device(self, queries, max_search=DEFAULT_MAX_SEARCH, num_elements=DEFAULT_NUM_ELEMENTS,
                            already_element=False, out=None, stats=None, stream=None):
        """Device-resident variant: `queries` is a CUDA torch tensor on this index's device; returns CUDA tensors
        (ids int32 view of u32 bits, dists float32, counts int32).  Asynchronous on the current torch stream."""
        import torch

        q = self._check_device_queries(queries)
        fmt = QUERY_ELEMENT if already_element or q.dtype == torch.int8 else QUERY_RAW_F32
        nq, k = q.shape[0], int(num_elements)
        if out is not None:
            for t, dt, shape in zip(out, (torch.int32, torch.float32, torch.int32), ((nq, k), (nq, k), (nq,))):
                if (not t.is_cuda or t.device != q.device or t.dtype != dt or tuple(t.shape) != shape
                        or not t.is_contiguous()):
                    raise ValueError("out tensors must be contiguous (ids int32 [nq,k], dists float32 [nq,k], counts "
                                     "int32 [nq]) on the queries' device")
        if stats is not None and (not stats.is_cuda or stats.device != q.device or stats.dtype != torch.int64
                                  or tuple(stats.shape) != (nq, STATS_PER_QUERY) or not stats.is_contiguous()):
            raise ValueError("stats must be a contiguous int64 [nq, %d] tensor on the queries' device" % STATS_PER_QUERY)
        if out is None:
            dev = q.device
            out = (torch.empty((nq, k), dtype=torch.int32, device=dev),
                   torch.empty((nq, k), dtype=torch.float32, device=dev),
                   torch.empty((nq,), dtype=torch.int32, device=dev))
        ids, dists, counts = out
        s = stream if stream is not None else torch.cuda.current_stream(q.device).cuda_stream
        _check(load_library().granne_b200_search_batch_device(
            self._h, C.c_void_p(q.data_ptr()), nq, fmt, int(max_search), k, C.c_void_p(ids.data_ptr()),
            C.c_void_p(dists.data_ptr()), C.c_void_p(counts.data_ptr()),
            C.c_void_p(stats.data_ptr()) if stats is not None else None, C.c_void_p(s)))
        return ids, dists, counts

    def search_batch_device_gather(self, queries, gather, max_search=DEFAULT_MAX_SEARCH,
                                   num_elements=DEFAULT_NUM_ELEMENTS, already_element=False, counts=None, stream=None):
        """Like search_batch_device, but the result rows are stored by the kernels straight into every peer's
        gathered buffer described by `gather` (a PeerGather)."""
        import torch

        q = self._check_device_queries(queries)
        fmt = QUERY_ELEMENT if already_element or q.dtype == torch.int8 else QUERY_RAW_F32
        s = stream if stream is not None else torch.cuda.current_stream(q.device).cuda_stream
        _check(load_library().granne_b200_search_batch_device_gather(
            self._h, C.c_void_p(q.data_ptr()), q.shape[0], fmt, int(max_search), int(num_elements), C.byref(gather),
            C.c_void_p(counts.data_ptr()) if counts is not None else None, None, C.c_void_p(s)))

    def _check_device_queries(self, q):
        """Same checks as the host path (search_batch): a wrong dtype, width, device or a strided view would make the
        kernels read out of bounds (they index queries as qi * dim)."""
        import torch

        if not isinstance(q, torch.Tensor) or not q.is_cuda:
            raise ValueError("queries must be a CUDA torch tensor")
        if q.device.index != self.device:
            raise ValueError("queries live on cuda:%s, the index on cuda:%d" % (q.device.index, self.device))
        if q.dim() != 2 or q.shape[1] != self.dim:
            raise ValueError("queries must have shape (nq, %d)" % self.dim)
        if q.dtype == torch.int8:
            if self.element_kind != ANGULAR_INT:
                raise ValueError("int8 queries need an angular_int index")
        elif q.dtype != torch.float32:
            raise ValueError("queries must be float32 (or int8 elements for angular_int)")
        return q.contiguous()

    def stream_status(self):
        _check(load_library().granne_b200_stream_status(self._h))

    def release_stream(self, stream):
        """Returns the workspace the library keeps for a caller stream (a raw cudaStream_t or a torch stream)."""
        raw = getattr(stream, "cuda_stream", stream)
        _check(load_library().granne_b200_release_stream(self._h, C.c_void_p(raw)))

    def launch_count(self):
        return int(load_library().granne_b200_launch_count(self._h))

    def device_bytes(self):
        return int(load_library().granne_b200_device_bytes(self._h))


MODE_REPLICATED, MODE_RANGE_PARTITIONED = 0, 1


class MultiGranne:
    """Several GPUs of THIS process behind one handle (granne_b200_multi_*): no torch, no NCCL.

    MultiGranne.replicated(index_bytes, element_type, elements_bytes, devices)       one index on every device,
                                                                                     query batches are sliced
    MultiGranne.partitioned([(index_bytes, elements_bytes), ...], element_type, devices)
                                                                                     one independent index per shard
    search_batch returns GLOBAL uint64 ids (0xFFFF_FFFF_FFFF_FFFF padded), float32 distances, counts."""

    def __init__(self, mode, shards, element_type, devices, embeddings_bytes=None):
        L = load_library()
        kind = _kind(element_type)
        n = len(shards)
        self._keep = [(np.frombuffer(i, dtype=np.uint8), np.frombuffer(e, dtype=np.uint8)) for i, e in shards]
        ip = (C.c_void_p * n)(*[a.ctypes.data for a, _ in self._keep])
        il = (C.c_size_t * n)(*[a.size for a, _ in self._keep])
        ep = (C.c_void_p * n)(*[b.ctypes.data for _, b in self._keep])
        el = (C.c_size_t * n)(*[b.size for _, b in self._keep])
        dv = (C.c_int * len(devices))(*[int(d) for d in devices])
        mb = np.frombuffer(embeddings_bytes, dtype=np.uint8) if embeddings_bytes is not None else None
        h = C.c_void_p()
        _check(L.granne_b200_multi_open(mode, dv, len(devices), kind, ip, il, ep, el, n,
                                        _ptr(mb) if mb is not None else None, mb.size if mb is not None else 0,
                                        C.byref(h)))
        self._h = h
        self._keep = None
        self.kind = kind
        self.mode = mode

    @classmethod
    def replicated(cls, index_bytes, element_type, elements_bytes, devices, embeddings_bytes=None):
        return cls(MODE_REPLICATED, [(index_bytes, elements_bytes)], element_type, devices, embeddings_bytes)

    @classmethod
    def partitioned(cls, shards, element_type, devices, embeddings_bytes=None):
        return cls(MODE_RANGE_PARTITIONED, shards, element_type, devices, embeddings_bytes)

    def __len__(self):
        return int(load_library().granne_b200_multi_len(self._h))

    def num_parts(self):
        return int(load_library().granne_b200_multi_num_parts(self._h))

    def shard_base(self, s):
        return int(load_library().granne_b200_multi_shard_base(self._h, s))

    def search_batch(self, queries, max_search=DEFAULT_MAX_SEARCH, num_elements=DEFAULT_NUM_ELEMENTS,
                     already_element=False):
        q = np.asarray(queries)
        fmt = QUERY_ELEMENT if already_element else QUERY_RAW_F32
        if q.dtype == np.int8:
            fmt = QUERY_ELEMENT
            q = np.ascontiguousarray(q)
        else:
            q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim != 2:
            raise ValueError("queries must have shape (nq, dim)")
        nq, k = q.shape[0], int(num_elements)
        ids = np.empty((nq, k), dtype=np.uint64)
        dists = np.empty((nq, k), dtype=np.float32)
        counts = np.empty(nq, dtype=np.uint32)
        _check(load_library().granne_b200_multi_search_batch(self._h, _ptr(q), nq, fmt, int(max_search), k, _ptr(ids),
                                                             _ptr(dists), _ptr(counts)))
        return ids, dists, counts

    def close(self):
        if getattr(self, "_h", None):
            load_library().granne_b200_multi_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PeerGather(C.Structure):
    """granne_b200_peer_gather: peer-mapped result buffers for the fused multi-GPU gather."""
    _fields_ = [("n_peers", C.c_uint32), ("my_rank", C.c_uint32), ("row_offset", C.c_uint64), ("seq", C.c_uint32),
                ("reserved", C.c_uint32), ("ids", C.c_void_p * 8), ("dists", C.c_void_p * 8),
                ("flags", C.c_void_p * 8)]


class BuildConfig(C.Structure):
    """granne::BuildConfig (src/index/mod.rs:198-291) as the C ABI struct granne_b200_build_config."""
    _fields_ = [("layer_multiplier", C.c_float), ("expected_num_elements", C.c_int64),
                ("num_neighbors", C.c_uint32), ("max_search", C.c_uint32), ("reinsert_elements", C.c_int32),
                ("show_progress", C.c_int32)]


def elements_from_raw(element_type, raw, device=0):
    """Builds an elements file image from raw f32 rows on the GPU exactly like `Vector::from(Vec<f32>)` per row
    (normalise for "angular", quantise for "angular_int")."""
    L = load_library()
    raw = np.ascontiguousarray(raw, dtype=np.float32)
    n, dim = raw.shape
    kind = _kind(element_type)
    need = C.c_size_t()
    _check(L.granne_b200_elements_from_raw(kind, None, n, dim, device, None, 0, C.byref(need)))
    out = np.empty(need.value, dtype=np.uint8)
    _check(L.granne_b200_elements_from_raw(kind, _ptr(raw), n, dim, device, _ptr(out), out.size, C.byref(need)))
    return out


def _write_index(fn, handle, layer_lens):
    """One encode instead of two: the first call gets a buffer sized for the common case (<= 32 neighbours per node:
    1 count byte + 4 bytes per id is the writer's worst case); only an index with wider rows needs the size query."""
    guess = 1024 + sum(8 + (1 + n // 60) * 128 + n * (1 + 4 * 32) for n in layer_lens)
    out = np.empty(guess, dtype=np.uint8)
    need = C.c_size_t()
    rc = fn(handle, _ptr(out), out.size, C.byref(need))
    if rc != 0 and need.value > out.size:
        out = np.empty(need.value, dtype=np.uint8)
        rc = fn(handle, _ptr(out), out.size, C.byref(need))
    _check(rc)
    return out[:need.value]


def _check_device_rows(kind, t):
    import torch

    want = torch.int8 if kind == ANGULAR_INT else torch.float32
    if kind not in (ANGULAR, ANGULAR_INT):
        raise ValueError("device-resident elements: angular or angular_int")
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dim() != 2 or t.dtype != want or not t.is_contiguous():
        raise ValueError("elements must be a contiguous CUDA tensor [n, dim] of %s" % want)


def elements_from_raw_device(element_type, raw, out=None, stream=None):
    """`Vector::from(Vec<f32>)` per row, device to device: raw = CUDA float32 [n, dim]; returns (or fills `out`) the
    elements as a CUDA tensor [n, dim] (float32 normalised for "angular", int8 for "angular_int")."""
    import torch

    kind = _kind(element_type)
    if not raw.is_cuda or raw.dtype != torch.float32 or raw.dim() != 2 or not raw.is_contiguous():
        raise ValueError("raw must be a contiguous CUDA float32 tensor [n, dim]")
    n, dim = raw.shape
    if out is None:
        out = torch.empty((n, dim), dtype=torch.int8 if kind == ANGULAR_INT else torch.float32, device=raw.device)
    _check_device_rows(kind, out)
    if tuple(out.shape) != (n, dim) or out.device != raw.device:
        raise ValueError("out must match raw's shape and device")
    s = stream if stream is not None else torch.cuda.current_stream(raw.device).cuda_stream
    _check(load_library().granne_b200_elements_from_raw_device(kind, C.c_void_p(raw.data_ptr()), n, dim,
                                                               raw.device.index, C.c_void_p(out.data_ptr()),
                                                               C.c_void_p(s)))
    return out


class GranneBuilder:
    """granne.GranneBuilder (py/src/lib.rs:346-579; src/index/mod.rs:295-531) on the GPU.

    GranneBuilder(element_type, elements_bytes, embeddings_bytes=None, num_neighbors=30, max_search=200,
                  layer_multiplier=15.0, reinsert_elements=True, expected_num_elements=None, device=0)
    `elements_bytes` is an elements file image (see elements_from_raw / Granne.save_elements).  num_neighbors may be
    anything the file format can hold (1..255: a list's length is one byte); show_progress is accepted and ignored."""

    def __init__(self, element_type, elements_bytes, embeddings_bytes=None, num_neighbors=30, max_search=200,
                 layer_multiplier=15.0, reinsert_elements=True, expected_num_elements=None, show_progress=False,
                 device=0):
        L = load_library()
        cfg = BuildConfig()
        L.granne_b200_build_config_default(C.byref(cfg))
        cfg.num_neighbors = num_neighbors
        cfg.max_search = max_search
        cfg.layer_multiplier = layer_multiplier
        cfg.reinsert_elements = int(bool(reinsert_elements))
        cfg.expected_num_elements = -1 if expected_num_elements is None else int(expected_num_elements)
        cfg.show_progress = int(bool(show_progress))
        eb = np.frombuffer(elements_bytes, dtype=np.uint8)
        mb = np.frombuffer(embeddings_bytes, dtype=np.uint8) if embeddings_bytes is not None else None
        h = C.c_void_p()
        _check(L.granne_b200_builder_new(C.byref(cfg), _kind(element_type), _ptr(eb), eb.size,
                                         _ptr(mb) if mb is not None else None, mb.size if mb is not None else 0,
                                         device, C.byref(h)))
        self._h = h
        self.device = device
        self._element_type = element_type
        self._elements_bytes = elements_bytes
        self._embeddings_bytes = embeddings_bytes
        self._pending = []

    @classmethod
    def from_device_elements(cls, element_type, elements, num_neighbors=30, max_search=200, layer_multiplier=15.0,
                             reinsert_elements=True, expected_num_elements=None, device=None):
        """GranneBuilder::new(config, elements) over a device-resident container (see Granne.from_device_elements)."""
        L = load_library()
        kind = _kind(element_type)
        _check_device_rows(kind, elements)
        device = elements.device.index if device is None else device
        cfg = BuildConfig()
        L.granne_b200_build_config_default(C.byref(cfg))
        cfg.num_neighbors = num_neighbors
        cfg.max_search = max_search
        cfg.layer_multiplier = layer_multiplier
        cfg.reinsert_elements = int(bool(reinsert_elements))
        cfg.expected_num_elements = -1 if expected_num_elements is None else int(expected_num_elements)
        self = cls.__new__(cls)
        h = C.c_void_p()
        _check(L.granne_b200_builder_new_device_elements(C.byref(cfg), kind, C.c_void_p(elements.data_ptr()),
                                                         elements.shape[0], elements.shape[1], device, C.byref(h)))
        self._h = h
        self.device = device
        self._element_type = element_type
        self._elements_bytes = None
        self._embeddings_bytes = None
        self._pending = []
        return self

    def append(self, element):
        """GranneBuilder.append(element) (py/src/lib.rs:474-476; py/src/variants/builder.rs:6-21): the raw vector
        becomes an element (`Vector::from`) and is pushed; it is indexed by the next build().  Rows are buffered and
        handed to the library in one batch."""
        if _kind(self._element_type) == EMBEDDINGS:
            # SumEmbeddings::push (src/elements/embeddings/mod.rs:97-100): an element is a list of embedding ids
            # (the reference's WordEmbeddingsBuilder maps a string to ids first, py/src/variants/builder.rs:84-93)
            if isinstance(element, str):
                raise ValueError("pass embedding ids (WordDict.get_word_ids(text)) — this builder holds no word list")
            self._pending.append([int(t) for t in element])
            return
        self._pending.append(np.asarray(element, dtype=np.float32))

    def _flush(self):
        if not self._pending:
            return
        if _kind(self._element_type) == EMBEDDINGS:
            from . import words as W

            fresh, self._pending = self._pending, []
            image = np.frombuffer(W.write_sum_terms(fresh), dtype=np.uint8)
            _check(load_library().granne_b200_builder_append(self._h, _ptr(image), image.size))
            if self._elements_bytes is not None:
                self._elements_bytes = W.write_sum_terms(W.read_sum_terms(bytes(self._elements_bytes)) + fresh)
            return
        raw = np.stack(self._pending)
        self._pending = []
        image = elements_from_raw(self._element_type, raw, self.device)
        _check(load_library().granne_b200_builder_append(self._h, _ptr(image), image.size))
        if self._elements_bytes is not None:
            old = np.frombuffer(self._elements_bytes, dtype=np.uint8)
            self._elements_bytes = np.concatenate([old, image[8:]])  # same u64 width prefix, more rows

    def build(self, num_elements=0):
        """Builder::build() / build_partial(num_elements)."""
        self._flush()
        _check(load_library().granne_b200_builder_build(self._h, int(num_elements)))

    def __len__(self):
        return int(load_library().granne_b200_builder_len(self._h))

    def num_layers(self):
        return int(load_library().granne_b200_builder_num_layers(self._h))

    def layer_len(self, layer):
        return int(load_library().granne_b200_builder_layer_len(self._h, layer))

    def index_bytes(self):
        """Index::write_index into memory: a granne index file image (uint8 array)."""
        return _write_index(load_library().granne_b200_builder_write_index, self._h,
                            [self.layer_len(l) for l in range(self.num_layers())])

    def save_index(self, path):
        """GranneBuilder.save_index(path) (py/src/lib.rs)."""
        with open(path, "wb") as f:
            f.write(self.index_bytes().tobytes())

    def get_index(self):
        """GranneBuilder::get_index: a searchable Granne sharing the staged elements."""
        self._flush()
        h = C.c_void_p()
        _check(load_library().granne_b200_builder_get_index(self._h, C.byref(h)))
        g = Granne.__new__(Granne)
        g._h = h
        g.device = self.device
        g._elements_src = ("bytes", self._elements_bytes) if self._elements_bytes is not None else None
        g._embeddings_src = ("bytes", self._embeddings_bytes) if self._embeddings_bytes is not None else None
        return g

    def save_elements(self, path):
        """GranneBuilder.save_elements(path) (py/src/lib.rs:518-522)."""
        self._flush()
        if self._elements_bytes is None:
            raise GranneError(-1, "this builder was made from device-resident elements: no host image to save")
        with open(path, "wb") as f:
            f.write(bytes(self._elements_bytes))

    def get_neighbors(self, idx, layer=None):
        """GranneBuilder.get_neighbors(idx, layer=last) (py/src/lib.rs:544-552)."""
        if layer is None:
            layer = self.num_layers() - 1
        out = np.empty(256, dtype=np.uint32)
        n = C.c_size_t()
        _check(load_library().granne_b200_builder_get_neighbors(self._h, idx, layer, _ptr(out), out.size,
                                                                C.byref(n)))
        return out[:n.value].tolist()

    def num_elements(self):
        """GranneBuilder.num_elements() (py/src/lib.rs:559-561): elements held, indexed or not."""
        self._flush()
        return int(load_library().granne_b200_builder_num_elements(self._h))

    def close(self):
        if getattr(self, "_h", None):
            load_library().granne_b200_builder_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def inspect_index(index_bytes):
    """Host-only: [(layer_len, max_degree, row_width)] per layer of a granne index image."""
    L = load_library()
    ib = np.frombuffer(index_bytes, dtype=np.uint8)
    nl = C.c_uint64()
    lens = np.zeros(64, dtype=np.uint64)
    degs = np.zeros(64, dtype=np.uint32)
    widths = np.zeros(64, dtype=np.uint32)
    _check(L.granne_b200_inspect_index(_ptr(ib), ib.size, C.byref(nl), _ptr(lens), _ptr(degs), _ptr(widths), 64))
    return [(int(lens[i]), int(degs[i]), int(widths[i])) for i in range(nl.value)]


def reencode_index(index_bytes):
    """Host-only: decode an index image and write it again with the library's writer (Index::write_index)."""
    L = load_library()
    ib = np.frombuffer(index_bytes, dtype=np.uint8)
    need = C.c_size_t()
    _check(L.granne_b200_reencode_index(_ptr(ib), ib.size, None, 0, C.byref(need)))
    out = np.empty(need.value, dtype=np.uint8)
    _check(L.granne_b200_reencode_index(_ptr(ib), ib.size, _ptr(out), out.size, C.byref(need)))
    return out.tobytes()


def compute_distances(element_type, a, b, device=0):
    """n pairwise distances Vector::from(a_i).dist(&Vector::from(b_i)) on the GPU; a, b: (n, dim) raw float32."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    if a.ndim != 2 or a.shape != b.shape:
        raise ValueError("a and b must both have shape (n, dim)")
    out = np.empty(a.shape[0], dtype=np.float32)
    _check(load_library().granne_b200_compute_distances(_kind(element_type), _ptr(a), _ptr(b), a.shape[0], a.shape[1],
                                                        device, _ptr(out)))
    return out


def compute_distance(element_type, a, b, device=0):
    """granne.compute_distance(element_type, a, b) (py/src/lib.rs:71-89)."""
    if element_type not in ("angular", "angular_int"):
        raise ValueError("Unsupported element type")
    return float(compute_distances(element_type, np.asarray(a, dtype=np.float32)[None, :],
                                   np.asarray(b, dtype=np.float32)[None, :], device)[0])


def order_from_trails(layer_lens, trails):
    """Host-only: the sort half of compute_order for trails (uint32 [n, 8]) computed elsewhere."""
    lens = np.ascontiguousarray(layer_lens, dtype=np.uint64)
    t = np.ascontiguousarray(trails, dtype=np.uint32)
    order = np.zeros(t.shape[0], dtype=np.uint64)
    _check(load_library().granne_b200_order_from_trails(_ptr(lens), lens.size, _ptr(t), t.shape[0], _ptr(order)))
    return order


def order_by_keys(index_bytes, keys):
    """Host-only: the layer-preserving sort of Granne::reorder_by_keys (reorder.rs:96-108)."""
    ib = np.frombuffer(index_bytes, dtype=np.uint8)
    k = np.ascontiguousarray(keys, dtype=np.uint64)
    k = k.reshape(k.shape[0], -1)
    order = np.zeros(k.shape[0], dtype=np.uint64)
    _check(load_library().granne_b200_order_by_keys(_ptr(ib), ib.size, _ptr(k), k.shape[0], k.shape[1], _ptr(order)))
    return order


def compute_keys_for_reordering(elements_bytes, embeddings_bytes):
    """Host-only: embeddings::compute_keys_for_reordering (embeddings/reorder.rs:31-58) -> uint64 [n, 8]."""
    eb = np.frombuffer(elements_bytes, dtype=np.uint8)
    mb = np.frombuffer(embeddings_bytes, dtype=np.uint8)
    n = int(np.frombuffer(eb[:8].tobytes(), dtype=np.uint64)[0]) if eb.size >= 8 else -1
    if n < 0 or 8 + (n + 1) * 5 > eb.size:  # u64 count | (count + 1) five-byte offsets | ids
        raise GranneError(-2, "embeddings elements file: offset table exceeds the file")
    keys = np.zeros((n, 8), dtype=np.uint64)
    _check(load_library().granne_b200_embedding_reorder_keys(_ptr(eb), eb.size, _ptr(mb), mb.size, _ptr(keys)))
    return keys


def apply_order(index_bytes, element_type, elements_bytes, order):
    """Host-only: reorder_layers + Permutable::permute -> (index file image, elements file image) as bytes."""
    L = load_library()
    kind = element_type if isinstance(element_type, int) else _kind(element_type)
    ib = np.frombuffer(index_bytes, dtype=np.uint8)
    eb = np.frombuffer(elements_bytes, dtype=np.uint8)
    o = np.ascontiguousarray(order, dtype=np.uint64)
    n_i, n_e = C.c_size_t(), C.c_size_t()
    _check(L.granne_b200_apply_order(_ptr(ib), ib.size, kind, _ptr(eb), eb.size, _ptr(o), o.size, None, 0,
                                     C.byref(n_i), None, 0, C.byref(n_e)))
    oi, oe = np.empty(n_i.value, dtype=np.uint8), np.empty(n_e.value, dtype=np.uint8)
    _check(L.granne_b200_apply_order(_ptr(ib), ib.size, kind, _ptr(eb), eb.size, _ptr(o), o.size, _ptr(oi), oi.size,
                                     C.byref(n_i), _ptr(oe), oe.size, C.byref(n_e)))
    return oi.tobytes(), oe.tobytes()


def decode_layer(index_bytes, layer):
    """Host-only: the fixed-width u32 rows (padded with 0xFFFFFFFF) the loader stages in HBM for `layer`."""
    shape = inspect_index(index_bytes)
    n, _, w = shape[layer]
    ib = np.frombuffer(index_bytes, dtype=np.uint8)
    rows = np.empty((n, w), dtype=np.uint32)
    _check(load_library().granne_b200_decode_layer(_ptr(ib), ib.size, layer, _ptr(rows), rows.size))
    return rows


def merge_topk_device(device, part_ids, part_dists, part_base, out_ids=None, out_dists=None, stream=None):
    """granne_b200_merge_topk_device over CUDA torch tensors: part_ids int32 [P, nq, k] (u32 bits), part_dists
    float32 [P, nq, k], part_base list of P global id offsets.  Returns (int64 [nq,k] global ids, float32 dists)."""
    import torch

    P, nq, k = part_ids.shape
    if out_ids is None:
        out_ids = torch.empty((nq, k), dtype=torch.int64, device=part_ids.device)
        out_dists = torch.empty((nq, k), dtype=torch.float32, device=part_ids.device)
    base = np.ascontiguousarray(part_base, dtype=np.uint64)
    s = stream if stream is not None else torch.cuda.current_stream(part_ids.device).cuda_stream
    _check(load_library().granne_b200_merge_topk_device(
        device, C.c_void_p(part_ids.data_ptr()), C.c_void_p(part_dists.data_ptr()), _ptr(base), P, nq, k,
        C.c_void_p(out_ids.data_ptr()), C.c_void_p(out_dists.data_ptr()), C.c_void_p(s)))
    return out_ids, out_dists
