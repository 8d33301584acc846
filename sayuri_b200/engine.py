"""ctypes host mirror of the C ABI (include/sayuri_b200.h) — what tests/ and bench.py drive.

The class mirrors the reference plugin interface for this path,
    NetworkForwardPipe / BatchForwardPipe   /root/reference/src/neural/network_basic.h:132-161,
                                            /root/reference/src/neural/batch_forward_pipe.h:13-62
    CudaForwardPipe                         /root/reference/src/neural/cuda/cuda_forward_pipe.h:20-36
(Initialize / Construct / Forward / BatchForward / Release / Destroy / Valid / GetNumWorkers), with the
same argument meaning and error behaviour (errors raise RuntimeError, as the C++ shim throws
std::runtime_error).  There is no CPU fallback: if libsayuri_b200.so is missing, or no B200 is present,
construction raises.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# SAYURI_B200_LIB selects an alternative build of the same library (kernel tuning experiments only)
LIB_PATH = os.environ.get("SAYURI_B200_LIB") or os.path.join(HERE, "libsayuri_b200.so")

MAX_INTERSECTIONS = 361
INPUT_CHANNELS = 43
PLANE_FLOATS = INPUT_CHANNELS * MAX_INTERSECTIONS

BLOCK_RESIDUAL = 0
BLOCK_BOTTLENECK = 1
BLOCK_NESTED_BOTTLENECK = 2
BLOCK_MIXER = 3
POLICY_HEAD_NORMAL = 0
POLICY_HEAD_REPLK = 1

PRECISION_FP32_SPLIT = 0
PRECISION_FP16 = 1
PRECISION_SIMT_DEBUG = 2

_F = ctypes.POINTER(ctypes.c_float)
_I = ctypes.POINTER(ctypes.c_int)


class SbNetDesc(ctypes.Structure):
    _fields_ = [("version", ctypes.c_int), ("input_channels", ctypes.c_int), ("blocks", ctypes.c_int),
                ("channels", ctypes.c_int), ("policy_channels", ctypes.c_int), ("value_channels", ctypes.c_int),
                ("activation", ctypes.c_int), ("se_sizes", _I), ("block_types", _I), ("inner_channels", _I),
                ("dw_kernels", _I), ("policy_head_type", ctypes.c_int), ("policy_dw_kernel", ctypes.c_int)]


class SbTensor(ctypes.Structure):
    _fields_ = [("data", _F), ("count", ctypes.c_longlong)]


class SbWeights(ctypes.Structure):
    _fields_ = [("tensors", ctypes.POINTER(SbTensor)), ("n_tensors", ctypes.c_int)]


class SbOutput(ctypes.Structure):
    _fields_ = [("probabilities", ctypes.c_float * MAX_INTERSECTIONS),
                ("ownership", ctypes.c_float * MAX_INTERSECTIONS),
                ("pass_probability", ctypes.c_float), ("wdl", ctypes.c_float * 3),
                ("stm_winrate", ctypes.c_float), ("final_score", ctypes.c_float), ("q_error", ctypes.c_float),
                ("score_error", ctypes.c_float), ("board_size", ctypes.c_int), ("offset", ctypes.c_int),
                ("fp16", ctypes.c_int)]


PACKED_WORDS = 12
PACKED_RAW = 1


class SbPackedPosition(ctypes.Structure):
    """sb_packed_position: exact 2.2 KB encoding of one InputData (include/sayuri_b200.h)."""
    _fields_ = [("bits", (ctypes.c_uint32 * PACKED_WORDS) * INPUT_CHANNELS), ("scale", ctypes.c_float * INPUT_CHANNELS),
                ("board_size", ctypes.c_int32), ("offset", ctypes.c_int32), ("flags", ctypes.c_int32),
                ("reserved", ctypes.c_int32)]


class SbSymm8Result(ctypes.Structure):
    _fields_ = [("probabilities", ctypes.c_float * MAX_INTERSECTIONS), ("ownership", ctypes.c_float * MAX_INTERSECTIONS),
                ("pass_probability", ctypes.c_float), ("wdl", ctypes.c_float * 3), ("wdl_winrate", ctypes.c_float),
                ("stm_winrate", ctypes.c_float), ("final_score", ctypes.c_float), ("q_error", ctypes.c_float),
                ("score_error", ctypes.c_float), ("board_size", ctypes.c_int)]


class SbEvalTicket(ctypes.Structure):
    _fields_ = [("owner", ctypes.c_void_p), ("batch", ctypes.c_void_p), ("seq", ctypes.c_uint32), ("index", ctypes.c_int32),
                ("lane", ctypes.c_int32), ("flags", ctypes.c_int32), ("board_size", ctypes.c_int32), ("offset", ctypes.c_int32)]


OUTPUT_DTYPE = np.dtype([("probabilities", np.float32, MAX_INTERSECTIONS), ("ownership", np.float32, MAX_INTERSECTIONS),
                         ("pass_probability", np.float32), ("wdl", np.float32, 3), ("stm_winrate", np.float32),
                         ("final_score", np.float32), ("q_error", np.float32), ("score_error", np.float32),
                         ("board_size", np.int32), ("offset", np.int32), ("fp16", np.int32)])
assert OUTPUT_DTYPE.itemsize == ctypes.sizeof(SbOutput)

# Every symbol include/sayuri_b200.h declares (checked by tests/test_abi.py).
ABI_SYMBOLS = [
    "sb_create", "sb_create_from_file", "sb_reconfigure", "sb_reload_weights", "sb_reload_weights_from_file",
    "sb_destroy", "sb_last_error", "sb_num_gpus", "sb_num_slots", "sb_max_batch", "sb_board_size", "sb_get_net_desc",
    "sb_forward_batch", "sb_submit", "sb_wait", "sb_host_alloc", "sb_host_free", "sb_weights_blob",
    "sb_weights_export", "sb_weights_import",
    "sb_weights_checksum", "sb_time_forward", "sb_launch_count", "sb_debug_read_trunk", "sb_conv_stats", "sb_set_option",
    "sb_get_block_desc", "sb_get_dw_desc", "sb_host_net_load", "sb_host_net_free", "sb_host_net_desc", "sb_host_net_tensor",
    "sb_pack_position", "sb_unpack_position", "sb_eval", "sb_batcher_config", "sb_batcher_stats", "sb_eval_throughput",
    "sb_eval_submit", "sb_eval_poll", "sb_eval_wait", "sb_eval_throughput_async", "sb_weights_broadcast", "sb_weights_stats",
    "sb_eval_symm8",
]

_lib = None


def load_library():
    """Load libsayuri_b200.so (built in-tree by __graft_entry__.build / sayuri_b200/csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("sayuri_b200: %s is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    if os.environ.get("SAYURI_B200_LIB"):
        # an older / experimental build for same-box A/B timing may lack newer entry points: give those a stub that raises
        def _missing(*_a, **_k):
            raise RuntimeError("this build of libsayuri_b200.so lacks the entry point")
        for name in ABI_SYMBOLS:
            if not hasattr(lib, name):
                setattr(lib, name, type("Stub", (), {"__call__": staticmethod(_missing), "argtypes": None, "restype": None})())
    vp = ctypes.c_void_p
    lib.sb_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(SbNetDesc), ctypes.POINTER(SbWeights), _I, ctypes.c_int,
                              ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.sb_create_from_file.argtypes = [ctypes.POINTER(vp), ctypes.c_char_p, _I, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.sb_reconfigure.argtypes = [vp, ctypes.c_int, ctypes.c_int]
    lib.sb_reload_weights.argtypes = [vp, ctypes.POINTER(SbNetDesc), ctypes.POINTER(SbWeights)]
    lib.sb_reload_weights_from_file.argtypes = [vp, ctypes.c_char_p]
    lib.sb_destroy.argtypes = [vp]
    lib.sb_destroy.restype = None
    lib.sb_last_error.argtypes = [vp]
    lib.sb_last_error.restype = ctypes.c_char_p
    for name in ("sb_num_gpus", "sb_num_slots", "sb_max_batch", "sb_board_size"):
        getattr(lib, name).argtypes = [vp]
    lib.sb_get_net_desc.argtypes = [vp, ctypes.POINTER(SbNetDesc), _I, ctypes.c_int]
    lib.sb_get_block_desc.argtypes = [vp, _I, _I, ctypes.c_int]
    lib.sb_get_dw_desc.argtypes = [vp, _I, ctypes.c_int, _I, _I]
    lib.sb_forward_batch.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_F), _I, _I, ctypes.c_void_p]
    lib.sb_submit.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong, _I, _I]
    lib.sb_wait.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.sb_host_alloc.argtypes = [ctypes.c_size_t]
    lib.sb_host_alloc.restype = ctypes.c_void_p
    lib.sb_host_free.argtypes = [ctypes.c_void_p]
    lib.sb_host_free.restype = None
    lib.sb_weights_blob.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]
    lib.sb_weights_export.argtypes = [vp, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
    lib.sb_weights_import.argtypes = [vp, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
    lib.sb_weights_checksum.argtypes = [vp, ctypes.c_int]
    lib.sb_weights_checksum.restype = ctypes.c_uint64
    lib.sb_time_forward.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _F, _F, _I]
    lib.sb_launch_count.argtypes = [vp]
    lib.sb_launch_count.restype = ctypes.c_longlong
    lib.sb_debug_read_trunk.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _F]
    lib.sb_conv_stats.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
    lib.sb_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_int]
    lib.sb_host_net_load.argtypes = [ctypes.POINTER(vp), ctypes.c_char_p]
    lib.sb_host_net_free.argtypes = [vp]
    lib.sb_host_net_free.restype = None
    lib.sb_host_net_desc.argtypes = [vp, ctypes.POINTER(SbNetDesc), _I, _I, _I, _I, ctypes.c_int]
    lib.sb_host_net_tensor.argtypes = [vp, ctypes.c_int, ctypes.POINTER(_F)]
    lib.sb_host_net_tensor.restype = ctypes.c_longlong
    lib.sb_pack_position.argtypes = [_F, ctypes.c_int, ctypes.c_int, ctypes.POINTER(SbPackedPosition)]
    lib.sb_unpack_position.argtypes = [ctypes.POINTER(SbPackedPosition), _F]
    lib.sb_eval.argtypes = [vp, _F, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.sb_batcher_config.argtypes = [vp, ctypes.c_int, ctypes.c_int]
    lib.sb_batcher_stats.argtypes = [vp, ctypes.POINTER(ctypes.c_longlong)]
    lib.sb_eval_throughput.argtypes = [vp, _F, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double]
    lib.sb_eval_throughput.restype = ctypes.c_double
    lib.sb_eval_throughput_async.argtypes = [vp, _F, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double]
    lib.sb_eval_throughput_async.restype = ctypes.c_double
    lib.sb_eval_submit.argtypes = [vp, _F, ctypes.c_int, ctypes.c_int, ctypes.POINTER(SbEvalTicket)]
    lib.sb_eval_poll.argtypes = [vp, ctypes.POINTER(SbEvalTicket), ctypes.c_void_p]
    lib.sb_eval_wait.argtypes = [vp, ctypes.POINTER(SbEvalTicket), ctypes.c_void_p]
    lib.sb_eval_symm8.argtypes = [vp, _F, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.POINTER(SbSymm8Result)]
    lib.sb_weights_broadcast.argtypes = [vp]
    lib.sb_weights_stats.argtypes = [vp, ctypes.POINTER(ctypes.c_longlong)]
    _lib = lib
    return lib


def load_weights_file(path):
    """sb_host_net_*: the engine's own weight-file reader (host only).  Returns (desc dict, [tensor, ...]) with the
    tensors in sb_weights order, BN folded — directly usable with B200ForwardPipe.initialize_from_tensors."""
    lib = load_library()
    h = ctypes.c_void_p(None)
    if lib.sb_host_net_load(ctypes.byref(h), str(path).encode()):
        raise RuntimeError(lib.sb_last_error(None).decode())
    try:
        d = SbNetDesc()
        se, ty, inner, dwk = ((ctypes.c_int * 1024)() for _ in range(4))
        if lib.sb_host_net_desc(h, ctypes.byref(d), se, ty, inner, dwk, 1024):
            raise RuntimeError("sb_host_net_desc failed")
        nb = d.blocks
        desc = dict(version=d.version, blocks=nb, channels=d.channels, P=d.policy_channels, V=d.value_channels,
                    activation=d.activation, se_sizes=[se[i] for i in range(nb)], block_types=[ty[i] for i in range(nb)],
                    inner_channels=[inner[i] for i in range(nb)], dw_kernels=[dwk[i] for i in range(nb)],
                    policy_head_type=d.policy_head_type, policy_dw_kernel=d.policy_dw_kernel)
        tensors = []
        idx = 0
        while True:
            p = _F()
            n = lib.sb_host_net_tensor(h, idx, ctypes.byref(p))
            if n < 0:
                break
            tensors.append(np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, dtype=np.float32))
            idx += 1
        return desc, tensors
    finally:
        lib.sb_host_net_free(h)


def pack_position(planes, board_size, offset=0):
    """sb_pack_position: (record, ok).  ok == False means the planes are not two-valued per plane (raw fallback)."""
    a = np.ascontiguousarray(planes, dtype=np.float32).ravel()
    if a.size < INPUT_CHANNELS * board_size * board_size:
        raise ValueError("planes array smaller than 43*bs*bs")
    rec = SbPackedPosition()
    ok = load_library().sb_pack_position(a.ctypes.data_as(_F), board_size, offset, ctypes.byref(rec))
    return rec, bool(ok)


def unpack_position(rec):
    """sb_unpack_position: fp32 planes [43, bs*bs] of a packed record (host inverse, for tests)."""
    n = rec.board_size * rec.board_size
    out = np.zeros(INPUT_CHANNELS * n, dtype=np.float32)
    if not load_library().sb_unpack_position(ctypes.byref(rec), out.ctypes.data_as(_F)):
        raise ValueError("record is flagged raw")
    return out.reshape(INPUT_CHANNELS, n)


class PinnedArray:
    """float32 numpy view over pinned host memory from sb_host_alloc (DMA'd in place by sb_submit)."""

    def __init__(self, shape, dtype=np.float32):
        lib = load_library()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = lib.sb_host_alloc(self.nbytes)
        if not self.ptr:
            raise RuntimeError("sb_host_alloc(%d) failed" % self.nbytes)
        buf = (ctypes.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            load_library().sb_host_free(self.ptr)
            self.ptr = None


class B200ForwardPipe:
    """Drop-in for the backend behind NetworkForwardPipe (see module docstring)."""

    def __init__(self):
        self._lib = load_library()
        self._h = ctypes.c_void_p(None)

    # ---- NetworkForwardPipe::Initialize / CudaForwardPipe::Construct ----------------------------
    def initialize(self, weights_path, board_size=19, batch_size=256, gpus=None, precision=PRECISION_FP32_SPLIT):
        self.destroy()
        ids = list(gpus) if gpus else []
        arr = (ctypes.c_int * max(len(ids), 1))(*ids)
        rc = self._lib.sb_create_from_file(ctypes.byref(self._h), str(weights_path).encode(), arr, len(ids), board_size,
                                           batch_size, precision)
        if rc:
            self._h = ctypes.c_void_p(None)
            raise RuntimeError(self._lib.sb_last_error(None).decode())
        return self

    def initialize_from_tensors(self, desc, tensors, board_size=19, batch_size=256, gpus=None,
                                precision=PRECISION_FP32_SPLIT):
        """desc: dict(version, blocks, channels, P, V, activation, se_sizes); tensors: list of float32 arrays in
        loader order with BN folded, or None to leave the weight blob to a later broadcast (weights_blob())."""
        self.destroy()
        nb = max(len(desc["se_sizes"]), 1)
        se = (ctypes.c_int * nb)(*desc["se_sizes"])
        types = (ctypes.c_int * nb)(*desc.get("block_types", [BLOCK_RESIDUAL] * len(desc["se_sizes"])))
        inner = (ctypes.c_int * nb)(*desc.get("inner_channels", [0] * len(desc["se_sizes"])))
        dwk = (ctypes.c_int * nb)(*desc.get("dw_kernels", [0] * len(desc["se_sizes"])))
        d = SbNetDesc(desc.get("version", 5), INPUT_CHANNELS, desc["blocks"], desc["channels"], desc["P"], desc["V"],
                      desc["activation"], se, types, inner, dwk, desc.get("policy_head_type", POLICY_HEAD_NORMAL),
                      desc.get("policy_dw_kernel", 0))
        wptr = None
        keep = []
        if tensors is not None:
            ts = (SbTensor * len(tensors))()
            for i, t in enumerate(tensors):
                a = np.ascontiguousarray(t, dtype=np.float32).ravel()
                keep.append(a)
                ts[i].data = a.ctypes.data_as(_F)
                ts[i].count = a.size
            w = SbWeights(ts, len(tensors))
            wptr = ctypes.byref(w)
        ids = list(gpus) if gpus else []
        arr = (ctypes.c_int * max(len(ids), 1))(*ids)
        rc = self._lib.sb_create(ctypes.byref(self._h), ctypes.byref(d), wptr, arr, len(ids), board_size, batch_size, precision)
        if rc:
            self._h = ctypes.c_void_p(None)
            raise RuntimeError(self._lib.sb_last_error(None).decode())
        return self

    def _check(self, rc):
        if rc:
            raise RuntimeError(self._lib.sb_last_error(self._h).decode())

    def construct(self, board_size=-1, batch_size=-1):
        """CudaForwardPipe::Construct(option, nullptr): non-positive keeps the current value."""
        self._check(self._lib.sb_reconfigure(self._h, board_size, batch_size))

    def reload(self, weights_path):
        self._check(self._lib.sb_reload_weights_from_file(self._h, str(weights_path).encode()))

    def valid(self):
        return bool(self._h)

    def get_num_workers(self):
        return self._lib.sb_num_gpus(self._h)

    num_slots = property(lambda self: self._lib.sb_num_slots(self._h))
    max_batch = property(lambda self: self._lib.sb_max_batch(self._h))
    board_size = property(lambda self: self._lib.sb_board_size(self._h))

    def net_desc(self):
        d = SbNetDesc()
        se = (ctypes.c_int * 1024)()
        self._check(self._lib.sb_get_net_desc(self._h, ctypes.byref(d), se, 1024))
        types = (ctypes.c_int * 1024)()
        inner = (ctypes.c_int * 1024)()
        self._check(self._lib.sb_get_block_desc(self._h, types, inner, 1024))
        dwk = (ctypes.c_int * 1024)()
        pht = ctypes.c_int(0)
        pdk = ctypes.c_int(0)
        self._check(self._lib.sb_get_dw_desc(self._h, dwk, 1024, ctypes.byref(pht), ctypes.byref(pdk)))
        return dict(dw_kernels=[dwk[i] for i in range(d.blocks)], policy_head_type=pht.value, policy_dw_kernel=pdk.value,
                    version=d.version, blocks=d.blocks, channels=d.channels, P=d.policy_channels, V=d.value_channels,
                    activation=d.activation, se_sizes=[se[i] for i in range(d.blocks)],
                    block_types=[types[i] for i in range(d.blocks)], inner_channels=[inner[i] for i in range(d.blocks)])

    def release(self):
        self.destroy()

    def destroy(self):
        if getattr(self, "_h", None):
            self._lib.sb_destroy(self._h)
            self._h = ctypes.c_void_p(None)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    # ---- BatchForwardPipe::BatchForward -------------------------------------------------------
    def batch_forward(self, gpu, planes_list, board_sizes, offsets=None):
        """planes_list[i]: float32 array with 43*bs_i*bs_i values (InputData.planes at native size).
        Returns a structured array of OutputResult mirrors."""
        n = len(planes_list)
        offsets = [0] * n if offsets is None else list(offsets)
        keep = [np.ascontiguousarray(p, dtype=np.float32).ravel() for p in planes_list]
        for a, bs in zip(keep, board_sizes):
            if a.size < INPUT_CHANNELS * bs * bs:
                raise ValueError("planes array smaller than 43*bs*bs")
        ptrs = (_F * n)(*[a.ctypes.data_as(_F) for a in keep])
        sizes = (ctypes.c_int * n)(*[int(b) for b in board_sizes])
        offs = (ctypes.c_int * n)(*[int(o) for o in offsets])
        out = np.zeros(n, dtype=OUTPUT_DTYPE)
        self._check(self._lib.sb_forward_batch(self._h, gpu, n, ptrs, sizes, offs, out.ctypes.data))
        return out

    def time_batch_forward_host(self, gpu, planes_list, board_sizes, offsets, seconds):
        """Wall-clock throughput of the blocking sb_forward_batch call with pageable host buffers (argument
        marshalling hoisted out of the loop).  Returns (evals/s, ms per call)."""
        import time
        n = len(planes_list)
        keep = [np.ascontiguousarray(p, dtype=np.float32).ravel() for p in planes_list]
        ptrs = (_F * n)(*[a.ctypes.data_as(_F) for a in keep])
        sizes = (ctypes.c_int * n)(*[int(b) for b in board_sizes])
        offs = (ctypes.c_int * n)(*[int(o) for o in offsets])
        out = np.zeros(n, dtype=OUTPUT_DTYPE)
        iters = 0
        t0 = time.perf_counter()
        while True:
            self._check(self._lib.sb_forward_batch(self._h, gpu, n, ptrs, sizes, offs, out.ctypes.data))
            iters += 1
            el = time.perf_counter() - t0
            if el >= seconds:
                break
        return iters * n / el, 1e3 * el / iters

    def forward(self, planes, board_size, offset=0, gpu=0):
        """NetworkForwardPipe::Forward for one InputData."""
        return self.batch_forward(gpu, [planes], [board_size], [offset])[0]

    # ---- the batcher: NetworkForwardPipe::Forward from any number of threads (sb_eval) ----------------
    def eval(self, planes, board_size, offset=0):
        """Blocking, thread-safe single-position evaluation through the engine's own batcher (ctypes releases the
        GIL during the call, so Python threads batch together like the front-end's search threads)."""
        a = np.ascontiguousarray(planes, dtype=np.float32).ravel()
        if a.size < INPUT_CHANNELS * board_size * board_size:
            raise ValueError("planes array smaller than 43*bs*bs")
        out = np.zeros(1, dtype=OUTPUT_DTYPE)
        self._check(self._lib.sb_eval(self._h, a.ctypes.data_as(_F), board_size, offset, out.ctypes.data))
        return out[0]

    def eval_submit(self, planes, board_size, offset=0):
        """sb_eval_submit: claims a batch entry and packs the position; returns a ticket for eval_poll / eval_wait."""
        a = np.ascontiguousarray(planes, dtype=np.float32).ravel()
        if a.size < INPUT_CHANNELS * board_size * board_size:
            raise ValueError("planes array smaller than 43*bs*bs")
        t = SbEvalTicket()
        self._check(self._lib.sb_eval_submit(self._h, a.ctypes.data_as(_F), board_size, offset, ctypes.byref(t)))
        return t

    def eval_poll(self, ticket):
        """None while the batch has not run, else the OutputResult mirror."""
        out = np.zeros(1, dtype=OUTPUT_DTYPE)
        rc = self._lib.sb_eval_poll(self._h, ctypes.byref(ticket), out.ctypes.data)
        if rc == 1:
            return None
        self._check(rc)
        return out[0]

    def eval_wait(self, ticket):
        out = np.zeros(1, dtype=OUTPUT_DTYPE)
        self._check(self._lib.sb_eval_wait(self._h, ctypes.byref(ticket), out.ctypes.data))
        return out[0]

    def eval_symm8(self, planes, board_size, offset=0, temperature=1.0):
        """Network::GetOutput(state, kAverage): the 8 symmetric views in one batch, post-processed and averaged."""
        a = np.ascontiguousarray(planes, dtype=np.float32).ravel()
        if a.size < INPUT_CHANNELS * board_size * board_size:
            raise ValueError("planes array smaller than 43*bs*bs")
        r = SbSymm8Result()
        self._check(self._lib.sb_eval_symm8(self._h, a.ctypes.data_as(_F), board_size, offset, temperature, ctypes.byref(r)))
        s = board_size * board_size
        return dict(probabilities=np.array(r.probabilities[:s], np.float32), ownership=np.array(r.ownership[:s], np.float32),
                    pass_probability=r.pass_probability, wdl=np.array(r.wdl[:], np.float32), wdl_winrate=r.wdl_winrate,
                    stm_winrate=r.stm_winrate, final_score=r.final_score, q_error=r.q_error, score_error=r.score_error)

    def batcher_config(self, batch_size=0, wait_us=-1):
        self._check(self._lib.sb_batcher_config(self._h, batch_size, wait_us))

    def batcher_stats(self):
        buf = (ctypes.c_longlong * 6)()
        self._check(self._lib.sb_batcher_stats(self._h, buf))
        return dict(zip(("batches", "positions", "full", "timer", "raw", "workers"), [int(v) for v in buf]))

    def eval_throughput(self, positions, board_size, threads, seconds):
        """positions: [n_pos, 43*361] float32 (records SB_PLANE_FLOATS apart).  Native host threads, wall clock."""
        a = np.ascontiguousarray(positions, dtype=np.float32)
        if a.ndim != 2 or a.shape[1] != PLANE_FLOATS:
            raise ValueError("positions must be [n_pos, 43*361]")
        v = self._lib.sb_eval_throughput(self._h, a.ctypes.data_as(_F), a.shape[0], board_size, threads, seconds)
        if v < 0:
            self._check(int(v))
        return float(v)

    def eval_throughput_async(self, positions, board_size, threads, depth, seconds):
        """Feeder threads with `depth` tickets in flight each (sb_eval_submit / sb_eval_wait)."""
        a = np.ascontiguousarray(positions, dtype=np.float32)
        v = self._lib.sb_eval_throughput_async(self._h, a.ctypes.data_as(_F), a.shape[0], board_size, threads, depth, seconds)
        if v < 0:
            self._check(int(v))
        return float(v)

    def submit(self, gpu, slot, planes, board_sizes, offsets, plane_stride=PLANE_FLOATS):
        """planes: contiguous float32 array (ideally a PinnedArray.array) of n records plane_stride apart."""
        n = len(board_sizes)
        sizes = np.ascontiguousarray(board_sizes, dtype=np.int32)
        offs = np.ascontiguousarray(offsets, dtype=np.int32)
        self._check(self._lib.sb_submit(self._h, gpu, slot, n, planes.ctypes.data, plane_stride,
                                        sizes.ctypes.data_as(_I), offs.ctypes.data_as(_I)))

    def wait(self, gpu, slot, out=None):
        self._check(self._lib.sb_wait(self._h, gpu, slot, out.ctypes.data if out is not None else None))
        return out

    # ---- weights blob / measurement / debug ---------------------------------------------------
    def weights_blob(self, gpu=0):
        p = ctypes.c_void_p()
        n = ctypes.c_size_t()
        self._check(self._lib.sb_weights_blob(self._h, gpu, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def weights_export(self, device_ptr, nbytes, gpu=0):
        self._check(self._lib.sb_weights_export(self._h, gpu, ctypes.c_void_p(device_ptr), nbytes))

    def weights_import(self, device_ptr, nbytes, gpu=0):
        self._check(self._lib.sb_weights_import(self._h, gpu, ctypes.c_void_p(device_ptr), nbytes))

    def weights_checksum(self, gpu=0):
        return int(self._lib.sb_weights_checksum(self._h, gpu))

    def weights_broadcast(self):
        """Replica 0 -> all other replicas of this process, device to device (ncclBroadcast / peer copy) + verification."""
        self._check(self._lib.sb_weights_broadcast(self._h))

    def weights_stats(self):
        buf = (ctypes.c_longlong * 6)()
        self._check(self._lib.sb_weights_stats(self._h, buf))
        d = dict(zip(("h2d_uploads", "d2d_fills", "method", "nccl_version", "verified", "broadcast_us"), [int(v) for v in buf]))
        d["method"] = {0: "single", 1: "nccl", 2: "peer"}[d["method"]]
        return d

    def time_forward(self, gpu, slot, iters, flush_l2=True, profile_conv=False):
        ms = np.zeros(max(iters, 1), dtype=np.float32)
        conv_ms = ctypes.c_float(0)
        conv_n = ctypes.c_int(0)
        self._check(self._lib.sb_time_forward(self._h, gpu, slot, iters, int(flush_l2), ms.ctypes.data_as(_F),
                                              ctypes.byref(conv_ms) if profile_conv else None,
                                              ctypes.byref(conv_n) if profile_conv else None))
        # with iters == 0 and profile_conv, ms[0] holds the whole-forward time of the profiling pass conv_ms comes from
        return (ms[:iters] if iters > 0 or not profile_conv else ms[:1]), conv_ms.value, conv_n.value

    def launch_count(self):
        return int(self._lib.sb_launch_count(self._h))

    def debug_read_trunk(self, gpu, slot, sample, board_size):
        c = self.net_desc()["channels"]
        out = np.zeros(c * board_size * board_size, dtype=np.float32)
        self._check(self._lib.sb_debug_read_trunk(self._h, gpu, slot, sample, out.ctypes.data_as(_F)))
        return out.reshape(c, board_size * board_size)

    def conv_stats(self, gpu=0, slot=0):
        """[n_ctas, 8] int64 cycle counters of the last conv3x3 launch (needs set_option("stats", 1))."""
        buf = np.zeros(8 * 1024, dtype=np.int64)
        n = self._lib.sb_conv_stats(self._h, gpu, slot, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), buf.size)
        if n < 0:
            self._check(n)
        return buf[:n].reshape(-1, 8)

    def set_option(self, key, value):
        self._check(self._lib.sb_set_option(self._h, key.encode(), int(value)))
