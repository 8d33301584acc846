"""TEST INFRASTRUCTURE — ctypes bindings for the CPU oracle (oracle/liboracle.so, the plain-C
restatement) and for the compiled UNMODIFIED reference (oracle/_ref/libsayuri_ref_*.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package (sayuri_b200/) must never import it.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_F = ctypes.POINTER(ctypes.c_float)


def _fp(a):
    return a.ctypes.data_as(_F)


def build(quiet=True):
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    cmd = ["make", "-C", HERE, "-j8", "all"]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL if quiet else None)


class Oracle:
    """The plain-C restatement (oracle_forward.c)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            path = os.path.join(HERE, "liboracle.so")
            if not os.path.exists(path):
                build()
            lib = ctypes.CDLL(path)
            lib.oracle_load.restype = ctypes.c_void_p
            lib.oracle_load.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
            lib.oracle_free.argtypes = [ctypes.c_void_p]
            lib.oracle_info.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
            lib.oracle_forward.argtypes = [ctypes.c_void_p, _F, ctypes.c_int, ctypes.c_int, _F]
            lib.oracle_forward_trace.argtypes = [ctypes.c_void_p, _F, ctypes.c_int, ctypes.c_int, _F, _F, _F]
            lib.oracle_get_tensor.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_F)]
            lib.oracle_canvas_place.argtypes = [_F, ctypes.c_int, ctypes.c_int, ctypes.c_int, _F]
            lib.oracle_canvas_crop.argtypes = [_F, ctypes.c_int, ctypes.c_int, _F]
            lib.oracle_symmetry_table.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
            cls._lib = lib
        return cls._lib

    def __init__(self, weights_path):
        lib = self.lib()
        err = ctypes.create_string_buffer(512)
        self._h = lib.oracle_load(weights_path.encode(), err, 512)
        if not self._h:
            raise RuntimeError("oracle_load: " + err.value.decode())
        info = (ctypes.c_int * 8)()
        lib.oracle_info(self._h, info)
        (self.version, self.input_channels, self.blocks, self.channels, self.P, self.V, self.act, self.n_se) = list(info)

    def __del__(self):
        if getattr(self, "_h", None):
            self.lib().oracle_free(self._h)
            self._h = None

    def forward(self, planes, board_size, offset=0):
        """planes: float32 [43*bs*bs]; returns dict(prob[s], own[s], misc[8])."""
        s = board_size * board_size
        planes = np.ascontiguousarray(planes, dtype=np.float32).ravel()
        assert planes.size >= 43 * s
        out = np.zeros(2 * s + 8, dtype=np.float32)
        rc = self.lib().oracle_forward(self._h, _fp(planes), board_size, int(offset), _fp(out))
        if rc:
            raise RuntimeError("oracle_forward rc=%d" % rc)
        return {"prob": out[:s].copy(), "own": out[s:2 * s].copy(), "misc": out[2 * s:].copy()}

    def forward_trace(self, planes, board_size, offset=0):
        s = board_size * board_size
        planes = np.ascontiguousarray(planes, dtype=np.float32).ravel()
        out = np.zeros(2 * s + 8, dtype=np.float32)
        trunk = np.zeros(self.channels * s, dtype=np.float32)
        allp = np.zeros(5 * s + 20, dtype=np.float32)
        rc = self.lib().oracle_forward_trace(self._h, _fp(planes), board_size, int(offset), _fp(out), _fp(trunk), _fp(allp))
        if rc:
            raise RuntimeError("oracle_forward_trace rc=%d" % rc)
        return {"prob": out[:s].copy(), "own": out[s:2 * s].copy(), "misc": out[2 * s:].copy(),
                "trunk": trunk.reshape(self.channels, s), "all_prob": allp[:5 * s].reshape(5, s).copy(),
                "all_pass": allp[5 * s:5 * s + 5].copy(), "all_misc": allp[5 * s + 5:].copy()}

    def tensors(self):
        """Folded (weights, biases) pairs in loader order."""
        res = []
        idx = 0
        while True:
            p = _F()
            n = self.lib().oracle_get_tensor(self._h, idx, 0, ctypes.byref(p))
            if n < 0:
                break
            w = np.ctypeslib.as_array(p, shape=(n,)).copy()
            n = self.lib().oracle_get_tensor(self._h, idx, 1, ctypes.byref(p))
            b = np.ctypeslib.as_array(p, shape=(n,)).copy()
            res.append((w, b))
            idx += 1
        return res

    @classmethod
    def canvas_place(cls, planes, channels, n, N):
        planes = np.ascontiguousarray(planes, dtype=np.float32).ravel()
        out = np.empty(channels * N * N, dtype=np.float32)
        cls.lib().oracle_canvas_place(_fp(planes), channels, n, N, _fp(out))
        return out

    @classmethod
    def canvas_crop(cls, canvas, n, N):
        canvas = np.ascontiguousarray(canvas, dtype=np.float32).ravel()
        out = np.empty(n * n, dtype=np.float32)
        cls.lib().oracle_canvas_crop(_fp(canvas), n, N, _fp(out))
        return out

    @classmethod
    def symmetry_table(cls, N, symm):
        t = np.empty(N * N, dtype=np.int32)
        cls.lib().oracle_symmetry_table(N, symm, t.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
        return t


def _cpu_has_avx512():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    fl = line.split()
                    return all(x in fl for x in ("avx512f", "avx512bw", "avx512vl", "avx512dq", "avx512cd"))
    except OSError:
        pass
    return False


class Reference:
    """The unmodified reference (loader + Eigen BlasForwardPipe) compiled into oracle/_ref.
    The process-global option map of the reference means ONE net per process."""

    _lib = None
    variant = None

    @classmethod
    def available(cls):
        return any(os.path.exists(os.path.join(HERE, "_ref", "libsayuri_ref_%s.so" % v)) for v in ("v3", "v4"))

    @classmethod
    def lib(cls):
        if cls._lib is None:
            order = ["v4", "v3"] if _cpu_has_avx512() else ["v3"]
            for v in order:
                path = os.path.join(HERE, "_ref", "libsayuri_ref_%s.so" % v)
                if os.path.exists(path):
                    lib = ctypes.CDLL(path)
                    cls.variant = v
                    break
            else:
                raise RuntimeError("oracle/_ref is not built (needs /root/reference; run make -C oracle)")
            lib.ref_init.argtypes = [ctypes.c_char_p, ctypes.c_int]
            lib.ref_net_info.argtypes = [ctypes.POINTER(ctypes.c_int)]
            lib.ref_forward.argtypes = [_F, ctypes.c_int, ctypes.c_float, ctypes.c_int, _F]
            lib.ref_time_forward.restype = ctypes.c_long
            lib.ref_time_forward.argtypes = [_F, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                             ctypes.POINTER(ctypes.c_double)]
            cls._lib = lib
        return cls._lib

    def __init__(self, weights_path, winograd=False):
        rc = self.lib().ref_init(weights_path.encode(), int(bool(winograd)))
        if rc:
            raise RuntimeError("ref_init rc=%d" % rc)
        info = (ctypes.c_int * 8)()
        self.lib().ref_net_info(info)
        (self.version, self.input_channels, self.blocks, self.channels, self.P, self.V, self.act, self.n_se) = list(info)

    def forward(self, planes, board_size, offset=0, komi=7.5):
        s = board_size * board_size
        planes = np.ascontiguousarray(planes, dtype=np.float32).ravel()
        out = np.zeros(2 * s + 8, dtype=np.float32)
        rc = self.lib().ref_forward(_fp(planes), board_size, komi, int(offset), _fp(out))
        if rc:
            raise RuntimeError("ref_forward rc=%d" % rc)
        return {"prob": out[:s].copy(), "own": out[s:2 * s].copy(), "misc": out[2 * s:].copy()}

    def time_forward(self, planes, n_pos, board_size, threads, seconds):
        planes = np.ascontiguousarray(planes, dtype=np.float32).ravel()
        el = ctypes.c_double(0)
        n = self.lib().ref_time_forward(_fp(planes), n_pos, board_size, threads, seconds, ctypes.byref(el))
        return n, el.value
