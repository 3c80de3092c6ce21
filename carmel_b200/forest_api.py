"""ctypes binding of the forest-em part of include/carmel_b200.h (cml_forests_*, cml_forest_job_*).

Plumbing for tests and bench.py; the product is the CUDA library.  No CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .api import ALLREDUCE_FN, CarmelB200Error, load_library

ZERO, SKIP, UNIFORM = 0, 1, 2
_vp = C.c_void_p
_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)
_f64p = C.POINTER(C.c_double)


class CmlForestBatch(C.Structure):
    _fields_ = [("n_forests", C.c_uint64), ("node_off", _u64p), ("next", _u32p), ("label", _u32p), ("backref", _u8p)]


class CmlForestEstimateResult(C.Structure):
    _fields_ = [("sum_ln_p", C.c_double), ("n_zero", C.c_uint64), ("n_forests", C.c_uint64)]


class CmlForestNormOpts(C.Structure):
    _fields_ = [("prior_total", C.c_double), ("add_k", C.c_double), ("zero_mode", C.c_int)]


class CmlForestJobInfo(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("forests", "nodes", "hyperedges", "links", "rulespace", "iterations")] + [
        ("best_avg_logprob", C.c_double)]


_bound = False


def _lib():
    global _bound
    lib = load_library()
    if not _bound:
        lib.cml_forests_create.argtypes = [C.POINTER(_vp), C.c_int, C.c_int]
        lib.cml_forests_destroy.argtypes = [_vp]
        lib.cml_forests_destroy.restype = None
        lib.cml_forests_last_error.argtypes = [_vp]
        lib.cml_forests_last_error.restype = C.c_char_p
        lib.cml_forests_set_stream.argtypes = [_vp, _vp]
        lib.cml_forests_set_layout.argtypes = [_vp, C.c_int]
        lib.cml_forests_layout_stats.argtypes = [_vp] + [_u64p] * 5
        lib.cml_forests_level_stats.argtypes = [_vp] + [_u64p] * 6
        lib.cml_forests_launch_count.argtypes = [_vp]
        lib.cml_forests_launch_count.restype = C.c_uint64
        lib.cml_forests_set_rules.argtypes = [_vp, C.c_uint64, C.c_uint64, _u64p, _u64p]
        lib.cml_forests_set_params.argtypes = [_vp, _f64p]
        lib.cml_forests_get_params.argtypes = [_vp, _f64p]
        lib.cml_forests_add.argtypes = [_vp, C.POINTER(CmlForestBatch)]
        lib.cml_forests_totals.argtypes = [_vp] + [_u64p] * 4
        lib.cml_forests_estimate.argtypes = [_vp, C.POINTER(CmlForestEstimateResult)]
        lib.cml_forests_estimate_launch.argtypes = [_vp]
        lib.cml_forests_estimate_finish.argtypes = [_vp, C.POINTER(CmlForestEstimateResult)]
        lib.cml_forests_last_time_ms.argtypes = [_vp, C.POINTER(C.c_float), _u32p]
        lib.cml_forests_get_inside.argtypes = [_vp, _f64p, C.c_uint64]
        lib.cml_forests_get_counts.argtypes = [_vp, _f64p, C.c_uint64]
        lib.cml_forests_reduce_buffer.argtypes = [_vp, C.POINTER(_vp), _u64p]
        lib.cml_forests_comm_init_rank.argtypes = [_vp, C.c_int, C.c_int, C.c_char_p]
        lib.cml_forests_allreduce_counts.argtypes = [_vp]
        lib.cml_forests_maximize.argtypes = [_vp, C.POINTER(CmlForestNormOpts), _f64p, _u64p]
        lib.cml_forests_normalize_params.argtypes = [_vp]
        lib.cml_forest_job_open.argtypes = [C.POINTER(_vp), C.c_int, C.POINTER(C.c_char_p)]
        lib.cml_forest_job_close.argtypes = [_vp]
        lib.cml_forest_job_close.restype = None
        lib.cml_forest_job_error.argtypes = [_vp]
        lib.cml_forest_job_error.restype = C.c_char_p
        lib.cml_forest_job_set_allreduce.argtypes = [_vp, ALLREDUCE_FN, _vp]
        lib.cml_forest_job_prepare.argtypes = [_vp]
        lib.cml_forest_job_context.argtypes = [_vp]
        lib.cml_forest_job_context.restype = _vp
        lib.cml_forest_job_train.argtypes = [_vp]
        lib.cml_forest_job_write.argtypes = [_vp]
        lib.cml_forest_job_stats.argtypes = [_vp, C.POINTER(CmlForestJobInfo)]
        _bound = True
    return lib


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Forests:
    """One cml_forests handle: rules + parameters + resident forests of one GPU."""

    def __init__(self, device: int = 0, precision: int = 32, handle=None):
        self.lib = _lib()
        self._own = handle is None
        if handle is None:
            h = _vp()
            rc = self.lib.cml_forests_create(C.byref(h), device, precision)
            if rc != 0:
                raise CarmelB200Error(rc, self.lib.cml_forests_last_error(None).decode())
            handle = h
        self.h = handle
        self.rulespace = 0

    def close(self):
        if self._own and self.h:
            self.lib.cml_forests_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ok(self, rc):
        if rc != 0:
            raise CarmelB200Error(rc, self.lib.cml_forests_last_error(self.h).decode())

    def set_stream(self, cuda_stream: int):
        self._ok(self.lib.cml_forests_set_stream(self.h, _vp(cuda_stream)))

    def set_layout(self, layout: int):
        """0 auto, 1 warp/CTA per forest, 2 thread per forest, 3 level-synchronous tiles (CTA per run of forests)"""
        self._ok(self.lib.cml_forests_set_layout(self.h, layout))

    def layout_stats(self) -> dict:
        v = [C.c_uint64() for _ in range(5)]
        self._ok(self.lib.cml_forests_layout_stats(self.h, *[C.byref(x) for x in v]))
        return dict(zip(("tile_forests", "tiles", "steps", "padded_steps", "padded_rows"), (x.value for x in v)))

    def level_stats(self) -> dict:
        v = [C.c_uint64() for _ in range(6)]
        self._ok(self.lib.cml_forests_level_stats(self.h, *[C.byref(x) for x in v]))
        return dict(zip(("level_forests", "level_tiles", "level_nodes", "level_links", "max_tile_nodes", "small_tiles"),
                        (x.value for x in v)))

    def set_rules(self, rulespace: int, group_off, group_members):
        go, gm = _arr(group_off, np.uint64), _arr(group_members, np.uint64)
        self._ok(self.lib.cml_forests_set_rules(self.h, rulespace, len(go) - 1, go.ctypes.data_as(_u64p), gm.ctypes.data_as(_u64p)))
        self.rulespace = rulespace

    def set_params(self, ln_w):
        w = _arr(ln_w, np.float64)
        assert len(w) == self.rulespace
        self._ok(self.lib.cml_forests_set_params(self.h, w.ctypes.data_as(_f64p)))

    def get_params(self) -> np.ndarray:
        w = np.empty(self.rulespace, np.float64)
        self._ok(self.lib.cml_forests_get_params(self.h, w.ctypes.data_as(_f64p)))
        return w

    def add(self, node_off, nxt, label, backref):
        no, nx, lb, br = _arr(node_off, np.uint64), _arr(nxt, np.uint32), _arr(label, np.uint32), _arr(backref, np.uint8)
        b = CmlForestBatch(len(no) - 1, no.ctypes.data_as(_u64p), nx.ctypes.data_as(_u32p), lb.ctypes.data_as(_u32p),
                           br.ctypes.data_as(_u8p))
        self._ok(self.lib.cml_forests_add(self.h, C.byref(b)))

    def totals(self):
        v = [C.c_uint64() for _ in range(4)]
        self._ok(self.lib.cml_forests_totals(self.h, *[C.byref(x) for x in v]))
        return dict(zip(("forests", "nodes", "hyperedges", "links"), (x.value for x in v)))

    def estimate(self):
        r = CmlForestEstimateResult()
        self._ok(self.lib.cml_forests_estimate(self.h, C.byref(r)))
        return r.sum_ln_p, r.n_zero, r.n_forests

    def estimate_launch(self):
        self._ok(self.lib.cml_forests_estimate_launch(self.h))

    def estimate_finish(self):
        r = CmlForestEstimateResult()
        self._ok(self.lib.cml_forests_estimate_finish(self.h, C.byref(r)))
        return r.sum_ln_p, r.n_zero, r.n_forests

    def last_time_ms(self):
        ms, nk = C.c_float(), C.c_uint32()
        self._ok(self.lib.cml_forests_last_time_ms(self.h, C.byref(ms), C.byref(nk)))
        return ms.value, nk.value

    def inside(self, n: int) -> np.ndarray:
        v = np.empty(n, np.float64)
        self._ok(self.lib.cml_forests_get_inside(self.h, v.ctypes.data_as(_f64p), n))
        return v

    def counts(self) -> np.ndarray:
        v = np.empty(self.rulespace, np.float64)
        self._ok(self.lib.cml_forests_get_counts(self.h, v.ctypes.data_as(_f64p), self.rulespace))
        return v

    def comm_init_rank(self, n_ranks: int, rank: int, token: bytes):
        assert len(token) == 128
        self._ok(self.lib.cml_forests_comm_init_rank(self.h, n_ranks, rank, token))

    def allreduce_counts(self):
        self._ok(self.lib.cml_forests_allreduce_counts(self.h))

    def reduce_buffer(self):
        p, n = _vp(), C.c_uint64()
        self._ok(self.lib.cml_forests_reduce_buffer(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def maximize(self, prior_total=0.0, add_k=0.0, zero_mode=UNIFORM):
        o = CmlForestNormOpts(prior_total, add_k, zero_mode)
        d, i = C.c_double(), C.c_uint64()
        self._ok(self.lib.cml_forests_maximize(self.h, C.byref(o), C.byref(d), C.byref(i)))
        return d.value, i.value

    def normalize_params(self):
        self._ok(self.lib.cml_forests_normalize_params(self.h))

    def launch_count(self) -> int:
        return int(self.lib.cml_forests_launch_count(self.h))


class ForestJob:
    """A whole forest-em run through the C ABI (what forest-em-b200 does)."""

    def __init__(self, argv: list[str]):
        self.lib = _lib()
        args = [b"forest-em-b200"] + [a.encode() for a in argv]
        arr = (C.c_char_p * len(args))(*args)
        h = _vp()
        rc = self.lib.cml_forest_job_open(C.byref(h), len(args), arr)
        self.h = h
        self._cb = None
        if rc != 0:
            msg = self.lib.cml_forest_job_error(h).decode()
            self.lib.cml_forest_job_close(h)
            self.h = None
            raise CarmelB200Error(rc, msg)

    def _ok(self, rc):
        if rc != 0:
            raise CarmelB200Error(rc, self.lib.cml_forest_job_error(self.h).decode())

    def set_allreduce(self, fn):
        self._cb = ALLREDUCE_FN(lambda user, ptr, n: fn(ptr, n))
        self._ok(self.lib.cml_forest_job_set_allreduce(self.h, self._cb, None))

    def prepare(self):
        self._ok(self.lib.cml_forest_job_prepare(self.h))

    def context(self) -> Forests:
        f = Forests(handle=_vp(self.lib.cml_forest_job_context(self.h)))
        return f

    def train(self):
        self._ok(self.lib.cml_forest_job_train(self.h))

    def write(self):
        self._ok(self.lib.cml_forest_job_write(self.h))

    def stats(self) -> dict:
        i = CmlForestJobInfo()
        self._ok(self.lib.cml_forest_job_stats(self.h, C.byref(i)))
        return {n: getattr(i, n) for n, _ in CmlForestJobInfo._fields_}

    def close(self):
        if self.h:
            self.lib.cml_forest_job_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
