"""ctypes binding of the CPU oracle (oracle/_build/libwrenc_oracle.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

RECORD_DTYPE = np.dtype([("split_mask", "<u4"), ("luma_mode", "u1", (64,)), ("chroma_mode", "u1", (16,)), ("cost", "<f4")])
assert RECORD_DTYPE.itemsize == 88


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ROOT, "oracle", "_build", "libwrenc_oracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        L = C.CDLL(so)
        L.wo_create.restype = C.c_void_p
        L.wo_create.argtypes = [C.c_int, C.c_int, C.c_char_p]
        L.wo_destroy.argtypes = [C.c_void_p]
        L.wo_encode_picture.restype = C.c_long
        L.wo_encode_picture.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 10 + [C.c_void_p, C.c_size_t]
        L.wo_num_pipelines.restype = C.c_uint64
        L.wo_num_pipelines.argtypes = [C.c_void_p]
        L.wo_num_predictions.restype = C.c_uint64
        L.wo_num_predictions.argtypes = [C.c_void_p]
        L.wo_rate.restype = C.c_int64
        L.wo_decode_picture.restype = C.c_int
        L.wo_decode_picture.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_size_t] + [C.c_void_p] * 7 + [C.POINTER(C.c_size_t)]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    def __init__(self, qp=32, max_depth=3, extra_params=None):
        self.L = lib()
        self.h = self.L.wo_create(qp, max_depth, extra_params.encode() if extra_params else None)
        if not self.h:
            raise ValueError("invalid extra params")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.wo_destroy(self.h)
            self.h = None

    def consts(self):
        lv = np.zeros(8, np.int64)
        dq = np.zeros(8, np.int64)
        lq = C.c_int64()
        lrd = C.c_float()
        ls = C.c_int32()
        self.L.wo_consts(C.c_void_p(self.h), _p(lv), _p(dq), C.byref(lq), C.byref(lrd), C.byref(ls))
        return dict(lv=lv, dq=dq, lambda_q=lq.value, lambda_rd=lrd.value, ls=ls.value)

    def encode_picture(self, y, cb, cr, want_slice_data=False):
        H, W = y.shape
        y, cb, cr = (np.ascontiguousarray(a, np.uint8) for a in (y, cb, cr))
        rec = [np.zeros_like(y), np.zeros_like(cb), np.zeros_like(cr)]
        coef = [np.zeros(y.shape, np.int16), np.zeros(cb.shape, np.int16), np.zeros(cr.shape, np.int16)]
        records = np.zeros((H // 32) * (W // 32), RECORD_DTYPE)
        cap = W * H * 4 + 4096
        sd = np.zeros(cap, np.uint8) if want_slice_data else None
        n = self.L.wo_encode_picture(C.c_void_p(self.h), W, H, _p(y), _p(cb), _p(cr), _p(rec[0]), _p(rec[1]), _p(rec[2]),
                                     _p(coef[0]), _p(coef[1]), _p(coef[2]), _p(records), _p(sd), cap)
        assert n >= 0
        out = dict(rec=rec, coef=coef, records=records)
        if want_slice_data:
            out["slice_data"] = sd[:n].tobytes()
        return out

    def predict(self, rec, x, y, w, tree, ar, bl, c, mode):
        H, W = rec[0].shape
        n = w if c == 0 else w // 2
        pred = np.zeros((n, n), np.uint8)
        r = [np.ascontiguousarray(a, np.uint8) for a in rec]
        self.L.wo_predict(W, H, _p(r[0]), _p(r[1]), _p(r[2]), x, y, w, tree, int(ar), int(bl), c, mode, _p(pred))
        return pred

    def fwd_dct(self, res):
        n = res.shape[0]
        out = np.zeros((n, n), np.int16)
        self.L.wo_fwd_dct(_p(np.ascontiguousarray(res, np.int16)), int(np.log2(n)), _p(out))
        return out

    def inv_dct(self, deq):
        n = deq.shape[0]
        out = np.zeros((n, n), np.int16)
        self.L.wo_inv_dct(_p(np.ascontiguousarray(deq, np.int16)), int(np.log2(n)), _p(out))
        return out

    def quantize(self, coef):
        n = coef.shape[0]
        out = np.zeros((n, n), np.int16)
        self.L.wo_quantize(C.c_void_p(self.h), _p(np.ascontiguousarray(coef, np.int16)), int(np.log2(n)), _p(out))
        return out

    def dequantize(self, q):
        n = q.shape[0]
        out = np.zeros((n, n), np.int16)
        self.L.wo_dequantize(C.c_void_p(self.h), _p(np.ascontiguousarray(q, np.int16)), int(np.log2(n)), _p(out))
        return out

    def rate(self, q):
        n = q.shape[0]
        return self.L.wo_rate(C.c_void_p(self.h), _p(np.ascontiguousarray(q, np.int16)), int(np.log2(n)))


def dct_matrix(log2n):
    n = 1 << log2n
    out = np.zeros((n, n), np.int16)
    lib().wo_dct_matrix(log2n, _p(out))
    return out


def scan(log2n):
    n = 1 << log2n
    out = np.zeros(n * n, np.uint16)
    lib().wo_scan(log2n, _p(out))
    return out


def decode_picture(qp, W, H, slice_data):
    """Minimal intra decoder of the emitted subset (oracle/wrenc_decode.cpp).  Returns dict(rec, coef, records) or None."""
    L = lib()
    rec = [np.zeros((H, W), np.uint8), np.zeros((H // 2, W // 2), np.uint8), np.zeros((H // 2, W // 2), np.uint8)]
    coef = [np.zeros((H, W), np.int16), np.zeros((H // 2, W // 2), np.int16), np.zeros((H // 2, W // 2), np.int16)]
    records = np.zeros((H // 32) * (W // 32), RECORD_DTYPE)
    used = C.c_size_t()
    rc = L.wo_decode_picture(qp, W, H, slice_data, len(slice_data), _p(rec[0]), _p(rec[1]), _p(rec[2]), _p(coef[0]), _p(coef[1]), _p(coef[2]),
                             _p(records), C.byref(used))
    if rc != 0:
        return None
    return dict(rec=rec, coef=coef, records=records, bits=used.value)
