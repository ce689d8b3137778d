"""ctypes binding of libwrenc_b200.so (include/wrenc_b200.h) — the host-side mirror of the reference's per-picture driver
(reference src/main.rs:294-402 -> SliceEncoder::encode -> CtuEncoder::encode -> BlockSplitter::split_ct).

There is no CPU fallback: importing works anywhere (so the symbol table can be checked on a CPU box), but creating a
`SearchEncoder` raises unless the CUDA library loads and finds a B200.  Nothing here touches oracle/.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WRENC_B200_LIB") or os.path.join(_HERE, "libwrenc_b200.so")  # env override: dev-time kernel variants

RECORD_DTYPE = np.dtype([("split_mask", "<u4"), ("luma_mode", "u1", (64,)), ("chroma_mode", "u1", (16,)), ("cost", "<f4")])
assert RECORD_DTYPE.itemsize == 88

SINGLE_TREE, DUAL_TREE_LUMA, DUAL_TREE_CHROMA = 0, 1, 2
MODE_LT_CCLM, MODE_L_CCLM, MODE_T_CCLM = 81, 82, 83

EXPORTS = [
    "wrenc_b200_create", "wrenc_b200_destroy", "wrenc_b200_last_error", "wrenc_b200_submit", "wrenc_b200_submit_pinned", "wrenc_b200_receive",
    "wrenc_b200_decisions", "wrenc_b200_flush", "wrenc_b200_pending", "wrenc_b200_search_resident", "wrenc_b200_code_resident",
    "wrenc_b200_workspace_bytes", "wrenc_b200_get_consts", "wrenc_b200_block_predict", "wrenc_b200_block_fwd_dct",
    "wrenc_b200_block_inv_dct", "wrenc_b200_block_quantize", "wrenc_b200_block_dequantize", "wrenc_b200_version",
    "wrenc_b200_measure_int32_peak", "wrenc_b200_derive_consts", "wrenc_b200_write_nal", "wrenc_b200_write_parameter_sets",
    "wrenc_b200_write_picture", "wrenc_b200_header_rbsp", "wrenc_b200_prepare", "wrenc_b200_code_resident_retry",
    "wrenc_b200_alloc_pinned", "wrenc_b200_free_pinned",
]


class Config(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("qp", C.c_int32), ("max_split_depth", C.c_int32),
                ("device", C.c_int32), ("pictures_in_flight", C.c_int32), ("want_recon", C.c_int32),
                ("want_decisions", C.c_int32), ("want_slice_data", C.c_int32), ("extra_params", C.c_char_p)]


class Consts(C.Structure):
    _fields_ = [("lambda_q", C.c_int64), ("lambda_rd", C.c_float), ("lambda_rd_chroma", C.c_float), ("ls", C.c_int32),
                ("lv", C.c_int64 * 8), ("dq", C.c_int64 * 8)]


class WrencB200Error(RuntimeError):
    pass


class WrencB200Full(WrencB200Error):
    """submit: every batch slot holds pictures that have not been received (WRENC_B200_EFULL)."""


_lib = None


def load_library():
    """Load the C-ABI library; raises (never falls back) when it is missing or does not load."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WrencB200Error(f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, u8p = C.c_void_p, C.c_int32, C.c_void_p
    L.wrenc_b200_create.restype = C.c_int
    L.wrenc_b200_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.wrenc_b200_destroy.restype = None
    L.wrenc_b200_destroy.argtypes = [vp]
    L.wrenc_b200_last_error.restype = C.c_char_p
    L.wrenc_b200_last_error.argtypes = [vp]
    L.wrenc_b200_submit.restype = C.c_int
    L.wrenc_b200_submit.argtypes = [vp, C.c_uint64, u8p, u8p, u8p]
    L.wrenc_b200_submit_pinned.restype = C.c_int
    L.wrenc_b200_submit_pinned.argtypes = [vp, C.c_uint64, u8p, u8p, u8p]
    L.wrenc_b200_receive.restype = C.c_int
    L.wrenc_b200_receive.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.wrenc_b200_decisions.restype = C.c_int
    L.wrenc_b200_decisions.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.wrenc_b200_flush.restype = C.c_int
    L.wrenc_b200_flush.argtypes = [vp]
    L.wrenc_b200_pending.restype = C.c_int
    L.wrenc_b200_pending.argtypes = [vp]
    L.wrenc_b200_search_resident.restype = C.c_int
    L.wrenc_b200_search_resident.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.wrenc_b200_code_resident.restype = C.c_int
    L.wrenc_b200_code_resident.argtypes = [vp, i32, vp, vp, vp, C.c_size_t, vp, vp]
    L.wrenc_b200_code_resident_retry.restype = C.c_int
    L.wrenc_b200_code_resident_retry.argtypes = [vp, i32, vp, vp, vp, C.c_size_t, vp, vp]
    L.wrenc_b200_prepare.restype = C.c_int
    L.wrenc_b200_prepare.argtypes = [vp, i32]
    L.wrenc_b200_workspace_bytes.restype = C.c_size_t
    L.wrenc_b200_workspace_bytes.argtypes = [vp, i32]
    L.wrenc_b200_get_consts.restype = C.c_int
    L.wrenc_b200_get_consts.argtypes = [vp, C.POINTER(Consts)]
    L.wrenc_b200_block_predict.restype = C.c_int
    L.wrenc_b200_block_predict.argtypes = [vp, vp] + [C.c_int] * 8 + [vp]
    for name in ("fwd_dct", "inv_dct", "dequantize"):
        f = getattr(L, "wrenc_b200_block_" + name)
        f.restype = C.c_int
        f.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    L.wrenc_b200_block_quantize.restype = C.c_int
    L.wrenc_b200_block_quantize.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
    L.wrenc_b200_version.restype = C.c_char_p
    L.wrenc_b200_write_nal.restype = C.c_int64
    L.wrenc_b200_write_nal.argtypes = [C.c_int32, C.c_int32, C.c_int32, vp, C.c_size_t, vp, C.c_size_t]
    L.wrenc_b200_write_parameter_sets.restype = C.c_int64
    L.wrenc_b200_write_parameter_sets.argtypes = [i32, i32, i32, vp, C.c_size_t]
    L.wrenc_b200_write_picture.restype = C.c_int64
    L.wrenc_b200_write_picture.argtypes = [i32, C.c_uint64, C.c_char_p, C.c_size_t, vp, C.c_size_t]
    L.wrenc_b200_header_rbsp.restype = C.c_int64
    L.wrenc_b200_header_rbsp.argtypes = [i32, i32, i32, i32, C.c_uint64, vp, C.c_size_t]
    L.wrenc_b200_derive_consts.restype = C.c_int
    L.wrenc_b200_derive_consts.argtypes = [C.c_int32, C.c_char_p, C.POINTER(Consts), vp, vp, vp]
    L.wrenc_b200_measure_int32_peak.restype = C.c_int
    L.wrenc_b200_measure_int32_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    _lib = L
    return L


def measure_int32_peak(device=0):
    """IMAD/s of the device (wrenc_b200_measure_int32_peak)."""
    v = C.c_double()
    rc = load_library().wrenc_b200_measure_int32_peak(int(device), C.byref(v))
    if rc != 0:
        raise WrencB200Error(f"wrenc_b200_measure_int32_peak failed ({rc})")
    return v.value


NAL_IDR_W_RADL, NAL_VPS, NAL_SPS, NAL_PPS, NAL_PH = 7, 14, 15, 16, 19  # nal.rs:10-42


def write_nal(payload, nal_unit_type=NAL_IDR_W_RADL, nuh_layer_id=9, nuh_temporal_id=0):
    """Byte-stream NAL unit around a byte-aligned payload (wrenc_b200_write_nal; host only).  For a picture the payload is
    the reference's slice header bytes followed by the slice_data() this library returns (main.rs:383-389)."""
    payload = bytes(payload)
    src = (C.c_uint8 * max(1, len(payload))).from_buffer_copy(payload or b"\0")
    cap = 8 + len(payload) + len(payload) // 2 + 1
    out = (C.c_uint8 * cap)()
    n = load_library().wrenc_b200_write_nal(int(nuh_layer_id), int(nal_unit_type), int(nuh_temporal_id), src, len(payload), out, cap)
    if n < 0:
        raise ValueError(f"wrenc_b200_write_nal failed ({n})")
    return bytes(out[:n])


def _sized_call(fn, *args, guess=4096):
    buf = C.create_string_buffer(guess)
    n = fn(*args, buf, guess)
    if n < -16:  # negated size needed
        buf = C.create_string_buffer(-n)
        n = fn(*args, buf, -n)
    if n < 0:
        raise ValueError(f"{fn.__name__} failed ({n})")
    return buf.raw[:n]


def write_parameter_sets(width, height, qp=None):
    """VPS + SPS + PPS byte-stream NAL units as reference main.rs:223-260 writes them (qp=None: no --qp flag)."""
    return _sized_call(load_library().wrenc_b200_write_parameter_sets, int(width), int(height), -1 if qp is None else int(qp))


def write_picture(qp, picture_index, slice_data):
    """PH NAL unit + IDR_W_RADL slice NAL unit (slice header + slice_data) of one picture, main.rs:297-316,380-389."""
    sd = bytes(slice_data)
    return _sized_call(load_library().wrenc_b200_write_picture, -1 if qp is None else int(qp), int(picture_index), sd, len(sd),
                       guess=len(sd) + len(sd) // 2 + 64)


def header_rbsp(which, width, height, qp=None, picture_index=0):
    """Raw header bits before NAL wrapping; which in ("vps", "sps", "pps", "ph", "sh")."""
    k = ("vps", "sps", "pps", "ph", "sh").index(which)
    return _sized_call(load_library().wrenc_b200_header_rbsp, k, int(width), int(height), -1 if qp is None else int(qp), int(picture_index))


def assemble_vvc(width, height, qp, slice_datas):
    """The complete .vvc byte stream of a sequence: parameter sets once, then PH + slice NAL units per picture."""
    return write_parameter_sets(width, height, qp) + b"".join(write_picture(qp, i, sd) for i, sd in enumerate(slice_datas))


def derive_consts(qp, extra_params=None):
    """Host-only derivation of the search constants (no GPU needed): wrenc_b200_derive_consts."""
    c = Consts()
    hs, hd, hc = np.zeros((67, 4), np.int64), np.zeros(67, np.int64), np.zeros(4, np.int64)
    rc = load_library().wrenc_b200_derive_consts(int(qp), extra_params.encode() if extra_params else None, C.byref(c),
                                                 hs.ctypes.data_as(C.c_void_p), hd.ctypes.data_as(C.c_void_p), hc.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise ValueError(f"wrenc_b200_derive_consts failed ({rc})")
    return dict(lambda_q=c.lambda_q, lambda_rd=c.lambda_rd, lambda_rd_chroma=c.lambda_rd_chroma, ls=c.ls,
                lv=np.array(list(c.lv), np.int64), dq=np.array(list(c.dq), np.int64), hdr_single=hs, hdr_dual=hd, hdr_chroma=hc)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class SearchEncoder:
    """One handle per GPU.  Mirrors the reference flags: --qp, --max-split-depth, --extra-params, --reconst."""

    def __init__(self, width, height, qp=26, max_split_depth=3, device=0, pictures_in_flight=8, want_recon=True,
                 want_decisions=True, extra_params=None, want_slice_data=True):
        self.L = load_library()
        self.h = C.c_void_p()
        self.width, self.height = int(width), int(height)
        self._extra = extra_params.encode() if extra_params else None
        cfg = Config(self.width, self.height, int(qp), int(max_split_depth), int(device), int(pictures_in_flight),
                     int(bool(want_recon)), int(bool(want_decisions)), int(bool(want_slice_data)), self._extra)
        rc = self.L.wrenc_b200_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            msg = self.L.wrenc_b200_last_error(None).decode()
            self.h = C.c_void_p()
            if rc == -1:
                raise ValueError(msg)
            raise WrencB200Error(f"wrenc_b200_create failed ({rc}): {msg}")
        self.want_recon, self.want_decisions = bool(want_recon), bool(want_decisions)
        self._arr_types = {}  # ctypes array types of receive()'s views, by size
        self.pictures_in_flight = max(1, int(pictures_in_flight))

    def close(self):
        if getattr(self, "h", None) and self.h:
            self.L.wrenc_b200_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close

    def _check(self, rc):
        if rc == -5:
            raise WrencB200Full(self.L.wrenc_b200_last_error(self.h).decode())
        if rc < 0:
            raise WrencB200Error(f"wrenc_b200 error {rc}: {self.L.wrenc_b200_last_error(self.h).decode()}")
        return rc

    def consts(self):
        c = Consts()
        self._check(self.L.wrenc_b200_get_consts(self.h, C.byref(c)))
        return dict(lambda_q=c.lambda_q, lambda_rd=c.lambda_rd, lambda_rd_chroma=c.lambda_rd_chroma, ls=c.ls,
                    lv=np.array(list(c.lv), np.int64), dq=np.array(list(c.dq), np.int64))

    # ---- host-plane path (the reference-facing call) ----
    def submit(self, pic_idx, y, cb, cr, pinned=False):
        """pinned=True: the planes lie in page-locked host memory (e.g. views of a torch pin_memory() tensor) and stay valid
        until the picture is received: wrenc_b200_submit_pinned, no staging copy."""
        y, cb, cr = (np.ascontiguousarray(a, np.uint8) for a in (y, cb, cr))
        assert y.shape == (self.height, self.width) and cb.shape == (self.height // 2, self.width // 2) and cr.shape == cb.shape
        fn = self.L.wrenc_b200_submit_pinned if pinned else self.L.wrenc_b200_submit
        self._check(fn(self.h, int(pic_idx), _ptr(y), _ptr(cb), _ptr(cr)))

    def pending(self):
        return self.L.wrenc_b200_pending(self.h)

    def flush(self):
        self._check(self.L.wrenc_b200_flush(self.h))

    def receive(self, copy=True):
        idx = C.c_uint64()
        sd, ry, rcb, rcr = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        n = C.c_size_t()
        self._check(self.L.wrenc_b200_receive(self.h, C.byref(idx), C.byref(sd), C.byref(n), C.byref(ry), C.byref(rcb), C.byref(rcr)))
        W, H = self.width, self.height
        out = {"pic_idx": idx.value, "slice_data": C.string_at(sd, n.value) if sd and n.value else b""}

        def view(p, shape, dt):
            # a numpy view of the handle's pinned host buffer (valid until the next receive of this slot): np.frombuffer over a cached
            # ctypes array type costs ~1 us, np.ctypeslib.as_array ~12 us, and receive runs once per picture on the e2e path
            n_bytes = int(np.prod(shape)) * np.dtype(dt).itemsize
            ty = self._arr_types.get(n_bytes)
            if ty is None:
                ty = self._arr_types[n_bytes] = C.c_uint8 * n_bytes
            a = np.frombuffer(ty.from_address(p.value), dtype=dt).reshape(shape)
            return a.copy() if copy else a

        if self.want_recon:
            out["rec"] = [view(ry, (H, W), np.uint8), view(rcb, (H // 2, W // 2), np.uint8), view(rcr, (H // 2, W // 2), np.uint8)]
        rec_p, ly, lcb, lcr = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self.L.wrenc_b200_decisions(self.h, C.byref(rec_p), C.byref(ly), C.byref(lcb), C.byref(lcr)))
        nctu = (H // 32) * (W // 32)
        out["records"] = view(rec_p, (nctu,), RECORD_DTYPE)
        if self.want_decisions:
            out["coef"] = [view(ly, (H, W), np.int16), view(lcb, (H // 2, W // 2), np.int16), view(lcr, (H // 2, W // 2), np.int16)]
        return out

    def encode_pictures(self, frames):
        """frames: iterable of (y, cb, cr).  Returns the per-picture results in order (batches of pictures_in_flight)."""
        results = []
        for i, (y, cb, cr) in enumerate(frames):
            while True:
                try:
                    self.submit(i, y, cb, cr)
                    break
                except WrencB200Full:  # every batch slot is in flight: take the oldest batch's pictures first
                    for _ in range(min(self.pending(), self.pictures_in_flight)):
                        results.append(self.receive())
        while self.pending():
            results.append(self.receive())
        return results

    # ---- device-resident path (torch tensors or raw device pointers) ----
    def search_resident(self, n_pictures, d_yuv, d_rec, d_levels, d_records, stream=None):
        def p(t):
            return C.c_void_p(t.data_ptr() if hasattr(t, "data_ptr") else int(t))
        st = C.c_void_p(int(stream)) if stream else None
        return self._check(self.L.wrenc_b200_search_resident(self.h, int(n_pictures), p(d_yuv), p(d_rec), p(d_levels), p(d_records), st))

    def code_resident(self, n_pictures, d_levels, d_records, d_out, out_cap, d_out_len, stream=None):
        def p(t):
            return C.c_void_p(t.data_ptr() if hasattr(t, "data_ptr") else int(t))
        st = C.c_void_p(int(stream)) if stream else None
        return self._check(self.L.wrenc_b200_code_resident(self.h, int(n_pictures), p(d_levels), p(d_records), p(d_out), int(out_cap), p(d_out_len), st))

    def code_resident_retry(self, n_pictures, d_levels, d_records, d_out, out_cap, d_out_len, stream=None):
        """After code_resident reported -2 lengths (bin arena too small for this batch): grow it and code again (blocks)."""
        def p(t):
            return C.c_void_p(t.data_ptr() if hasattr(t, "data_ptr") else int(t))
        st = C.c_void_p(int(stream)) if stream else None
        return self._check(self.L.wrenc_b200_code_resident_retry(self.h, int(n_pictures), p(d_levels), p(d_records), p(d_out), int(out_cap), p(d_out_len), st))

    def prepare(self, n_pictures):
        """Allocate the resident workspace and upload the work list for batches of n_pictures (blocking), so that the resident calls only enqueue."""
        return self._check(self.L.wrenc_b200_prepare(self.h, int(n_pictures)))

    # ---- per-block entry points ----
    def block_predict(self, rec, x, y, w, tree, ar, bl, c, mode):
        i420 = np.concatenate([np.ascontiguousarray(a, np.uint8).ravel() for a in rec])
        n = w if c == 0 else w // 2
        out = np.zeros((n, n), np.uint8)
        self._check(self.L.wrenc_b200_block_predict(self.h, _ptr(i420), x, y, w, tree, int(ar), int(bl), c, mode, _ptr(out)))
        return out

    def _block16(self, fn, blocks):
        b = np.ascontiguousarray(blocks, np.int16)
        if b.ndim == 2:
            b = b[None]
        n = b.shape[-1]
        out = np.zeros_like(b)
        self._check(fn(self.h, _ptr(b), int(np.log2(n)), b.shape[0], _ptr(out)))
        return out

    def block_fwd_dct(self, res):
        return self._block16(self.L.wrenc_b200_block_fwd_dct, res)

    def block_inv_dct(self, deq):
        return self._block16(self.L.wrenc_b200_block_inv_dct, deq)

    def block_dequantize(self, lev):
        return self._block16(self.L.wrenc_b200_block_dequantize, lev)

    def block_quantize(self, coef):
        b = np.ascontiguousarray(coef, np.int16)
        if b.ndim == 2:
            b = b[None]
        n = b.shape[-1]
        out = np.zeros_like(b)
        rates = np.zeros(b.shape[0], np.int32)
        self._check(self.L.wrenc_b200_block_quantize(self.h, _ptr(b), int(np.log2(n)), b.shape[0], _ptr(out), _ptr(rates)))
        return out, rates
