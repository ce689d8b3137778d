#!/usr/bin/env python3
"""Decode the reference's two test clips (assets/bus_352x288_30fps_30fr.mp4, assets/mobile_...mp4) to raw I420.

The reference's own evaluation feeds `ffmpeg -i clip.mp4 -f rawvideo -` into wrenc
(/root/reference/tools/evaluation/wrenc_fixed_qp.sh:3, scripts/intergration_test.sh:6), i.e. the H.264-decoded yuv420p
planes of the clip.  H.264 decoding is bit-exact by specification, so any conformant decoder gives the same planes.
This image has no ffmpeg binary and cv2.VideoCapture only hands out the luma plane un-converted, but the cv2 wheel
bundles libavformat/libavcodec (FFmpeg 8): this script drives them through ctypes and reads all three planes of every
AVFrame as they are (AV_PIX_FMT_YUV420P, no colour conversion, no scaling) — luma AND chroma exact.
The luma is cross-checked against cv2's CAP_PROP_CONVERT_RGB=0 output.

Usage: decode_assets.py [--assets DIR] [--out DIR]  -> DIR/bus_cif.yuv, DIR/mobile_cif.yuv + prints sha256.
Nothing here is product code: it prepares inputs for the oracle-pinning tests (tests/golden/README.md).
"""
import argparse
import ctypes as C
import glob
import hashlib
import os

import numpy as np

CLIPS = {"bus": "bus_352x288_30fps_30fr.mp4", "mobile": "mobile_352x288_30fps_30fr.mp4"}
W, H, FRAMES = 352, 288, 30


def _libs():
    import cv2
    d = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
    if not os.path.isdir(d):
        d = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python.libs")

    def load(stem):
        c = sorted(glob.glob(os.path.join(d, stem + "-*.so*")))
        if not c:
            raise RuntimeError("no bundled %s next to cv2 (%s)" % (stem, d))
        return C.CDLL(c[0], mode=C.RTLD_GLOBAL)
    avutil = load("libavutil")
    try:
        load("libswresample")
    except Exception:
        pass
    avcodec = load("libavcodec")
    avformat = load("libavformat")
    return avutil, avcodec, avformat


class _AVFrameHead(C.Structure):  # stable prefix of AVFrame (libavutil/frame.h): data, linesize, extended_data, width, height, nb_samples, format
    _fields_ = [("data", C.c_void_p * 8), ("linesize", C.c_int * 8), ("extended_data", C.c_void_p),
                ("width", C.c_int), ("height", C.c_int), ("nb_samples", C.c_int), ("format", C.c_int)]


class _AVPacketHead(C.Structure):  # stable prefix of AVPacket (libavcodec/packet.h)
    _fields_ = [("buf", C.c_void_p), ("pts", C.c_int64), ("dts", C.c_int64), ("data", C.c_void_p),
                ("size", C.c_int), ("stream_index", C.c_int)]


class _AVFormatContextHead(C.Structure):  # stable prefix of AVFormatContext (libavformat/avformat.h)
    _fields_ = [("av_class", C.c_void_p), ("iformat", C.c_void_p), ("oformat", C.c_void_p), ("priv_data", C.c_void_p),
                ("pb", C.c_void_p), ("ctx_flags", C.c_int), ("nb_streams", C.c_uint), ("streams", C.POINTER(C.c_void_p))]


class _AVStreamHead(C.Structure):  # FFmpeg >= 5: av_class, index, id, codecpar
    _fields_ = [("av_class", C.c_void_p), ("index", C.c_int), ("id", C.c_int), ("codecpar", C.c_void_p)]


def decode_clip(path, frames=FRAMES):
    """-> list of (Y, Cb, Cr) uint8 arrays, exactly the decoder's yuv420p output."""
    avutil, avcodec, avformat = _libs()
    vp = C.c_void_p
    avformat.avformat_open_input.argtypes = [C.POINTER(vp), C.c_char_p, vp, vp]
    avformat.avformat_find_stream_info.argtypes = [vp, vp]
    avformat.av_find_best_stream.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp), C.c_int]
    avformat.av_read_frame.argtypes = [vp, vp]
    avformat.avformat_close_input.argtypes = [C.POINTER(vp)]
    avcodec.avcodec_alloc_context3.argtypes = [vp]
    avcodec.avcodec_alloc_context3.restype = vp
    avcodec.avcodec_parameters_to_context.argtypes = [vp, vp]
    avcodec.avcodec_open2.argtypes = [vp, vp, vp]
    avcodec.avcodec_send_packet.argtypes = [vp, vp]
    avcodec.avcodec_receive_frame.argtypes = [vp, vp]
    avcodec.av_packet_alloc.restype = vp
    avcodec.av_packet_unref.argtypes = [vp]
    avcodec.avcodec_free_context.argtypes = [C.POINTER(vp)]
    avutil.av_frame_alloc.restype = vp
    avutil.av_frame_unref.argtypes = [vp]
    avutil.av_opt_set.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int]

    fmt = vp()
    if avformat.avformat_open_input(C.byref(fmt), path.encode(), None, None) < 0:
        raise RuntimeError("cannot open " + path)
    if avformat.avformat_find_stream_info(fmt, None) < 0:
        raise RuntimeError("no stream info")
    dec = vp()
    si = avformat.av_find_best_stream(fmt, 0, -1, -1, C.byref(dec), 0)  # AVMEDIA_TYPE_VIDEO = 0
    if si < 0 or not dec:
        raise RuntimeError("no video stream / decoder")
    head = C.cast(fmt, C.POINTER(_AVFormatContextHead)).contents
    st = C.cast(head.streams[si], C.POINTER(_AVStreamHead)).contents
    assert st.index == si, "AVStream layout mismatch"
    ctx = vp(avcodec.avcodec_alloc_context3(dec))
    if avcodec.avcodec_parameters_to_context(ctx, st.codecpar) < 0:
        raise RuntimeError("parameters_to_context")
    avutil.av_opt_set(ctx, b"threads", b"1", 0)
    if avcodec.avcodec_open2(ctx, dec, None) < 0:
        raise RuntimeError("avcodec_open2")
    pkt = vp(avcodec.av_packet_alloc())
    frm = vp(avutil.av_frame_alloc())
    out = []

    def drain():
        while len(out) < frames and avcodec.avcodec_receive_frame(ctx, frm) == 0:
            f = C.cast(frm, C.POINTER(_AVFrameHead)).contents
            assert f.format == 0, "expected AV_PIX_FMT_YUV420P (0), got %d" % f.format
            planes = []
            for i, (w, h) in enumerate(((f.width, f.height), (f.width // 2, f.height // 2), (f.width // 2, f.height // 2))):
                ls = f.linesize[i]
                buf = (C.c_uint8 * (ls * h)).from_address(f.data[i])
                planes.append(np.frombuffer(buf, dtype=np.uint8).reshape(h, ls)[:, :w].copy())
            out.append(tuple(planes))
            avutil.av_frame_unref(frm)

    while len(out) < frames and avformat.av_read_frame(fmt, pkt) >= 0:
        if C.cast(pkt, C.POINTER(_AVPacketHead)).contents.stream_index == si:
            if avcodec.avcodec_send_packet(ctx, pkt) < 0:
                raise RuntimeError("send_packet")
            drain()
        avcodec.av_packet_unref(pkt)
    avcodec.avcodec_send_packet(ctx, None)
    drain()
    avcodec.avcodec_free_context(C.byref(ctx))
    avformat.avformat_close_input(C.byref(fmt))
    if len(out) != frames:
        raise RuntimeError("decoded %d of %d frames" % (len(out), frames))
    return out


def cv2_luma(path, frames=FRAMES):
    import cv2
    cap = cv2.VideoCapture(path)
    cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
    ys = []
    for _ in range(frames):
        ok, f = cap.read()
        assert ok
        ys.append(f.reshape(-1)[: W * H].reshape(H, W).copy())
    return ys


def to_i420_bytes(fr):
    return b"".join(p.tobytes() for f in fr for p in f)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--assets", default="/root/reference/assets")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "assets"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    for name, fn in CLIPS.items():
        p = os.path.join(a.assets, fn)
        fr = decode_clip(p)
        ys = cv2_luma(p)
        assert all(np.array_equal(f[0], y) for f, y in zip(fr, ys)), "libavcodec luma != cv2 luma"
        raw = to_i420_bytes(fr)
        assert len(raw) == W * H * 3 // 2 * FRAMES
        dst = os.path.join(a.out, name + "_cif.yuv")
        with open(dst, "wb") as f:
            f.write(raw)
        print(name, dst, len(raw), "sha256", hashlib.sha256(raw).hexdigest())


if __name__ == "__main__":
    main()
