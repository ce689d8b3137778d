"""Byte-stream NAL wrapper (host only): wrenc_b200_write_nal against a restatement of nal.rs:210-299 and hand-derived vectors."""
import numpy as np

import wrenc_b200


def ref_write_nal(payload, nal_unit_type=7, nuh_layer_id=9, nuh_temporal_id=0):
    """nal.rs:210-299, byte for byte: 00 00 00 | 00 00 01 | header | payload with the idx + 3 < len emulation-prevention scan."""
    out = bytearray([0, 0, 0, 0, 0, 1])
    bits = [0, 0] + [(nuh_layer_id >> i) & 1 for i in range(5, -1, -1)] + [(nal_unit_type >> i) & 1 for i in range(4, -1, -1)] + \
        [((nuh_temporal_id + 1) >> i) & 1 for i in range(2, -1, -1)]
    for b in range(2):
        v = 0
        for bit in bits[8 * b: 8 * b + 8]:
            v = (v << 1) | bit
        out.append(v)
    idx = 0
    while idx + 3 < len(payload):
        if payload[idx] == 0 and payload[idx + 1] == 0 and payload[idx + 2] <= 3:
            out += bytes(payload[idx:idx + 2])
            idx += 2
            out.append(3)
        else:
            out.append(payload[idx])
            idx += 1
    out += bytes(payload[idx:])
    return bytes(out)


def test_known_vectors():
    hdr = bytes([0, 0, 0, 0, 0, 1, 0x09, 0x39])  # layer 9, IDR_W_RADL (7) << 3 | temporal id + 1
    assert wrenc_b200.write_nal(b"") == hdr
    assert wrenc_b200.write_nal(b"\x12\x34") == hdr + b"\x12\x34"
    # 00 00 01 inside the payload is escaped ...
    assert wrenc_b200.write_nal(bytes([0, 0, 1, 255, 255, 255, 255])) == hdr + bytes([0, 0, 3, 1, 255, 255, 255, 255])
    # ... but the scan stops three bytes before the end: the trailing 00 00 01 stays as it is
    assert wrenc_b200.write_nal(bytes([0, 0, 0, 0, 1])) == hdr + bytes([0, 0, 3, 0, 0, 1])
    assert wrenc_b200.write_nal(bytes([7, 0, 0, 2])) == hdr + bytes([7, 0, 0, 2])
    # parameter-set NAL units of main.rs:223-260: VPS with layer id 1, SPS / PPS with layer id 9
    assert wrenc_b200.write_nal(b"\xaa", nal_unit_type=14, nuh_layer_id=1)[6:8] == bytes([0x01, (14 << 3) | 1])
    assert wrenc_b200.write_nal(b"\xaa", nal_unit_type=19)[6:8] == bytes([0x09, (19 << 3) | 1])


def test_random_payloads_match_restatement():
    rng = np.random.default_rng(7)
    for n in list(range(0, 12)) + [64, 257, 4096]:
        for p_zero in (0.0, 0.5, 0.9):
            p = rng.integers(0, 5, n).astype(np.uint8)
            p[rng.random(n) < p_zero] = 0
            for t, lid, tid in ((7, 9, 0), (15, 9, 0), (0, 0, 6), (31, 63, 3)):
                assert wrenc_b200.write_nal(p.tobytes(), t, lid, tid) == ref_write_nal(p.tobytes(), t, lid, tid)
