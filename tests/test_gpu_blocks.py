"""GPU (-m gpu): every per-block kernel is bit-exact against the oracle on random blocks (north_star correctness level 1).
Calls go through the C ABI (wrenc_b200_block_*)."""
import numpy as np
import pytest

import wrenc_b200
from oracle_lib import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[22, 32, 37])
def pair(request):
    qp = request.param
    enc = wrenc_b200.SearchEncoder(96, 64, qp=qp)
    yield enc, Oracle(qp)
    enc.close()


@pytest.mark.parametrize("n", [4, 8, 16, 32])
def test_forward_inverse_dct(pair, n):
    enc, ora = pair
    rng = np.random.default_rng(n)
    res = rng.integers(-255, 256, (24, n, n)).astype(np.int16)
    res[0] = 0; res[1] = 255; res[2] = -255
    res[3] = np.where((np.add.outer(np.arange(n), np.arange(n)) & 1) == 0, 255, -255)
    g = enc.block_fwd_dct(res)
    o = np.stack([ora.fwd_dct(b) for b in res])
    assert np.array_equal(g, o)
    deq = rng.integers(-32768, 32768, (24, n, n)).astype(np.int16)  # extreme inputs exercise the int16 clamp of the first stage
    deq[:12] = o[:12]
    assert np.array_equal(enc.block_inv_dct(deq), np.stack([ora.inv_dct(b) for b in deq]))


@pytest.mark.parametrize("n,general", [(4, 0), (8, 0), (8, 1), (16, 0), (16, 1), (32, 0)])
def test_dep_quant_trellis_rate_dequant(pair, n, general, monkeypatch):
    """Dependent quantisation + rate + dequantisation of random, sparse, DCT-like and adversarial (DC-leaf wrap) blocks.  8x8 and 16x16
    TBs have two routines in the kernel: the state-per-lane chain the search uses and the general chunk-matrix routine; both are checked."""
    enc, ora = pair
    monkeypatch.setenv("WRENC_B200_BLOCK_GENERAL", str(general))
    rng = np.random.default_rng(100 + n)
    blocks = []
    for sigma in (1, 3, 10, 30, 100, 300, 1000):
        blocks.append(rng.normal(0, sigma, (n, n)))
        blocks.append(rng.normal(0, sigma, (n, n)) * (rng.random((n, n)) < 0.15))
        b = rng.normal(0, sigma, (n, n)) / (1 + np.add.outer(np.arange(n), np.arange(n)))  # energy compacted like a DCT
        blocks.append(b * 8)
    z = np.zeros((n, n)); blocks.append(z)
    for v in (1, -1, 7, -7, 40, -40, 455, -455, 457, -457, 3000):
        b = z.copy(); b[0, 0] = v; blocks.append(b)          # DC only: exercises the H3 leaf
        b = z.copy(); b[n - 1, n - 1] = v; blocks.append(b)  # last scan position only
        b = z.copy(); b[0, 0] = v; b[1, 1] = 900; blocks.append(b)
    coef = np.clip(np.rint(np.stack(blocks)), -32768, 32767).astype(np.int16)
    gq, gr = enc.block_quantize(coef)
    oq = np.stack([ora.quantize(b) for b in coef])
    assert np.array_equal(gq, oq)
    assert np.array_equal(gr.astype(np.int64), np.array([ora.rate(b) for b in oq], np.int64))
    assert np.array_equal(enc.block_dequantize(oq), np.stack([ora.dequantize(b) for b in oq]))


def test_prediction_all_modes_sizes_and_availability(pair):
    enc, ora = pair
    W, H = 96, 64
    rec = list(wrenc_b200.random_frame(W, H, 5))
    rec[0][:, 40:44] = 255; rec[0][30:33, :] = 0   # saturating edges exercise the clips
    bad = []
    cases = [(32, 32, 32), (32, 32, 16), (48, 32, 16), (32, 48, 16), (48, 48, 16), (40, 40, 8), (56, 40, 8), (36, 36, 4), (60, 60, 4), (0, 0, 32),
             (0, 32, 16), (64, 0, 8), (64, 32, 32), (88, 56, 8), (92, 60, 4), (0, 60, 4)]
    for (x0, y0, w) in cases:
        for ar in (0, 1):
            for bl in (0, 1):
                for c in ((0, 1, 2) if w >= 8 else (0,)):
                    for m in list(range(67)) + ([81, 82, 83] if c else []):
                        tree = 0 if w >= 8 else 1
                        g = enc.block_predict(rec, x0, y0, w, tree, ar, bl, c, m)
                        o = ora.predict(rec, x0, y0, w, tree, ar, bl, c, m)
                        if not np.array_equal(g, o):
                            bad.append((x0, y0, w, ar, bl, c, m))
    assert not bad, bad[:10]
