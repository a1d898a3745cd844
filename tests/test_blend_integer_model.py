"""Host-side proof of the arithmetic blend16_int_kernel (csrc/vu_composite.cu) relies on, over ALL 2^24 (alpha, fg, bg)
byte triples of the replace / fuse blend (tools/replace/replace.py:74-76, unscreen/utils/visualize.py:7-24):

* trunc of the reference's float64 expression equals floor(t / 255), t = c a + q (255 - a), except on some triples where 255
  divides t, and there it is exactly one less;
* the kernel's division, (t + 1 + (t >> 8)) >> 8, is floor(t / 255), and the LOW byte of that sum is zero exactly when
  255 | t and t > 0 - so "low byte zero on a soft pixel" flags every triple the float64 sequence rounds down;
* the loop's helpers: (t * 0x8081) >> 23 == t // 255 and (i * 171) >> 9 == i // 3 on their ranges.

The GPU test test_gpu_parity.py::test_replace_blend_every_triple_on_every_lane checks the kernel itself; this one pins the
model on the CPU, where the driver runs it every round."""
import numpy as np

from oracle import refport as R


def test_integer_blend_model_covers_every_triple():
    a = np.arange(256, dtype=np.int64)[:, None, None]
    c = np.arange(256, dtype=np.int64)[None, :, None]
    q = np.arange(256, dtype=np.int64)[None, None, :]
    m = a.astype(np.float64) / 255
    ref = (c.astype(np.float64) * m + q.astype(np.float64) * (1 - m)).astype(np.uint8).astype(np.int64)
    # the oracle's own function on a slice, so that `ref` is the expression the parity tests use
    al = np.full((256, 256), 77, np.uint8)
    fg = np.repeat(np.arange(256, dtype=np.uint8)[:, None, None], 256, 1).repeat(3, 2)
    bg = np.repeat(np.arange(256, dtype=np.uint8)[None, :, None], 256, 0).repeat(3, 2)
    assert np.array_equal(R.replace_blend(fg, al, bg)[..., 0], ref[77])
    t = c * a + q * (255 - a)
    s = t + 1 + (t >> 8)
    k = s >> 8
    assert np.array_equal(k, t // 255)
    low_zero = (s & 255) == 0
    assert np.array_equal(low_zero, (t % 255 == 0) & (t > 0))
    down = ref != k
    assert np.array_equal(ref[down], k[down] - 1)
    soft = np.broadcast_to((a > 0) & (a < 255), t.shape)
    assert not (down & ~(low_zero & soft)).any()          # every rounded-down triple is a candidate
    assert int(down.sum()) == 12397                       # the count DESIGN.md quotes
    assert 0.018 < float((low_zero & soft).mean()) < 0.020


def test_integer_blend_helpers():
    t = np.arange(65026, dtype=np.int64)
    assert np.array_equal((t * 0x8081) >> 23, t // 255)
    i = np.arange(48)
    assert np.array_equal((i * 171) >> 9, i // 3)
    # two 16-bit lanes in one register: [c0, c1] * a + [q0, q1] * (255 - a) never carries from the low lane into the high one
    assert 255 * 255 + 254 + 1 < 65536
