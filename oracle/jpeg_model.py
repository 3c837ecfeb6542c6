"""jpeg_model.py -- TEST INFRASTRUCTURE (oracle): the pixel arithmetic of a baseline JPEG round trip
as `cv.imencode('.jpeg', mat, [IMWRITE_JPEG_QUALITY, q])` + `cv.imdecode` perform it through
libjpeg-turbo (vkit's jpeg_quality, vkit/mechanism/distortion/photometric/effect.py:26-55).

Entropy coding is lossless and left out; what changes pixels is restated with libjpeg's integer
arithmetic (libjpeg-turbo 3.x is not vendored in /root/reference; opencv-python-headless 4.13.0.92
bundles it):

  encoder  jccolor.c rgb_ycc_convert (16-bit fixed point) -> edge replication to whole MCUs ->
           jcsample.c h2v2_downsample (2x2 box, bias 1,2,1,2 ..) -> jfdctint.c (islow, x8 scaled)
           -> jcdctmgr.c quantisation with the jpeg_set_quality tables (Annex K scaled)
  decoder  dequantise -> jidctint.c (islow) + range limit -> jdsample.c h2v2_fancy_upsample
           (triangle filter) -> jdcolor.c ycc_rgb_convert
cv2 hands libjpeg its channels as B, G, R: vkit passes an RGB array, so the first channel plays
blue (kept as is: the reference does exactly that).

Pinned against the installed wheel by tests/test_oracle_cv2_model.py::test_jpeg_model_vs_cv2.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this package.
"""
import numpy as np

# Annex K tables (jcparam.c std_luminance_quant_tbl / std_chrominance_quant_tbl), natural order
STD_LUMA = np.array([
    16, 11, 10, 16, 24, 40, 51, 61,
    12, 12, 14, 19, 26, 58, 60, 55,
    14, 13, 16, 24, 40, 57, 69, 56,
    14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77,
    24, 35, 55, 64, 81, 104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101,
    72, 92, 95, 98, 112, 100, 103, 99], dtype=np.int64).reshape(8, 8)
STD_CHROMA = np.array([
    17, 18, 24, 47, 99, 99, 99, 99,
    18, 21, 26, 66, 99, 99, 99, 99,
    24, 26, 56, 99, 99, 99, 99, 99,
    47, 66, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99], dtype=np.int64).reshape(8, 8)


def quant_table(base: np.ndarray, quality: int) -> np.ndarray:
    """jpeg_set_quality(quality, force_baseline = TRUE)."""
    quality = min(max(int(quality), 1), 100)
    scale = 5000 // quality if quality < 50 else 200 - quality * 2
    return np.clip((base * scale + 50) // 100, 1, 255)


def _fix(x: float) -> int:
    return int(x * 65536 + 0.5)


def rgb_to_ycc(r, g, b):
    """jccolor.c rgb_ycc_convert: int64 arrays in, (y, cb, cr) uint8-valued int64 arrays out."""
    half, offset = 1 << 15, 128 << 16
    y = (_fix(0.29900) * r + _fix(0.58700) * g + _fix(0.11400) * b + half) >> 16
    cb = (-_fix(0.16874) * r - _fix(0.33126) * g + _fix(0.50000) * b + offset + half - 1) >> 16
    cr = (_fix(0.50000) * r - _fix(0.41869) * g - _fix(0.08131) * b + offset + half - 1) >> 16
    return y, cb, cr


def ycc_to_rgb(y, cb, cr):
    """jdcolor.c ycc_rgb_convert (tables built by build_ycc_rgb_table)."""
    half = 1 << 15
    x_cb, x_cr = cb - 128, cr - 128
    r = y + ((_fix(1.40200) * x_cr + half) >> 16)
    g = y + ((-_fix(0.34414) * x_cb + half - _fix(0.71414) * x_cr) >> 16)
    b = y + ((_fix(1.77200) * x_cb + half) >> 16)
    return np.clip(r, 0, 255), np.clip(g, 0, 255), np.clip(b, 0, 255)


def _pad_edges(plane: np.ndarray, multiple_h: int, multiple_w: int) -> np.ndarray:
    h, w = plane.shape
    ph, pw = -h % multiple_h, -w % multiple_w
    return np.pad(plane, ((0, ph), (0, pw)), mode='edge')


def h2v2_downsample(plane: np.ndarray) -> np.ndarray:
    """jcsample.c h2v2_downsample: 2x2 box with the alternating bias 1, 2, 1, 2, ..."""
    a = plane[0::2, 0::2] + plane[0::2, 1::2] + plane[1::2, 0::2] + plane[1::2, 1::2]
    bias = np.where(np.arange(a.shape[1]) % 2 == 0, 1, 2)[None, :]
    return (a + bias) >> 2


# jfdctint.c / jidctint.c constants (CONST_BITS = 13, PASS1_BITS = 2)
_C = {name: int(v * (1 << 13) + 0.5) for name, v in {
    'F_0_298': 0.298631336, 'F_0_390': 0.390180644, 'F_0_541': 0.541196100, 'F_0_765': 0.765366865,
    'F_0_899': 0.899976223, 'F_1_175': 1.175875602, 'F_1_501': 1.501321110, 'F_1_847': 1.847759065,
    'F_1_961': 1.961570560, 'F_2_053': 2.053119869, 'F_2_562': 2.562915447, 'F_3_072': 3.072711026}.items()}


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def fdct_islow(block: np.ndarray) -> np.ndarray:
    """jpeg_fdct_islow on (..., 8, 8) int64 samples already centred (sample - 128): x8 output."""
    c = _C
    d = block.astype(np.int64)

    def one_pass(d, first):
        t0, t7 = d[..., 0] + d[..., 7], d[..., 0] - d[..., 7]
        t1, t6 = d[..., 1] + d[..., 6], d[..., 1] - d[..., 6]
        t2, t5 = d[..., 2] + d[..., 5], d[..., 2] - d[..., 5]
        t3, t4 = d[..., 3] + d[..., 4], d[..., 3] - d[..., 4]
        t10, t13 = t0 + t3, t0 - t3
        t11, t12 = t1 + t2, t1 - t2
        out = [None] * 8
        if first:
            out[0] = (t10 + t11) << 2
            out[4] = (t10 - t11) << 2
            n = 13 - 2
        else:
            out[0] = _descale(t10 + t11, 2)
            out[4] = _descale(t10 - t11, 2)
            n = 13 + 2
        z1 = (t12 + t13) * c['F_0_541']
        out[2] = _descale(z1 + t13 * c['F_0_765'], n)
        out[6] = _descale(z1 + t12 * (-c['F_1_847']), n)
        z1, z2, z3, z4 = t4 + t7, t5 + t6, t4 + t6, t5 + t7
        z5 = (z3 + z4) * c['F_1_175']
        t4 = t4 * c['F_0_298']
        t5 = t5 * c['F_2_053']
        t6 = t6 * c['F_3_072']
        t7 = t7 * c['F_1_501']
        z1 = z1 * (-c['F_0_899'])
        z2 = z2 * (-c['F_2_562'])
        z3 = z3 * (-c['F_1_961']) + z5
        z4 = z4 * (-c['F_0_390']) + z5
        out[7] = _descale(t4 + z1 + z3, n)
        out[5] = _descale(t5 + z2 + z4, n)
        out[3] = _descale(t6 + z2 + z3, n)
        out[1] = _descale(t7 + z1 + z4, n)
        return np.stack(out, axis=-1)

    rows = one_pass(d, True)                                        # pass 1: rows
    cols = one_pass(np.swapaxes(rows, -1, -2), False)                # pass 2: columns
    return np.swapaxes(cols, -1, -2)


def quantize(coef: np.ndarray, qtbl: np.ndarray) -> np.ndarray:
    """jcdctmgr.c quantize for the islow DCT: divisor = q << 3, round half away from zero."""
    qval = qtbl << 3
    mag = (np.abs(coef) + (qval >> 1)) // qval
    return np.where(coef < 0, -mag, mag)


def idct_islow(coef: np.ndarray) -> np.ndarray:
    """jpeg_idct_islow on dequantised (..., 8, 8) coefficients: samples 0..255."""
    c = _C
    d = coef.astype(np.int64)

    def one_pass(d, first):
        z2, z3 = d[..., 2], d[..., 6]
        z1 = (z2 + z3) * c['F_0_541']
        t2 = z1 + z3 * (-c['F_1_847'])
        t3 = z1 + z2 * c['F_0_765']
        z2, z3 = d[..., 0], d[..., 4]
        t0 = (z2 + z3) << 13
        t1 = (z2 - z3) << 13
        t10, t13 = t0 + t3, t0 - t3
        t11, t12 = t1 + t2, t1 - t2
        t0, t1, t2, t3 = d[..., 7], d[..., 5], d[..., 3], d[..., 1]
        z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
        z5 = (z3 + z4) * c['F_1_175']
        t0 = t0 * c['F_0_298']
        t1 = t1 * c['F_2_053']
        t2 = t2 * c['F_3_072']
        t3 = t3 * c['F_1_501']
        z1 = z1 * (-c['F_0_899'])
        z2 = z2 * (-c['F_2_562'])
        z3 = z3 * (-c['F_1_961']) + z5
        z4 = z4 * (-c['F_0_390']) + z5
        t0 = t0 + z1 + z3
        t1 = t1 + z2 + z4
        t2 = t2 + z2 + z3
        t3 = t3 + z1 + z4
        n = 13 - 2 if first else 13 + 2 + 3
        out = [_descale(t10 + t3, n), _descale(t11 + t2, n), _descale(t12 + t1, n),
               _descale(t13 + t0, n), _descale(t13 - t0, n), _descale(t12 - t1, n),
               _descale(t11 - t2, n), _descale(t10 - t3, n)]
        return np.stack(out, axis=-1)

    cols = one_pass(np.swapaxes(d, -1, -2), True)                    # pass 1: columns
    rows = one_pass(np.swapaxes(cols, -1, -2), False)                # pass 2: rows
    # range_limit[(x) & RANGE_MASK] with the table centred on 128: clamp(x + 128)
    return np.clip(rows + 128, 0, 255)


def _blocks(plane: np.ndarray) -> np.ndarray:
    h, w = plane.shape
    return plane.reshape(h // 8, 8, w // 8, 8).swapaxes(1, 2)


def _unblocks(blocks: np.ndarray) -> np.ndarray:
    bh, bw = blocks.shape[:2]
    return blocks.swapaxes(1, 2).reshape(bh * 8, bw * 8)


def codec_plane(plane: np.ndarray, qtbl: np.ndarray) -> np.ndarray:
    """FDCT -> quantise -> dequantise -> IDCT of a plane whose sides are multiples of 8."""
    blocks = _blocks(plane.astype(np.int64)) - 128
    coef = quantize(fdct_islow(blocks), qtbl)
    return _unblocks(idct_islow(coef * qtbl))


def h2v2_fancy_upsample(plane: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """jdsample.c h2v2_fancy_upsample on the REAL chroma samples (ceil(h/2) x ceil(w/2)); rows above
    the first / below the last are the edge rows themselves (jdmainct.c context rows)."""
    p = plane.astype(np.int64)
    if p.shape[1] <= 2:
        # jinit_upsampler: the triangle filter needs downsampled_width > 2, narrower components
        # are box-replicated (h2v2_upsample)
        return np.repeat(np.repeat(p, 2, axis=0), 2, axis=1)[:out_h, :out_w]
    up = np.concatenate([p[:1], p[:-1]], axis=0)
    down = np.concatenate([p[1:], p[-1:]], axis=0)
    rows = np.empty((p.shape[0] * 2, p.shape[1]), dtype=np.int64)
    rows[0::2] = p * 3 + up      # upper output row of a pair: nearer = this, further = above
    rows[1::2] = p * 3 + down
    last = np.concatenate([rows[:, :1], rows[:, :-1]], axis=1)
    nxt = np.concatenate([rows[:, 1:], rows[:, -1:]], axis=1)
    out = np.empty((rows.shape[0], rows.shape[1] * 2), dtype=np.int64)
    out[:, 0::2] = (rows * 3 + last + 8) >> 4
    out[:, 1::2] = (rows * 3 + nxt + 7) >> 4
    # first / last column: the missing neighbour is the column itself (thiscolsum * 4)
    out[:, 0] = (rows[:, 0] * 4 + 8) >> 4
    out[:, -1] = (rows[:, -1] * 4 + 7) >> 4
    return out[:out_h, :out_w]


def jpeg_round_trip(mat: np.ndarray, quality: int) -> np.ndarray:
    """cv.imdecode(cv.imencode('.jpeg', mat, [IMWRITE_JPEG_QUALITY, quality])[1], IMREAD_UNCHANGED)
    for a uint8 HxWx3 (cv2 channel order: blue first) or HxW array."""
    q_luma, q_chroma = quant_table(STD_LUMA, quality), quant_table(STD_CHROMA, quality)
    if mat.ndim == 2:
        h, w = mat.shape
        y = _pad_edges(mat.astype(np.int64), 8, 8)
        return codec_plane(y, q_luma)[:h, :w].astype(np.uint8)
    h, w = mat.shape[:2]
    b, g, r = (mat[..., i].astype(np.int64) for i in range(3))  # cv2: channel 0 is blue
    y, cb, cr = rgb_to_ycc(r, g, b)
    # right edge: the full-resolution rows are replicated out to whole MCUs before the
    # downsampling (jcsample.c expand_right_edge); bottom edge: the colour buffer is replicated
    # to a whole row GROUP (2 rows), and the DOWNSAMPLED rows are then replicated to a whole iMCU
    # (jcprepct.c pre_process_data) -- so the last chroma row repeats, not the last image row
    y = _pad_edges(y, 16, 16)
    cb, cr = (_pad_edges(h2v2_downsample(_pad_edges(p, 2, 16)), 8, 8) for p in (cb, cr))
    y = codec_plane(y, q_luma)
    cb, cr = codec_plane(cb, q_chroma), codec_plane(cr, q_chroma)
    ch, cw = (h + 1) // 2, (w + 1) // 2
    cb = h2v2_fancy_upsample(cb[:ch, :cw], h, w)
    cr = h2v2_fancy_upsample(cr[:ch, :cw], h, w)
    r, g, b = ycc_to_rgb(y[:h, :w], cb, cr)
    return np.stack([b, g, r], axis=-1).astype(np.uint8)
