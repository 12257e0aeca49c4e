"""ORACLE (test infrastructure only): numpy restatement of the baseline-JPEG decode that Pillow performs for
`Image.open(path).convert("RGB")` (/root/reference/main.py:330-334 `load_image`, main.py:165 / 412 the callers) - row N2
of SURVEY.md 8(f), image ingest.

The algorithm lives in a third-party dependency that is absent from /root/reference: Pillow (12.2.0 in this image) linked
against libjpeg-turbo (API level 6.2).  Pillow's JpegDecode.c calls jpeg_read_header / jpeg_start_decompress with the
library defaults: dct_method = JDCT_ISLOW, do_fancy_upsampling = TRUE, out_color_space = JCS_RGB.  What is restated here is
libjpeg-turbo's published algorithm for that configuration (its SIMD paths are bit-identical to the C ones by design):

  entropy decoding   jdhuff.c      sequential Huffman, DC prediction per component, restart intervals (ITU T.81 F.2.2)
  dequant + IDCT     jidctint.c    jpeg_idct_islow: 13-bit constants, PASS1_BITS = 2, range-limit table with & RANGE_MASK
  upsampling         jdsample.c    h2v1_fancy_upsample / h2v2_fancy_upsample (triangle filter), fullsize copy;
                     jdmainct.c    context rows at the image top / bottom replicate the first / last REAL sample row
  colour             jdcolor.c     ycc_rgb_convert with the 16.16 fixed-point tables;  grayscale -> R = G = B = Y

Pinned by (tests/test_oracle.py::test_jpeg_oracle_*, oracle/pin_jpeg.py): bit-exact equality with Pillow itself - the real
reference decoder, present in this image and on the GPU box - on the JPEG fixtures under tests/golden/jpeg/, on files Pillow
encodes on the fly (4:4:4 / 4:2:2 / 4:2:0 / grayscale, restart intervals, odd sizes, optimised tables), and, in the build
container, on every baseline file of the reference's own dataset (140 of 151; the other 11 are progressive and stay on the
host path in the product too).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

ZIGZAG = np.array([
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55,
    62, 63], dtype=np.int64)   # jutils.c jpeg_natural_order


class Unsupported(Exception):
    """not baseline / extended-sequential Huffman 8-bit, or a sampling layout outside {1x1, 2x1, 2x2 luma} (product: host path)"""


class Header:
    def __init__(self):
        self.width = self.height = 0
        self.comps: List[Tuple[int, int, int, int]] = []       # (id, h, v, tq)
        self.qt: Dict[int, np.ndarray] = {}                    # natural order, int32 [64]
        self.huff: Dict[Tuple[int, int], Tuple[np.ndarray, np.ndarray]] = {}   # (class, id) -> (bits[1..16], vals)
        self.restart_interval = 0
        self.scan_comps: List[Tuple[int, int, int]] = []       # (component index, dc table, ac table)
        self.scan_offset = 0                                   # first entropy-coded byte
        self.adobe_transform: Optional[int] = None
        self.jfif = False


def parse_header(data: bytes) -> Header:
    """marker segments up to and including the first SOS (jdmarker.c read_markers)"""
    if data[:2] != b"\xff\xd8":
        raise Unsupported("not a JPEG")
    h = Header()
    pos = 2
    n = len(data)
    while True:
        while pos < n and data[pos] != 0xFF:
            pos += 1                                   # next_marker skips garbage
        while pos < n and data[pos] == 0xFF:
            pos += 1
        if pos >= n:
            raise Unsupported("no SOS")
        m = data[pos]
        pos += 1
        if m == 0xD8 or (0xD0 <= m <= 0xD7) or m == 0x01:
            continue
        if m == 0xD9:
            raise Unsupported("EOI before SOS")
        (seglen,) = struct.unpack(">H", data[pos:pos + 2])
        seg = data[pos + 2:pos + seglen]
        if m == 0xDB:                                  # DQT
            q = 0
            while q < len(seg):
                pq, tq = seg[q] >> 4, seg[q] & 15
                q += 1
                t = np.zeros(64, np.int32)
                for i in range(64):
                    if pq:
                        t[ZIGZAG[i]] = (seg[q] << 8) | seg[q + 1]
                        q += 2
                    else:
                        t[ZIGZAG[i]] = seg[q]
                        q += 1
                h.qt[tq] = t
        elif m in (0xC0, 0xC1):                        # SOF0 / SOF1: sequential Huffman
            if seg[0] != 8:
                raise Unsupported("sample precision")
            h.height, h.width = struct.unpack(">HH", seg[1:5])
            for c in range(seg[5]):
                cid, hv, tq = seg[6 + 3 * c:9 + 3 * c]
                h.comps.append((cid, hv >> 4, hv & 15, tq))
        elif 0xC2 <= m <= 0xCF and m not in (0xC4, 0xC8, 0xCC):
            raise Unsupported("SOF%d (progressive / lossless / arithmetic)" % (m - 0xC0))
        elif m == 0xCC:
            raise Unsupported("arithmetic conditioning")
        elif m == 0xC4:                                # DHT
            q = 0
            while q < len(seg):
                tc, th = seg[q] >> 4, seg[q] & 15
                bits = np.frombuffer(seg[q + 1:q + 17], np.uint8).astype(np.int64)
                cnt = int(bits.sum())
                vals = np.frombuffer(seg[q + 17:q + 17 + cnt], np.uint8).astype(np.int64)
                h.huff[(tc, th)] = (bits, vals)
                q += 17 + cnt
        elif m == 0xDD:
            (h.restart_interval,) = struct.unpack(">H", seg[:2])
        elif m == 0xE0 and seg[:5] == b"JFIF\0":
            h.jfif = True
        elif m == 0xEE and seg[:5] == b"Adobe" and len(seg) >= 12:
            h.adobe_transform = seg[11]
        elif m == 0xDA:                                # SOS
            ns = seg[0]
            ids = [cc[0] for cc in h.comps]
            for s in range(ns):
                cs, tt = seg[1 + 2 * s], seg[2 + 2 * s]
                h.scan_comps.append((ids.index(cs), tt >> 4, tt & 15))
            h.scan_offset = pos + seglen
            return h
        pos += seglen


def _luts(bits: np.ndarray, vals: np.ndarray):
    """canonical code tables of ITU T.81 C.2 / F.2.2.3 (jdhuff.c jpeg_make_d_derived_tbl): mincode / maxcode / valptr per length"""
    code = 0
    k = 0
    mincode = [0] * 17
    maxcode = [-1] * 18
    valptr = [0] * 17
    for l in range(1, 17):
        valptr[l] = k
        mincode[l] = code
        code += int(bits[l - 1])
        k += int(bits[l - 1])
        maxcode[l] = code - 1 if bits[l - 1] else -1
        code <<= 1
    return mincode, maxcode, valptr, [int(v) for v in vals]


class _Bits:
    """bit reader over the entropy-coded segment: FF00 -> FF, stops at a marker (jdhuff.c jpeg_fill_bit_buffer)"""

    def __init__(self, data: bytes, pos: int):
        self.d, self.p, self.acc, self.n = data, pos, 0, 0
        self.marker = 0

    def _fill(self):
        while self.n <= 24:
            if self.marker or self.p >= len(self.d):
                b = 0                                   # libjpeg feeds zeros after a marker / EOF
            else:
                b = self.d[self.p]
                self.p += 1
                if b == 0xFF:
                    b2 = self.d[self.p] if self.p < len(self.d) else 0xD9
                    while b2 == 0xFF:                   # fill bytes
                        self.p += 1
                        b2 = self.d[self.p] if self.p < len(self.d) else 0xD9
                    self.p += 1
                    if b2 == 0:
                        b = 0xFF
                    else:
                        self.marker = b2
                        b = 0
            self.acc = ((self.acc << 8) | b) & 0xFFFFFFFFFF
            self.n += 8

    def get(self, k: int) -> int:
        if k == 0:
            return 0
        if self.n < k:
            self._fill()
        self.n -= k
        return (self.acc >> self.n) & ((1 << k) - 1)

    def decode(self, tbl) -> int:
        mincode, maxcode, valptr, vals = tbl
        code = 0
        for l in range(1, 17):
            code = (code << 1) | self.get(1)
            if maxcode[l] >= 0 and code <= maxcode[l] and code >= mincode[l]:
                return vals[valptr[l] + code - mincode[l]]
        return 0                                       # corrupt data: libjpeg warns and returns 0

    def restart(self):
        """byte-align, consume the RSTn marker (jdhuff.c process_restart)"""
        self.n = 0
        self.acc = 0
        if not self.marker:                            # marker not yet reached: scan for it (jdmarker.c next_marker)
            while self.p + 1 < len(self.d) and not (self.d[self.p] == 0xFF and self.d[self.p + 1] not in (0, 0xFF)):
                self.p += 1
            self.p += 2
        self.marker = 0


def _extend(v: int, s: int) -> int:
    return v if v >= (1 << (s - 1)) else v - (1 << s) + 1      # HUFF_EXTEND


def decode_coefficients(data: bytes, h: Header):
    """-> per component int16 [blocks_h, blocks_w, 64] in NATURAL order, quantised (jdhuff.c decode_mcu, jdcoefct.c decompress_onepass)"""
    hmax = max(c[1] for c in h.comps)
    vmax = max(c[2] for c in h.comps)
    mcux = -(-h.width // (8 * hmax))
    mcuy = -(-h.height // (8 * vmax))
    single = len(h.scan_comps) == 1
    if len(h.scan_comps) != len(h.comps):
        raise Unsupported("multi-scan sequential file")
    if single:   # non-interleaved scan: MCU = one block, the grid is the component's own block grid (jdinput.c per_scan_setup)
        mcux = -(-h.width // 8)
        mcuy = -(-h.height // 8)
    coefs = []
    for (_, ch, cv, _) in h.comps:
        bw, bh = (mcux, mcuy) if single else (mcux * ch, mcuy * cv)
        coefs.append(np.zeros((bh, bw, 64), np.int16))
    br = _Bits(data, h.scan_offset)
    tabs = {k: _luts(*v) for k, v in h.huff.items()}
    pred = [0] * len(h.comps)
    todo = h.restart_interval
    for my in range(mcuy):
        for mx in range(mcux):
            if h.restart_interval and todo == 0:
                br.restart()
                pred = [0] * len(h.comps)
                todo = h.restart_interval
            for (ci, td, ta) in h.scan_comps:
                _, ch, cv, _ = h.comps[ci]
                nh, nv = (1, 1) if single else (ch, cv)
                for by in range(nv):
                    for bx in range(nh):
                        blk = coefs[ci][my * nv + by, mx * nh + bx]
                        s = br.decode(tabs[(0, td)])
                        diff = _extend(br.get(s), s) if s else 0
                        pred[ci] += diff
                        blk[0] = np.int16(pred[ci])    # libjpeg stores (JCOEF) s
                        k = 1
                        tac = tabs[(1, ta)]
                        while k < 64:
                            rs = br.decode(tac)
                            r, s = rs >> 4, rs & 15
                            if s:
                                k += r
                                v = _extend(br.get(s), s)
                                if k < 64:
                                    blk[ZIGZAG[k]] = np.int16(v)
                                k += 1
                            else:
                                if r != 15:
                                    break
                                k += 16
            todo -= 1
    return coefs


# ---- jidctint.c ----------------------------------------------------------------------------------------------------
_F = dict(f0_298=2446, f0_390=3196, f0_541=4433, f0_765=6270, f0_899=7373, f1_175=9633, f1_501=12299, f1_847=15137,
          f1_961=16069, f2_053=16819, f2_562=20995, f3_072=25172)


def _idct_1d(c, shift_even_in: int, descale: int):
    """one pass of jpeg_idct_islow on arrays [..., 8] (int64 arithmetic stands in for the 32-bit JLONG: no overflow occurs on
    either side for dequantised 8-bit JPEG data, and arithmetic right shifts agree)"""
    z2, z3 = c[..., 2], c[..., 6]
    z1 = (z2 + z3) * _F["f0_541"]
    tmp2 = z1 + z3 * (-_F["f1_847"])
    tmp3 = z1 + z2 * _F["f0_765"]
    z2, z3 = c[..., 0], c[..., 4]
    tmp0 = (z2 + z3) << shift_even_in
    tmp1 = (z2 - z3) << shift_even_in
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = c[..., 7], c[..., 5], c[..., 3], c[..., 1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * _F["f1_175"]
    tmp0 = tmp0 * _F["f0_298"]
    tmp1 = tmp1 * _F["f2_053"]
    tmp2 = tmp2 * _F["f3_072"]
    tmp3 = tmp3 * _F["f1_501"]
    z1 = z1 * (-_F["f0_899"])
    z2 = z2 * (-_F["f2_562"])
    z3 = z3 * (-_F["f1_961"]) + z5
    z4 = z4 * (-_F["f0_390"]) + z5
    tmp0 += z1 + z3
    tmp1 += z2 + z4
    tmp2 += z2 + z3
    tmp3 += z1 + z4
    rnd = 1 << (descale - 1)
    out = np.stack([tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2,
                    tmp10 - tmp3], axis=-1)
    return (out + rnd) >> descale


def idct_islow(coef: np.ndarray, qt: np.ndarray) -> np.ndarray:
    """coef int16 [..., 64] natural order, qt [64] -> uint8 samples [..., 8, 8]  (jpeg_idct_islow incl. dequantisation and the
    range-limit table: index = value & RANGE_MASK into the post-IDCT table of jdmaster.c prepare_range_limit_table)"""
    x = (coef.astype(np.int64) * qt.astype(np.int64)).reshape(coef.shape[:-1] + (8, 8))   # [row, col]
    ws = _idct_1d(np.swapaxes(x, -1, -2), 13, 13 - 2)          # pass 1: columns -> ws[col][row-out]
    ws = np.swapaxes(ws, -1, -2)                                # back to [row][col]
    y = _idct_1d(ws, 13, 13 + 2 + 3)                            # pass 2: rows
    idx = y & 1023
    out = np.where(idx < 128, idx + 128, np.where(idx < 512, 255, np.where(idx < 896, 0, idx - 896)))
    return out.astype(np.uint8)


# ---- jdsample.c ----------------------------------------------------------------------------------------------------
def _h2v1_fancy(p: np.ndarray) -> np.ndarray:
    """[rows, w] -> [rows, 2w]"""
    p = p.astype(np.int32)
    w = p.shape[1]
    out = np.empty((p.shape[0], 2 * w), np.int32)
    left = np.concatenate([p[:, :1], p[:, :-1]], axis=1)
    right = np.concatenate([p[:, 1:], p[:, -1:]], axis=1)
    out[:, 0::2] = (3 * p + left + 1) >> 2
    out[:, 1::2] = (3 * p + right + 2) >> 2
    out[:, 0] = p[:, 0]
    out[:, -1] = p[:, -1]
    return out.astype(np.uint8)


def _h2v2_fancy(p: np.ndarray) -> np.ndarray:
    """[rows, w] (REAL rows only; the context rows above the first / below the last replicate them, jdmainct.c) -> [2 rows, 2w]"""
    p = p.astype(np.int32)
    rows, w = p.shape
    up = np.concatenate([p[:1], p[:-1]], axis=0)
    dn = np.concatenate([p[1:], p[-1:]], axis=0)
    out = np.empty((2 * rows, 2 * w), np.int32)
    for v, other in ((0, up), (1, dn)):
        cs = 3 * p + other                                          # thiscolsum
        last = np.concatenate([cs[:, :1], cs[:, :-1]], axis=1)
        nxt = np.concatenate([cs[:, 1:], cs[:, -1:]], axis=1)
        o = np.empty((rows, 2 * w), np.int32)
        o[:, 0::2] = (3 * cs + last + 8) >> 4
        o[:, 1::2] = (3 * cs + nxt + 7) >> 4
        o[:, 0] = (cs[:, 0] * 4 + 8) >> 4
        o[:, -1] = (cs[:, -1] * 4 + 7) >> 4
        out[v::2] = o
    return out.astype(np.uint8)


def _ycc_tables():
    """jdcolor.c build_ycc_rgb_table"""
    x = np.arange(256, dtype=np.int64) - 128
    fix = lambda v: int(v * 65536 + 0.5)
    half = 1 << 15
    return ((fix(1.40200) * x + half) >> 16, (fix(1.77200) * x + half) >> 16, -fix(0.71414) * x, -fix(0.34414) * x + half)


def decode_rgb(data: bytes) -> np.ndarray:
    """JPEG bytes -> uint8 [H, W, 3], what `Image.open(io.BytesIO(data)).convert("RGB")` returns"""
    h = parse_header(data)
    nc = len(h.comps)
    if nc not in (1, 3):
        raise Unsupported("%d components" % nc)
    hmax = max(c[1] for c in h.comps)
    vmax = max(c[2] for c in h.comps)
    if nc == 3:
        if (h.comps[1][1:3], h.comps[2][1:3]) != ((1, 1), (1, 1)) or (hmax, vmax) not in ((1, 1), (2, 1), (2, 2)):
            raise Unsupported("sampling layout")
        ids = [c[0] for c in h.comps]
        # jdapimin.c default_decompress_parms: JFIF or ids 1,2,3 -> YCbCr; Adobe transform 0 / ids 'R','G','B' -> RGB
        if not h.jfif and (h.adobe_transform == 0 or (h.adobe_transform is None and ids == [82, 71, 66])):
            raise Unsupported("RGB-coded JPEG")
    elif (hmax, vmax) != (1, 1):
        hmax = vmax = 1                                      # a lone component is always treated as 1x1 (jdinput.c)
    coefs = decode_coefficients(data, h)
    planes = []
    for ci, (_, ch, cv, tq) in enumerate(h.comps):
        s = idct_islow(coefs[ci], h.qt[tq])                   # [bh, bw, 8, 8]
        bh, bw = s.shape[:2]
        plane = s.transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)
        if nc == 1:
            ch = cv = 1
        dw = -(-h.width * ch // hmax)                         # downsampled_width / height: the REAL samples
        dh = -(-h.height * cv // vmax)
        plane = plane[:dh, :dw]
        # jdsample.c jinit_upsampler: the triangle filters are chosen only when downsampled_width > 2, else box replication
        if ch == hmax and cv == vmax:
            up = plane
        elif ch * 2 == hmax and cv == vmax:
            up = _h2v1_fancy(plane) if dw > 2 else np.repeat(plane, 2, axis=1)
        elif ch * 2 == hmax and cv * 2 == vmax:
            up = _h2v2_fancy(plane) if dw > 2 else np.repeat(np.repeat(plane, 2, axis=0), 2, axis=1)
        else:
            raise Unsupported("sampling layout")
        planes.append(up[:h.height, :h.width])
    if nc == 1:
        return np.repeat(planes[0][:, :, None], 3, axis=2)
    y = planes[0].astype(np.int64)
    cb, cr = planes[1].astype(np.int64), planes[2].astype(np.int64)
    cr_r, cb_b, cr_g, cb_g = _ycc_tables()
    r = np.clip(y + cr_r[cr], 0, 255)
    g = np.clip(y + ((cb_g[cb] + cr_g[cr]) >> 16), 0, 255)
    b = np.clip(y + cb_b[cb], 0, 255)
    return np.stack([r, g, b], axis=2).astype(np.uint8)
