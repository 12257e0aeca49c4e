"""Test helper: a tiny baseline-JPEG *stream writer* that turns arbitrary quantised coefficient blocks into valid files with
features Pillow's own encoder never emits but third-party encoders (and 10 + 13 of the reference's 140 baseline dataset files)
do: Huffman tables using the full 16-bit code space, restart intervals of any length with fill bytes before the RSTn markers,
16-bit quantisation tables, tables redefined before the scan, unusual component / table ids, Adobe + JFIF marker combinations,
blocks that end exactly at coefficient 63, long zero runs (ZRL), large coefficient magnitudes.  Pillow (libjpeg-turbo) decodes
these files exactly as it decodes any other, so it stays the reference; the oracle and the CUDA decoder are compared with it."""
import struct

import numpy as np

ZIGZAG = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
          35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55,
          62, 63]


def deep_table(symbols):
    """canonical Huffman table over `symbols` (most frequent first) that reaches into the long codes a 10/11-bit look-up cannot
    hold: the first eight symbols get the lengths 2..9, the others are spread over the lengths 12, 14 and 16 (Kraft sum ~0.52:
    an incomplete code, which T.81 and libjpeg accept)"""
    symbols = list(symbols)
    bits = [0] * 16
    head = min(8, len(symbols))
    for i in range(head):
        bits[1 + i] = 1                      # lengths 2..9
    rest = len(symbols) - head
    third = rest // 3
    if rest:
        bits[11] = third                     # length 12
        bits[13] = third                     # length 14
        bits[15] = rest - 2 * third          # length 16
    return bits, symbols


def flat_table(symbols, length=8):
    symbols = list(symbols)
    assert len(symbols) < (1 << length)
    bits = [0] * 16
    bits[length - 1] = len(symbols)
    return bits, symbols


def codes_of(bits, vals):
    out, code, k = {}, 0, 0
    for length in range(1, 17):
        for _ in range(bits[length - 1]):
            out[vals[k]] = (code, length)
            code += 1
            k += 1
        code <<= 1
    return out


class BitWriter:
    def __init__(self):
        self.buf = bytearray()
        self.acc = 0
        self.n = 0

    def put(self, value, length):
        if length == 0:
            return
        self.acc = (self.acc << length) | (value & ((1 << length) - 1))
        self.n += length
        while self.n >= 8:
            b = (self.acc >> (self.n - 8)) & 0xFF
            self.buf.append(b)
            if b == 0xFF:
                self.buf.append(0)
            self.n -= 8
        self.acc &= (1 << self.n) - 1

    def flush(self):
        if self.n:
            self.put((1 << (8 - self.n)) - 1, 8 - self.n)      # pad with ones


def _size_bits(v):
    a = abs(int(v))
    s = a.bit_length()
    return s, (v if v >= 0 else v + (1 << s) - 1)


def write_jpeg(coefs, width, height, sampling=(2, 2), qtabs=None, tables=None, restart=0, comp_ids=(1, 2, 3), table_ids=((0, 0), (1, 1), (1, 1)),
               jfif=True, adobe=None, dqt16=False, fill_before_rst=0, redefine=False, gray=False):
    """coefs: list per component of int arrays [blocks_h, blocks_w, 64] in ZIG-ZAG order (quantised values).  tables: dict
    {(class, id): (bits, vals)}.  Returns the file bytes."""
    ncomp = 1 if gray else 3
    hs, vs = (1, 1) if gray else sampling
    samp = [(hs, vs), (1, 1), (1, 1)][:ncomp]
    mcux = -(-width // (8 * hs))
    mcuy = -(-height // (8 * vs))
    qtabs = qtabs or [np.full(64, 3, np.int64), np.full(64, 5, np.int64)]
    out = bytearray(b"\xff\xd8")
    if jfif:
        out += b"\xff\xe0" + struct.pack(">H", 16) + b"JFIF\0\x01\x01\x00\x00\x01\x00\x01\x00\x00"
    if adobe is not None:
        out += b"\xff\xee" + struct.pack(">H", 14) + b"Adobe" + struct.pack(">HHHB", 100, 0, 0, adobe)
    out += b"\xff\xfe" + struct.pack(">H", 2 + 7) + b"crafted"

    def dqt(tid, q):
        seg = bytearray([(16 if dqt16 else 0) | tid])
        for k in range(64):
            seg += struct.pack(">H", int(q[k])) if dqt16 else bytes([int(q[k])])
        return b"\xff\xdb" + struct.pack(">H", 2 + len(seg)) + bytes(seg)

    if redefine:                                                # a first definition that must be overridden by the second
        out += dqt(0, np.full(64, 99, np.int64))
    out += dqt(0, qtabs[0])
    if not gray:
        out += dqt(1, qtabs[1])
    out += b"\xff\xc0" + struct.pack(">HBHHB", 8 + 3 * ncomp, 8, height, width, ncomp)
    for c in range(ncomp):
        out += bytes([comp_ids[c], (samp[c][0] << 4) | samp[c][1], 0 if c == 0 else 1])

    def dht(cls, tid, bv):
        bits, vals = bv
        return b"\xff\xc4" + struct.pack(">H", 2 + 17 + len(vals)) + bytes([(cls << 4) | tid]) + bytes(bits) + bytes(vals)

    used = sorted({(0, table_ids[c][0]) for c in range(ncomp)} | {(1, table_ids[c][1]) for c in range(ncomp)})
    if redefine:
        out += dht(0, table_ids[0][0], flat_table(range(12), 6))
    for (cls, tid) in used:
        out += dht(cls, tid, tables[(cls, tid)])
    if restart:
        out += b"\xff\xdd" + struct.pack(">HH", 4, restart)
    out += b"\xff\xda" + struct.pack(">HB", 6 + 2 * ncomp, ncomp)
    for c in range(ncomp):
        out += bytes([comp_ids[c], (table_ids[c][0] << 4) | table_ids[c][1]])
    out += bytes([0, 63, 0])
    enc = {k: codes_of(*v) for k, v in tables.items()}
    bw = BitWriter()
    pred = [0] * ncomp
    count = 0
    rst = 0
    for my in range(mcuy):
        for mx in range(mcux):
            if restart and count and count % restart == 0:
                bw.flush()
                out += bytes(bw.buf) + b"\xff" * fill_before_rst + bytes([0xFF, 0xD0 + rst])
                rst = (rst + 1) & 7
                bw = BitWriter()
                pred = [0] * ncomp
            count += 1
            for c in range(ncomp):
                dc_codes, ac_codes = enc[(0, table_ids[c][0])], enc[(1, table_ids[c][1])]
                for by in range(samp[c][1]):
                    for bx in range(samp[c][0]):
                        blk = coefs[c][my * samp[c][1] + by, mx * samp[c][0] + bx]
                        s, v = _size_bits(int(blk[0]) - pred[c])
                        pred[c] = int(blk[0])
                        bw.put(*dc_codes[s])
                        bw.put(v, s)
                        run = 0
                        last = max([k for k in range(1, 64) if blk[k] != 0], default=0)
                        for k in range(1, last + 1):
                            if blk[k] == 0:
                                run += 1
                                continue
                            while run > 15:
                                bw.put(*ac_codes[0xF0])
                                run -= 16
                            s, v = _size_bits(int(blk[k]))
                            bw.put(*ac_codes[(run << 4) | s])
                            bw.put(v, s)
                            run = 0
                        if last < 63:
                            bw.put(*ac_codes[0x00])
    bw.flush()
    out += bytes(bw.buf) + b"\xff\xd9"
    return bytes(out)


def random_case(rng, width, height, sampling=(2, 2), gray=False, deep=True, **kw):
    """random but LEGIT coefficient blocks (dequantised values small enough that every decoder agrees on the samples): mostly
    sparse blocks with short runs, some dense ones, some with runs > 16 (ZRL), some filled up to coefficient 63"""
    hs, vs = (1, 1) if gray else sampling
    mcux = -(-width // (8 * hs))
    mcuy = -(-height // (8 * vs))
    ncomp = 1 if gray else 3
    shapes = [(mcuy * vs, mcux * hs)] + [(mcuy, mcux)] * (ncomp - 1)
    coefs = []
    for (bh, bwid) in shapes:
        c = np.zeros((bh, bwid, 64), np.int64)
        dc = np.cumsum(rng.integers(-6, 7, size=bh * bwid)).reshape(bh, bwid)
        c[..., 0] = np.clip(dc, -120, 120)
        kind = rng.integers(0, 10, size=(bh, bwid))
        for y in range(bh):
            for x in range(bwid):
                k = kind[y, x]
                if k < 5:                                   # sparse low frequencies
                    idx = rng.integers(1, 12, size=rng.integers(0, 5))
                    c[y, x, idx] = rng.integers(-9, 10, size=len(idx))
                elif k < 7:                                 # long runs: ZRL symbols
                    idx = rng.integers(20, 64, size=2)
                    c[y, x, idx] = rng.integers(-3, 4, size=2)
                elif k < 8:                                 # dense, ends at coefficient 63: no EOB
                    c[y, x, 1:] = rng.integers(-2, 3, size=63)
                    c[y, x, 63] = 1
                elif k < 9:                                 # one large magnitude (size 7-8 values)
                    c[y, x, rng.integers(1, 6)] = int(rng.integers(-130, 131))
        coefs.append(c)
    # every (run, size) symbol that can occur: sizes up to 8 for AC, up to 9 for DC differences
    ac_syms = [0x00, 0xF0] + [(r << 4) | s for s in range(1, 9) for r in range(16)]
    order = np.array(ac_syms)
    rng.shuffle(order[2:])
    dc_syms = list(range(0, 10))
    make = deep_table if deep else (lambda s: flat_table(s, 9))
    tables = {(0, 0): make(dc_syms), (0, 1): make(dc_syms[::-1]), (1, 0): make(list(order)), (1, 1): make(list(order[::-1]))}
    data = write_jpeg(coefs, width, height, sampling=sampling, tables=tables, gray=gray, **kw)
    return data
