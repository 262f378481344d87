"""csrc/hbt_inflate.cpp — the reader's own gzip decoder — against zlib (Python's gzip / zlib modules) on every kind of
stream the format allows: text like the reader's inputs, incompressible bytes (stored blocks), long runs (distance-1
matches, maximal lengths), every compression level and strategy (fixed and dynamic Huffman blocks, Z_HUFFMAN_ONLY,
Z_RLE), header fields (name, comment, extra, header CRC), concatenated members, plain files; and that damaged or
truncated streams are reported as errors, never returned as data."""
import ctypes
import gzip
import io
import os
import struct
import zlib

import numpy as np
import pytest

from hadronic_afterburner_toolkit_b200 import capi


def _lib():
    L = capi.lib()
    L.hbt_gz_open.restype = ctypes.c_void_p
    L.hbt_gz_open.argtypes = [ctypes.c_char_p]
    L.hbt_gz_read.restype = ctypes.c_long
    L.hbt_gz_read.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
    L.hbt_gz_error.restype = ctypes.c_char_p
    L.hbt_gz_error.argtypes = [ctypes.c_void_p]
    L.hbt_gz_close.argtypes = [ctypes.c_void_p]
    return L


def inflate_file(path, piece=1 << 16):
    """(data, error): everything hbt_gz_read returns for the file, in calls of `piece` bytes"""
    L = _lib()
    g = L.hbt_gz_open(str(path).encode())
    assert g
    buf = ctypes.create_string_buffer(piece)
    out = io.BytesIO()
    err = None
    while True:
        k = L.hbt_gz_read(g, buf, piece)
        if k < 0:
            err = L.hbt_gz_error(g).decode()
            break
        if k == 0:
            break
        out.write(buf.raw[:k])
    L.hbt_gz_close(g)
    return out.getvalue(), err


def deflate_raw(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=-15, memlevel=8):
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, memlevel, strategy)
    return c.compress(data) + c.flush()


def gz_wrap(raw_deflate, data, flags=0, extra=b"", name=b"", comment=b""):
    hdr = b"\x1f\x8b\x08" + bytes([flags]) + b"\0\0\0\0\0\x03"
    if flags & 4:
        hdr += struct.pack("<H", len(extra)) + extra
    if flags & 8:
        hdr += name + b"\0"
    if flags & 16:
        hdr += comment + b"\0"
    if flags & 2:
        hdr += struct.pack("<H", zlib.crc32(hdr) & 0xffff)
    return hdr + raw_deflate + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data) & 0xffffffff)


def particle_text(rng, n):
    rows = rng.normal(0, 1.3, (n, 8))
    return "".join("211 0.13800000000000001 " + " ".join("%.17g" % v for v in r) + "\n" for r in rows).encode()


def samples():
    rng = np.random.default_rng(5)
    text = b"1500\n" + particle_text(rng, 6000)
    return {
        "empty": b"",
        "one_byte": b"x",
        "particle_text": text,
        "incompressible": rng.integers(0, 256, 700000, dtype=np.uint8).tobytes(),
        "runs": b"\0" * 300000 + b"ab" * 150000 + b"abc" * 70000 + bytes(range(256)) * 900,
        "mixed": text[:200000] + rng.integers(0, 256, 100000, dtype=np.uint8).tobytes() + b"7" * 70000 + text[:50000],
        "long_codes": bytes(rng.choice(256, 400000, p=np.array([0.5 ** min(k + 1, 24) for k in range(256)]) /
                                       sum(0.5 ** min(k + 1, 24) for k in range(256))).astype(np.uint8)),
    }


SAMPLES = samples()


@pytest.mark.parametrize("name", sorted(SAMPLES))
@pytest.mark.parametrize("level,strategy", [(1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                                             (0, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE),
                                             (9, zlib.Z_FILTERED)])
def test_same_bytes_as_zlib(tmp_path, name, level, strategy):
    data = SAMPLES[name]
    f = tmp_path / "a.gz"
    f.write_bytes(gz_wrap(deflate_raw(data, level, strategy), data))
    assert gzip.decompress(f.read_bytes()) == data  # the stream is what zlib itself accepts
    for piece in (1 << 16, 977):
        got, err = inflate_file(f, piece)
        assert err is None, err
        assert got == data


def test_small_windows_and_memlevels(tmp_path):
    data = SAMPLES["mixed"]
    for wbits, memlevel in ((-9, 1), (-12, 4), (-15, 9)):
        f = tmp_path / f"w{-wbits}.gz"
        f.write_bytes(gz_wrap(deflate_raw(data, 6, zlib.Z_DEFAULT_STRATEGY, wbits, memlevel), data))
        got, err = inflate_file(f)
        assert err is None and got == data


def test_header_fields_members_and_plain_files(tmp_path):
    a, b = SAMPLES["particle_text"][:300000], SAMPLES["runs"][:200000]
    m1 = gz_wrap(deflate_raw(a), a, flags=4 | 8 | 16 | 2, extra=b"\x01\x02" * 300, name=b"particle_samples", comment=b"made by a test")
    m2 = gz_wrap(deflate_raw(b, 9), b, flags=8, name=b"second member")
    m3 = gz_wrap(deflate_raw(b""), b"")
    f = tmp_path / "members.gz"
    f.write_bytes(m1 + m2 + m3 + m1)
    got, err = inflate_file(f)
    assert err is None and got == a + b + a
    assert got == gzip.decompress(f.read_bytes())
    # python's gzip module as the writer
    g = tmp_path / "py.gz"
    with gzip.open(g, "wb", compresslevel=6) as h:
        h.write(a)
    got, err = inflate_file(g)
    assert err is None and got == a
    # not a gzip file: the bytes as they are (gzread's transparent mode)
    p = tmp_path / "plain.dat"
    p.write_bytes(a[:100001])
    got, err = inflate_file(p, 4099)
    assert err is None and got == a[:100001]
    # bytes that are no gzip member after a complete one are ignored
    t = tmp_path / "trailing.gz"
    t.write_bytes(m2 + b"\0" * 37)
    got, err = inflate_file(t)
    assert err is None and got == b


def test_a_member_may_not_reach_back_into_the_previous_one(tmp_path):
    a = SAMPLES["particle_text"][:100000]
    m1 = gz_wrap(deflate_raw(a), a)
    # a hand-made fixed-Huffman block whose first symbol is a match (length 3, distance 1): nothing to copy from
    bits = "1" + "10" + "0000001" + "00000"  # BFINAL = 1, BTYPE = 1 (two bits, LSB first), length code 257 = 0000001, distance code 0 = 00000
    bits += "0000000"                         # end of block (256)
    bits += "0" * (-len(bits) % 8)
    raw = bytes(int(bits[i:i + 8][::-1], 2) for i in range(0, len(bits), 8))
    f = tmp_path / "back.gz"
    f.write_bytes(m1 + gz_wrap(raw, b"xxx"))
    got, err = inflate_file(f)
    assert err is not None and "distance" in err
    with pytest.raises(zlib.error):
        gzip.decompress(f.read_bytes())


def test_damaged_streams_are_errors(tmp_path):
    rng = np.random.default_rng(9)
    data = SAMPLES["particle_text"][:400000]
    good = gz_wrap(deflate_raw(data), data)
    cases = {
        "truncated_in_the_data": good[: len(good) // 2],
        "truncated_in_the_trailer": good[:-3],
        "truncated_after_the_header": good[:10],
        "wrong_crc": good[:-8] + struct.pack("<I", (zlib.crc32(data) ^ 1) & 0xffffffff) + good[-4:],
        "wrong_length": good[:-4] + struct.pack("<I", len(data) + 1),
        "bad_method": good[:2] + b"\x07" + good[3:],
        "reserved_flag": good[:3] + b"\x20" + good[4:],
        "reserved_block_type": good[:10] + bytes([good[10] | 0x06]) + good[11:],
    }
    for name, blob in cases.items():
        f = tmp_path / (name + ".gz")
        f.write_bytes(blob)
        got, err = inflate_file(f)
        assert err is not None, name
        assert data.startswith(got), name  # whatever came out before the error is a prefix of the real data
    # random damage inside the compressed data: an error, or — when the flipped bits still form a valid stream — at
    # least never more bytes than zlib would accept (the CRC catches the rest)
    n_err = 0
    for k in range(40):
        blob = bytearray(good)
        pos = int(rng.integers(10, len(good) - 8))
        blob[pos] ^= 1 << int(rng.integers(0, 8))
        f = tmp_path / f"flip{k}.gz"
        f.write_bytes(bytes(blob))
        got, err = inflate_file(f)
        try:
            ref = gzip.decompress(bytes(blob))
        except Exception:
            ref = None
        if ref is None:
            assert err is not None, (k, pos)
            n_err += 1
        else:
            assert err is None and got == ref
    assert n_err >= 35


def test_stored_blocks_across_input_refills(tmp_path):
    """level 0 writes 65535-byte stored blocks; the decoder's input buffer (1 MiB) and output pieces (256 KiB) both
    end inside blocks many times over"""
    rng = np.random.default_rng(2)
    data = rng.integers(0, 256, 5_000_000, dtype=np.uint8).tobytes()
    f = tmp_path / "stored.gz"
    f.write_bytes(gz_wrap(deflate_raw(data, 0), data))
    got, err = inflate_file(f, 1 << 20)
    assert err is None and got == data


def test_large_text_file_matches_gzip(tmp_path):
    rng = np.random.default_rng(3)
    data = b"".join(b"1500\n" + particle_text(rng, 1500) for _ in range(8))
    f = tmp_path / "big.gz"
    with gzip.open(f, "wb", compresslevel=6) as h:
        h.write(data)
    got, err = inflate_file(f, 4 << 20)
    assert err is None and got == data


def test_randomised_sweep_against_zlib(tmp_path):
    """random sizes, entropies, levels, strategies, window sizes, memory levels and read sizes"""
    rng = np.random.default_rng(123)
    f = tmp_path / "x.gz"
    for case in range(90):
        kind = case % 6
        size = int(rng.integers(0, 400000))
        if kind == 0:
            data = rng.integers(0, 256, size, dtype=np.uint8).tobytes()
        elif kind == 1:
            data = bytes(rng.integers(48, 58, size, dtype=np.uint8))  # digits: one short match after the other
        elif kind == 2:
            data = (b"abcdefgh" * (size // 8 + 1))[:size]
        elif kind == 3:
            data = bytes(rng.integers(0, 4, size, dtype=np.uint8) * int(rng.integers(1, 60)))
        elif kind == 4:
            words = [bytes(rng.integers(97, 123, int(rng.integers(1, 12)), dtype=np.uint8)) for _ in range(200)]
            data = b" ".join(words[i] for i in rng.integers(0, 200, size // 6 + 1))[:size]
        else:
            data = bytes(np.repeat(rng.integers(0, 256, size // 50 + 1, dtype=np.uint8), rng.integers(1, 100, size // 50 + 1)))[:size]
        level = int(rng.integers(0, 10))
        strategy = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED][int(rng.integers(0, 5))]
        f.write_bytes(gz_wrap(deflate_raw(data, level, strategy, -int(rng.integers(9, 16)), int(rng.integers(1, 10))), data))
        got, err = inflate_file(f, int(rng.integers(1, 1 << 20)))
        assert err is None and got == data, (case, kind, size, level, strategy, err)
