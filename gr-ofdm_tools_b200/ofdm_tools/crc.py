"""MAC-level CRC-32 used by the reference's benchmark / MAC scripts through gnuradio.digital.crc
(examples/benchmarks.py:347, python/ofdm_cr_tools.py:1760 `digital.crc.gen_and_append_crc32`,
python/ofdm_cr_tools.py:1759 `digital.crc.check_crc32`).

This is NOT the in-graph crc32_bb (zlib, little-endian, computed on the GPU in the TX/RX kernels): it is
the MSB-first CRC-32 of gr-digital/lib/crc32.cc (poly 0x04C11DB7, init and final XOR 0xFFFFFFFF, check value
0xFC891918 for b"123456789"), appended big-endian.  It runs on the host, as in the
reference: it belongs to the MAC layer that builds the byte strings handed to payload_source.
"""
import struct


def _make_table():
    tab = []
    for i in range(256):
        c = i << 24
        for _ in range(8):
            c = ((c << 1) ^ 0x04C11DB7) & 0xFFFFFFFF if c & 0x80000000 else (c << 1) & 0xFFFFFFFF
        tab.append(c)
    return tab


_TABLE = _make_table()


def _b(s):
    return s.encode('latin-1') if isinstance(s, str) else bytes(s)


def crc32(s):
    """digital.crc32(s): unsigned 32-bit value."""
    crc = 0xFFFFFFFF
    for byte in _b(s):
        crc = ((crc << 8) & 0xFFFFFFFF) ^ _TABLE[((crc >> 24) ^ byte) & 0xFF]
    return crc ^ 0xFFFFFFFF


def gen_and_append_crc32(s):
    s = _b(s)
    return s + struct.pack(">I", crc32(s))


def check_crc32(s):
    s = _b(s)
    if len(s) < 4:
        return (False, b'')
    msg = s[:-4]
    (expected,) = struct.unpack(">I", s[-4:])
    return (crc32(msg) == expected, msg)
