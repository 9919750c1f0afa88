"""B200 drop-in for python/payload_sink_pdu.py: tagged_stream_to_pdu -> crc32_async_bb(True) ->
message handler calling `callback(addr, tpe, nr, payload)` (:33-44, python/ofdm_cr_tools.py:2139-2151).

crc32_async_bb(True) drops PDUs whose trailing CRC-32 (zlib, little-endian) does not match and strips the
4 bytes from the ones that pass.  deliver() takes the packets an RX call decoded (RxResult.payloads() of a
PHY with crc_mode=0, i.e. the CRC still attached) and checks them on the GPU through OfdmPhy.crc32 when a
`phy` is given, else with zlib on the host.
"""
import struct
import zlib


class payload_sink_pdu(object):
    def __init__(self, callback=None, phy=None):
        self.callback = callback
        self.phy = phy
        self.n_rcvd = 0
        self.n_right = 0

    def deliver(self, packets):
        packets = [bytes(p) for p in packets]
        self.n_rcvd += len(packets)
        body = [p[:-4] for p in packets if len(p) >= 4]
        tail = [struct.unpack("<I", p[-4:])[0] for p in packets if len(p) >= 4]
        if self.phy is not None and body:
            crcs = [int(c) for c in self.phy.crc32(body)]
        else:
            crcs = [zlib.crc32(b) & 0xFFFFFFFF for b in body]
        out = []
        for b, c, t in zip(body, crcs, tail):
            if c != t or len(b) < 3:
                continue
            self.n_right += 1
            out.append(b)
            if self.callback is not None:
                self.callback(b[0], b[1], b[2], b[3:])
        return out
