"""B200 drop-in for python/payload_source_pdu.py: message handler -> crc32_async_bb(False) ->
pdu_to_tagged_stream(byte, "packet_len") (:38-45).

`post_message(addr, tpe, nr, payload)` keeps the signature of ofdm_cr_tools.message_handler.post_message
(python/ofdm_cr_tools.py:2124-2137): the PDU is [addr, tpe, nr] + payload bytes.  crc32_async_bb(False)
appends the in-graph CRC-32 (zlib, 4 bytes little-endian) -- computed on the GPU by OfdmPhy.crc32 when a
`phy` is given (batched over the queued PDUs), else with zlib on the host.  PDUs keep their own length:
pop_packets() returns one byte string per PDU (variable "packet_len" tags in the reference).
"""
import struct
import threading
import zlib


class payload_source_pdu(object):
    def __init__(self, callback=None, phy=None):
        self.callback = callback
        self.phy = phy
        self._q = []
        self._lock = threading.Lock()

    def post_message(self, addr, tpe, nr, payload):
        if isinstance(payload, str):
            payload = payload.encode('latin-1')
        with self._lock:
            self._q.append(bytes(bytearray([addr & 0xFF, tpe & 0xFF, nr & 0xFF])) + bytes(payload))

    def pop_packets(self, max_packets=None):
        with self._lock:
            k = len(self._q) if max_packets is None else min(max_packets, len(self._q))
            pdus, self._q = self._q[:k], self._q[k:]
        if not pdus:
            return []
        if self.phy is not None:
            crcs = self.phy.crc32(pdus)
        else:
            crcs = [zlib.crc32(p) & 0xFFFFFFFF for p in pdus]
        return [p + struct.pack("<I", int(c)) for p, c in zip(pdus, crcs)]
