"""B200 drop-in for python/payload_source.py: byte queue -> fixed-length tagged packets.

`send_pkt_s(payload, eof)` keeps the reference signature (python/payload_source.py:49-54).  As in
the reference (message_source -> stream_to_tagged_stream(char, 1, packet_len, "packet_len"), :41-42)
the byte stream is cut every `packet_len` bytes regardless of message boundaries.
"""
import threading


class payload_source(object):
    def __init__(self, packet_len=500):
        self.packet_len = packet_len
        self._buf = bytearray()
        self._eof = False
        self._lock = threading.Lock()

    def send_pkt_s(self, payload='', eof=False):
        with self._lock:
            if eof:
                self._eof = True
            else:
                self._buf += payload.encode('latin-1') if isinstance(payload, str) else bytes(payload)

    def pop_packets(self, max_packets=None):
        """Packets ready for the TX chain (each exactly packet_len bytes)."""
        out = []
        with self._lock:
            while len(self._buf) >= self.packet_len and (max_packets is None or len(out) < max_packets):
                out.append(bytes(self._buf[:self.packet_len]))
                del self._buf[:self.packet_len]
        return out

    def eof(self):
        return self._eof

    def get_packet_len(self):
        return self.packet_len

    def set_packet_len(self, packet_len):
        self.packet_len = packet_len
