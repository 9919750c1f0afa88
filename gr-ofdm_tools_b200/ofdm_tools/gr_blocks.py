"""GNU Radio leaf blocks around libofdmx.so: `ofdm_rx_b200` and `ofdm_tx_b200`.

These are the two blocks a maintainer drops into python/ofdm_txrx_modules.py of the reference in place of
everything `ofdm_rx` (:295-426) / `ofdm_tx` (:159-254) wire internally: same ports
(`gr.io_signature(1, 1, gr.sizeof_gr_complex)` -> `gr.io_signature(1, 1, gr.sizeof_char)` and back, :156-158,
:292-294), same length-tag convention (`packet_length_tag_key` on the first item of every packet,
python/payload_source.py:42), leaf-block shape as python/papr_sink.py:28-44 (`work()` over numpy views).

GNU Radio calls a block with arbitrary chunk sizes; the GPU chain wants long buffers and a frame may straddle
two calls.  The stream logic therefore lives in two plain classes that only see numpy arrays --

    RxStreamer.push(samples) -> [(absolute trigger index, record, payload bytes)]
    TxStreamer.push(bytes, tag offsets/lengths) -> [numpy complex64 burst]

-- whose results do not depend on how the stream was cut into calls (tests/test_gr_blocks.py drives them and
the block classes below with random chunk sizes).  `make_blocks(gr, pmt)` builds the gr.basic_block classes
from any module pair with GNU Radio's block API; at import time it is applied to the real `gnuradio` if that
can be imported (it cannot in this image: the classes are then None and `make_blocks` is what the tests call
with a stub).
"""
import numpy as np

from . import _lib
from .phy import FRAME_DTYPE


class RxStreamer(object):
    """Chunk-invariant receiver over `phy.rx_host` (ofdmx_rx_host).

    The sample stream is processed in windows.  A window is loaded with `lead` samples of history in front of
    the part it owns (so that the Schmidl & Cox window sums and the plateau detector are in the same state as
    in the unsplit stream: lead = fft_len + 2 cp_len + 2, as ofdm_tools.dist.plan_segments) and `tail` samples
    behind it (the longest configured frame plus one symbol, so that an owned trigger normally sees its whole
    frame; a header that announces more than that makes the streamer wait for those samples before it goes
    on, exactly as the demux does).  The library returns one record per plateau trigger (ofdmx_set_emit_all); the
    header_payload_demux rule (which triggers the demux examines) is then resumed on the host
    from the position the previous window left it at.  The emitted frames are exactly those of one call over
    the whole stream."""

    def __init__(self, phy, chunk=1 << 22, tail=None):
        self.phy = phy
        self.D = phy.fft_len + phy.cp_len
        self.lead = phy.fft_len + 2 * phy.cp_len + 2
        self.n_pre = 2 if getattr(phy, "n_sync_words", 2) == 2 else 1          # sync symbols in front of the header
        self.tail = int(phy.frame_samples(max(1, phy.max_pkt_bytes - (4 if phy.crc_mode else 0)))) + self.D
        if tail is not None:    # any value >= 3 symbols is exact (longer frames are waited for); it only sets how
            self.tail = max(int(tail), 3 * self.D + 2 * phy.cp_len)     # often a window has to be cut short
        self.min_end = 0        # absolute index the buffer must reach before the next window (a longer frame pending)
        self.holdoff = int(phy.params.demux_holdoff)
        self.chunk = max(int(chunk), self.D)
        self.buf = np.zeros(0, np.complex64)
        self.base = 0           # absolute index of buf[0]
        self.own_from = 0       # absolute index where the not yet processed part of the stream starts
        self.pos = 0            # demux resume position (absolute)
        self.stalled = False
        phy.set_emit_all(True)

    def pending(self):
        return self.base + len(self.buf) - self.own_from

    def push(self, samples, flush=False):
        """Append samples; returns the frames that became final: list of (trigger, record, payload-or-None).
        payload is None for a frame whose CRC-32 failed (the flowgraph drops it) or that is oversize."""
        if len(samples):
            self.buf = np.concatenate([self.buf, np.asarray(samples, np.complex64)])
        out = []
        while True:
            avail = self.base + len(self.buf) - self.own_from
            if flush:
                if avail <= 0:
                    break
                own_stop = self.base + len(self.buf)
            else:
                if avail < self.chunk + self.tail or self.base + len(self.buf) < self.min_end:
                    break
                own_stop = self.base + len(self.buf) - self.tail
            out.extend(self._window(own_stop, final=flush))
            if flush:
                break
        return out

    def _window(self, own_stop, final):
        L = len(self.buf)
        # one record per plateau trigger: sized for frame-dense input first, for the worst case (a trigger every
        # cp_len + 1 samples) if that overflows
        try:
            res = self.phy.rx_host(self.buf, max_frames=2 * self.phy.default_max_frames(1, L) + 64)
        except BufferError:
            res = self.phy.rx_host(self.buf, max_frames=int(L // (self.phy.cp_len + 1) + 16))
        rec = res.frames
        trig = rec["trigger"].astype(np.int64) + self.base
        frames = []
        for i in np.nonzero((trig >= self.own_from) & (trig < own_stop))[0]:
            t = int(trig[i])
            if t < self.pos or self.stalled:
                continue
            f = rec[i]
            fl = int(f["flags"])
            if not (fl & _lib.F_HDR_SEEN):
                # only at the very end of a flushed stream: the demux waits for samples that never come
                self.stalled = True
                continue
            if not (fl & _lib.F_HDR_OK):
                self.pos = t + 1
                continue
            nsy = int(f["frame_syms"])
            if not (fl & _lib.F_COMPLETE):
                if final:
                    self.stalled = True          # the demux waits for payload samples that never come
                    continue
                # the header announces a frame that ends behind this buffer: take the window up to this trigger
                # only and come back when those samples have arrived
                own_stop = t
                self.min_end = t + (self.n_pre + 2 + nsy) * self.D
                break
            self.pos = t + (self.n_pre + 1 + nsy) * self.D - self.holdoff if nsy > 0 else t + (self.n_pre + 1) * self.D
            payload = None
            n = int(f["pkt_len"])
            ok = not (fl & _lib.F_OVERSIZE) and (not self.phy.crc_mode or (fl & _lib.F_CRC_OK))
            if ok:
                if self.phy.crc_mode:
                    n -= 4
                payload = bytes(np.asarray(res.slots[int(f["slot"])][:n]))
            g = f.copy()
            g["trigger"] = t
            g["flags"] = fl | _lib.F_ACCEPTED
            frames.append((t, g, payload))
        if not final:
            self.own_from = own_stop
            keep_from = own_stop - self.lead            # absolute index of the first sample the next window loads
            if keep_from > self.base:
                self.buf = self.buf[keep_from - self.base:]
                self.base = keep_from
        else:
            self.own_from = self.base + L
        return frames


class TxStreamer(object):
    """Tagged byte stream -> bursts.  Bytes arrive in arbitrary pieces; a packet is `length` bytes starting at a
    tagged item.  Whole packets are modulated in one ofdmx_tx call per push (the header counter continues
    across calls, packet_header_default)."""

    def __init__(self, phy):
        self.phy = phy
        self.bytes = bytearray()
        self.base = 0               # absolute item index of self.bytes[0]
        self.starts = []            # (absolute offset, length) of announced packets, in order
        self.n_sent = 0

    def push(self, data, tags=()):
        """data: bytes-like; tags: iterable of (absolute item offset, packet length).  Returns a list of numpy
        complex64 bursts, one per packet completed by this call."""
        self.bytes += bytes(bytearray(data))
        for off, ln in tags:
            self.starts.append((int(off), int(ln)))
        pk = []
        while self.starts:
            off, ln = self.starts[0]
            if off + ln > self.base + len(self.bytes):
                break
            a = off - self.base
            pk.append(bytes(self.bytes[a:a + ln]))
            del self.bytes[:a + ln]
            self.base = off + ln
            self.starts.pop(0)
        if not pk:
            return []
        s, soff = self.phy.tx(pk, first_pkt_num=self.n_sent)
        self.n_sent = (self.n_sent + len(pk)) & 0xFFF
        s = s.cpu().numpy() if hasattr(s, "cpu") else np.asarray(s)
        soff = soff.cpu().numpy() if hasattr(soff, "cpu") else np.asarray(soff)
        return [s[int(soff[i]):int(soff[i + 1])] for i in range(len(pk))]


def make_blocks(gr, pmt):
    """Build (ofdm_rx_b200, ofdm_tx_b200) on the block API of the given `gr` / `pmt` modules."""
    from .phy import OfdmPhy

    class ofdm_rx_b200(gr.basic_block):
        """complex stream in -> tagged byte stream out; same ports as ofdm_rx
        (python/ofdm_txrx_modules.py:292-294).  Every delivered packet starts with a `packet_length_tag_key`
        tag (value = bytes) and a `packet_num_tag_key` tag, as the stock chain's output does (:401-406,:363-372)."""

        def __init__(self, chunk=1 << 22, packet_length_tag_key="packet_len", packet_num_tag_key="packet_num",
                     phy=None, **phy_kwargs):
            gr.basic_block.__init__(self, name="ofdm_rx_b200", in_sig=[np.complex64], out_sig=[np.uint8])
            self.phy = phy if phy is not None else OfdmPhy(**phy_kwargs)
            self.streamer = RxStreamer(self.phy, chunk)
            self.len_key = pmt.intern(packet_length_tag_key)
            self.num_key = pmt.intern(packet_num_tag_key)
            self.queue = []             # [pkt_num, bytes, bytes already written]
            self.n_frames = self.n_dropped = 0

        def forecast(self, noutput_items, ninput_items_required):
            # output is produced from queued packets; any amount of input is welcome
            for i in range(len(ninput_items_required)):
                ninput_items_required[i] = 0 if self.queue else 1

        def _accept(self, frames):
            for _, rec, payload in frames:
                self.n_frames += 1
                if payload is None:
                    self.n_dropped += 1          # crc32_bb(True) drops the packet (python/ofdm_radio_hier.py:122)
                elif len(payload):
                    self.queue.append([int(rec["pkt_num"]), payload, 0])

        def general_work(self, input_items, output_items):
            x = input_items[0]
            if len(x):
                self._accept(self.streamer.push(x))
                self.consume(0, len(x))
            out = output_items[0]
            n_out = 0
            while self.queue and n_out < len(out):
                num, data, done = self.queue[0]
                if done == 0:
                    off = self.nitems_written(0) + n_out
                    self.add_item_tag(0, off, self.len_key, pmt.from_long(len(data)))
                    self.add_item_tag(0, off, self.num_key, pmt.from_long(num))
                k = min(len(data) - done, len(out) - n_out)
                out[n_out:n_out + k] = np.frombuffer(data, np.uint8, k, done)
                n_out += k
                if done + k == len(data):
                    self.queue.pop(0)
                else:
                    self.queue[0][2] = done + k
            return n_out

        def flush(self):
            """End of stream (a flowgraph calls this from stop()): decode what is still buffered."""
            self._accept(self.streamer.push(np.zeros(0, np.complex64), flush=True))

        def stop(self):
            self.flush()
            return True

    class ofdm_tx_b200(gr.basic_block):
        """tagged byte stream in -> complex stream out; same ports as ofdm_tx
        (python/ofdm_txrx_modules.py:156-158).  The first sample of every burst carries the length tag with the
        burst length in samples (what ofdm_cyclic_prefixer leaves on its output, :247-253)."""

        def __init__(self, packet_length_tag_key="packet_len", phy=None, **phy_kwargs):
            gr.basic_block.__init__(self, name="ofdm_tx_b200", in_sig=[np.uint8], out_sig=[np.complex64])
            self.phy = phy if phy is not None else OfdmPhy(**phy_kwargs)
            self.streamer = TxStreamer(self.phy)
            self.len_key = pmt.intern(packet_length_tag_key)
            self.n_read = 0
            self.queue = []             # [burst, samples already written]

        def forecast(self, noutput_items, ninput_items_required):
            for i in range(len(ninput_items_required)):
                ninput_items_required[i] = 0 if self.queue else 1

        def general_work(self, input_items, output_items):
            x = input_items[0]
            if len(x):
                tags = [(t.offset, pmt.to_long(t.value)) for t in self.get_tags_in_window(0, 0, len(x))
                        if pmt.eq(t.key, self.len_key)]
                for b in self.streamer.push(x.tobytes(), tags):
                    self.queue.append([b, 0])
                self.n_read += len(x)
                self.consume(0, len(x))
            out = output_items[0]
            n_out = 0
            while self.queue and n_out < len(out):
                b, done = self.queue[0]
                if done == 0:
                    self.add_item_tag(0, self.nitems_written(0) + n_out, self.len_key, pmt.from_long(len(b)))
                k = min(len(b) - done, len(out) - n_out)
                out[n_out:n_out + k] = b[done:done + k]
                n_out += k
                if done + k == len(b):
                    self.queue.pop(0)
                else:
                    self.queue[0][1] = done + k
            return n_out

    return ofdm_rx_b200, ofdm_tx_b200


try:                                    # the real thing, when GNU Radio is installed
    from gnuradio import gr as _gr
    import pmt as _pmt
    ofdm_rx_b200, ofdm_tx_b200 = make_blocks(_gr, _pmt)
    HAVE_GNURADIO = True
except ImportError:
    ofdm_rx_b200 = ofdm_tx_b200 = None
    HAVE_GNURADIO = False

__all__ = ["RxStreamer", "TxStreamer", "make_blocks", "ofdm_rx_b200", "ofdm_tx_b200", "HAVE_GNURADIO", "FRAME_DTYPE"]
