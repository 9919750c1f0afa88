"""GNU Radio adaptor (ofdm_tools/gr_blocks.py): the stream logic and the gr.basic_block classes, driven through a
stub of GNU Radio's block API with random chunk sizes.

CPU tests put a stand-in for OfdmPhy behind the blocks whose rx_host / tx are served by the oracle (the adaptor
only sees records, slots and sample arrays, so its chunk logic is what is under test); the GPU test runs the same
drive with the real library and compares against ONE big ofdmx_rx_host call."""
import types

import numpy as np
import pytest

import common as cm

from ofdm_tools import gr_blocks, _lib
from ofdm_tools.phy import FRAME_DTYPE, RxResult


# ---------------------------------------------------------------------------------------------------------
# a stub of the parts of gnuradio.gr / pmt the blocks use
class _Tag(object):
    def __init__(self, offset, key, value):
        self.offset, self.key, self.value = offset, key, value


class _BasicBlock(object):
    def __init__(self, name, in_sig, out_sig):
        self._name, self._in_sig, self._out_sig = name, in_sig, out_sig
        self._written, self._read, self._consumed = [0] * len(out_sig), [0] * len(in_sig), [0] * len(in_sig)
        self._in_tags, self.out_tags = [[] for _ in in_sig], [[] for _ in out_sig]

    def consume(self, port, n):
        self._consumed[port] += n

    def nitems_written(self, port):
        return self._written[port]

    def nitems_read(self, port):
        return self._read[port]

    def add_item_tag(self, port, offset, key, value):
        self.out_tags[port].append(_Tag(offset, key, value))

    def get_tags_in_window(self, port, start, end):
        a, b = self._read[port] + start, self._read[port] + end
        return [t for t in self._in_tags[port] if a <= t.offset < b]


gr_stub = types.SimpleNamespace(basic_block=_BasicBlock)
pmt_stub = types.SimpleNamespace(intern=lambda s: ("sym", s), from_long=lambda v: ("long", int(v)),
                                 to_long=lambda v: v[1], eq=lambda a, b: a == b)


def drive(block, data, rng, in_tags=(), max_chunk=5000, out_room=(1, 4000)):
    """What the GNU Radio scheduler does: hand the block pieces of the input and an output buffer of some size,
    advance the read / write counters by what it consumed / produced.  Returns (output items, output tags)."""
    block._in_tags[0] = [_Tag(o, k, v) for o, k, v in in_tags]
    out_dt = block._out_sig[0]
    outs = []
    pos = 0
    idle = 0
    while pos < len(data) or idle < 3:
        n_in = min(len(data) - pos, int(rng.integers(0, max_chunk + 1)))
        need = [0]
        block.forecast(1, need)
        room = np.zeros(int(rng.integers(out_room[0], out_room[1] + 1)), out_dt)
        block._consumed = [0]
        n_out = block.general_work([data[pos:pos + n_in]], [room])
        assert block._consumed[0] <= n_in
        pos += block._consumed[0]
        block._read[0] += block._consumed[0]
        block._written[0] += n_out
        outs.append(room[:n_out].copy())
        idle = idle + 1 if (n_out == 0 and pos >= len(data)) else 0
    return np.concatenate(outs) if outs else np.zeros(0, out_dt), block.out_tags[0]


def packets_of(stream, tags, len_key):
    lens = [(t.offset, t.value[1]) for t in tags if t.key == len_key]
    return [bytes(stream[o:o + n]) for o, n in lens]


# ---------------------------------------------------------------------------------------------------------
class OraclePhy(object):
    """Duck-typed OfdmPhy for the CPU tests: the calls the adaptor makes, answered by the oracle."""

    def __init__(self, cfg, max_pkt_bytes=4095):
        self.o = cm.make_oracle(cfg)
        self.fft_len, self.cp_len, self.crc_mode = self.o.fft_len, self.o.cp_len, self.o.crc_mode
        self.max_pkt_bytes = max_pkt_bytes
        self.byte_stride = (max_pkt_bytes + 15) // 16 * 16
        self.params = types.SimpleNamespace(demux_holdoff=self.o.p.demux_holdoff)
        self.emit_all = False
        self.calls = 0

    def set_emit_all(self, on=True):
        self.emit_all = bool(on)

    def frame_samples(self, nbytes):
        return self.o.frame_samples(int(nbytes))

    def default_max_frames(self, n_streams, n):
        return int(n_streams * (n // (3 * (self.fft_len + self.cp_len)) + 4))

    def rx_host(self, samples, max_frames=None):
        assert self.emit_all
        self.calls += 1
        r = self.o.rx_all(samples, byte_stride=self.byte_stride)
        if max_frames is not None and len(r["frames"]) > max_frames:
            raise BufferError("more triggers than max_frames")
        f = np.zeros(len(r["frames"]), FRAME_DTYPE)
        for k in ("trigger", "cfo", "flags", "pkt_len", "pkt_num", "frame_syms", "carr_offset", "slot"):
            f[k] = r["frames"][k]
        return RxResult(f, r["bytes"], None, None, self.crc_mode, len(f))

    def tx(self, packets, first_pkt_num=0):
        return self.o.tx(packets, first_pkt_num=first_pkt_num)


def _stream(cfg, rng, n_pk=9, plen=96, snr=25.0, gaps=(0, 700)):
    orc = cm.make_oracle(cfg)
    pk = cm.rand_packets(rng, n_pk, plen)
    s, off = orc.tx(pk)
    x = cm.channel(cm.split_frames(s, off), rng, gaps=gaps, snr_db=snr, cfo=0.15, lead=300, tail=900)
    return orc, pk, x


@pytest.mark.parametrize("seed,chunk,tail", [(0, 700, None), (1, 2500, None), (2, 64, None), (6, 300, 0), (7, 1000, 400)])
def test_rx_streamer_is_chunk_invariant(seed, chunk, tail):
    """RxStreamer over random pushes == the oracle's one-shot receiver (records and payloads), including
    back-to-back frames, noise triggers and frames that straddle the windows; with a tail shorter than a frame
    the streamer has to wait for the announced payload before it goes on."""
    rng = np.random.default_rng(seed)
    cfg = cm.cfg_c1(2, True, 1)
    orc, pk, x = _stream(cfg, rng)
    ref = orc.rx(x, want_z=False)
    st = gr_blocks.RxStreamer(OraclePhy(cfg, max_pkt_bytes=104), chunk=chunk, tail=tail)
    got = []
    pos = 0
    while pos < len(x):
        k = int(rng.integers(0, 3000))
        got += st.push(x[pos:pos + k])
        pos += k
    got += st.push(x[:0], flush=True)
    assert st.phy.calls > 2
    assert [g[0] for g in got] == list(ref["frames"]["trigger"])
    assert [int(g[1]["pkt_num"]) for g in got] == list(ref["frames"]["pkt_num"])
    assert [g[2] for g in got if g[2] is not None] == orc.payloads(ref) == pk


def test_rx_block_general_work_chunk_invariance():
    """ofdm_rx_b200.general_work under a stub scheduler: the tagged byte stream it writes carries exactly the
    packets of the one-shot receiver, each behind a packet_len tag, whatever the chunking."""
    rx_cls, _ = gr_blocks.make_blocks(gr_stub, pmt_stub)
    cfg = cm.cfg_c1(2, True, 1)
    for seed in (3, 4):
        rng = np.random.default_rng(seed)
        orc, pk, x = _stream(cfg, rng, n_pk=7)
        ref = orc.rx(x, want_z=False)
        blk = rx_cls(chunk=1500, phy=OraclePhy(cfg, max_pkt_bytes=104))
        out1, tags1 = drive(blk, x, rng)
        blk.flush()
        out2, tags2 = drive(blk, x[:0], rng)
        out = np.concatenate([out1, out2])
        got = packets_of(out, tags2, ("sym", "packet_len"))
        assert got == orc.payloads(ref) == pk
        nums = [t.value[1] for t in tags2 if t.key == ("sym", "packet_num")]
        assert nums == list(ref["frames"]["pkt_num"])
        assert blk.n_frames == len(ref["frames"]) and blk.n_dropped == 0


def test_tx_block_general_work():
    """ofdm_tx_b200.general_work: packets cut at the length tags, bursts equal to the oracle's ofdm_tx, one
    length tag (in samples) per burst; header counter continuous across calls."""
    _, tx_cls = gr_blocks.make_blocks(gr_stub, pmt_stub)
    cfg = cm.cfg_c1(2, True, 1)
    rng = np.random.default_rng(5)
    orc = cm.make_oracle(cfg)
    lens = [96, 10, 200, 96, 1]
    pk = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in lens]
    s_ref, off_ref = orc.tx(pk)
    data = np.frombuffer(b"".join(pk), np.uint8)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    tags = [(int(o), ("sym", "packet_len"), ("long", n)) for o, n in zip(offs, lens)]
    blk = tx_cls(phy=OraclePhy(cfg))
    out, otags = drive(blk, data, rng, in_tags=tags, max_chunk=70, out_room=(1, 900))
    assert np.array_equal(out, s_ref)
    assert [(t.offset, t.value[1]) for t in otags] == [(int(off_ref[i]), int(off_ref[i + 1] - off_ref[i])) for i in range(len(pk))]


def test_module_without_gnuradio():
    # in this image GNU Radio is absent: the module imports, says so, and the factory is the way in
    if gr_blocks.HAVE_GNURADIO:
        assert gr_blocks.ofdm_rx_b200 is not None
    else:
        assert gr_blocks.ofdm_rx_b200 is None and callable(gr_blocks.make_blocks)


@pytest.mark.gpu
def test_rx_block_on_gpu_equals_one_big_rx_host():
    """The same drive with the real library: ofdm_rx_b200 over random chunks == ONE ofdmx_rx_host call over the
    whole stream (records, payload bytes), on the fft_len 64 hier-block plan and on the fft_len 1024 plan."""
    rx_cls, tx_cls = gr_blocks.make_blocks(gr_stub, pmt_stub)
    for cfg, plen, n_pk, chunk, kw in ((cm.cfg_c1(2, True, 1), 96, 40, 6000, {}),
                                       (cm.cfg_c3(), 1500, 12, 30000, dict(fft_len=1024, taps=cm.MULTIPATH))):
        rng = np.random.default_rng(11)
        orc = cm.make_oracle(cfg)
        pk = cm.rand_packets(rng, n_pk, plen)
        s, off = orc.tx(pk)
        x = cm.channel(cm.split_frames(s, off), rng, gaps=(0, 900), snr_db=40.0, cfo=0.2, lead=300, tail=3000, **kw)
        one = cm.make_phy(cfg).rx_host(x)
        want = one.payloads()
        assert want == pk
        blk = rx_cls(chunk=chunk, phy=cm.make_phy(cfg))
        out1, _ = drive(blk, x, rng, max_chunk=20000, out_room=(1, 20000))
        blk.flush()
        out2, tags = drive(blk, x[:0], rng, out_room=(1000, 20000))
        got = packets_of(np.concatenate([out1, out2]), tags, ("sym", "packet_len"))
        assert got == want
        assert [t.value[1] for t in tags if t.key == ("sym", "packet_num")] == list(one.frames["pkt_num"])
        # and the transmitter block against the library's one-shot ofdm_tx
        data = np.frombuffer(b"".join(pk), np.uint8)
        itags = [(i * plen, ("sym", "packet_len"), ("long", plen)) for i in range(n_pk)]
        tb = tx_cls(phy=cm.make_phy(cfg))
        so, _ = drive(tb, data, rng, in_tags=itags, max_chunk=4000, out_room=(1, 50000))
        s_gpu, _ = cm.make_phy(cfg).tx(pk)
        assert np.array_equal(so, s_gpu.cpu().numpy())
