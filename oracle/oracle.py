"""ctypes front end of the CPU oracle (oracle/ofdm_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (gr-ofdm_tools_b200/) never imports it.

Also holds numpy restatements of the reference's host-side generators:
  * _make_sync_word1/2      python/ofdm_txrx_modules.py:75-104 (sqrt(2) amplitude)
  * _make_sync_word1 (1.42) python/ofdm_cr_tools.py:262-279
  * spectrum_enforcer       python/ofdm_cr_tools.py:348-378
PARITY: partially pinned (see ofdm_oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("ofdm_oracle.c", "ofdm_oracle.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


class _Params(C.Structure):
    _fields_ = [
        ("fft_len", C.c_int32), ("cp_len", C.c_int32),
        ("n_occ_sets", C.c_int32), ("occ_sizes", C.c_void_p), ("occ_carriers", C.c_void_p),
        ("n_pilot_sets", C.c_int32), ("pilot_sizes", C.c_void_p), ("pilot_carriers", C.c_void_p),
        ("n_pilot_sym_sets", C.c_int32), ("pilot_sym_sizes", C.c_void_p), ("pilot_symbols", C.c_void_p),
        ("sync_word1", C.c_void_p), ("sync_word2", C.c_void_p),
        ("bps_header", C.c_int32), ("bps_payload", C.c_int32),
        ("scramble_header", C.c_int32), ("scramble_seed", C.c_int32),
        ("crc_mode", C.c_int32), ("threshold", C.c_float), ("max_carr_offset", C.c_int32),
        ("alpha", C.c_float), ("tx_scale", C.c_float), ("demux_holdoff", C.c_int32), ("tx_clip", C.c_float), ("rolloff", C.c_int32),
        ("qam_normalization", C.c_int32),
    ]


FRAME_DTYPE = np.dtype([
    ("trigger", "<i8"), ("cfo", "<f4"), ("carr_offset", "<i4"), ("flags", "<u4"),
    ("pkt_len", "<u2"), ("pkt_num", "<u2"), ("frame_syms", "<u4"), ("slot", "<u4"),
])
assert FRAME_DTYPE.itemsize == 32
F_HDR_OK, F_CRC_OK, F_COMPLETE, F_ACCEPTED = 1, 2, 4, 8


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.orc_crc32.restype = C.c_uint32
        L.orc_crc32.argtypes = [C.c_void_p, C.c_int64]
        L.orc_crc8.restype = C.c_uint8
        L.orc_crc8.argtypes = [C.c_void_p, C.c_int64]
        L.orc_lfsr_bits.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int64]
        L.orc_scramble.argtypes = [C.c_void_p, C.c_int64, C.c_uint32]
        L.orc_repack.restype = C.c_int64
        L.orc_repack.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_header_len.argtypes = [C.c_void_p]
        L.orc_header_format.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_header_parse.argtypes = [C.c_void_p, C.c_void_p] + [C.c_void_p] * 4
        L.orc_constellation.argtypes = [C.c_int, C.c_void_p]
        L.orc_decide.argtypes = [C.c_int, C.c_double, C.c_double]
        L.orc_fft.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_tx.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                             C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_agc2_v.restype = None
        L.orc_agc2_v.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                 C.c_int]
        L.orc_iir_ccd_v.restype = None
        L.orc_iir_ccd_v.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                    C.c_void_p, C.c_int]
        L.orc_set_threads.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_constellation_n.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.orc_decide_n.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
        L.orc_crc32_mac.restype = C.c_uint32
        L.orc_crc32_mac.argtypes = [C.c_void_p, C.c_int64]
        L.orc_tx_frame_samples.restype = C.c_int64
        L.orc_tx_frame_samples.argtypes = [C.c_void_p, C.c_int64]
        L.orc_sync.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_sync_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                   C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_rx.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                             C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_int64, C.c_void_p]
        L.orc_rx_all.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                 C.c_void_p]
        L.orc_rx_baseline.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                      C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------------------------
# host-side generators (numpy restatements)
# ---------------------------------------------------------------------------------------------
def _active_carriers(fft_len, occupied_carriers, pilot_carriers):
    # python/ofdm_txrx_modules.py:66-73
    act = []
    for c in list(occupied_carriers[0]) + list(pilot_carriers[0]):
        act.append(c + fft_len if c < 0 else c)
    return act


def make_sync_word1(fft_len, occupied_carriers, pilot_carriers, amplitude=None):
    """python/ofdm_txrx_modules.py:75-91 (amplitude sqrt 2) / python/ofdm_cr_tools.py:262-279 (1.42)."""
    amp = np.sqrt(2) if amplitude is None else amplitude
    act = set(_active_carriers(fft_len, occupied_carriers, pilot_carriers))
    np.random.seed(42)
    bpsk = {0: amp, 1: -amp}
    sw1 = [bpsk[np.random.randint(2)] if (x in act and x % 2) else 0 for x in range(fft_len)]
    return np.fft.fftshift(sw1)


def make_sync_word2(fft_len, occupied_carriers, pilot_carriers):
    """python/ofdm_txrx_modules.py:93-104."""
    act = set(_active_carriers(fft_len, occupied_carriers, pilot_carriers))
    np.random.seed(42)
    bpsk = {0: 1, 1: -1}
    sw2 = [bpsk[np.random.randint(2)] if x in act else 0 for x in range(fft_len)]
    sw2[0] = 0j
    return np.fft.fftshift(sw2)


def spectrum_enforcer(fft_len, spectrum_constraint_fft, lobe_len):
    """python/ofdm_cr_tools.py:348-378 (integer division as in Python 2)."""
    usable = list(range(-fft_len // 2, fft_len // 2, 1))
    usable.remove(0)
    del usable[0:lobe_len]
    del usable[-lobe_len:]
    for carr in spectrum_constraint_fft:
        if carr in usable:
            usable.remove(carr)
    space = len(usable) // 8
    middle = len(usable) // 2
    pilots = ((usable[middle - 3 * space], usable[middle - space],
               usable[middle + space], usable[middle + 3 * space]),)
    for carr in pilots[0]:
        usable.remove(carr)
    occ = (usable,)
    sw1 = make_sync_word1(fft_len, occ, pilots, amplitude=1.42)
    sw2 = make_sync_word2(fft_len, occ, pilots)
    return occ, pilots, ((1, 1, 1, -1),), sw1.tolist(), sw2.tolist()


# ---------------------------------------------------------------------------------------------
class Oracle:
    """One PHY configuration of the CPU oracle (argument names follow ofdm_tx/ofdm_rx,
    python/ofdm_txrx_modules.py:143-155,278-291)."""

    def __init__(self, fft_len=64, cp_len=16, occupied_carriers=None, pilot_carriers=None,
                 pilot_symbols=None, sync_word1=None, sync_word2=None, bps_header=1, bps_payload=1,
                 scramble_bits=False, scramble_header=None, crc_mode=0, threshold=0.9,
                 max_carr_offset=-1, alpha=0.1, tx_scale=1.0, demux_holdoff=None, tx_clip=0.0, rolloff=0,
                 qam_normalization=0):
        self.fft_len, self.cp_len = int(fft_len), int(cp_len)
        self.occ = [list(map(int, s)) for s in occupied_carriers]
        self.pil = [list(map(int, s)) for s in pilot_carriers]
        self.pls = [list(map(complex, s)) for s in pilot_symbols]
        if sync_word1 is None:
            sync_word1 = make_sync_word1(fft_len, self.occ, self.pil)
        if sync_word2 is None:
            sync_word2 = make_sync_word2(fft_len, self.occ, self.pil)
        if len(sync_word1) != fft_len or len(sync_word2) not in (0, fft_len):
            raise ValueError("Length of sync sequence(s) must be FFT length.")
        self.n_sync_words = 2 if len(sync_word2) else 1      # sync_word2=(): python/ofdm_txrx_modules.py:174-183
        self.sw1 = np.asarray(sync_word1, dtype=np.complex64)
        self.sw2 = np.asarray(sync_word2, dtype=np.complex64)
        self.bps_header, self.bps_payload = int(bps_header), int(bps_payload)
        self.crc_mode = int(crc_mode)
        self._keep = [
            np.array([len(s) for s in self.occ], np.int32), np.array(sum(self.occ, []), np.int32),
            np.array([len(s) for s in self.pil], np.int32), np.array(sum(self.pil, []), np.int32),
            np.array([len(s) for s in self.pls], np.int32),
            np.array(sum(self.pls, []), np.complex64), self.sw1, self.sw2,
        ]
        k = self._keep
        p = _Params()
        p.fft_len, p.cp_len = self.fft_len, self.cp_len
        p.n_occ_sets, p.occ_sizes, p.occ_carriers = len(self.occ), k[0].ctypes.data, k[1].ctypes.data
        p.n_pilot_sets, p.pilot_sizes, p.pilot_carriers = len(self.pil), k[2].ctypes.data, k[3].ctypes.data
        p.n_pilot_sym_sets, p.pilot_sym_sizes, p.pilot_symbols = len(self.pls), k[4].ctypes.data, k[5].ctypes.data
        p.sync_word1, p.sync_word2 = k[6].ctypes.data, (k[7].ctypes.data if self.n_sync_words == 2 else None)
        p.bps_header, p.bps_payload = self.bps_header, self.bps_payload
        p.scramble_header = int(scramble_bits if scramble_header is None else scramble_header)
        p.scramble_seed = 0x7F if scramble_bits else 0x00
        p.crc_mode = self.crc_mode
        p.threshold, p.max_carr_offset, p.alpha, p.tx_scale = threshold, max_carr_offset, alpha, tx_scale
        p.demux_holdoff = (self.fft_len + self.cp_len) if demux_holdoff is None else int(demux_holdoff)
        p.tx_clip = float(tx_clip)
        p.rolloff = int(rolloff)
        p.qam_normalization = int(qam_normalization)
        self.p = p
        self.L = lib()

    @property
    def _pp(self):
        return C.byref(self.p)

    def header_len(self):
        return self.L.orc_header_len(self._pp)

    def header_format(self, pkt_len, pkt_num):
        out = np.zeros(self.header_len(), np.uint8)
        self.L.orc_header_format(self._pp, pkt_len, pkt_num, _ptr(out))
        return out

    def header_parse(self, items):
        items = np.ascontiguousarray(items, np.uint8)
        v = [C.c_int() for _ in range(4)]
        ok = self.L.orc_header_parse(self._pp, _ptr(items), *[C.byref(x) for x in v])
        return bool(ok), v[0].value, v[1].value, v[2].value, v[3].value

    def frame_samples(self, payload_bytes):
        return self.L.orc_tx_frame_samples(self._pp, payload_bytes)

    def tx(self, packets, first_pkt_num=0):
        """packets: list of bytes/uint8 arrays -> (complex64 samples, int64 offsets[n+1])."""
        lens = [len(b) for b in packets]
        off = np.zeros(len(packets) + 1, np.int64)
        off[1:] = np.cumsum(lens)
        flat = np.frombuffer(b"".join(bytes(bytearray(b)) for b in packets), np.uint8).copy() \
            if packets else np.zeros(0, np.uint8)
        cap = sum(self.frame_samples(n) for n in lens)
        out = np.zeros(max(cap, 1), np.complex64)
        soff = np.zeros(len(packets) + 1, np.int64)
        rc = self.L.orc_tx(self._pp, _ptr(flat), _ptr(off), len(packets), first_pkt_num,
                           _ptr(out), cap, _ptr(soff))
        if rc:
            raise RuntimeError("orc_tx failed: %d" % rc)
        return out[:cap], soff

    def sync(self, samples, want_detect=False, f32=False, max_trig=None):
        s = np.ascontiguousarray(samples, np.complex64)
        n = s.shape[0]
        max_trig = max_trig or max(16, n // max(1, self.cp_len) + 16)
        trig = np.zeros(max_trig, np.int64)
        cfo = np.zeros(max_trig, np.float32)
        nt = C.c_int64()
        det = np.zeros(n, np.uint8) if want_detect else None
        if f32:
            rc = self.L.orc_sync_f32(self._pp, _ptr(s), n, _ptr(trig), _ptr(cfo), max_trig, C.byref(nt))
        else:
            rc = self.L.orc_sync(self._pp, _ptr(s), n, _ptr(det) if want_detect else None,
                                 _ptr(trig), _ptr(cfo), max_trig, C.byref(nt))
        if rc:
            raise RuntimeError("orc_sync failed: %d" % rc)
        k = min(nt.value, max_trig)
        return (trig[:k].copy(), cfo[:k].copy(), det) if want_detect else (trig[:k].copy(), cfo[:k].copy())

    def rx(self, samples, max_frames=None, byte_stride=4096, want_z=True, max_pkt_syms=None, want_taps=False):
        s = np.ascontiguousarray(samples, np.complex64)
        n = s.shape[0]
        D = self.fft_len + self.cp_len
        max_frames = max_frames or (n // ((self.n_sync_words + 1) * D) + 4)
        max_trig = max(16, n // max(1, self.cp_len) + 16)
        recs = np.zeros(max_frames, FRAME_DTYPE)
        by = np.zeros((max_frames, byte_stride), np.uint8)
        zs = self.header_len() + (max_pkt_syms or (byte_stride * 8 // self.bps_payload + 1))
        z = np.zeros((max_frames, zs), np.complex64) if want_z else None
        trig = np.zeros(max_trig, np.int64)
        cfo = np.zeros(max_trig, np.float32)
        nf, nt = C.c_int64(), C.c_int64()
        taps = np.zeros((max_frames, self.fft_len), np.complex64) if (want_taps and want_z) else None
        self.L.orc_set_taps_out.argtypes = [C.c_void_p]
        self.L.orc_set_taps_out(_ptr(taps) if taps is not None else None)
        try:
            rc = self.L.orc_rx(self._pp, _ptr(s), n, _ptr(recs), max_frames, _ptr(by), byte_stride,
                               _ptr(z) if want_z else None, zs, C.byref(nf), _ptr(trig), _ptr(cfo),
                               max_trig, C.byref(nt))
        finally:
            self.L.orc_set_taps_out(None)
        if rc:
            raise RuntimeError("orc_rx failed: %d" % rc)
        k = nf.value
        return {"frames": recs[:k].copy(), "bytes": by[:k], "z": z[:k] if want_z else None,
                "triggers": trig[:nt.value].copy(), "cfo": cfo[:nt.value].copy(),
                "taps": taps[:k] if taps is not None else None}

    def rx_all(self, samples, byte_stride=4096):
        """One record per raw plateau trigger, each decoded on its own (the counterpart of ofdmx_set_emit_all):
        returns {"frames": records (flags: 1 hdr ok, 2 crc ok, 4 complete, 16 hdr seen), "bytes": [n, stride]}."""
        s = np.ascontiguousarray(samples, np.complex64)
        n = s.shape[0]
        cap = n // max(1, self.cp_len) + 16
        recs = np.zeros(cap, FRAME_DTYPE)
        by = np.zeros((cap, byte_stride), np.uint8)
        k = C.c_int64()
        rc = self.L.orc_rx_all(self._pp, _ptr(s), n, _ptr(recs), cap, _ptr(by), byte_stride, C.byref(k))
        if rc:
            raise RuntimeError("orc_rx_all failed: %d" % rc)
        return {"frames": recs[:k.value].copy(), "bytes": by[:k.value]}

    def rx_baseline(self, samples, max_frames=None, byte_stride=4096):
        """RX with the float32 FIR sync port (CPU-baseline timing only). Returns frame records."""
        s = np.ascontiguousarray(samples, np.complex64)
        n = s.shape[0]
        D = self.fft_len + self.cp_len
        max_frames = max_frames or (n // (3 * D) + 4)
        max_trig = max(16, n // max(1, self.cp_len) + 16)
        recs = np.zeros(max_frames, FRAME_DTYPE)
        by = np.zeros((max_frames, byte_stride), np.uint8)
        trig = np.zeros(max_trig, np.int64)
        cfo = np.zeros(max_trig, np.float32)
        nf, nt = C.c_int64(), C.c_int64()
        rc = self.L.orc_rx_baseline(self._pp, _ptr(s), n, _ptr(recs), max_frames, _ptr(by), byte_stride,
                                    C.byref(nf), _ptr(trig), _ptr(cfo), max_trig, C.byref(nt))
        if rc:
            raise RuntimeError("orc_rx_baseline failed: %d" % rc)
        return recs[:nf.value].copy()

    def payloads(self, res):
        """Byte strings the flowgraph would deliver (CRC-failed packets dropped, CRC stripped
        when crc_mode, python/ofdm_radio_hier.py:122,222-226)."""
        out = []
        for f, b in zip(res["frames"], res["bytes"]):
            n = int(f["pkt_len"])
            if self.crc_mode:
                if not (f["flags"] & F_CRC_OK):
                    continue
                n -= 4
            out.append(bytes(b[:n]))
        return out


def crc32(data):
    a = np.frombuffer(bytes(data), np.uint8)
    return lib().orc_crc32(_ptr(a) if len(a) else None, len(a))


def crc8(data):
    a = np.frombuffer(bytes(data), np.uint8)
    return lib().orc_crc8(_ptr(a) if len(a) else None, len(a))


def crc32_mac(data):
    """MSB-first CRC-32 of digital.crc.gen_and_append_crc32."""
    a = np.frombuffer(bytes(data), np.uint8)
    return lib().orc_crc32_mac(_ptr(a) if len(a) else None, len(a))


def set_threads(n=0):
    """OpenMP team size of the oracle's parallel loops (0: query).  Returns the size in effect."""
    return int(lib().orc_set_threads(int(n)))


def agc2(samples, gain=1.0, attack=1e-1, decay=1e-2, reference=1.0, max_gain=65536.0, abs_rate=False):
    """analog.agc2_cc over one stream (or each row of a 2-D array).  Returns (out, final gain(s))."""
    x = np.ascontiguousarray(samples, np.complex64)
    one = x.ndim == 1
    x2 = x[None, :] if one else x
    out = np.empty_like(x2)
    g = np.broadcast_to(np.asarray(gain, np.float32), (x2.shape[0],)).copy()
    for s in range(x2.shape[0]):
        gs = np.array([g[s]], np.float32)
        lib().orc_agc2_v(_ptr(x2[s]), _ptr(out[s]), x2.shape[1], attack, decay, reference, max_gain, _ptr(gs),
                         int(bool(abs_rate)))
        g[s] = gs[0]
    return (out[0], float(g[0])) if one else (out, g)


def iir_ccd(samples, fftaps, fbtaps, state=None, oldstyle=False):
    """filter.iir_filter_ccd(fftaps, fbtaps, oldstyle=False) over one stream.  state: None (fresh block) or the
    float64 array a previous call returned.  Returns (out complex64, state)."""
    x = np.ascontiguousarray(samples, np.complex64)
    ff = np.ascontiguousarray(fftaps, np.float64)
    fb = np.ascontiguousarray(fbtaps, np.float64)
    ns = 2 * (len(ff) - 1) + 2 * max(len(fb) - 1, 0)
    st = np.zeros(max(ns, 1), np.float64) if state is None else np.array(state, np.float64)
    out = np.empty_like(x)
    lib().orc_iir_ccd_v(_ptr(x), _ptr(out), len(x), _ptr(ff), len(ff), _ptr(fb), len(fb), _ptr(st), int(bool(oldstyle)))
    return out, st


def papr(block):
    """papr_sink.set_papr (python/papr_sink.py:46-50): max(x * conj(x)) / (vdot(x, x) / len(x)), real part."""
    m = np.asarray(block, np.complex64)
    mean_square = np.vdot(m, np.transpose(m)) / len(m)
    peak = max(m * np.conjugate(m))
    return float((peak / mean_square).real)


def lfsr_bits(mask, seed, reg_len, n):
    out = np.zeros(n, np.uint8)
    lib().orc_lfsr_bits(mask, seed, reg_len, _ptr(out), n)
    return out


def scramble(data, seed):
    a = np.frombuffer(bytes(data), np.uint8).copy()
    lib().orc_scramble(_ptr(a), len(a), seed)
    return a


def repack(items, k, l, align_output):
    a = np.ascontiguousarray(items, np.uint8)
    out = np.zeros(len(a) * k // l + 2, np.uint8)
    n = lib().orc_repack(_ptr(a), len(a), k, l, int(align_output), _ptr(out))
    return out[:n]


def constellation(bps, norm=0):
    pts = np.zeros(1 << bps, np.complex64)
    lib().orc_constellation_n(bps, int(norm), _ptr(pts))
    return pts


def decide(bps, z, norm=0):
    return lib().orc_decide_n(bps, int(norm), float(np.real(z)), float(np.imag(z)))


def fft(x, forward=True):
    a = np.ascontiguousarray(x, np.complex128)
    out = np.zeros_like(a)
    lib().orc_fft(len(a), int(forward), _ptr(a), _ptr(out))
    return out
