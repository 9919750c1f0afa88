"""Host-side driver of the B200 OFDM PHY: one `OfdmPhy` = one configured TX/RX chain.

PyTorch is used only as the device-buffer carrier (allocation, streams); all arithmetic happens in
libofdmx.so (gr-ofdm_tools_b200/csrc).  Argument names follow the reference's hier blocks
(python/ofdm_txrx_modules.py:143-155,278-291; python/ofdm_radio_hier.py:34-39).
"""
import ctypes as C

import numpy as np

from . import _lib

FRAME_DTYPE = np.dtype([
    ("trigger", "<i8"), ("cfo", "<f4"), ("stream", "<i4"), ("flags", "<u4"),
    ("pkt_len", "<u2"), ("pkt_num", "<u2"), ("frame_syms", "<u2"), ("carr_offset", "<i2"),
    ("slot", "<u4"),
])
assert FRAME_DTYPE.itemsize == 32

_SEQ_SEED = 42


def _get_active_carriers(fft_len, occupied_carriers, pilot_carriers):
    """python/ofdm_txrx_modules.py:66-73."""
    active = []
    for carrier in list(occupied_carriers[0]) + list(pilot_carriers[0]):
        active.append(carrier + fft_len if carrier < 0 else carrier)
    return active


def _make_sync_word1(fft_len, occupied_carriers, pilot_carriers, amplitude=None):
    """First Schmidl & Cox preamble symbol: seeded random BPSK on the odd active carriers
    (python/ofdm_txrx_modules.py:75-91; amplitude 1.42 variant python/ofdm_cr_tools.py:262-279)."""
    amp = float(np.sqrt(2)) if amplitude is None else float(amplitude)
    active = set(_get_active_carriers(fft_len, occupied_carriers, pilot_carriers))
    rs = np.random.RandomState(_SEQ_SEED)
    vals = []
    for x in range(fft_len):
        if x in active and x % 2:
            vals.append(amp if rs.randint(2) == 0 else -amp)
        else:
            vals.append(0.0)
    return np.fft.fftshift(vals)


def _make_sync_word2(fft_len, occupied_carriers, pilot_carriers):
    """Second preamble symbol: seeded random BPSK on every active carrier, DC forced to zero
    (python/ofdm_txrx_modules.py:93-104)."""
    active = set(_get_active_carriers(fft_len, occupied_carriers, pilot_carriers))
    rs = np.random.RandomState(_SEQ_SEED)
    vals = []
    for x in range(fft_len):
        if x in active:
            vals.append(1 + 0j if rs.randint(2) == 0 else -1 + 0j)
        else:
            vals.append(0j)
    vals[0] = 0j
    return np.fft.fftshift(vals)


class RxResult(object):
    """Output of one RX call: frame records (numpy, FRAME_DTYPE), payload slots, counts."""

    def __init__(self, frames, slots, counts, z, crc_mode, n_triggers):
        self.frames = frames
        self.slots = slots          # uint8 [max_frames, byte_stride] (torch tensor or numpy)
        self.counts = counts
        self.z = z
        self.crc_mode = crc_mode
        self.n_triggers = n_triggers

    def slot_bytes(self, f):
        row = self.slots[int(f["slot"])]
        row = row.cpu().numpy() if hasattr(row, "cpu") else row
        return row[: int(f["pkt_len"])]

    def payloads(self):
        """Packets as the flowgraph delivers them: with crc_mode, failed packets are dropped and the
        4 CRC bytes stripped (python/ofdm_radio_hier.py:122,222-226)."""
        slots = self.slots.cpu().numpy() if hasattr(self.slots, "cpu") else self.slots
        out = []
        for f in self.frames:
            n = int(f["pkt_len"])
            if self.crc_mode:
                if not (int(f["flags"]) & _lib.F_CRC_OK):
                    continue
                n -= 4
            if int(f["flags"]) & _lib.F_OVERSIZE:
                continue        # longer than max_pkt_bytes: consumed by the demux, not decoded (no slot content)
            out.append(bytes(slots[int(f["slot"]), :n]))
        return out


class OfdmPhy(object):
    def __init__(self, fft_len=64, cp_len=16, occupied_carriers=None, pilot_carriers=None,
                 pilot_symbols=None, sync_word1=None, sync_word2=None, bps_header=1, bps_payload=1,
                 scramble_bits=False, scramble_header=None, crc_mode=0, threshold=0.9,
                 max_carr_offset=-1, alpha=0.1, tx_scale=1.0, demux_holdoff=None,
                 max_pkt_bytes=4095, device=0, tx_clip=0.0, rolloff=0, qam_normalization=0):
        self._cfg = dict(fft_len=fft_len, cp_len=cp_len, occupied_carriers=occupied_carriers,
                         pilot_carriers=pilot_carriers, pilot_symbols=pilot_symbols, sync_word1=sync_word1,
                         sync_word2=sync_word2, bps_header=bps_header, bps_payload=bps_payload,
                         scramble_bits=scramble_bits, scramble_header=scramble_header, crc_mode=crc_mode,
                         threshold=threshold, max_carr_offset=max_carr_offset, alpha=alpha, tx_scale=tx_scale,
                         demux_holdoff=demux_holdoff, max_pkt_bytes=max_pkt_bytes, device=device, tx_clip=tx_clip,
                         rolloff=rolloff, qam_normalization=qam_normalization)
        self.fft_len, self.cp_len = int(fft_len), int(cp_len)
        self.occupied_carriers = [list(map(int, s)) for s in occupied_carriers]
        self.pilot_carriers = [list(map(int, s)) for s in pilot_carriers]
        self.pilot_symbols = [list(map(complex, s)) for s in pilot_symbols]
        if sync_word1 is None:
            sync_word1 = _make_sync_word1(fft_len, self.occupied_carriers, self.pilot_carriers)
        elif len(sync_word1) != self.fft_len:
            raise ValueError("Length of sync sequence(s) must be FFT length.")
        if sync_word2 is None:
            sync_word2 = _make_sync_word2(fft_len, self.occupied_carriers, self.pilot_carriers)
        elif len(sync_word2) not in (0, self.fft_len):
            raise ValueError("Length of sync sequence(s) must be FFT length.")
        # sync_word2=(): one sync word, two OFDM symbols before the payload (python/ofdm_txrx_modules.py:174-183,
        # 311-329); the library takes a NULL sync_word2 for it
        self.n_sync_words = 2 if len(sync_word2) else 1
        self.sync_word1 = np.asarray(sync_word1, dtype=np.complex64)
        self.sync_word2 = np.asarray(sync_word2, dtype=np.complex64)
        self.bps_header, self.bps_payload = int(bps_header), int(bps_payload)
        self.crc_mode = int(crc_mode)
        self.max_pkt_bytes = int(max_pkt_bytes)
        self.device = int(device)
        self.byte_stride = (self.max_pkt_bytes + 15) // 16 * 16

        def flat(sets, dt):
            return np.array([v for s in sets for v in s], dtype=dt)

        self._keep = [
            np.array([len(s) for s in self.occupied_carriers], np.int32), flat(self.occupied_carriers, np.int32),
            np.array([len(s) for s in self.pilot_carriers], np.int32), flat(self.pilot_carriers, np.int32),
            np.array([len(s) for s in self.pilot_symbols], np.int32), flat(self.pilot_symbols, np.complex64),
            self.sync_word1, self.sync_word2,
        ]
        k = self._keep
        p = _lib.Params()
        p.fft_len, p.cp_len = self.fft_len, self.cp_len
        p.n_occ_sets, p.occ_sizes, p.occ_carriers = len(self.occupied_carriers), k[0].ctypes.data, k[1].ctypes.data
        p.n_pilot_sets, p.pilot_sizes, p.pilot_carriers = len(self.pilot_carriers), k[2].ctypes.data, k[3].ctypes.data
        p.n_pilot_sym_sets, p.pilot_sym_sizes, p.pilot_symbols = len(self.pilot_symbols), k[4].ctypes.data, k[5].ctypes.data
        p.sync_word1, p.sync_word2 = k[6].ctypes.data, (k[7].ctypes.data if self.n_sync_words == 2 else None)
        p.bps_header, p.bps_payload = self.bps_header, self.bps_payload
        p.scramble_header = int(scramble_bits if scramble_header is None else scramble_header)
        p.scramble_seed = 0x7F if scramble_bits else 0x00
        p.crc_mode = self.crc_mode
        p.threshold, p.max_carr_offset, p.alpha, p.tx_scale = threshold, max_carr_offset, alpha, tx_scale
        p.demux_holdoff = (self.fft_len + self.cp_len) if demux_holdoff is None else int(demux_holdoff)
        p.max_pkt_bytes = self.max_pkt_bytes
        p.tx_clip = float(tx_clip)
        p.rolloff = int(rolloff)
        p.qam_normalization = int(qam_normalization)
        self.rolloff = int(rolloff)
        self.params = p
        self._ctx = None

    # ------------------------------------------------------------------ plumbing
    @property
    def ctx(self):
        if self._ctx is None:
            L = _lib.load()
            h = C.c_void_p()
            _lib.check(L.ofdmx_create(C.byref(self.params), self.device, C.byref(h)))
            self._ctx = h
        return self._ctx

    def reconfigure(self, **changes):
        """Run-time reconfiguration: new constructor arguments -- typically the
        occupied_carriers / pilot_carriers / pilot_symbols / sync_word1 / sync_word2 that
        ofdm_cr_tools.spectrum_enforcer returns (python/ofdm_cr_tools.py:348-378,
        python/cognitive_engine_mac.py:278-285) -- replace the PHY tables of the live context
        (ofdmx_reconfigure); workspace, staging buffers, stream and counters are kept.  When the carrier plan
        changes and no sync words are given they are regenerated as the constructor would."""
        cfg = dict(self._cfg)
        if any(k in changes for k in ("fft_len", "occupied_carriers", "pilot_carriers")):
            cfg["sync_word1"] = cfg["sync_word2"] = None
        cfg.update(changes)
        if int(cfg["device"]) != self.device:
            raise ValueError("reconfigure cannot move a context to another device")
        fresh = OfdmPhy(**cfg)          # host-side validation and parameter block only (no context yet)
        if self._ctx is not None:
            _lib.check(_lib.load().ofdmx_reconfigure(C.byref(self._ctx), C.byref(fresh.params)), self._ctx)
        ctx = self._ctx
        self.__dict__.pop("_host_bufs", None)     # sized by the old max_pkt_bytes: rx_host re-allocates them
        self.__dict__.update(fresh.__dict__)
        self._ctx = ctx
        return self

    def close(self):
        if self._ctx is not None:
            _lib.load().ofdmx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _torch(self):
        import torch
        return torch

    def _dev(self):
        return self._torch().device("cuda", self.device)

    def _stream(self):
        return C.c_void_p(self._torch().cuda.current_stream(self._dev()).cuda_stream)

    def header_len(self):
        return _lib.load().ofdmx_header_len(self.ctx)

    def frame_samples(self, payload_bytes):
        return _lib.load().ofdmx_tx_frame_samples(self.ctx, int(payload_bytes))

    def launch_count(self):
        return _lib.load().ofdmx_launch_count(self.ctx)

    def counter(self, which):
        """Life-time counters of the context (ofdmx_counter): _lib.CNT_LAUNCHES / CNT_DEVICE_ALLOCS /
        CNT_HOST_SYNCS / CNT_RECONFIGS."""
        return int(_lib.load().ofdmx_counter(self.ctx, int(which)))

    def set_debug_taps(self, taps=None):
        """Channel-estimate debug tap (ofdmx_set_debug_taps): `taps` is a zeroed cuda complex64 [max_frames, >= fft_len]
        tensor that later rx calls fill with the initial channel taps of every trigger slot (shifted bin order, the
        ofdm_sync_chan_taps tag of the reference chain); None switches the tap off."""
        if taps is None:
            _lib.check(_lib.load().ofdmx_set_debug_taps(self.ctx, None, 0), self.ctx)
            self._taps = None
            return
        assert taps.is_cuda and taps.dim() == 2 and taps.stride(1) == 1 and taps.shape[1] >= self.fft_len
        self._taps = taps
        _lib.check(_lib.load().ofdmx_set_debug_taps(self.ctx, taps.data_ptr(), taps.stride(0)), self.ctx)

    def set_emit_all(self, enable=True):
        """RX then returns one record per plateau trigger (see ofdmx_set_emit_all)."""
        _lib.check(_lib.load().ofdmx_set_emit_all(self.ctx, int(bool(enable))), self.ctx)

    def profile(self, enable=True):
        """Start/stop recording CUDA events around every kernel this context launches."""
        _lib.check(_lib.load().ofdmx_profile(self.ctx, int(bool(enable))), self.ctx)

    def profile_read(self):
        """{kernel name: (total ms, launches)} since profile(True); synchronises the events."""
        L = _lib.load()
        n = L.ofdmx_profile_slots()
        ms = (C.c_float * n)()
        calls = (C.c_int64 * n)()
        _lib.check(L.ofdmx_profile_read(self.ctx, ms, calls), self.ctx)
        return {L.ofdmx_profile_name(i).decode(): (float(ms[i]), int(calls[i])) for i in range(n) if calls[i]}

    # ------------------------------------------------------------------ TX
    def tx(self, packets, first_pkt_num=0, out=None, soff=None):
        """packets: list of bytes-like (or (uint8 cuda tensor, int64 cuda offsets)).
        Returns (complex64 cuda tensor of samples, int64 cuda tensor of n+1 frame offsets).
        With device packets and a pre-allocated `out` (complex64 cuda tensor; `soff`: int64 [n+1], optional) the
        call only enqueues work: nothing is copied to the host, the whole `out` tensor is returned and
        soff[-1] tells how much of it was written (packets that would not fit are left out)."""
        torch = self._torch()
        dev = self._dev()
        if isinstance(packets, tuple):
            payload, off = packets
            n = int(off.numel()) - 1
            lens = None if out is not None else (off[1:] - off[:-1]).cpu().numpy()
        else:
            lens = np.array([len(b) for b in packets], np.int64)
            flat = np.frombuffer(b"".join(bytes(bytearray(b)) for b in packets), np.uint8)
            payload = torch.from_numpy(flat.copy() if len(flat) else np.zeros(1, np.uint8)).to(dev)
            o = np.zeros(len(lens) + 1, np.int64)
            o[1:] = np.cumsum(lens)
            off = torch.from_numpy(o).to(dev)
            n = len(lens)
        extra = 4 if self.crc_mode else 0
        if lens is not None:
            if n and int(lens.max()) + extra > self.max_pkt_bytes:
                raise ValueError("packet longer than max_pkt_bytes")
            vals, cnt = np.unique(lens, return_counts=True)
            need = int(sum(self.frame_samples(int(v)) * int(c) for v, c in zip(vals, cnt)))
        if out is None:
            cap = need
            out = torch.empty(max(cap, 1), dtype=torch.complex64, device=dev)
        else:
            assert out.is_cuda and out.dtype == torch.complex64 and out.is_contiguous()
            cap = int(out.numel())
            if lens is not None and need > cap:
                raise BufferError("output tensor too small: %d samples needed, %d given" % (need, cap))
        if soff is None:
            soff = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        if n:
            _lib.check(_lib.load().ofdmx_tx(self.ctx, payload.data_ptr(), off.data_ptr(), n, int(first_pkt_num),
                                            out.data_ptr(), cap, soff.data_ptr(), self._stream()), self.ctx)
        return (out[:cap] if lens is not None else out), soff

    # ------------------------------------------------------------------ RX
    def default_max_frames(self, n_streams, n):
        return int(n_streams * (n // ((self.n_sync_words + 1) * (self.fft_len + self.cp_len)) + 4))

    def rx_buffers(self, max_frames, device, want_z=False, max_pkt_syms=None):
        """Pre-allocated output buffers for rx_enqueue (reusable across calls)."""
        torch = self._torch()
        b = {"max_frames": int(max_frames),
             "frames": torch.empty(max_frames * 32, dtype=torch.uint8, device=device),
             "slots": torch.empty((max_frames, self.byte_stride), dtype=torch.uint8, device=device),
             "counts": torch.zeros(4, dtype=torch.int32, device=device), "z": None, "zs": 0}
        if want_z:
            b["zs"] = self.header_len() + (max_pkt_syms or (self.max_pkt_bytes * 8 // self.bps_payload + 1))
            b["z"] = torch.zeros((max_frames, b["zs"]), dtype=torch.complex64, device=device)
        return b

    def rx_enqueue(self, samples, bufs):
        """Launch the RX chain on the current stream; no host synchronisation."""
        torch = self._torch()
        s = samples if samples.dim() == 2 else samples.unsqueeze(0)
        assert s.dtype == torch.complex64 and s.is_cuda and s.stride(1) == 1
        n_streams, n = s.shape
        z = bufs["z"]
        _lib.check(_lib.load().ofdmx_rx(
            self.ctx, s.data_ptr(), n_streams, n, s.stride(0), bufs["frames"].data_ptr(), bufs["max_frames"],
            bufs["slots"].data_ptr(), self.byte_stride, z.data_ptr() if z is not None else None, bufs["zs"],
            bufs["counts"].data_ptr(), self._stream()), self.ctx)

    def rx_collect(self, bufs):
        """Synchronise and read the records of the last rx_enqueue on these buffers."""
        c = bufs["counts"].cpu().numpy()
        if c[2]:
            raise BufferError("more triggers (%d) than max_frames (%d)" % (c[0], bufs["max_frames"]))
        nf = int(c[1])
        rec = np.frombuffer(bufs["frames"][: nf * 32].cpu().numpy().tobytes(), FRAME_DTYPE).copy()
        return RxResult(rec, bufs["slots"], c, bufs["z"], self.crc_mode, int(c[0]))

    def rx(self, samples, max_frames=None, want_z=False, max_pkt_syms=None):
        """samples: complex64 cuda tensor [n] or [n_streams, n].  Returns RxResult (synchronises)."""
        s = samples if samples.dim() == 2 else samples.unsqueeze(0)
        n_streams, n = s.shape
        if max_frames is None:
            max_frames = self.default_max_frames(n_streams, n)
        bufs = self.rx_buffers(max_frames, s.device, want_z, max_pkt_syms)
        self.rx_enqueue(s, bufs)
        return self.rx_collect(bufs)

    def rx_host(self, samples, max_frames=None):
        """Same through ofdmx_rx_host: numpy complex64 in, numpy out (H2D/D2H inside the call).  The
        output staging buffers are pinned and reused across calls (the returned arrays are views that
        stay valid until the next rx_host call on this object)."""
        s = np.ascontiguousarray(samples, np.complex64)
        if s.ndim == 1:
            s = s[None, :]
        n_streams, n = s.shape
        if max_frames is None:
            max_frames = self.default_max_frames(n_streams, n)
        hb = getattr(self, "_host_bufs", None)
        if hb is None or hb[0] < max_frames:
            torch = self._torch()
            fr = torch.empty(max_frames * 32, dtype=torch.uint8, pin_memory=True)
            sl = torch.empty((max_frames, self.byte_stride), dtype=torch.uint8, pin_memory=True)
            hb = self._host_bufs = (max_frames, fr, sl, fr.numpy().view(FRAME_DTYPE), sl.numpy())
        _, _, _, frames, slots = hb
        cnt = _lib.Counts()
        _lib.check(_lib.load().ofdmx_rx_host(self.ctx, s.ctypes.data, n_streams, n, frames.ctypes.data, max_frames,
                                             slots.ctypes.data, self.byte_stride, C.byref(cnt)), self.ctx)
        if cnt.overflow:
            raise BufferError("more triggers (%d) than max_frames (%d)" % (cnt.n_triggers, max_frames))
        c = np.array([cnt.n_triggers, cnt.n_frames, cnt.overflow, 0], np.int32)
        return RxResult(frames[: cnt.n_frames], slots, c, None, self.crc_mode, cnt.n_triggers)

    def sync(self, samples, max_trig=None):
        """Schmidl & Cox only: returns (trigger indices int64, cfo float32, stream int32) as numpy."""
        torch = self._torch()
        s = samples if samples.dim() == 2 else samples.unsqueeze(0)
        n_streams, n = s.shape
        if max_trig is None:
            max_trig = int(n_streams * (n // max(1, self.cp_len) + 16))
        dev = s.device
        trig = torch.zeros(max_trig, dtype=torch.int64, device=dev)
        cfo = torch.zeros(max_trig, dtype=torch.float32, device=dev)
        st = torch.zeros(max_trig, dtype=torch.int32, device=dev)
        counts = torch.zeros(4, dtype=torch.int32, device=dev)
        _lib.check(_lib.load().ofdmx_sync(self.ctx, s.data_ptr(), n_streams, n, s.stride(0), trig.data_ptr(),
                                          cfo.data_ptr(), st.data_ptr(), max_trig, counts.data_ptr(),
                                          self._stream()), self.ctx)
        c = counts.cpu().numpy()
        if c[2]:
            raise BufferError("more triggers (%d) than max_trig (%d)" % (c[0], max_trig))
        k = int(c[0])
        return trig[:k].cpu().numpy(), cfo[:k].cpu().numpy(), st[:k].cpu().numpy()

    # ------------------------------------------------------------------ single blocks
    def fft(self, x, forward=True):
        """fft.fft_vcc(fft_len, forward, (), True) on a [n_syms, fft_len] complex64 cuda tensor."""
        torch = self._torch()
        x = x.contiguous()
        out = torch.empty_like(x)
        _lib.check(_lib.load().ofdmx_fft(self.ctx, x.data_ptr(), out.data_ptr(), x.numel() // self.fft_len,
                                         int(bool(forward)), self._stream()), self.ctx)
        return out

    def crc32(self, packets):
        torch = self._torch()
        dev = self._dev()
        lens = np.array([len(b) for b in packets], np.int64)
        flat = np.frombuffer(b"".join(bytes(bytearray(b)) for b in packets), np.uint8)
        payload = torch.from_numpy(flat.copy() if len(flat) else np.zeros(1, np.uint8)).to(dev)
        o = np.zeros(len(lens) + 1, np.int64)
        o[1:] = np.cumsum(lens)
        off = torch.from_numpy(o).to(dev)
        out = torch.zeros(max(len(lens), 1), dtype=torch.int64, device=dev).to(torch.int32)
        _lib.check(_lib.load().ofdmx_crc32(self.ctx, payload.data_ptr(), off.data_ptr(), len(lens), out.data_ptr(),
                                           self._stream()), self.ctx)
        return out[: len(lens)].cpu().numpy().astype(np.uint32)

    def agc2(self, samples, gain=None, attack=1e-1, decay=1e-2, reference=1.0, max_gain=65536.0, out=None,
             abs_rate=False):
        """analog.agc2_cc(attack, decay, reference, 1.0) + set_max_gain, the block in front of ofdm_rx in both
        hier surfaces (python/ofdm_tx_rx_hier.py:75-76, python/ofdm_radio_hier.py:180-181).  samples: cuda
        complex64 [n] or [n_streams, n]; gain: None (1.0 per stream, a fresh block) or a cuda float32
        [n_streams] tensor holding the loop gain left by the previous call -- it is updated in place.
        Returns (out, gain).  Streams run in parallel, samples of one stream sequentially."""
        torch = self._torch()
        if samples.dim() == 1:
            samples = samples.unsqueeze(0)
        assert samples.is_cuda and samples.dtype == torch.complex64 and samples.stride(1) == 1
        n_streams, n = samples.shape
        if gain is None:
            gain = torch.ones(n_streams, dtype=torch.float32, device=samples.device)
        assert gain.is_cuda and gain.dtype == torch.float32 and gain.numel() == n_streams and gain.is_contiguous()
        if out is None:
            out = torch.empty_like(samples)
        assert out.shape == samples.shape and out.stride() == samples.stride()
        _lib.check(_lib.load().ofdmx_agc2(self.ctx, samples.data_ptr(), out.data_ptr(), n_streams, n,
                                          samples.stride(0) if n_streams > 1 else max(n, 1), attack, decay, reference,
                                          max_gain, gain.data_ptr(), _lib.AGC2_ABS_RATE if abs_rate else 0,
                                          self._stream()), self.ctx)
        return out, gain

    def iir_ccd(self, samples, fftaps, fbtaps, state=None, span=0, out=None, oldstyle=False):
        """filter.iir_filter_ccd(fftaps, fbtaps, oldstyle=False): the out-of-band filter behind the TX chain of
        ofdm_radio_hier when filter_mode=1 (python/ofdm_radio_hier.py:83-84,93,232-237).  samples: cuda complex64
        [n] or [n_streams, n]; state: None (a fresh block) or the cuda float64 [n_streams, 32] tensor a previous
        call returned -- updated in place, so consecutive calls continue each stream.  Returns (out, state).
        Each stream is cut into spans that run in parallel (ofdmx_iir_ccd in include/ofdmx.h)."""
        import ctypes as C
        torch = self._torch()
        one = samples.dim() == 1
        if one:
            samples = samples.unsqueeze(0)
        assert samples.is_cuda and samples.dtype == torch.complex64 and samples.stride(1) == 1
        n_streams, n = samples.shape
        lib = _lib.load()
        nd = int(lib.ofdmx_iir_state_doubles())
        if state is None:
            state = torch.zeros((n_streams, nd), dtype=torch.float64, device=samples.device)
        assert state.is_cuda and state.dtype == torch.float64 and state.numel() == n_streams * nd and state.is_contiguous()
        if out is None:
            out = torch.empty_strided(samples.shape, samples.stride(), dtype=samples.dtype, device=samples.device)
        elif out.dim() == 1:
            out = out.unsqueeze(0)
        assert out.shape == samples.shape and (n_streams == 1 or out.stride() == samples.stride())
        ff = (C.c_double * len(fftaps))(*[float(t) for t in fftaps])
        fb = (C.c_double * max(len(fbtaps), 1))(*[float(t) for t in fbtaps])
        _lib.check(lib.ofdmx_iir_ccd(self.ctx, samples.data_ptr(), out.data_ptr(), n_streams, n,
                                     samples.stride(0) if n_streams > 1 else max(n, 1), ff, len(fftaps), fb,
                                     len(fbtaps), int(span), state.data_ptr(), _lib.IIR_OLDSTYLE if oldstyle else 0,
                                     self._stream()), self.ctx)
        return (out[0] if one else out), state

    def papr(self, block):
        """papr_sink.level() (python/papr_sink.py:46-54) of a cuda complex64 block: peak |x|^2 over mean |x|^2.
        Returns a cuda float32 [3] tensor {papr, peak, mean square} (no host synchronisation)."""
        torch = self._torch()
        assert block.is_cuda and block.dtype == torch.complex64 and block.is_contiguous() and block.numel() > 0
        out = torch.empty(3, dtype=torch.float32, device=block.device)
        _lib.check(_lib.load().ofdmx_papr(self.ctx, block.data_ptr(), block.numel(), out.data_ptr(), self._stream()),
                   self.ctx)
        return out
