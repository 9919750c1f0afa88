"""Shared test fixtures: PHY configurations, the synthetic channel, and builders
that configure the CPU oracle and the CUDA library from the same keyword arguments."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_vectors.json")))

_tm = GOLD["ofdm_txrx_modules"]
OCC64 = [list(_tm["def_occupied_carriers"][0])]
PIL64 = [list(_tm["def_pilot_carriers"][0])]
PLS64 = [(x, x, x, -x) for x in _tm["pilot_sym_scramble_seq"]]


def cfg_c1(bps=2, scramble=False, crc=0, **kw):
    """config 1: fft 64, cp 16, 802.11a carriers, BPSK header (python/ofdm_tx_rx_hier.py:55-73)."""
    d = dict(fft_len=64, cp_len=16, occupied_carriers=OCC64, pilot_carriers=PIL64, pilot_symbols=PLS64,
             bps_header=1, bps_payload=bps, scramble_bits=scramble, crc_mode=crc)
    d.update(kw)
    return d


def cfg_radio128(bps=2, scramble=1, crc=1, **kw):
    """ofdm_radio_hier defaults (python/ofdm_radio_hier.py:34-39,73-86,106)."""
    g = GOLD["ofdm_radio_hier_defaults"]
    c = lambda v: [complex(a, b) for a, b in v]
    d = dict(fft_len=128, cp_len=32, occupied_carriers=g["occupied_carriers"], pilot_carriers=g["pilot_carriers"],
             pilot_symbols=g["pilot_symbols"], sync_word1=c(g["sync_word1"]), sync_word2=c(g["sync_word2"]),
             bps_header=1, bps_payload=bps, scramble_bits=bool(scramble), scramble_header=True, crc_mode=crc,
             max_carr_offset=3)
    d.update(kw)
    return d


def cfg_c3(**kw):
    """config 3: fft 1024, cp 72, 600 data carriers, 16-QAM, CRC + scrambler."""
    occ = [[k for k in range(-302, 303) if k not in (0, 150, -150, 300, -300)]]
    d = dict(fft_len=1024, cp_len=72, occupied_carriers=occ, pilot_carriers=[[-300, -150, 150, 300]],
             pilot_symbols=[[1, 1, 1, -1]], bps_header=1, bps_payload=4, scramble_bits=True, crc_mode=1,
             max_carr_offset=3)
    d.update(kw)
    return d


def cfg_c4(**kw):
    """config 4: fft 2048, cp 144, 1200 data carriers, 64-QAM (extension)."""
    pil = [-600, -300, 300, 600]
    occ = [[k for k in range(-602, 603) if k != 0 and k not in pil]]
    d = dict(fft_len=2048, cp_len=144, occupied_carriers=occ, pilot_carriers=[pil], pilot_symbols=[[1, 1, 1, -1]],
             bps_header=1, bps_payload=6, scramble_bits=True, crc_mode=1, max_carr_offset=3)
    d.update(kw)
    return d


def make_oracle(cfg):
    import oracle as O
    return O.Oracle(**cfg)


def make_phy(cfg, **extra):
    from ofdm_tools import OfdmPhy
    d = dict(cfg)
    d.update(extra)
    return OfdmPhy(**d)


def channel(frames, rng, gaps=(200, 2000), snr_db=20.0, cfo=0.0, fft_len=64, taps=None, lead=None, tail=None,
            scale=1.0):
    """Concatenate frames with random zero gaps, apply multipath FIR, CFO (in subcarrier spacings)
    and AWGN at snr_db relative to the mean frame power.  Returns complex64."""
    parts = []
    lead = int(rng.integers(gaps[0], gaps[1] + 1)) if lead is None else lead
    parts.append(np.zeros(lead, np.complex64))
    for i, f in enumerate(frames):
        parts.append(np.asarray(f, np.complex64))
        if i + 1 < len(frames):
            g = int(rng.integers(gaps[0], gaps[1] + 1)) if gaps[1] > 0 else 0
            parts.append(np.zeros(g, np.complex64))
    tail = int(rng.integers(max(gaps[0], 1), gaps[1] + 2)) if tail is None else tail
    parts.append(np.zeros(tail, np.complex64))
    x = np.concatenate(parts).astype(np.complex128) * scale
    if taps is not None:
        h = np.zeros(max(d for d, _ in taps) + 1, np.complex128)
        for d, v in taps:
            h[d] = v
        x = np.convolve(x, h)[: len(x)]
    if cfo:
        x = x * np.exp(2j * np.pi * cfo / fft_len * np.arange(len(x)))
    p = np.mean(np.abs(np.concatenate([np.asarray(f) for f in frames])) ** 2) * scale ** 2 if frames else 1.0
    sigma = np.sqrt(p / (10 ** (snr_db / 10.0)) / 2.0)
    x = x + sigma * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x)))
    return x.astype(np.complex64)


MULTIPATH = [(0, 1.0), (3, 0.4 * np.exp(1j * np.deg2rad(40))), (9, 0.2 * np.exp(-1j * np.deg2rad(70))),
             (20, 0.1 * np.exp(1j * np.deg2rad(120)))]


def rand_packets(rng, n, length):
    return [rng.integers(0, 256, length, dtype=np.uint8).tobytes() for _ in range(n)]


def split_frames(samples, offsets):
    return [samples[offsets[i]:offsets[i + 1]] for i in range(len(offsets) - 1)]


def rel_evm(a, b):
    a = np.asarray(a, np.complex128)
    b = np.asarray(b, np.complex128)
    return float(np.sqrt(np.sum(np.abs(a - b) ** 2) / max(np.sum(np.abs(b) ** 2), 1e-300)))
