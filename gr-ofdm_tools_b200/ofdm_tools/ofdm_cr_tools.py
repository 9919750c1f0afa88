"""Hot-path part of python/ofdm_cr_tools.py: the carrier-plan / sync-word generators, plus the MAC framing
helpers either side of the PHY.

`_make_sync_word1` (1.42-amplitude variant, :262-279), `_make_sync_word2` (:282-293) and
`spectrum_enforcer` (:348-378) are on the OFDM PHY path; `make_packet` / `unmake_packet` (:1741-1773) are
the next row of SURVEY.md 8(f) (rank 3).  The rest of that file (PSD helpers, loggers, cognitive engine) is
out of scope.
"""
from . import phy

_seq_seed = 42
_1024_pilot_carriers = ((-300, -150, 150, 300,),)
_128_pilot_carriers = ((-35, -20, 20, 35,),)
_64_pilot_carriers = ((-21, -7, 7, 21,),)
_pilot_symbols = ((1, 1, 1, -1,),)


def _get_active_carriers(fft_len, occupied_carriers, pilot_carriers):
    return phy._get_active_carriers(fft_len, occupied_carriers, pilot_carriers)


def _make_sync_word1(fft_len, occupied_carriers, pilot_carriers):
    return phy._make_sync_word1(fft_len, occupied_carriers, pilot_carriers, amplitude=1.42)


def _make_sync_word2(fft_len, occupied_carriers, pilot_carriers):
    return phy._make_sync_word2(fft_len, occupied_carriers, pilot_carriers)


def spectrum_enforcer(fft_len, spectrum_constraint_fft, lobe_len):
    """Derives occupied carriers, 4 pilots and both sync words from a spectrum mask."""
    usable = list(range(-fft_len // 2, fft_len // 2, 1))
    usable.remove(0)                      # DC
    del usable[0:lobe_len]                # side lobes
    del usable[-lobe_len:]
    for carr in spectrum_constraint_fft:  # constrained carriers
        if carr in usable:
            usable.remove(carr)
    space = len(usable) // 8
    middle = len(usable) // 2
    pilot_carriers = ((usable[middle - 3 * space], usable[middle - space],
                       usable[middle + space], usable[middle + 3 * space]),)
    for carr in pilot_carriers[0]:
        usable.remove(carr)
    occupied_carriers = ((usable),)
    sync_word1 = _make_sync_word1(fft_len, occupied_carriers, pilot_carriers)
    sync_word2 = _make_sync_word2(fft_len, occupied_carriers, pilot_carriers)
    return occupied_carriers, pilot_carriers, _pilot_symbols, sync_word1.tolist(), sync_word2.tolist()


def find_nearest_l(lst, value):
    """Nearest entry of a list (python/ofdm_cr_tools.py:451-452); the first one wins a tie."""
    return min(lst, key=lambda x: abs(x - value))


def spectrum_translator(spectrum_constraint_hz, fc, sf, fft_len, canc_bins):
    """Hz -> FFT bins (python/ofdm_cr_tools.py:455-469): every constrained frequency is snapped to the nearest bin
    centre of an fft_len-point grid of sample rate `sf` around `fc`, and canc_bins / 2 bins on either side of it
    are put on the list that spectrum_enforcer removes from the carrier plan.  Same output as the reference,
    element for element (bins as floats, the centre bin listed twice) -- cognitive_engine_mac.generate_sync_data
    (python/cognitive_engine_mac.py:278-285) feeds it straight into spectrum_enforcer."""
    resolution = float(sf) / float(fft_len)                       # Hz per bin
    grid_hz = [(k - fft_len // 2) * resolution + fc for k in range(fft_len)]
    half = canc_bins // 2 if isinstance(canc_bins, int) else canc_bins / 2
    out = []
    for f in spectrum_constraint_hz:
        centre = (find_nearest_l(grid_hz, f) - fc) / resolution
        x = 0
        while x < half:
            out.extend((centre + x, centre - x))
            x += 1
    return out


# ---- MAC framing either side of the PHY -----------------------------------
# Host-side byte-string helpers, as in the reference; bytes in / bytes out (the reference is Python 2 str).
from . import crc as _crc  # noqa: E402


def _b(s):
    return s.encode('latin-1') if isinstance(s, str) else bytes(s)


def make_packet(payload, pkt_size, type_pkt):
    """python/ofdm_cr_tools.py:1741-1748: 4 ASCII digits of payload length + 1 type byte + payload +
    0x55 padding up to pkt_size bytes."""
    payload, type_pkt = _b(payload), _b(type_pkt)
    length = len(payload)
    digits = str(length).encode()
    header = (4 - len(digits)) * b'0' + digits
    packing = (pkt_size - (5 + length)) * b'\x55'
    return header + type_pkt + payload + packing


def unmake_packet(frame, w_crc):
    """python/ofdm_cr_tools.py:1750-1773.  w_crc true: the in-flowgraph CRC already validated the frame;
    false: the frame ends in the MAC-level CRC-32 (digital.crc.check_crc32).  Returns (payload, type, ok);
    ('BAD', 'BAD', False) on a CRC failure, as the reference does."""
    frame = _b(frame)
    if w_crc:
        try:
            pld_len = int(frame[0:4])
            tpe = frame[4:5]
            return frame[5:5 + pld_len], tpe, True
        except Exception:
            return None
    ok, frame = _crc.check_crc32(frame)
    if not ok:
        return 'BAD', 'BAD', ok
    try:
        pld_len = int(frame[0:4])
        tpe = frame[4:5]
        return frame[5:5 + pld_len], tpe, ok
    except Exception:
        return None
