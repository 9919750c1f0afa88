"""Hot-path part of python/ofdm_cr_tools.py: the carrier-plan / sync-word generators.

Only `_make_sync_word1` (1.42-amplitude variant, :262-279), `_make_sync_word2` (:282-293) and
`spectrum_enforcer` (:348-378) are on the OFDM PHY path; the rest of that file (PSD helpers, MAC
framing, loggers) is out of scope.
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
