"""ctypes binding of libofdmx.so (include/ofdmx.h).

There is no CPU fallback: if the CUDA library has not been built, importing this module raises.
Build it with ``python __graft_entry__.py`` (or ``build.build()`` in this package).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "lib", "libofdmx.so"))

OK, ERR_PARAM, ERR_CUDA, ERR_CAPACITY, ERR_NOMEM = 0, -1, -2, -3, -4
F_HDR_OK, F_CRC_OK, F_COMPLETE, F_ACCEPTED, F_HDR_SEEN, F_OVERSIZE = 1, 2, 4, 8, 16, 32
ABI_VERSION = 4
AGC2_ABS_RATE, IIR_OLDSTYLE = 1, 1
CNT_LAUNCHES, CNT_DEVICE_ALLOCS, CNT_HOST_SYNCS, CNT_RECONFIGS = 0, 1, 2, 3


class Params(C.Structure):
    """ofdmx_params (include/ofdmx.h)."""
    _fields_ = [
        ("fft_len", C.c_int32), ("cp_len", C.c_int32),
        ("n_occ_sets", C.c_int32), ("occ_sizes", C.c_void_p), ("occ_carriers", C.c_void_p),
        ("n_pilot_sets", C.c_int32), ("pilot_sizes", C.c_void_p), ("pilot_carriers", C.c_void_p),
        ("n_pilot_sym_sets", C.c_int32), ("pilot_sym_sizes", C.c_void_p), ("pilot_symbols", C.c_void_p),
        ("sync_word1", C.c_void_p), ("sync_word2", C.c_void_p),
        ("bps_header", C.c_int32), ("bps_payload", C.c_int32),
        ("scramble_header", C.c_int32), ("scramble_seed", C.c_int32),
        ("crc_mode", C.c_int32), ("threshold", C.c_float), ("max_carr_offset", C.c_int32),
        ("alpha", C.c_float), ("tx_scale", C.c_float), ("demux_holdoff", C.c_int32),
        ("max_pkt_bytes", C.c_int32), ("tx_clip", C.c_float), ("rolloff", C.c_int32),
        ("qam_normalization", C.c_int32), ("reserved", C.c_int32 * 7),
    ]


class Counts(C.Structure):
    _fields_ = [("n_triggers", C.c_int32), ("n_frames", C.c_int32), ("overflow", C.c_int32),
                ("reserved", C.c_int32)]


# every symbol include/ofdmx.h declares: (restype, argtypes)
_P, _I64, _I32 = C.c_void_p, C.c_int64, C.c_int32
SYMBOLS = {
    "ofdmx_abi_version": (C.c_int, []),
    "ofdmx_params_size": (C.c_int, []),
    "ofdmx_frame_size": (C.c_int, []),
    "ofdmx_create": (C.c_int, [_P, C.c_int, _P]),
    "ofdmx_destroy": (None, [_P]),
    "ofdmx_last_error": (C.c_char_p, [_P]),
    "ofdmx_reserve": (C.c_int, [_P, _I64, _I64, _I64]),
    "ofdmx_header_len": (C.c_int, [_P]),
    "ofdmx_tx_frame_samples": (_I64, [_P, _I64]),
    "ofdmx_launch_count": (_I64, [_P]),
    "ofdmx_tx": (C.c_int, [_P, _P, _P, _I64, _I32, _P, _I64, _P, _P]),
    "ofdmx_rx": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _P]),
    "ofdmx_set_emit_all": (C.c_int, [_P, C.c_int]),
    "ofdmx_set_debug_taps": (C.c_int, [_P, _P, _I64]),
    "ofdmx_rx_host": (C.c_int, [_P, _P, _I64, _I64, _P, _I64, _P, _I64, _P]),
    "ofdmx_sync": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _P, _P, _I64, _P, _P]),
    "ofdmx_profile": (C.c_int, [_P, C.c_int]),
    "ofdmx_profile_slots": (C.c_int, []),
    "ofdmx_profile_name": (C.c_char_p, [C.c_int]),
    "ofdmx_profile_read": (C.c_int, [_P, _P, _P]),
    "ofdmx_fft": (C.c_int, [_P, _P, _P, _I64, C.c_int, _P]),
    "ofdmx_crc32": (C.c_int, [_P, _P, _P, _I64, _P, _P]),
    "ofdmx_agc2": (C.c_int, [_P, _P, _P, _I64, _I64, _I64, C.c_float, C.c_float, C.c_float, C.c_float, _P, _I32, _P]),
    "ofdmx_iir_state_doubles": (_I64, []),
    "ofdmx_iir_ccd": (C.c_int, [_P, _P, _P, _I64, _I64, _I64, _P, _I32, _P, _I32, _I64, _P, _I32, _P]),
    "ofdmx_papr": (C.c_int, [_P, _P, _I64, _P, _P]),
    "ofdmx_reconfigure": (C.c_int, [_P, _P]),
    "ofdmx_counter": (_I64, [_P, C.c_int]),
}

_lib = None


def load():
    """Load libofdmx.so and declare every exported entry point.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "CUDA library %s not built (run `python __graft_entry__.py`); there is no CPU fallback"
            % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)    # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    # a stale .so with another ofdmx_params layout must not be driven with this binding's structs
    if lib.ofdmx_abi_version() != ABI_VERSION or lib.ofdmx_params_size() != C.sizeof(Params) \
            or lib.ofdmx_frame_size() != 32 or C.sizeof(Counts) != 16:
        raise ImportError("%s has ABI version %d / ofdmx_params of %d bytes; this binding needs version %d / %d bytes "
                          "(rebuild with `python __graft_entry__.py`)"
                          % (LIB_PATH, lib.ofdmx_abi_version(), lib.ofdmx_params_size(), ABI_VERSION, C.sizeof(Params)))
    _lib = lib
    return lib


def check(rc, ctx=None):
    if rc == OK:
        return
    msg = load().ofdmx_last_error(ctx)
    msg = msg.decode() if msg else "ofdmx error %d" % rc
    if rc == ERR_PARAM:
        raise ValueError(msg)
    if rc == ERR_CAPACITY:
        raise BufferError(msg)
    if rc == ERR_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)
