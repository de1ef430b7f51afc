"""ctypes binding of include/fastsk_b200.h.  No torch types cross this boundary."""
from __future__ import annotations

import ctypes
import os

from .build import LIB_PATH

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_voidpp = ctypes.POINTER(ctypes.c_void_p)

FSK_OK, FSK_EINVAL, FSK_ECUDA, FSK_ENOMEM, FSK_ESTATE = 0, 1, 2, 3, 4
FSK_DT_I64, FSK_DT_F64 = 0, 1
FSK_IPC_HANDLE_BYTES = 64


class FskStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in ("n_seq", "nfeat", "n_pairs", "n_combos_total", "combos_done",
                                              "pair_updates", "entries", "runs", "kernel_launches")] + \
               [(n, ctypes.c_int32) for n in ("key_bits", "id_bits", "record_bytes", "sort_passes", "alphabet",
                                              "bits_per_char", "batch", "acc_bytes")] + \
               [(n, ctypes.c_double) for n in ("ms_pack", "ms_sort", "ms_segment", "ms_accumulate", "ms_welford",
                                               "ms_normalise", "ms_total")] + \
               [(n, ctypes.c_int32) for n in ("acc_path", "heavy_tau")] + [("heavy_runs", ctypes.c_int64)] + \
               [(n, ctypes.c_int32) for n in ("n_devices", "seg_mode", "dense_mode")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/fastsk_b200.h declares: (restype, argtypes)
_H = ctypes.c_void_p
SIGNATURES = {
    "fsk_create": (ctypes.c_int, [c_voidpp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                  ctypes.c_int, ctypes.c_int]),
    "fsk_destroy": (None, [_H]),
    "fsk_last_error": (ctypes.c_char_p, [_H]),
    "fsk_version": (ctypes.c_char_p, []),
    "fsk_set_device": (ctypes.c_int, [_H, ctypes.c_int]),
    "fsk_set_devices": (ctypes.c_int, [_H, ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "fsk_ipc_export_partial": (ctypes.c_int, [_H, ctypes.c_void_p]),
    "fsk_set_peer_partials": (ctypes.c_int, [_H, ctypes.c_void_p, ctypes.c_int]),
    "fsk_set_peer_pointers": (ctypes.c_int, [_H, c_voidpp, ctypes.c_int]),
    "fsk_release_peers": (ctypes.c_int, [_H]),
    "fsk_set_output_weights": (ctypes.c_int, [_H, c_f64p, ctypes.c_int]),
    "fsk_probe_d2h": (ctypes.c_int, [ctypes.c_int, ctypes.c_size_t, c_f64p]),
    "fsk_output_rows": (ctypes.c_int, [_H, c_i64p, c_i64p, c_i64p, c_i64p]),
    "fsk_host_alloc": (ctypes.c_int, [c_voidpp, ctypes.c_size_t]),
    "fsk_host_free": (ctypes.c_int, [ctypes.c_void_p]),
    "fsk_host_register": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t]),
    "fsk_host_unregister": (ctypes.c_int, [ctypes.c_void_p]),
    "fsk_trim_cache": (ctypes.c_int, []),
    "fsk_selftest_division": (ctypes.c_int, [ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]),
    "fsk_set_seed": (ctypes.c_int, [_H, ctypes.c_uint64]),
    "fsk_set_combo_sequence": (ctypes.c_int, [_H, c_i32p, ctypes.c_int64]),
    "fsk_set_shard": (ctypes.c_int, [_H, ctypes.c_int, ctypes.c_int]),
    "fsk_set_option": (ctypes.c_int, [_H, ctypes.c_char_p, ctypes.c_int64]),
    "fsk_compute": (ctypes.c_int, [_H, c_i32p, c_i64p, ctypes.c_int64, ctypes.c_int64]),
    "fsk_upload": (ctypes.c_int, [_H, c_i32p, c_i64p, ctypes.c_int64, ctypes.c_int64]),
    "fsk_upload_split": (ctypes.c_int, [_H, c_i32p, c_i64p, ctypes.c_int64, c_i32p, c_i64p, ctypes.c_int64]),
    "fsk_compute_split": (ctypes.c_int, [_H, c_i32p, c_i64p, ctypes.c_int64, c_i32p, c_i64p, ctypes.c_int64]),
    "fsk_build_partial": (ctypes.c_int, [_H]),
    "fsk_partial_buffer": (ctypes.c_int, [_H, c_voidpp, c_i64p, ctypes.POINTER(ctypes.c_int)]),
    "fsk_finalize": (ctypes.c_int, [_H]),
    "fsk_accumulate_combos": (ctypes.c_int, [_H, c_i32p, ctypes.c_int64, ctypes.c_int]),
    "fsk_reset_partial": (ctypes.c_int, [_H]),
    "fsk_stream": (ctypes.c_int, [_H, c_voidpp]),
    "fsk_synchronize": (ctypes.c_int, [_H]),
    "fsk_shape": (ctypes.c_int, [_H, c_i64p, c_i64p, c_i64p, c_i64p]),
    "fsk_get_train_kernel": (ctypes.c_int, [_H, c_f64p]),
    "fsk_get_test_kernel": (ctypes.c_int, [_H, c_f64p]),
    "fsk_get_kernel_packed": (ctypes.c_int, [_H, c_f64p]),
    "fsk_get_unnormalised_i64": (ctypes.c_int, [_H, c_i64p]),
    "fsk_get_unnormalised_f64": (ctypes.c_int, [_H, c_f64p]),
    "fsk_train_kernel_device": (ctypes.c_int, [_H, c_voidpp]),
    "fsk_test_kernel_device": (ctypes.c_int, [_H, c_voidpp]),
    "fsk_get_stdevs": (ctypes.c_int, [_H, c_f64p, ctypes.c_int64, c_i64p]),
    "fsk_save_kernel": (ctypes.c_int, [_H, ctypes.c_char_p]),
    "fsk_get_queue": (ctypes.c_int, [_H, c_i32p, ctypes.c_int64, c_i64p]),
    "fsk_get_shard_work": (ctypes.c_int, [_H, c_i32p, ctypes.c_int64, c_i64p]),
    "fsk_get_stats": (ctypes.c_int, [_H, ctypes.POINTER(FskStats)]),
}

_lib = None


def load():
    """Load libfastsk_b200.so.  There is no fallback: a missing extension is an error."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA extension first "
                "(python -c 'import __graft_entry__ as g; g.build()' or python -m fastsk_b200.build)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(lib, handle, rc):
    if rc == FSK_OK:
        return
    msg = lib.fsk_last_error(handle)
    msg = msg.decode() if msg else f"error {rc}"
    if rc == FSK_EINVAL:
        raise ValueError(msg)
    if rc == FSK_ENOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)
