"""ctypes binding of libmdb_b200.so (the C ABI declared in include/mdb_b200.h).

There is no CPU fallback: if the CUDA library has not been built, or no sm_100a device is
present, every entry point raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"``
or ``mddatasetbuilder_b200/csrc/build.sh``.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MDB_B200_LIB", os.path.join(_HERE, "libmdb_b200.so"))  # override: A/B runs of two builds
_lib = None

c_ctx = C.c_void_p
_vp = C.c_void_p

# name -> (restype, argtypes); mirrors include/mdb_b200.h one to one
SIGNATURES = {
    "mdb_version": (C.c_int, []),
    "mdb_last_error": (C.c_char_p, []),
    "mdb_create": (C.c_int, [C.c_int, C.POINTER(c_ctx)]),
    "mdb_destroy": (None, [c_ctx]),
    "mdb_set_elements": (C.c_int, [c_ctx, C.c_int, C.POINTER(C.c_int)]),
    "mdb_set_option": (C.c_int, [c_ctx, C.c_char_p, C.c_double]),
    "mdb_launch_count": (C.c_uint64, [c_ctx]),
    "mdb_arena_bytes": (C.c_size_t, [c_ctx]),
    "mdb_bonds_from_coords": (C.c_int, [c_ctx, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp]),
    "mdb_ell_to_csr": (C.c_int, [c_ctx, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, _vp, C.c_int, _vp]),
    "mdb_bond_keys": (C.c_int, [c_ctx, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "mdb_bond_keys_bo": (C.c_int, [c_ctx, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "mdb_bond_keys_csr": (C.c_int, [c_ctx, C.c_int, C.c_int, _vp, _vp, C.c_longlong, _vp, _vp, _vp]),
    "mdb_components_csr": (C.c_int, [c_ctx, C.c_int, C.c_int, _vp, _vp, C.c_longlong, _vp, _vp]),
    "mdb_group_keys": (C.c_int, [c_ctx, C.c_int, C.c_int, _vp, _vp, C.c_longlong, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp,
                                 _vp]),
    "mdb_type_maxcount": (C.c_int, [c_ctx, _vp, C.c_longlong, _vp, _vp, _vp, _vp]),
    "mdb_minmax_accumulate": (C.c_int, [c_ctx, C.c_longlong, C.c_int, _vp, C.c_longlong, _vp, _vp, _vp]),
    "mdb_gather_rows": (C.c_int, [c_ctx, C.c_longlong, C.c_int, _vp, C.c_longlong, _vp, _vp, _vp, C.c_longlong, _vp]),
    "mdb_label_hist": (C.c_int, [c_ctx, C.c_longlong, C.c_int, _vp, _vp, _vp]),
    "mdb_pick_rows": (C.c_int, [c_ctx, _vp, C.c_int, _vp, _vp, _vp]),
    "mdb_components": (C.c_int, [c_ctx, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp]),
    "mdb_dps": (C.c_int, [c_ctx, C.c_int, _vp, _vp, _vp, _vp, C.POINTER(C.c_int), _vp]),
    "mdb_analyze_frames": (C.c_int, [c_ctx, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp]),
    "mdb_analyze_frames_host_edges": (C.c_int, [c_ctx, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_longlong, _vp, _vp, _vp,
                                                _vp]),
    "mdb_analyze_frames_host": (C.c_int, [c_ctx, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp]),
    "mdb_analyze_frames_host_coded": (C.c_int, [c_ctx, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp,
                                                _vp]),
    "mdb_compositions": (C.c_int, [c_ctx, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, C.c_double, _vp, C.c_longlong,
                                   _vp, _vp, _vp]),
    "mdb_sphere_members": (C.c_int, [c_ctx, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, C.c_double, _vp, C.c_longlong,
                                     C.c_int, _vp, _vp, _vp]),
    "mdb_descriptors": (C.c_int, [c_ctx, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, C.c_double, _vp, C.c_longlong,
                                  _vp, _vp, _vp, _vp, C.c_int, _vp, _vp]),
    "mdb_compositions_members": (C.c_int, [c_ctx, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, C.c_double, _vp, C.c_longlong,
                                           _vp, _vp, C.c_int, _vp, _vp]),
    "mdb_descriptors_members": (C.c_int, [c_ctx, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_longlong, _vp, _vp,
                                          _vp, C.c_int, _vp, _vp, C.c_int, _vp, _vp]),
    "mdb_descriptors_members_compact": (C.c_int, [c_ctx, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_longlong, _vp, _vp,
                                                  _vp, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "mdb_feature_rows_compact": (C.c_int, [c_ctx, C.c_longlong, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mdb_feature_rows": (C.c_int, [c_ctx, C.c_longlong, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "mdb_minmax_fit": (C.c_int, [c_ctx, C.c_longlong, C.c_int, _vp, C.c_longlong, _vp, _vp, _vp]),
    "mdb_minmax_transform": (C.c_int, [c_ctx, C.c_longlong, C.c_int, _vp, C.c_longlong, _vp, _vp, _vp]),
    "mdb_kmeans_assign": (C.c_int, [c_ctx, C.c_longlong, C.c_int, C.c_int, _vp, C.c_longlong, _vp, _vp, _vp, _vp, _vp,
                                    C.POINTER(C.c_double), _vp]),
    "mdb_kmeans_assign_tc": (C.c_int, [c_ctx, C.c_longlong, C.c_int, C.c_int, _vp, C.c_longlong, _vp, _vp, _vp,
                                       C.POINTER(C.c_longlong), _vp, _vp]),
    "mdb_kmeans_inertia": (C.c_int, [c_ctx, C.c_longlong, C.c_int, _vp, C.c_longlong, _vp, _vp, _vp, _vp,
                                     C.POINTER(C.c_double), _vp]),
    "mdb_kmeans_partial_sums": (C.c_int, [c_ctx, C.c_longlong, C.c_int, C.c_int, _vp, C.c_longlong, _vp, _vp, _vp, _vp,
                                          _vp, _vp]),
    "mdb_kmeans_update_exact": (C.c_int, [c_ctx, C.c_longlong, C.c_int, _vp, C.c_longlong, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mdb_kmeans_minibatch_step": (C.c_int, [c_ctx, C.c_longlong, C.c_int, C.c_int, _vp, C.c_longlong, _vp, _vp, _vp,
                                            C.POINTER(C.c_double), C.POINTER(C.c_int), _vp]),
    "mdb_kmeans_apply_update": (C.c_int, [c_ctx, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "mdb_kpp_step": (C.c_int, [c_ctx, C.c_longlong, C.c_int, _vp, C.c_longlong, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp,
                               _vp]),
    "mdb_kmeans_plusplus": (C.c_int, [c_ctx, C.c_int, C.c_longlong, C.c_int, _vp, C.c_longlong, _vp, C.c_int, C.c_int, _vp,
                                      _vp, _vp, _vp]),
    "mdb_reader_open": (C.c_int, [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.POINTER(_vp)]),
    "mdb_reader_close": (None, [_vp]),
    "mdb_reader_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_long), C.POINTER(C.c_int)]),
    "mdb_reader_types": (C.c_int, [_vp, _vp]),
    "mdb_reader_rewind": (C.c_int, [_vp]),
    "mdb_reader_next": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.POINTER(C.c_int)]),
    "mdb_reader_next_raw": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_size_t, _vp, _vp, _vp, _vp, C.POINTER(C.c_int)]),
    "mdb_parse_dump_text": (C.c_int, [c_ctx, _vp, C.c_int, C.c_int, _vp, _vp, C.c_longlong, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mdb_parse_bonds_text": (C.c_int, [c_ctx, _vp, C.c_int, C.c_int, _vp, _vp, C.c_longlong, C.c_int, _vp, _vp, _vp, _vp,
                                       _vp]),
    "mdb_reader_parse_buffer": (C.c_int, [_vp, C.c_char_p, C.c_size_t, _vp, _vp, _vp, _vp, C.c_int]),
    "mdb_write_xyz_batch": (C.c_int, [C.c_int, C.POINTER(C.c_char_p), _vp, _vp, _vp, C.c_int]),
    "mdb_write_gjf_batch": (C.c_int, [C.c_int, C.POINTER(C.c_char_p), _vp, _vp, _vp, _vp, _vp, _vp, C.c_int,
                                      C.POINTER(C.c_char_p), C.c_int, C.c_int]),
    "mdb_fp64_peak": (C.c_int, [c_ctx, C.POINTER(C.c_double), _vp]),
    "mdb_profile_enable": (C.c_int, [c_ctx, C.c_int]),
    "mdb_profile_read": (C.c_int, [c_ctx, C.c_char_p, C.c_size_t]),
    "mdb_perceive_bond_orders": (C.c_int, [c_ctx, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "mdb_perceive_stats": (C.c_int, [c_ctx, C.POINTER(C.c_longlong)]),
    "mdb_bond_stats": (C.c_int, [c_ctx, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
}


def lib():
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA library of mddatasetbuilder_b200 has not been built "
                "(run mddatasetbuilder_b200/csrc/build.sh). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here = header and library out of sync
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return lib().mdb_last_error().decode(errors="replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {last_error()}")
