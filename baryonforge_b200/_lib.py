"""
ctypes binding of libbfg_b200.so (C ABI in include/bfg_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes as C
import os

from ._build import LIB_PATH

_LIB = None

c_i64 = C.c_int64
c_dbl = C.c_double
c_ptr = C.c_void_p

HALO_STRIDE = 16
TABLE_LOG_VALUES = 1
TABLE_RDELTA = 2

# shell record fields (include/bfg_b200.h)
HS_VX, HS_VY, HS_VZ, HS_THETA, HS_PHI, HS_D, HS_A, HS_RADIUS, HS_LNZ, HS_LNM, HS_RCUT, HS_LNRCOM, HS_SCALE, \
    HS_THETA_LL, HS_PHI_LL, HS_SKIP = range(16)
# box record fields
HB_X, HB_Y, HB_Z, HB_RQ, HB_NSIZE, HB_CX, HB_CY, HB_CZ, HB_LNZ, HB_LNM, HB_RCUT, HB_LNRCOM, HB_DX, HB_DY, HB_DZ, \
    HB_PAINTCUT = range(16)

_SIGNATURES = {
    "bfg_abi_version": ([], C.c_int),
    "bfg_last_error": ([], C.c_char_p),
    "bfg_device_count": ([], C.c_int),
    "bfg_device_info": ([C.c_int, C.POINTER(C.c_int), C.POINTER(c_i64), C.POINTER(c_i64)], C.c_int),
    "bfg_table_create": ([C.POINTER(c_ptr), C.c_int, C.POINTER(c_i64), C.POINTER(c_ptr), c_ptr, C.c_int, C.c_int], C.c_int),
    "bfg_table_destroy": ([c_ptr], C.c_int),
    "bfg_table_info": ([c_ptr, C.POINTER(C.c_int), C.POINTER(c_i64), C.POINTER(C.c_int), C.POINTER(C.c_int),
                        C.POINTER(C.c_int)], C.c_int),
    "bfg_table_readout": ([c_ptr, c_dbl, c_dbl, c_ptr, c_i64, c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_healpix_disc_counts": ([C.c_int, c_i64, c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_healpix_query_disc": ([C.c_int, c_ptr, c_ptr, c_i64, c_ptr, c_ptr], C.c_int),
    "bfg_healpix_pix2vec": ([C.c_int, c_i64, c_i64, c_ptr, c_ptr], C.c_int),
    "bfg_healpix_interp_weights": ([C.c_int, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_healpix_reorder": ([C.c_int, C.c_int, c_i64, c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_healpix_ang2pix": ([C.c_int, c_i64, c_ptr, c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_shell_offsets": ([c_ptr, C.c_int, c_i64, c_ptr, c_ptr, C.c_int, c_ptr, c_i64, c_i64, c_ptr, c_ptr], C.c_int),
    "bfg_shell_paint": ([c_ptr, C.c_int, c_i64, c_ptr, c_ptr, C.c_int, c_ptr, c_i64, c_i64, c_ptr, c_ptr], C.c_int),
    "bfg_shell_paint_anis": ([c_ptr, c_ptr, C.c_int, c_i64, c_ptr, c_ptr, C.c_int, c_ptr, c_dbl, c_ptr, c_ptr, c_i64, c_i64,
                              c_ptr, c_ptr], C.c_int),
    "bfg_anis_background": ([c_i64, c_ptr, c_dbl, c_ptr, c_dbl, c_dbl, c_ptr, c_ptr], C.c_int),
    "bfg_shell_regrid": ([C.c_int, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_ptr], C.c_int),
    "bfg_shell_regrid_range": ([C.c_int, c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr], C.c_int),
    "bfg_offsets_max_norm2": ([c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr], C.c_int),
    "bfg_halo_band_bounds": ([c_i64, c_ptr, c_dbl, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_shell_records": ([c_i64, c_ptr, C.c_int, c_dbl, c_dbl, c_dbl, C.c_int, c_ptr, c_ptr, C.c_int, c_ptr, c_ptr, c_ptr,
                           c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_shell_regrid_p2p": ([C.c_int, c_ptr, c_ptr, c_i64, c_i64, C.c_int, C.c_int, C.POINTER(c_i64), C.POINTER(c_ptr),
                              c_ptr, c_ptr], C.c_int),
    "bfg_shell_regrid_p2p_range": ([C.c_int, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i64, C.c_int, C.c_int, C.POINTER(c_i64),
                                    C.POINTER(c_ptr), c_ptr, c_ptr], C.c_int),
    "bfg_shared_alloc": ([C.POINTER(c_ptr), c_i64, C.c_int], C.c_int),
    "bfg_shared_free": ([c_ptr], C.c_int),
    "bfg_ipc_export": ([c_ptr, c_ptr], C.c_int),
    "bfg_ipc_import": ([c_ptr, C.POINTER(c_ptr)], C.c_int),
    "bfg_ipc_close": ([c_ptr], C.c_int),
    "bfg_host_register": ([c_ptr, c_i64], C.c_int),
    "bfg_host_unregister": ([c_ptr], C.c_int),
    "bfg_copy_to_host_async": ([c_ptr, c_ptr, c_i64, c_ptr], C.c_int),
    "bfg_box_records": ([c_i64, c_ptr, C.c_int, C.c_int, C.c_int, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl,
                         c_i64, c_ptr, c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_grid_offsets": ([c_ptr, C.c_int, c_i64, c_dbl, c_i64, c_ptr, c_ptr, C.c_int, C.c_int, c_ptr, c_i64, c_i64, c_ptr,
                          c_ptr], C.c_int),
    "bfg_grid_paint": ([c_ptr, C.c_int, c_i64, c_dbl, c_dbl, c_i64, c_ptr, c_ptr, C.c_int, C.c_int, c_ptr, c_i64, c_i64,
                        c_ptr, c_ptr], C.c_int),
    "bfg_grid_paint_anis": ([c_ptr, c_ptr, c_i64, c_dbl, c_i64, c_ptr, c_ptr, C.c_int, C.c_int, c_ptr, c_dbl, c_ptr, c_ptr,
                             c_i64, c_i64, c_ptr, c_ptr], C.c_int),
    "bfg_grid_regrid": ([C.c_int, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_ptr], C.c_int),
    "bfg_snap_build_cells": ([C.c_int, c_i64, c_ptr, c_ptr, c_ptr, c_dbl, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_snap_build_cells_strided": ([C.c_int, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_dbl, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr,
                                      c_ptr, c_ptr], C.c_int),
    "bfg_snap_apply_records": ([C.c_int, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int,
                                c_ptr], C.c_int),
    "bfg_snap_offsets": ([c_ptr, C.c_int, c_i64, c_ptr, c_ptr, c_ptr, c_dbl, C.c_int, c_ptr, c_i64, c_ptr, c_ptr, C.c_int,
                          c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_snap_apply": ([C.c_int, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_snap_deposit_ngp": ([C.c_int, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_dbl, c_i64, c_ptr, c_ptr], C.c_int),
    "bfg_snap_apply_deposit": ([C.c_int, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_dbl, c_dbl, c_i64, c_ptr, c_ptr],
                               C.c_int),
    "bfg_snap_deposit_folded": ([c_i64, c_ptr, c_ptr, c_ptr, c_dbl, c_i64, c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_snap_apply_deposit_folded": ([c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_dbl, c_dbl, c_i64, c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_power_bin_spectrum": ([c_i64, c_ptr, c_ptr, c_dbl, c_dbl, c_i64, c_ptr, c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_grid_power_spectrum": ([c_i64, c_ptr, c_ptr, c_dbl, c_dbl, c_i64, c_ptr, c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_sht_workspace_elems": ([C.c_int, C.c_int], c_i64),
    "bfg_sht_map2alm_pass": ([C.c_int, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_sht_alm2map": ([C.c_int, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_sht_alm2cl": ([C.c_int, c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_test_sht_ring_host": ([c_i64, C.c_int, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_test_sht_legendre_host": ([C.c_int, C.c_int, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_test_sht_lambda_host": ([C.c_int, C.c_int, c_dbl, c_dbl, c_dbl, c_ptr], C.c_int),
    "bfg_halo_sort": ([C.c_int, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, c_dbl, c_dbl, C.c_int, c_ptr], C.c_int),
    "bfg_halo_sort_owned": ([C.c_int, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, c_dbl, c_ptr], C.c_int),
    "bfg_test_regrid_target_host": ([C.c_int, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_test_shell_records_host": ([c_i64, c_ptr, C.c_int, c_dbl, c_dbl, c_dbl, C.c_int, c_ptr, c_ptr, C.c_int, c_ptr, c_ptr, c_ptr,
                                     c_ptr, c_ptr], C.c_int),
    "bfg_test_table_readout_host": ([C.c_int, C.POINTER(c_i64), C.POINTER(c_ptr), c_ptr, C.c_int, C.c_int, c_dbl, c_dbl, c_ptr, c_i64,
                                     c_ptr, c_ptr], C.c_int),
    "bfg_test_shell_update_host": ([C.c_int, C.POINTER(c_i64), C.POINTER(c_ptr), c_ptr, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr,
                                    c_i64, c_ptr, c_ptr], C.c_int),
    "bfg_test_axis_deposit_host": ([c_i64, c_ptr, c_i64, c_ptr, c_ptr], C.c_int),
    "bfg_test_row_at_r2_host": ([C.c_int, C.POINTER(c_i64), C.POINTER(c_ptr), c_ptr, C.c_int, c_dbl, c_dbl, c_ptr, c_dbl, c_i64, c_ptr,
                                 c_ptr, c_ptr], C.c_int),
    "bfg_test_index_helpers_host": ([C.c_int, c_i64, c_ptr, c_dbl, c_i64, c_ptr, c_ptr], C.c_int),
    "bfg_test_healpix_host": ([C.c_int, C.c_int, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr], C.c_int),
    "bfg_test_fast_log2_host": ([c_i64, c_ptr, c_ptr], C.c_int),
    "bfg_test_fast_log2": ([c_i64, c_ptr, c_ptr, c_ptr], C.c_int),
    "bfg_sum_f64": ([c_ptr, c_i64, c_ptr, c_ptr], C.c_int),
    "bfg_transpose_offsets": ([c_ptr, c_ptr, c_i64, C.c_int, c_ptr], C.c_int),
    "bfg_shell_baryonify_host": ([c_ptr, C.c_int, c_i64, c_ptr, c_ptr, C.c_int, c_ptr, c_ptr, C.POINTER(c_i64),
                                  C.POINTER(c_dbl)], C.c_int),
    "bfg_shell_paint_host": ([c_ptr, C.c_int, c_i64, c_ptr, c_ptr, C.c_int, c_ptr, C.POINTER(c_i64)], C.c_int),
}


class BFGError(RuntimeError):
    pass


def exported_symbols():
    return sorted(_SIGNATURES)


def lib():
    """Load libbfg_b200.so; raises (never falls back) when it is missing."""
    global _LIB
    if _LIB is None:
        path = os.environ.get("BFG_LIB", LIB_PATH)   # BFG_LIB: an alternative build of the same ABI (kernel tuning runs)
        if not os.path.exists(path):
            raise BFGError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). baryonforge_b200 has no CPU fallback.")
        L = C.CDLL(path)
        for name, (args, res) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        msg = lib().bfg_last_error()
        raise BFGError(f"libbfg_b200 call failed ({rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
