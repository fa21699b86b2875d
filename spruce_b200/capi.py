"""ctypes binding of the C ABI in include/spruce_b200.h (one function per exported symbol, same names).

This is the product-side Python binding (test infrastructure is never imported here).  No CPU fallback: if the shared
library is missing or no CUDA device is usable the calls raise SpruceError.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "lib" / "libspruce_b200.so"
ABI_VERSION = 1

BC = {"periodic": 0, "open": 1, "fixed": 2, "reflect": 3, "open_moc": 4, "open_ucnp": 5}    # plasmadomain.hpp:22-26
TI = {"euler": 0, "rk2": 1, "rk4": 2}                                                        # plasmadomain.hpp:29-32
EQS = {"ideal_mhd": 0, "ideal_mhd_2E": 2, "ideal_2F": 3}                                                        # equationset.hpp:22


class SpruceError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("equation_set", C.c_int32), ("xdim", C.c_int32), ("ydim", C.c_int32),
                ("x_bound_1", C.c_int32), ("x_bound_2", C.c_int32), ("y_bound_1", C.c_int32), ("y_bound_2", C.c_int32),
                ("time_integrator", C.c_int32), ("device", C.c_int32),
                ("row0", C.c_int32), ("nx_local", C.c_int32), ("rank", C.c_int32), ("n_ranks", C.c_int32),
                ("ion_mass", C.c_double), ("adiabatic_index", C.c_double), ("epsilon", C.c_double),
                ("density_min", C.c_double), ("temp_min", C.c_double), ("thermal_energy_min", C.c_double),
                ("open_boundary_strength", C.c_double), ("open_boundary_decay_base", C.c_double), ("time", C.c_double)]


# name -> (restype, argtypes); every symbol include/spruce_b200.h declares
_DP = C.POINTER(C.c_double)
_VPP = C.POINTER(C.c_void_p)
SYMBOLS = {
    "spruce_last_error": (C.c_char_p, []),
    "spruce_abi_version": (C.c_int, []),
    "spruce_domain_create": (C.c_int, [C.POINTER(Config), _VPP]),
    "spruce_domain_destroy": (None, [C.c_void_p]),
    "spruce_set_cell_sizes": (C.c_int, [C.c_void_p, _DP, C.c_size_t, _DP, C.c_size_t]),
    "spruce_grid_upload": (C.c_int, [C.c_void_p, C.c_char_p, _DP, C.c_size_t]),
    "spruce_grid_download": (C.c_int, [C.c_void_p, C.c_char_p, _DP, C.c_size_t]),
    "spruce_eqs_setup": (C.c_int, [C.c_void_p]),
    "spruce_eqs_propagate_changes": (C.c_int, [C.c_void_p]),
    "spruce_next_step_size": (C.c_int, [C.c_void_p, _DP]),
    "spruce_advance": (C.c_int, [C.c_void_p, C.c_int, C.c_double, _DP, C.POINTER(C.c_int)]),
    "spruce_get_time": (C.c_int, [C.c_void_p, _DP, C.POINTER(C.c_int64)]),
    "spruce_eqs_time_derivatives": (C.c_int, [C.c_void_p, _DP, C.c_size_t]),
    "spruce_operator": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, _DP, _DP, _DP, C.c_size_t]),
    "spruce_module_thermal_conduction": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]),
    "spruce_module_radiative_losses": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int]),
    "spruce_module_ambient_heating": (C.c_int, [C.c_void_p, _DP, C.c_size_t]),
    "spruce_module_viscosity": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int]),
    "spruce_module_viscosity_term": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double, C.c_char_p, C.c_char_p, C.c_char_p, _DP, C.c_size_t]),
    "spruce_module_physical_viscosity": (C.c_int, [C.c_void_p, C.c_double, C.POINTER(C.c_double), C.c_size_t, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "spruce_module_ambient_heating_sink": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "spruce_module_localized_heating": (C.c_int, [C.c_void_p] + [C.c_double] * 8),
    "spruce_module_mass_injection": (C.c_int, [C.c_void_p] + [C.c_double] * 7),
    "spruce_module_momentum_injection": (C.c_int, [C.c_void_p] + [C.c_double] * 10 + [C.c_int, C.c_double]),
    "spruce_module_div_cleaning": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "spruce_module_field_heating": (C.c_int, [C.c_void_p] + [C.c_double] * 5 + [C.c_int]),
    "spruce_module_boundary_outflow": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                                 C.c_double, C.c_double]),
    "spruce_module_boundary_outflow_state": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "spruce_module_anomalous_resistivity": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]),
    "spruce_module_anomalous_resistivity_state": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "spruce_module_output_to_file": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "spruce_module_output": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]),
    "spruce_multispecies_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "spruce_module_inactive_mode": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "spruce_multispecies_reset": (C.c_int, [C.c_void_p]),
    "spruce_module_ms_fraction": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "spruce_eqs_ideal_mhd_options": (C.c_int, [C.c_void_p, C.c_double]),
    "spruce_eqs_ideal_mhd_moc_limiting": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double]),
    "spruce_eqs_ideal2f_options": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "spruce_module_eic_thermalization": (C.c_int, [C.c_void_p]),
    "spruce_module_subcycles": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]),
    "spruce_halo_buffers": (C.c_int, [C.c_void_p, _VPP, _VPP, _VPP, _VPP, C.POINTER(C.c_size_t)]),
    "spruce_mgpu_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "spruce_mgpu_ipc_connect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "spruce_mgpu_initial_exchange": (C.c_int, [C.c_void_p]),
    "spruce_mgpu_pack": (C.c_int, [C.c_void_p, C.c_int]),
    "spruce_mgpu_unpack": (C.c_int, [C.c_void_p, C.c_int]),
    "spruce_mgpu_stage": (C.c_int, [C.c_void_p, C.c_int]),
    "spruce_mgpu_n_stages": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "spruce_mgpu_stage_output": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "spruce_mgpu_dt_min_ptr": (C.c_int, [C.c_void_p, _VPP]),
    "spruce_mgpu_begin_step": (C.c_int, [C.c_void_p]),
    "spruce_mgpu_end_step": (C.c_int, [C.c_void_p]),
    "spruce_operator2": (C.c_int, [C.c_void_p, C.c_char_p, _DP, _DP, _DP, _DP, C.c_size_t]),
    "spruce_plane_activity": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.c_int]),
    "spruce_stream": (C.c_int, [C.c_void_p, _VPP]),
    "spruce_synchronize": (C.c_int, [C.c_void_p]),
    "spruce_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "spruce_time_stage_kernel": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library.  Raises SpruceError (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise SpruceError("CUDA library %s is missing: run `python -m spruce_b200.build` (there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        if L.spruce_abi_version() != ABI_VERSION:
            raise SpruceError("ABI mismatch")
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise SpruceError("spruce_b200 error %d: %s" % (rc, load().spruce_last_error().decode()))
