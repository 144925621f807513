"""ctypes binding of the C ABI declared in include/adfvm_b200.h.

The product library is `adfvm_b200/csrc/libadfvm_b200.so` (CUDA, sm_100a). There is NO CPU fallback:
if it is missing or no CUDA device is usable, loading / context creation raises. (tests/hostsim builds a
CPU simulator of the device code with the same ABI; only tests pass its path explicitly.)
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(HERE, "csrc", "libadfvm_b200.so")

# enums of include/adfvm_b200.h
MU_CONSTANT, MU_SUTHERLAND = 0, 1
RIEMANN = {"eulerRoe": 0, "eulerLaxFriedrichs": 1}
PATCH_WALL, PATCH_CYCLIC, PATCH_SYMMETRY, PATCH_EMPTY, PATCH_CHARACTERISTIC, PATCH_PROCESSOR, PATCH_PROCESSOR_CYCLIC = range(7)
BC_CALCULATED, BC_CYCLIC, BC_ZEROGRADIENT, BC_FIXEDVALUE, BC_SYMMETRY, BC_CBC_UPT, BC_CBC_TOTAL_PT, BC_PROCESSOR = range(8)
KEY_VALUE_U, KEY_VALUE_T, KEY_VALUE_P, KEY_U0, KEY_T0, KEY_P0, KEY_TT, KEY_PT, KEY_DIRECTION = range(9)
OBJ_NONE, OBJ_CELL_TV, OBJ_PATCH_PA, OBJ_DRAG = range(4)
OBJ_PLANE_PTLOSS, OBJ_CELL_T, OBJ_CALLBACK = 4, 5, 6
OBJECTIVE_FN = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_void_p)
RETURN_STATIC, ZERO_STATIC, REPLACE_STATIC, RETURN_REUSABLE, REPLACE_REUSABLE, VISCOUS = 1, 2, 4, 8, 16, 32
VISC = {None: 0, "abarbanel": 1, "turkel": 2, "uniform": 3}


class Patch(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("startFace", "nFaces", "cellStartFace", "mesh_type", "bc_U", "bc_T", "bc_p",
                                         "neighbour_patch", "peer_rank", "tag")]


class MetricPatch(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("startFace", "nFaces", "kind", "nbrStartFace")]


class AdfvmError(RuntimeError):
    pass


EXPORTS = ["adfvm_last_error", "adfvm_version", "adfvm_is_cuda", "adfvm_create", "adfvm_destroy", "adfvm_set_physics",
           "adfvm_set_mesh", "adfvm_set_bc_value", "adfvm_set_objective", "adfvm_set_source", "adfvm_primal",
           "adfvm_primal_grad", "adfvm_primal_step_resident", "adfvm_adjoint_step_resident", "adfvm_get_dtc_obj",
           "adfvm_get_state", "adfvm_sync", "adfvm_launch_count", "adfvm_device_bytes", "adfvm_comm_unique_id",
           "adfvm_comm_init", "adfvm_kernel_timing", "adfvm_kernel_report", "adfvm_set_tile_cells", "adfvm_tile_stats", "adfvm_tile_halo_stats",
           "adfvm_host_alloc", "adfvm_host_free", "adfvm_tile_rounds", "adfvm_graph_replays",
           "adfvm_set_state", "adfvm_mesh_metrics", "adfvm_set_objective_plane", "adfvm_set_parameter_bc", "adfvm_set_parameter_mesh", "adfvm_get_mesh_grad", "adfvm_primal_block", "adfvm_adjoint_block", "adfvm_set_adjoint", "adfvm_get_adjoint",
           "adfvm_set_objective_callback", "adfvm_get_cell_perm", "adfvm_init_fields", "adfvm_get_dtc_global",
           "adfvm_set_adjoint_viscosity", "adfvm_adjoint_viscous_resident", "adfvm_get_adjoint_viscosity", "adfvm_viscosity_iterations",
           "adfvm_adjoint_block_viscous", "adfvm_state_cache_reserve", "adfvm_state_cache_put", "adfvm_state_cache_select",
           "adfvm_state_cache_hits"]


class Lib:
    def __init__(self, path=None):
        path = path or DEFAULT_LIB
        if not os.path.exists(path):
            raise AdfvmError("native library %s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback)" % path)
        self.path = path
        self.dll = C.CDLL(path)
        d = self.dll
        vp, i32, f64 = C.c_void_p, C.c_int32, C.c_double
        d.adfvm_last_error.restype = C.c_char_p
        d.adfvm_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, vp]
        d.adfvm_destroy.argtypes = [vp]
        d.adfvm_set_physics.argtypes = [vp, f64, f64, f64, C.c_int, f64, C.c_int, C.c_int]
        d.adfvm_set_mesh.argtypes = [vp, C.POINTER(i32)] + [vp] * 10 + [vp] * 5 + [i32, C.POINTER(Patch)]
        d.adfvm_set_bc_value.argtypes = [vp, i32, i32, vp]
        d.adfvm_set_objective.argtypes = [vp, i32, i32, i32]
        d.adfvm_set_source.argtypes = [vp, vp, vp, vp]
        d.adfvm_set_parameter_bc.argtypes = [vp, i32, i32]
        d.adfvm_set_parameter_mesh.argtypes = [vp]
        d.adfvm_get_mesh_grad.argtypes = [vp] + [vp] * 10 + [i32]
        d.adfvm_set_objective_plane.argtypes = [vp, i32, vp, vp, f64, C.POINTER(f64), f64]
        d.adfvm_primal.argtypes = [vp, vp, vp, vp, f64, i32, vp, vp, vp, vp, vp]
        d.adfvm_primal_grad.argtypes = [vp, vp, vp, vp, f64, vp, vp, vp, f64, f64, i32, vp, vp, vp, vp, vp, vp]
        d.adfvm_primal_step_resident.argtypes = [vp, f64]
        d.adfvm_set_adjoint_viscosity.argtypes = [vp, i32, f64, f64, i32]
        d.adfvm_adjoint_viscous_resident.argtypes = [vp, f64]
        d.adfvm_state_cache_reserve.argtypes = [vp, i32]
        d.adfvm_state_cache_put.argtypes = [vp, C.c_int64]
        d.adfvm_state_cache_select.argtypes = [vp, C.c_int64, C.POINTER(i32)]
        d.adfvm_state_cache_hits.argtypes = [vp]
        d.adfvm_state_cache_hits.restype = C.c_int64
        d.adfvm_adjoint_block_viscous.argtypes = [vp, i32, C.POINTER(f64), f64]
        d.adfvm_get_adjoint_viscosity.argtypes = [vp, vp, vp, vp, vp]
        d.adfvm_viscosity_iterations.argtypes = [vp]
        d.adfvm_viscosity_iterations.restype = C.c_int64
        d.adfvm_adjoint_step_resident.argtypes = [vp, f64, f64, i32]
        d.adfvm_get_dtc_obj.argtypes = [vp, C.POINTER(f64), C.POINTER(f64)]
        d.adfvm_get_state.argtypes = [vp, vp, vp, vp]
        d.adfvm_sync.argtypes = [vp]
        d.adfvm_launch_count.argtypes = [vp]; d.adfvm_launch_count.restype = C.c_int64
        d.adfvm_device_bytes.argtypes = [vp]; d.adfvm_device_bytes.restype = C.c_int64
        d.adfvm_kernel_timing.argtypes = [vp, i32]
        d.adfvm_kernel_report.argtypes = [vp, C.c_char_p, i32]
        d.adfvm_set_tile_cells.argtypes = [vp, i32]
        d.adfvm_tile_stats.argtypes = [vp, C.POINTER(f64), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
        d.adfvm_tile_halo_stats.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), i32]
        d.adfvm_comm_unique_id.argtypes = [vp]
        d.adfvm_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
        d.adfvm_host_free.argtypes = [vp]
        d.adfvm_primal_block.argtypes = [vp, i32, C.POINTER(f64), C.POINTER(f64), C.POINTER(f64)]
        d.adfvm_adjoint_block.argtypes = [vp, i32, C.POINTER(f64), f64]
        d.adfvm_set_adjoint.argtypes = [vp, vp, vp, vp]
        d.adfvm_set_state.argtypes = [vp, vp, vp, vp]
        d.adfvm_mesh_metrics.argtypes = [vp, i32, vp, i32, i32, i32, vp, vp, vp, vp, i32, C.POINTER(MetricPatch), vp] + [vp] * 10
        d.adfvm_get_adjoint.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32]
        d.adfvm_graph_replays.argtypes = [vp]; d.adfvm_graph_replays.restype = C.c_int64
        d.adfvm_tile_rounds.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(i32)]
        d.adfvm_comm_init.argtypes = [vp, vp, i32, i32]
        d.adfvm_set_objective_callback.argtypes = [vp, OBJECTIVE_FN, vp]
        d.adfvm_get_cell_perm.argtypes = [vp, C.POINTER(i32)]
        d.adfvm_init_fields.argtypes = [vp] + [vp] * 6
        d.adfvm_get_dtc_global.argtypes = [vp, C.POINTER(f64)]

    @property
    def is_cuda(self):
        return bool(self.dll.adfvm_is_cuda())

    def check(self, rc):
        if rc != 0:
            raise AdfvmError(self.dll.adfvm_last_error().decode())


_default = None


def default_lib():
    """The CUDA product library; raises if it has not been built."""
    global _default
    if _default is None:
        _default = Lib(os.environ.get("ADFVM_B200_LIB") or DEFAULT_LIB)     # (override: kernel experiments with a second build of the same sources)
    return _default
