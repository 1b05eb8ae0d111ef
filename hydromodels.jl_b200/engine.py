"""Engine: the drop-in functor `component(input, params, config; initstates)` on the GPU.

`B200Model(component)` is the wrapper component of SURVEY.md §8(b): same call signature and
result layout as the reference functors (`/root/reference/src/model.jl:241-309`,
`src/bucket.jl:169-256`, `src/flux.jl:157-203`, `src/uh.jl:209-274`), executed by the
NVRTC-compiled stepper behind the C-ABI `libhydrob200.so` (`include/hydrob200.h`, bound here
with `ctypes`).  There is NO CPU path: if the library is missing, or there is no CUDA device,
calls raise.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .components import (Component, ConstantInterpolation, D8_CODES, GridAggregation, HydroBucket, HydroConfig, HydroFlux,
                         HydroModel, HydroRoute, LinearInterpolation, NeuralFlux, SolverType, UnitHydrograph, normalize_config)
from .lowering import GridProgram, LoweringError, RouteProgram, StepperProgram, lower_grid, lower_route, lower_stepper
from .symbolic import evaluate_np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhydrob200.so")

HB_MAX_UH = 8
HB_OUT_ROWS, HB_OUT_OBJECTIVE = 0, 1
METRICS = {"mse": 0, "nse": 1, "kge": 2, "logkge": 3, "sse": 4}
HB_MAX_TANGENTS = 16


class HydroB200Error(RuntimeError):
    pass


class NoDeviceError(HydroB200Error):
    pass


class StepperSpec(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("n_inputs", C.c_int32), ("n_rows", C.c_int32), ("n_states", C.c_int32),
                ("n_params", C.c_int32), ("n_uh", C.c_int32), ("uh_kmax", C.c_int32 * HB_MAX_UH),
                ("nn_words", C.c_int32), ("output_mode", C.c_int32), ("metric", C.c_int32),
                ("shared_forcing", C.c_int32), ("block_threads", C.c_int32), ("stages", C.c_int32),
                ("max_registers", C.c_int32), ("precision", C.c_int32), ("nn_tensor", C.c_int32), ("nodes_per_warp", C.c_int32),
                ("n_tangents", C.c_int32), ("n_aux_inputs", C.c_int32), ("nn_tc_bytes", C.c_int32)]


class RunDesc(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("n_htype_sets", C.c_int32), ("n_nodes", C.c_int64), ("n_steps", C.c_int64),
                ("input", C.c_void_p), ("params", C.c_void_p), ("n_param_values", C.c_int64),
                ("param_offset", C.c_void_p), ("param_mode", C.c_void_p), ("htypes", C.c_void_p),
                ("initstates", C.c_void_p), ("uh_weights", C.c_void_p), ("uh_hist_in", C.c_void_p),
                ("nn_params", C.c_void_p), ("target", C.c_void_p), ("target_per_node", C.c_int32),
                ("warm_up", C.c_int32), ("min_value", C.c_double), ("out", C.c_void_p), ("objective", C.c_void_p),
                ("final_states", C.c_void_p), ("uh_hist_out", C.c_void_p), ("elapsed_ms", C.c_void_p),
                ("gradient", C.c_void_p), ("aux_input", C.c_void_p), ("n_compact_rows", C.c_int32), ("compact_rows", C.c_int32 * 4)]


class KernelInfo(C.Structure):
    _fields_ = [("registers_per_thread", C.c_int32), ("static_smem_bytes", C.c_int32), ("dynamic_smem_bytes", C.c_int32),
                ("local_bytes_per_thread", C.c_int32), ("block_threads", C.c_int32), ("max_blocks_per_sm", C.c_int32),
                ("from_cache", C.c_int32), ("cubin_bytes", C.c_int32)]


class RouteSpec(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("n_inputs", C.c_int32), ("n_params", C.c_int32), ("block_threads", C.c_int32)]


class RouteDesc(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("n_htype_sets", C.c_int32), ("n_nodes", C.c_int64), ("n_steps", C.c_int64),
                ("records", C.c_void_p), ("n_rows", C.c_int32), ("s_row", C.c_int32), ("o_row", C.c_int32),
                ("in_row", C.c_void_p), ("params", C.c_void_p), ("n_param_values", C.c_int64),
                ("param_offset", C.c_void_p), ("param_mode", C.c_void_p), ("htypes", C.c_void_p),
                ("initstates", C.c_void_p), ("up_ptr", C.c_void_p), ("up_idx", C.c_void_p), ("min_value", C.c_double),
                ("final_states", C.c_void_p), ("elapsed_ms", C.c_void_p), ("route_inputs", C.c_void_p), ("route_out", C.c_void_p)]


class GridSpec(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("n_inputs", C.c_int32), ("n_rows", C.c_int32), ("n_states", C.c_int32),
                ("n_params", C.c_int32), ("n_params_stepper", C.c_int32), ("n_route_inputs", C.c_int32),
                ("s_row", C.c_int32), ("o_row", C.c_int32), ("steps_per_chunk", C.c_int32), ("block_w", C.c_int32),
                ("block_h", C.c_int32), ("variant", C.c_int32), ("stages", C.c_int32), ("max_registers", C.c_int32)]


class GridDesc(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("n_htype_sets", C.c_int32), ("rows", C.c_int32), ("cols", C.c_int32),
                ("n_steps", C.c_int64), ("input", C.c_void_p), ("out", C.c_void_p), ("params", C.c_void_p),
                ("n_param_values", C.c_int64), ("param_offset", C.c_void_p), ("param_mode", C.c_void_p),
                ("htypes", C.c_void_p), ("init_bucket", C.c_void_p), ("init_route", C.c_void_p), ("flwdir", C.c_void_p),
                ("min_value", C.c_double), ("final_bucket", C.c_void_p), ("final_route", C.c_void_p),
                ("elapsed_ms", C.c_void_p), ("out_row0", C.c_int32), ("out_rows", C.c_int32), ("out_col0", C.c_int32),
                ("out_cols", C.c_int32), ("in_t_stride", C.c_int64), ("out_t_stride", C.c_int64), ("in_pitch", C.c_int32),
                ("in_col0", C.c_int32), ("out_pitch", C.c_int32), ("out_base", C.c_int32)]


class AdjointSpec(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("n_inputs", C.c_int32), ("n_states", C.c_int32), ("n_params", C.c_int32),
                ("nn_words", C.c_int32), ("block_threads", C.c_int32), ("acc_doubles_per_warp", C.c_int32)]


class AdjointDesc(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("n_htype_sets", C.c_int32), ("n_nodes", C.c_int64), ("n_steps", C.c_int64),
                ("input", C.c_void_p), ("trajectory", C.c_void_p), ("params", C.c_void_p), ("n_param_values", C.c_int64),
                ("param_offset", C.c_void_p), ("param_mode", C.c_void_p), ("htypes", C.c_void_p), ("initstates", C.c_void_p),
                ("nn_params", C.c_void_p), ("target", C.c_void_p), ("target_per_node", C.c_int32), ("warm_up", C.c_int32),
                ("min_value", C.c_double), ("scale", C.c_void_p), ("grad_params", C.c_void_p), ("grad_init", C.c_void_p),
                ("grad_nn", C.c_void_p), ("loss", C.c_void_p), ("elapsed_ms", C.c_void_p)]


class HaloEdge(C.Structure):
    _fields_ = [("peer", C.c_int32), ("s_row0", C.c_int32), ("s_col0", C.c_int32), ("r_row0", C.c_int32), ("r_col0", C.c_int32),
                ("nrows", C.c_int32), ("ncols", C.c_int32), ("peer_slot", C.c_int32)]


_lib = None


def lib():
    """Load `libhydrob200.so` (built in-tree by `__graft_entry__.build()`); fail loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HydroB200Error(f"{LIB_PATH} not found — run `python -c 'import __graft_entry__ as g; g.build()'`; "
                             "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.hb_last_error.restype = C.c_char_p
    L.hb_version.restype = C.c_int
    for name, args in {
        "hb_device_count": [C.POINTER(C.c_int)], "hb_init": [C.c_int], "hb_set_cache_dir": [C.c_char_p],
        "hb_compile_stepper": [C.POINTER(StepperSpec), C.c_char_p, C.POINTER(C.c_void_p)],
        "hb_model_source": [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_int64)],
        "hb_model_cubin": [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)],
        "hb_model_info": [C.c_void_p, C.POINTER(KernelInfo)], "hb_free_model": [C.c_void_p],
        "hb_run": [C.c_void_p, C.POINTER(RunDesc)], "hb_run_device": [C.c_void_p, C.POINTER(RunDesc), C.c_void_p],
        "hb_compile_route": [C.POINTER(RouteSpec), C.c_char_p, C.POINTER(C.c_void_p)],
        "hb_compile_adjoint": [C.POINTER(AdjointSpec), C.c_char_p, C.POINTER(C.c_void_p)],
        "hb_run_adjoint_device": [C.c_void_p, C.POINTER(AdjointDesc), C.c_void_p],
        "hb_compile_grid": [C.POINTER(GridSpec), C.c_char_p, C.POINTER(C.c_void_p)],
        "hb_run_grid_device": [C.c_void_p, C.POINTER(GridDesc), C.c_void_p],
        "hb_grid_status": [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)],
        "hb_run_route_device": [C.c_void_p, C.POINTER(RouteDesc), C.c_void_p],
        "hb_route_step_device": [C.c_void_p, C.POINTER(RouteDesc), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p],
        "hb_d8_flow_direction_device": [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_void_p],
        "hb_pack_forcing_device": [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                   C.c_void_p, C.c_void_p],
        "hb_pack_forcing_dense_device": [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                         C.c_void_p, C.c_void_p],
        "hb_ingest_forcing": [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                              C.c_int64, C.c_void_p, C.c_void_p],
        "hb_de_trial_device": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                               C.c_uint64, C.c_uint32, C.c_void_p],
        "hb_de_select_device": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                C.c_void_p],
        "hb_irf_compose_device": [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                  C.c_void_p],
        "hb_irf_reduce_segments_device": [C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
        "hb_irf_gather_rows_device": [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p],
        "hb_de_best_device": [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p],
        "hb_de_trial_sharded_device": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_double, C.c_double, C.c_uint64, C.c_uint32, C.c_void_p],
        "hb_de_select_sharded_device": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_int32, C.c_void_p],
        "hb_ade_init_device": [C.c_void_p, C.c_void_p, C.c_int32, C.c_uint64, C.c_void_p],
        "hb_ade_trial_device": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_int32, C.c_uint64, C.c_uint32, C.c_void_p],
        "hb_ade_select_device": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                 C.c_uint64, C.c_uint32, C.c_void_p, C.c_int32, C.c_void_p],
        "hb_malloc_device": [C.POINTER(C.c_void_p), C.c_int64], "hb_free_device": [C.c_void_p],
        "hb_malloc_host": [C.POINTER(C.c_void_p), C.c_int64], "hb_free_host": [C.c_void_p],
        "hb_memcpy_h2d": [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p],
        "hb_memcpy_d2h": [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p],
        "hb_memset_device": [C.c_void_p, C.c_int, C.c_int64, C.c_void_p], "hb_stream_synchronize": [C.c_void_p],
        "hb_convert_f64_f32_device": [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p],
        "hb_stream_create": [C.POINTER(C.c_void_p)], "hb_stream_destroy": [C.c_void_p],
        "hb_probe_tc_layer": [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p],
        "hb_probe_occupy_sms": [C.c_int, C.c_int, C.c_double, C.c_void_p],
        "hb_probe_fp64_tflops": [C.POINTER(C.c_double)], "hb_probe_dmma_tflops": [C.POINTER(C.c_double)],
        "hb_fill_device": [C.c_void_p, C.c_int64, C.c_double, C.c_void_p],
        "hb_comm_unique_id": [C.c_void_p, C.c_int32], "hb_comm_init": [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)],
        "hb_comm_rank": [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)], "hb_comm_free": [C.c_void_p],
        "hb_comm_all_gather": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p],
        "hb_comm_alloc": [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int32)],
        "hb_comm_peer_ptr": [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)],
        "hb_halo_enable_peer": [C.c_void_p, C.c_int64],
        "hb_halo_exchange_device": [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.POINTER(HaloEdge), C.c_int32, C.c_int32,
                                    C.c_void_p],
        "hb_halo_status": [C.c_void_p, C.c_void_p],
        "hb_probe_pcie_gbs": [C.c_int64, C.c_int, C.POINTER(C.c_double)], "hb_host_threads": [C.POINTER(C.c_int)],
        "hb_host_copy": [C.c_void_p, C.c_void_p, C.c_int64],
    }.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    if L.hb_version() != 1:
        raise HydroB200Error("libhydrob200.so ABI version mismatch")
    _lib = L
    return L


def check(rc: int):
    if rc == 0:
        return
    msg = lib().hb_last_error().decode("utf-8", "replace")
    if rc == 2:
        raise NoDeviceError(msg)
    if rc == 1:
        raise ValueError(msg)
    if rc == 5:
        raise MemoryError(msg)
    raise HydroB200Error(msg)


def device_count() -> int:
    n = C.c_int(0)
    check(lib().hb_device_count(C.byref(n)))
    return n.value


def init(device: int = 0):
    check(lib().hb_init(int(device)))


# ---------------------------------------------------------------------------------------------------
# compiled kernels
# ---------------------------------------------------------------------------------------------------
class CompiledStepper:
    def __init__(self, program: StepperProgram, *, metric: int = 0, shared_forcing: bool = False,
                 block_threads: int = 0, stages: int = 0, max_registers: int = 0, nodes_per_warp: int = 0):
        self.program = program
        spec = StepperSpec()
        spec.struct_size = C.sizeof(StepperSpec)
        spec.n_inputs = len(program.input_names)
        spec.n_rows = 0 if program.objective else len(program.row_names)
        spec.n_states = len(program.state_names)
        spec.n_params = len(program.param_slots)
        spec.n_uh = len(program.uh)
        if spec.n_uh > HB_MAX_UH:
            raise LoweringError(f"at most {HB_MAX_UH} unit hydrographs per model")
        for i, k in enumerate(program.uh_kmax):
            spec.uh_kmax[i] = k
        spec.nn_words = program.nn_words
        spec.output_mode = HB_OUT_OBJECTIVE if program.objective else HB_OUT_ROWS
        spec.metric = metric
        spec.shared_forcing = 1 if shared_forcing else 0
        spec.block_threads = block_threads
        spec.stages = stages
        spec.max_registers = max_registers
        spec.nn_tensor = 1 if getattr(program, "nn_tensor", False) else 0
        spec.nodes_per_warp = nodes_per_warp if spec.nn_tensor else 0
        spec.precision = 1 if getattr(program, "precision", "f64") == "f32" else 0
        spec.n_tangents = len(getattr(program, "tangents", ()) or ())
        spec.nn_tc_bytes = int(getattr(program, "nn_tc_bytes", 0)) if spec.precision else 0
        spec.n_aux_inputs = len(getattr(program, "aux_names", ()) or ())
        self.spec = spec
        h = C.c_void_p()
        check(lib().hb_compile_stepper(C.byref(spec), program.body.encode(), C.byref(h)))
        self.handle = h

    def source(self) -> str:
        s = C.c_char_p()
        n = C.c_int64()
        check(lib().hb_model_source(self.handle, C.byref(s), C.byref(n)))
        return s.value.decode()

    def cubin(self) -> bytes:
        p = C.c_void_p()
        n = C.c_int64()
        check(lib().hb_model_cubin(self.handle, C.byref(p), C.byref(n)))
        return C.string_at(p, n.value)

    def info(self) -> Dict[str, int]:
        k = KernelInfo()
        check(lib().hb_model_info(self.handle, C.byref(k)))
        return {f: getattr(k, f) for f, _ in KernelInfo._fields_}

    def __del__(self):
        try:
            if self.handle and _lib is not None:
                _lib.hb_free_model(self.handle)
                self.handle = None
        except Exception:
            pass


class CompiledAdjoint:
    """The reverse-mode kernel of a per-node model (`lowering.lower_stepper(adjoint=True)` spliced into templates/adjoint.cu)."""

    def __init__(self, program, block_threads: int = 0):
        self.program = program
        spec = AdjointSpec(C.sizeof(AdjointSpec), len(program.input_names), len(program.state_names), len(program.param_slots),
                           program.nn_words, block_threads, program.n_groups)
        self.spec = spec
        h = C.c_void_p()
        check(lib().hb_compile_adjoint(C.byref(spec), program.body.encode(), C.byref(h)))
        self.handle = h

    def source(self) -> str:
        s = C.c_char_p()
        n = C.c_int64()
        check(lib().hb_model_source(self.handle, C.byref(s), C.byref(n)))
        return s.value.decode()

    def cubin(self) -> bytes:
        p = C.c_void_p()
        n = C.c_int64()
        check(lib().hb_model_cubin(self.handle, C.byref(p), C.byref(n)))
        return C.string_at(p, n.value)

    def info(self) -> Dict[str, int]:
        k = KernelInfo()
        check(lib().hb_model_info(self.handle, C.byref(k)))
        return {f: getattr(k, f) for f, _ in KernelInfo._fields_}

    def __del__(self):
        try:
            if self.handle and _lib is not None:
                _lib.hb_free_model(self.handle)
                self.handle = None
        except Exception:
            pass


class CompiledRoute:
    def __init__(self, program: RouteProgram, block_threads: int = 0):
        self.program = program
        spec = RouteSpec(C.sizeof(RouteSpec), len(program.input_names), len(program.param_slots), block_threads)
        self.spec = spec
        h = C.c_void_p()
        check(lib().hb_compile_route(C.byref(spec), program.body.encode(), C.byref(h)))
        self.handle = h

    def source(self) -> str:
        s = C.c_char_p()
        n = C.c_int64()
        check(lib().hb_model_source(self.handle, C.byref(s), C.byref(n)))
        return s.value.decode()

    def __del__(self):
        try:
            if self.handle and _lib is not None:
                _lib.hb_free_model(self.handle)
                self.handle = None
        except Exception:
            pass


def upstream_csr(aggr, n_nodes: int):
    """CSR list of upstream nodes (0-based) — the gather form of `build_aggr_func`
    (`/root/reference/src/route.jl:336-365`).  Grid: upstream nodes ordered by their D8 code, the order
    in which the reference adds its eight shifted arrays."""
    if isinstance(aggr, GridAggregation):
        down = aggr.downstream()
        pos0 = aggr.positions - 1
        codes = aggr.flwdir[pos0[:, 0], pos0[:, 1]]
        lut = np.full(256, 99, dtype=np.int64)
        lut[list(D8_CODES)] = np.arange(8)
        rank = np.where((codes >= 0) & (codes < 256), lut[np.clip(codes, 0, 255)], 99)
        src = np.nonzero(down >= 0)[0]
        order = np.lexsort((rank[src], down[src]))
        src, dst = src[order], down[src][order]
    else:
        if aggr.n_nodes != n_nodes:
            raise ValueError("graph aggregation node count does not match the input")
        e = np.asarray(aggr.edges, dtype=np.int64).reshape(-1, 2) - 1
        order = np.argsort(e[:, 1], kind="stable") if len(e) else np.zeros(0, dtype=np.int64)
        src, dst = e[order, 0], e[order, 1]
    ptr = np.zeros(n_nodes + 1, dtype=np.int64)
    np.add.at(ptr, dst + 1, 1)
    return np.cumsum(ptr).astype(np.int32), np.ascontiguousarray(src.astype(np.int32) if len(src) else np.zeros(1, dtype=np.int32))


# ---------------------------------------------------------------------------------------------------
# argument preparation (host side of the functor)
# ---------------------------------------------------------------------------------------------------
def _as_componentvector(params):
    if isinstance(params, dict) and "params" not in params and "nns" in params:
        params = dict(params, params={})      # `ComponentVector(nns=(...))`: what a stand-alone NeuralFlux is called with
    if not isinstance(params, dict) or "params" not in params:
        # the reference has no method for plain vectors either (`src/utils.jl:15,22`)
        raise TypeError("params must be a dict {'params': {...}, 'nns': {...}} (ComponentVector stand-in)")
    return params


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class PreparedRun:
    """Host arrays of one call, in the C-ABI layouts."""

    def __init__(self):
        self.keep: List[np.ndarray] = []


def _check_config(cfg: HydroConfig, T: int):
    if cfg.solver == SolverType.ODESolver:
        raise NotImplementedError("ODESolver is not part of the B200 engine (explicit MutableSolver stepping only)")
    if cfg.interpolator not in (ConstantInterpolation, LinearInterpolation):
        raise NotImplementedError("only Constant/Linear interpolation at integer time indices is supported")
    if len(cfg.timeidx) and list(cfg.timeidx) != list(range(1, T + 1)):
        # values index `input` directly and pass 2 uses the whole input (`src/interpolate.jl:39,49`,
        # `src/bucket.jl:190`): only 1:T is well-formed in the reference
        raise ValueError("timeidx must be empty or 1:size(input, end)")


def _uh_weight_table(uh: UnitHydrograph, pvals: Dict[str, np.ndarray]) -> np.ndarray:
    """[K_max, n] normalised ordinates for `n` parameter sets (zero-padded; K = 0 → identity).
    Semantics: `UnitHydrograph.weights` (`/root/reference/src/uh.jl:215-229,188`)."""
    n = len(next(iter(pvals.values()))) if pvals else 1
    lag = np.broadcast_to(np.asarray(evaluate_np(uh.uh_conds[0][0], pvals), dtype=np.float64), (n,))
    K = np.ceil(lag).astype(np.int64)
    K = np.maximum(K, 0)
    kmax = max(1, int(K.max()) if n else 1)
    thr = [np.broadcast_to(np.asarray(evaluate_np(c, pvals), dtype=np.float64), (n,)) for c, _ in uh.uh_conds]
    S = np.ones((kmax + 1, n))
    S[0] = 0.0
    for k in range(1, kmax + 1):
        env = dict(pvals)
        env[uh.tvar] = np.float64(k)
        val = np.ones(n)                     # S = 1 beyond the largest threshold
        best = np.full(n, np.inf)
        for (c, v), th in zip(uh.uh_conds, thr):   # the condition with the smallest threshold >= k wins
            sel = (k <= th) & (th < best)
            if sel.any():
                cand = np.broadcast_to(np.asarray(evaluate_np(v, env), dtype=np.float64), (n,))
                val = np.where(sel, cand, val)
                best = np.where(sel, th, best)
        S[k] = val
    w = S[1:] - S[:-1]
    kk = np.arange(1, kmax + 1)[:, None]
    w = np.where(kk <= K[None, :], w, 0.0)
    tot = w.sum(axis=0)
    with np.errstate(all="ignore"):
        w = np.where(K[None, :] > 0, w / tot, 0.0)
    w[0, K == 0] = 1.0                      # K == 0: the reference returns the input unchanged
    return np.ascontiguousarray(w)


class B200Model:
    """Wrapper component: `B200Model(model)(input, params, config; initstates)`."""

    def __init__(self, component: Component, *, fast: bool = True, block_threads: int = 0, stages: int = 0,
                 max_registers: int = 0, nn_tensor: Optional[bool] = None, precision: str = "f64"):
        """`precision="f32"` is the opt-in Float32 mode: forcing and results are float32 arrays, the body
        computes in float (stated tolerance: DESIGN.md §5, `tests/test_gpu_parity.py::test_float32_mode`)."""
        if precision not in ("f64", "f32"):
            raise ValueError("precision must be 'f64' or 'f32'")
        self.component = component
        self.precision = precision
        self.dtype = np.float64 if precision == "f64" else np.float32
        self.fast = fast
        # dense chains on the tensor cores where their shape fits: Float64 -> FP64 MMA (DMMA) fragments per warp; Float32 ->
        # tcgen05.mma kind::tf32 with the CTA's 128 basins as M (templates/hb_tc.cuh).  None = on: both tensor paths measured faster
        # than their per-thread twins (M50 50 k x 3650: Float64 50.9 -> 41 ms round 1; Float32 33.5 -> 32.9 ms, +8 % at >= 4 CTAs/SM)
        self.nn_tensor = True if nn_tensor is None else bool(nn_tensor)
        # [buckets…, HydroRoute] on a dense row-major D8 grid: "wave" = single-pass kernel whose warps exchange
        # outflows through L2 (templates/gridwave.cu, default); "tile" = tile + halo kernel (templates/grid.cu);
        # "generic" = stepper over all T, then one routing launch per step (any node list / graph)
        self.nodes_per_warp = 0           # tensor-path MLP bodies: 0 = default (32), 16 = two lanes per node (stepper.cu HB_NPW)
        self.grid_kernel = "wave"
        self.generic_route_passes = 3        # generic routed path: 3 = stepper / route on compact planes / stepper (default);
                                             #   2 = round 1's stepper + in-place route patches into the records
        self.grid_steps_per_launch = 0       # wave kernel: steps per launch (0 = chosen by the library from the grid width)
        self.grid_panel_min_cols = 3072      # wave kernel: grids at least this wide run as column panels (parallel.PanelledGridRun)
        self.block_threads, self.stages, self.max_registers = block_threads, stages, max_registers
        self._cache: Dict[Tuple, CompiledStepper] = {}
        self.last_kernel: Optional[CompiledStepper] = None
        if isinstance(component, HydroModel):
            comps = component.components
        else:
            comps = (component,)
        self.route: Optional[HydroRoute] = None
        self.lumped: Optional[HydroModel] = None
        n_routes = sum(isinstance(c, HydroRoute) for c in comps)
        if n_routes:
            if n_routes > 1 or not isinstance(comps[-1], HydroRoute):
                raise NotImplementedError("one HydroRoute, as the last component, is supported")
            self.route = comps[-1]
            if len(comps) > 1:
                lumped = HydroModel(components=comps[:-1], name="##lumped_part")
                if set(lumped.input_names) != set(component.input_names):
                    raise NotImplementedError("HydroRoute inputs must be produced by the preceding components")
                # keep the model's own input row order
                self.lumped = HydroModel(components=comps[:-1], name="##lumped_part", input_names=component.input_names)
            self._route_kernel: Optional[CompiledRoute] = None
        # a NeuralFlux has no htypes (its chain is shared by all nodes): stand-alone it takes 2D or 3D input
        # (`/root/reference/src/nn.jl:140-165`), inside a model it follows the other components
        self.any_rank = all(isinstance(c, NeuralFlux) for c in comps)
        self.multinode = all(c.htypes is not None for c in comps if not isinstance(c, NeuralFlux)) and not self.any_rank

    # -- name getters forward to the wrapped component ----------------------------------------------
    @property
    def use_fused_grid(self) -> bool:
        return self.grid_kernel != "generic"

    @use_fused_grid.setter
    def use_fused_grid(self, on: bool):     # round-1 switch: True = the tile + halo kernel, False = the generic path
        self.grid_kernel = "tile" if on else "generic"

    def get_input_names(self): return self.component.get_input_names()
    def get_output_names(self): return self.component.get_output_names()
    def get_state_names(self): return self.component.get_state_names()
    def get_param_names(self): return self.component.get_param_names()
    def get_nn_names(self): return self.component.get_nn_names()

    # -- compile ------------------------------------------------------------------------------------
    def compiled(self, *, rows=None, objective=False, metric=0, shared_forcing=False,
                 uh_kmax: Sequence[int] = (), uh_on_device: bool = False, n_nodes: int = 0,
                 component_min: Optional[Sequence[float]] = None, tangents: Sequence[str] = (),
                 aux_rows: Sequence[str] = ()) -> CompiledStepper:
        # tensor-path MLP bodies: `nodes_per_warp = 16` doubles the warps of a small run (stepper.cu HB_NPW).  Measured on M50
        # 50 k x 3650: 48.6 ms vs 46.1 ms at 32 — the kernel is FP64-pipe bound (math_pipe_throttle), not short of warps, and
        # the duplicated scalar work costs more than the extra warps hide — so 32 stays the default and 16 is opt-in.
        npw = self.nodes_per_warp or 32
        key = (tuple(rows) if rows is not None else None, objective, metric, shared_forcing, tuple(uh_kmax), bool(uh_on_device), npw,
               tuple(component_min) if component_min is not None else None, tuple(tangents), tuple(aux_rows))
        if key not in self._cache:
            prog = lower_stepper(self._stepper_component(), rows=rows, objective=objective,
                                 uh_kmax=list(uh_kmax) if len(uh_kmax) else None, fast=self.fast,
                                 holes=self._route_rows(), nn_tensor=self.nn_tensor,
                                 precision=self.precision, uh_on_device=uh_on_device,
                                 component_min_values=component_min, tangents=list(tangents) or None, aux_rows=list(aux_rows))
            # MLP-fused bodies keep 16+16 activations per net in registers: give them 168 (3 blocks/SM);
            # gradient bodies carry 1 + D doubles per value: all 255
            maxreg = self.max_registers or (255 if tangents else (168 if prog.nn_words else 0))
            bt, st = self.block_threads, self.stages
            if self.route is not None:
                # the stepper passes of a routed model (compact route inputs / records with the route planes as a second
                # input stream, typically short series): 128 threads x 4 stages measured better than the library's default for
                # plain record-writing bodies, 64 x 8 (2048^2 x 60 d: 13.5 vs 14.2 ms)
                bt, st = bt or 128, st or 4
            self._cache[key] = CompiledStepper(prog, metric=metric, shared_forcing=shared_forcing,
                                               block_threads=bt, stages=st,
                                               max_registers=maxreg, nodes_per_warp=npw)
        self.last_kernel = self._cache[key]
        return self._cache[key]

    def _stepper_component(self) -> Component:
        return self.lumped if self.route is not None else self.component

    def _route_rows(self) -> List[str]:
        return [] if self.route is None else self.route.state_names + self.route.output_names

    # -- host argument marshalling ------------------------------------------------------------------
    def _dims(self, input: np.ndarray, shared_forcing: bool, n_members: Optional[int]) -> Tuple[int, int]:
        comp = self.component
        V = len(comp.input_names)
        kind = type(comp).__name__
        if self.any_rank:
            if input.ndim not in (2, 3):
                raise ValueError(f"{kind} accepts 2D (variables × time) or 3D (variables × nodes × time) input, got shape {input.shape}")
            if input.shape[0] != V:
                raise AssertionError("Input variables length mismatch")
            if shared_forcing:
                raise ValueError("ensembles are built from single-node models with parameters")
            return (input.shape[1] if input.ndim == 3 else 1), input.shape[-1]
        if self.multinode:
            if input.ndim != 3:
                raise ValueError(f"{kind} with htypes only accepts 3D input (variables × nodes × time).\n"
                                 f"For single-node computation, omit htypes.\nGot input shape: {input.shape}")
            N = input.shape[1]
        else:
            if input.ndim != 2:
                raise ValueError(f"{kind} without htypes only accepts 2D input (variables × time).\n"
                                 f"For multi-node computation, provide htypes.\nGot input shape: {input.shape}")
            N = 1
        if input.shape[0] != V:
            raise AssertionError("Input variables length mismatch")
        if shared_forcing:
            if N != 1:
                raise ValueError("shared forcing expects a single forcing column")
            N = int(n_members)
        return N, input.shape[-1]

    def prepare(self, prog: StepperProgram, params, N: int, *, ensemble: bool = False,
                node_range: Optional[Tuple[int, int]] = None):
        """Flat parameter table + modes + htypes + UH ordinates + MLP parameters.  `node_range=(n0, n1)`:
        this process owns nodes `[n0, n1)` of a model described globally (multi-GPU shards): htypes are
        sliced, per-node parameter vectors are sliced, shared (gathered) vectors are kept whole."""
        params = _as_componentvector(params)
        n0, n1 = node_range if node_range is not None else (0, N)
        if node_range is not None and n1 - n0 != N:
            raise ValueError("node_range does not match the local node count")
        pdict = params["params"]
        names: List[str] = []
        for nm, _ in prog.param_slots:
            if nm not in names:
                names.append(nm)
        for u in prog.uh:
            for nm in u.param_names:
                if nm not in names:
                    names.append(nm)
        vecs: Dict[str, np.ndarray] = {}
        for nm in names:
            if nm not in pdict:
                raise KeyError(f"Parameter {nm} does not exist in params")
            vecs[nm] = np.atleast_1d(np.asarray(pdict[nm], dtype=np.float64)).reshape(-1)
        sets0, ident = [], []
        for h in prog.htype_sets:
            if node_range is None and len(h) != N:
                raise ValueError(f"htypes has {len(h)} entries but the input has {N} nodes")
            if node_range is not None and len(h) < n1:
                raise ValueError("node_range exceeds htypes")
            h0 = np.asarray(h, dtype=np.int64) - 1
            ident.append(bool(np.array_equal(h0[n0:n1], np.arange(n0, n1))))
            sets0.append(h0[n0:n1] if node_range is not None else h0)
        n_global = max([len(h) for h in prog.htype_sets] or [N])

        def mode_of(nm: str, hs: Optional[int]) -> int:
            v = vecs[nm]
            if ensemble:                       # one parameter set per ensemble member
                if v.size == N:
                    return 1
                if v.size == 1:
                    return 0
                raise ValueError(f"ensemble parameter {nm} must have 1 or {N} values")
            if hs is None:
                if v.size != 1:
                    raise ValueError(f"single-node component expects scalar parameter {nm}")
                return 0
            h0 = sets0[hs]
            if h0.size and (h0.min() < 0 or h0.max() >= v.size):
                raise IndexError(f"BoundsError: attempt to access {v.size}-element parameter {nm} at index htypes")
            if v.size == n_global and ident[hs]:
                return 1
            return 2 + hs
        modes = [mode_of(nm, hs) for nm, hs in prog.param_slots]
        # a name used per-node by every slot is uploaded as its local slice; otherwise whole
        per_node_only = {nm: all(m == 1 for (n2, _), m in zip(prog.param_slots, modes) if n2 == nm) for nm in names}
        offs, chunks, pos = {}, [], 0
        for nm in names:
            v = vecs[nm]
            if node_range is not None and per_node_only.get(nm, False) and v.size == n_global and any(n2 == nm for n2, _ in prog.param_slots):
                v = v[n0:n1]
            offs[nm] = pos
            chunks.append(v)
            pos += v.size
        flat = np.ascontiguousarray(np.concatenate(chunks)) if chunks else np.zeros(1)
        p_off = np.array([offs[nm] for nm, _ in prog.param_slots] or [0], dtype=np.int32)
        p_mode = np.array(modes or [0], dtype=np.int32)
        htypes = np.ascontiguousarray(np.stack(sets0).astype(np.int32)) if sets0 else None
        # unit hydrographs: ordinates per node (host-evaluated N×K table)
        uhw = None
        kmax: List[int] = []
        if prog.uh:
            tabs = []
            for u, hs in zip(prog.uh, prog.uh_htype_set):
                pv = {nm: vecs[nm] for nm in u.param_names}
                if ensemble or hs is None:
                    n = max([v.size for v in pv.values()] or [1])
                    pv = {k: np.broadcast_to(v, (n,)) for k, v in pv.items()}
                    w = _uh_weight_table(u, pv)
                    w = np.broadcast_to(w, (w.shape[0], N)) if w.shape[1] == 1 else w
                else:
                    h0 = sets0[hs]
                    for nm, v in pv.items():
                        if h0.size and h0.max() >= v.size:
                            raise IndexError(f"BoundsError: parameter {nm} indexed by htypes")
                    w = _uh_weight_table(u, pv)[:, h0]
                tabs.append(np.ascontiguousarray(w))
                kmax.append(w.shape[0])
            uhw = tabs
        nn = None
        if prog.nn_layout:
            nns = params.get("nns", {})
            parts = []
            for name, off, n in prog.nn_layout:
                if name not in nns:
                    raise KeyError(f"Neural network parameters {name} do not exist in params[:nns]")
                v = np.asarray(nns[name], dtype=np.float64).reshape(-1)
                if v.size != n:
                    raise ValueError(f"chain {name} expects {n} parameters, got {v.size}")
                parts.append(v)
            nn = np.ascontiguousarray(np.concatenate(parts))
        return flat, p_off, p_mode, htypes, uhw, kmax, nn

    def _initstates(self, prog: StepperProgram, initstates, N: int,
                    node_range: Optional[Tuple[int, int]] = None) -> Optional[np.ndarray]:
        S = len(prog.state_names)
        if S == 0 or initstates is None:
            return None                       # zeros (`model._defaultstates`, `src/model.jl:51-52`)
        if isinstance(initstates, dict):
            out = np.empty((S, N), dtype=np.float64)
            for i, s in enumerate(prog.state_names):
                if s not in initstates:
                    raise KeyError(f"State {s} not found in initstates")
                v = np.asarray(initstates[s], dtype=np.float64)
                if node_range is not None and v.ndim == 1 and v.size != N:
                    v = v[node_range[0]:node_range[1]]
                out[i] = np.broadcast_to(v, (N,))
            return out
        if node_range is not None:
            raise NotImplementedError("sharded runs take initstates as a dict of per-state vectors")
        v = np.asarray(initstates, dtype=np.float64).reshape(-1)
        if v.size != S * N:
            raise ValueError(f"initstates has {v.size} elements, expected {S}*{N} (state-major, node-fastest)")
        return np.ascontiguousarray(v.reshape(S, N))

    # -- the functor ----------------------------------------------------------------------------------
    def __call__(self, input, params, config=None, *, initstates=None, rows: Optional[Sequence[str]] = None,
                 return_final_states: bool = False, out: Optional[np.ndarray] = None, timing: Optional[dict] = None):
        # per-component configs (`/root/reference/src/model.jl:140-152`): a tuple, or {"components": tuple}
        if isinstance(config, dict) and "components" in config:
            if not isinstance(config["components"], (tuple, list)):
                raise ValueError("components field in config must be a Tuple")
            config = tuple(config["components"])
        component_min = None
        if isinstance(config, (tuple, list)):
            ncomp = len(self.component.components) if isinstance(self.component, HydroModel) else 1
            if len(config) != ncomp:
                raise ValueError("Component configs length must equal components length")
            cfgs = [normalize_config(c) for c in config]
            cfg = cfgs[0]
            for c in cfgs[1:]:
                if c.solver == SolverType.ODESolver or c.interpolator not in (ConstantInterpolation, LinearInterpolation) \
                        or list(c.timeidx) != list(cfg.timeidx):
                    _check_config(c, len(cfg.timeidx))
                    raise NotImplementedError("per-component timeidx must agree across components")
            if any(c.min_value != cfg.min_value for c in cfgs):
                if self.route is not None:
                    raise NotImplementedError("per-component min_value is not supported for models with a HydroRoute")
                component_min = tuple(float(c.min_value) for c in cfgs)     # baked into the body as literals
        else:
            cfg = normalize_config(config)
        input = np.asarray(input)
        if input.dtype != self.dtype:
            input = input.astype(self.dtype)
        if self.route is not None:
            if self.precision != "f64":
                raise NotImplementedError("Float32 mode is not available for models with a HydroRoute")
            if rows is not None or out is not None:
                raise NotImplementedError("rows= / out= are not supported for models with a HydroRoute")
            if input.ndim != 3:
                raise ValueError("HydroRoute only accepts 3D input (variables x nodes x time).\n"
                                 "Routing is inherently spatial and requires multiple nodes.\n"
                                 f"Got input shape: {input.shape}")
            return self._call_routed(input, params, cfg, initstates, return_final_states, timing)
        N, T = self._dims(input, False, None)
        _check_config(cfg, T)
        inp = input if input.flags.f_contiguous else np.asfortranarray(input)   # memory = [T][N][V]
        # first pass over the parameters to size the unit hydrographs (compile-time K_max)
        prog0 = self.compiled(rows=rows, n_nodes=N, component_min=component_min).program if not self._has_uh() else None
        if prog0 is None:
            probe = lower_stepper(self._stepper_component(), rows=rows, fast=self.fast)
            _, _, _, _, _, kmax, _ = self.prepare(probe, params, N)
            ck = self.compiled(rows=rows, uh_kmax=kmax, n_nodes=N, component_min=component_min)
        else:
            ck = self.compiled(rows=rows, n_nodes=N, component_min=component_min)
        prog = ck.program
        flat, p_off, p_mode, htypes, uhw, kmax, nn = self.prepare(prog, params, N)
        uhw_flat = np.ascontiguousarray(np.concatenate(uhw, axis=0)) if uhw else None
        init = self._initstates(prog, initstates, N)
        R = len(prog.row_names)
        shape = (R, N, T) if (self.multinode or (self.any_rank and input.ndim == 3)) else (R, T)
        if out is None:
            out = np.empty(shape, dtype=self.dtype, order="F")                  # memory = [T][N][R]
        elif out.shape != shape or not out.flags.f_contiguous or out.dtype != self.dtype:
            raise ValueError(f"out must be a Fortran-ordered {np.dtype(self.dtype).name} array of the result shape")
        S = len(prog.state_names)
        final = np.empty((S, N), dtype=np.float64) if (return_final_states and S) else None
        elapsed = C.c_float(0.0)
        d = RunDesc()
        d.struct_size = C.sizeof(RunDesc)
        d.n_htype_sets = 0 if htypes is None else htypes.shape[0]
        d.n_nodes, d.n_steps = N, T
        d.input = _ptr(inp)
        d.params, d.n_param_values = _ptr(flat), flat.size
        d.param_offset, d.param_mode = _ptr(p_off), _ptr(p_mode)
        d.htypes = _ptr(htypes)
        d.initstates = _ptr(init)
        d.uh_weights = _ptr(uhw_flat)
        d.nn_params = _ptr(nn)
        d.min_value = cfg.min_value
        d.warm_up = 1
        d.out = _ptr(out)
        d.final_states = _ptr(final)
        d.elapsed_ms = C.cast(C.byref(elapsed), C.c_void_p)
        check(lib().hb_run(ck.handle, C.byref(d)))
        if timing is not None:
            timing["kernel_ms"] = elapsed.value
        if return_final_states:
            return out, final
        return out

    # -- models ending in a HydroRoute: per-node stepper, then one routing launch per time step ---------
    def _call_routed(self, input, params, cfg, initstates, return_final_states, timing):
        params = _as_componentvector(params)
        comp, route = self.component, self.route
        all_rows = comp.state_names + comp.output_names if not isinstance(comp, HydroModel) else comp.var_names
        V, N, T = input.shape
        if V != len(comp.input_names):
            raise AssertionError("Input variables length mismatch")
        if len(route.htypes) != N:
            raise ValueError(f"htypes has {len(route.htypes)} entries but the input has {N} nodes")
        _check_config(cfg, T)
        R = len(all_rows)
        # route inputs that are model inputs (stand-alone HydroRoute) ride along as extra record rows
        extra = [nm for nm in route.input_names if nm not in all_rows]
        rec_rows = all_rows + extra
        Rr = len(rec_rows)
        s_name, o_name = route.state_names[0], route.output_names[0]
        state_names = comp.state_names
        init_all = None
        if initstates is not None:
            if isinstance(initstates, dict):
                init_all = {k: np.broadcast_to(np.asarray(v, dtype=np.float64), (N,)) for k, v in initstates.items()}
            else:
                v = np.asarray(initstates, dtype=np.float64).reshape(-1)
                if v.size != len(state_names) * N:
                    raise ValueError(f"initstates has {v.size} elements, expected {len(state_names)}*{N}")
                init_all = {s: v[i * N:(i + 1) * N] for i, s in enumerate(state_names)}
        # dense row-major grid, bucket/flux components only: fused single-pass kernel (templates/grid.cu)
        if (self.lumped is not None and not extra and self.use_fused_grid and dense_grid_shape(route.aggr, N) is not None
                and not self._has_uh() and not self.lumped.nn_names and N * T > 0):
            shape = dense_grid_shape(route.aggr, N)
            if self.grid_kernel == "wave" and shape[1] >= self.grid_panel_min_cols:
                # a wide grid runs faster as column panels, one after the other, in place on the whole grid's arrays
                from .parallel import PanelledGridRun
                run = PanelledGridRun(self, shape[0], shape[1], T, params, config=cfg, initstates=init_all,
                                      steps_per_launch=self.grid_steps_per_launch, block_threads=self.block_threads,
                                      stages=self.stages, max_registers=self.max_registers)
                self._cache[("grid", "wave", "panels", run.K)] = run.ck
            elif self.grid_kernel == "wave":
                run = GridDeviceRun(self, N, T, params, config=cfg, initstates=init_all, variant="wave",
                                    steps_per_chunk=self.grid_steps_per_launch, block_w=self.block_threads, stages=self.stages,
                                    max_registers=self.max_registers)
            else:
                run = GridDeviceRun(self, N, T, params, config=cfg, initstates=init_all)
            inp = input if input.flags.f_contiguous else np.asfortranarray(input)
            din = DeviceBuffer.from_host(inp.ravel(order="K"))
            rec = DeviceBuffer(T * N * R * 8)
            try:
                if isinstance(run, GridDeviceRun):
                    ms = run.launch(din.ptr, rec.ptr, None, timed=True)
                else:
                    run.launch(din.ptr, rec.ptr, None)
                    ms = None
                run.check()
                full = rec.to_host((R, N, T), order="F")
            except HydroB200Error as ex:
                # a grid wider than the co-resident blocks can cover keeps the wave kernel's forward-progress guarantee
                # from holding: the library refuses it and the generic two-pass path takes over.  The same happens when a
                # warp's bounded poll expired (a co-tenant kernel held SMs the dependency cone needed): the records are
                # invalid, the generic path recomputes them
                if "too wide for the wave kernel" not in str(ex) and "gave up waiting" not in str(ex):
                    raise
                self.wave_fallbacks = getattr(self, "wave_fallbacks", 0) + 1
                full = None
            din.free()
            rec.free()
            if full is not None:
                if timing is not None:
                    timing["kernel_ms"] = ms
                if return_final_states:
                    return full, np.ascontiguousarray(full[:len(state_names), :, -1])
                return full
        if self.lumped is not None and not extra and self.generic_route_passes == 3:
            # generic path for any node list / graph: stepper (route inputs only) -> route on compact planes -> stepper
            # (all rows, route rows from the auxiliary stream): every record written once (RoutedDeviceRun)
            inp = input if input.flags.f_contiguous else np.asfortranarray(input)
            run3 = RoutedDeviceRun(self, N, T, params, config=cfg, initstates=init_all)
            din = DeviceBuffer.from_host(inp.ravel(order="K"))
            rec = DeviceBuffer(T * N * R * 8)
            ms = run3.launch(din.ptr, rec.ptr, None, timed=True)
            full = rec.to_host((R, N, T), order="F")
            for bfr in (din, rec):
                bfr.free()
            run3.free()
            if timing is not None:
                timing["kernel_ms"] = ms
            if return_final_states:
                return full, np.ascontiguousarray(full[:len(state_names), :, -1])
            return full
        rec = DeviceBuffer(T * N * Rr * 8)
        kernel_ms = 0.0
        if self.lumped is not None:
            inp = input if input.flags.f_contiguous else np.asfortranarray(input)
            if self._has_uh():
                probe = lower_stepper(self.lumped, rows=rec_rows, fast=self.fast, holes=self._route_rows())
                kmax = self.prepare(probe, params, N)[5]
            else:
                kmax = []
            ck = self.compiled(rows=rec_rows, uh_kmax=kmax)
            run = DeviceRun(self, N, T, params, config=cfg,
                            initstates=None if init_all is None else {k: v for k, v in init_all.items() if k != s_name},
                            rows=rec_rows, _kmax=kmax)
            din = DeviceBuffer.from_host(inp.ravel(order="K"))
            kernel_ms += run.launch(din.ptr, rec.ptr, None, timed=True)
            din.free()
        else:
            host = np.zeros((T, N, Rr))
            for nm in extra:
                host[:, :, rec_rows.index(nm)] = input[comp.input_names.index(nm)].T
            check(lib().hb_memcpy_h2d(rec.ptr, host.ctypes.data_as(C.c_void_p), host.nbytes, None))
            check(lib().hb_stream_synchronize(None))
        # ---- route ----
        stage = RouteStage(self, N, T, params, cfg, rec_rows,
                           None if init_all is None or s_name not in init_all else init_all[s_name])
        kernel_ms += stage.launch(rec.ptr, None, timed=True)
        stage.free()
        full = rec.to_host((Rr, N, T), order="F")
        rec.free()
        if timing is not None:
            timing["kernel_ms"] = kernel_ms
        out = full if not extra else np.asfortranarray(full[:R])
        if return_final_states:
            return out, np.ascontiguousarray(out[:len(state_names), :, -1])
        return out

    def _has_uh(self) -> bool:
        c0 = self._stepper_component()
        if c0 is None:
            return False
        comps = c0.components if isinstance(c0, HydroModel) else (c0,)
        return any(isinstance(c, UnitHydrograph) for c in comps)

    # -- calibration ensembles: fused objective --------------------------------------------------------
    def objective(self, input, params, target, *, metric: str = "NSE", warm_up: int = 1, config=None,
                  initstates=None, n_members: Optional[int] = None, timing: Optional[dict] = None,
                  _tangents: Sequence[str] = ()) -> np.ndarray:
        """`loss(target[warm_up:end], model(input, p_i, config; initstates)[end, warm_up:end])` for
        every parameter set / node `i` in one launch
        (`/root/reference/ext/HydroModelsOptimizationExt/HydroModelsOptimizationExt.jl:153-175`).

        * single-node model + `n_members`: a calibration ensemble — `input` is `[V, T]`, shared by all
          members, every parameter is a scalar or a vector of `n_members` values;
        * multi-node model: one objective per node, `target` is `[T]` (shared) or `[N, T]`."""
        m = metric.lower()
        if m not in METRICS:
            raise ValueError(f"Unknown metric: {metric}. Supported metrics: KGE, NSE, LogKGE, MSE, SSE")
        if self.route is not None:
            raise NotImplementedError("fused objectives are for per-node models (no HydroRoute)")
        cfg = normalize_config(config)
        input = np.asarray(input, dtype=self.dtype)
        ensemble = n_members is not None
        if ensemble and self.multinode:
            raise ValueError("ensembles are built from single-node models")
        N, T = self._dims(input, ensemble, n_members)
        _check_config(cfg, T)
        inp = input if input.flags.f_contiguous else np.asfortranarray(input)
        if self._has_uh():
            probe = lower_stepper(self.component, objective=True, fast=self.fast)
            kmax = self.prepare(probe, params, N, ensemble=ensemble)[5]
        else:
            kmax = []
        ck = self.compiled(objective=True, metric=METRICS[m], shared_forcing=ensemble, uh_kmax=kmax,
                           uh_on_device=bool(_tangents) and bool(kmax), tangents=_tangents)
        prog = ck.program
        flat, p_off, p_mode, htypes, uhw, kmax, nn = self.prepare(prog, params, N, ensemble=ensemble)
        uhw_flat = np.ascontiguousarray(np.concatenate(uhw, axis=0)) if uhw else None
        init = self._initstates(prog, initstates, N)
        tgt = np.zeros(T) if (target is None and m == "sum") else np.asarray(target, dtype=np.float64)
        per_node = 0
        if tgt.ndim == 2:
            if tgt.shape != (N, T):
                raise ValueError("target must be [T] or [N, T]")
            tgt = np.ascontiguousarray(tgt.T)
            per_node = 1
        elif tgt.shape != (T,):
            raise ValueError("target must be [T] or [N, T]")
        tgt = np.ascontiguousarray(tgt)
        obj = np.empty(N, dtype=np.float64)
        elapsed = C.c_float(0.0)
        d = RunDesc()
        d.struct_size = C.sizeof(RunDesc)
        d.n_htype_sets = 0 if htypes is None else htypes.shape[0]
        d.n_nodes, d.n_steps = N, T
        d.input = _ptr(inp)
        d.params, d.n_param_values = _ptr(flat), flat.size
        d.param_offset, d.param_mode = _ptr(p_off), _ptr(p_mode)
        d.htypes = _ptr(htypes)
        d.initstates = _ptr(init)
        d.uh_weights = _ptr(uhw_flat)
        d.nn_params = _ptr(nn)
        d.target, d.target_per_node, d.warm_up = _ptr(tgt), per_node, int(warm_up)
        d.min_value = cfg.min_value
        d.objective = _ptr(obj)
        grad = np.empty((len(_tangents), N), dtype=np.float64) if _tangents else None
        d.gradient = _ptr(grad)
        d.elapsed_ms = C.cast(C.byref(elapsed), C.c_void_p)
        check(lib().hb_run(ck.handle, C.byref(d)))
        if timing is not None:
            timing["kernel_ms"] = timing.get("kernel_ms", 0.0) + elapsed.value if _tangents else elapsed.value
            timing["kernel"] = ck
        if _tangents:
            return obj, grad, prog, p_mode, htypes
        return obj

    # -- gradients: forward-mode tangents through the same fused stepper ------------------------------
    def gradient(self, input, params, target, *, metric: str = "SSE", warm_up: int = 1, config=None, initstates=None,
                 wrt: Optional[Sequence[str]] = None, n_members: Optional[int] = None, max_directions: int = 6,
                 timing: Optional[dict] = None, mode: Optional[str] = None):
        """`(loss, grads)`: the objective of `objective()` and its derivative w.r.t. the physical parameters — what the
        reference obtains with `Zygote.gradient` through its ImmutableSolver (`/root/reference/README.md:161-187`; loss
        `sum((sim - obs).^2)` = metric "SSE").  Forward mode: the stepper body is re-lowered with value + D tangents per
        quantity (`hb_dual`, hb_math.cuh) and ONE launch returns the objective and D derivative rows per node; more than
        `max_directions` parameters take several launches.  `wrt` may also name STATES: the derivative w.r.t. that state's
        initial value (one per node).  `grads[name]` has the shape of `params["params"][name]`
        (scalar, `[ntypes]`, or `[n_members]` for ensembles) and is the derivative of `sum(loss)`: node types shared by
        several nodes sum their nodes' derivatives.  `loss` is the per-node / per-member objective vector.
        Models with neural components are not covered (hundreds of weights need the reverse-mode adjoint)."""
        if mode is None:                    # models with neural fluxes: hundreds of weights -> the reverse-mode adjoint
            mode = "reverse" if self.component.nn_names else "forward"
        if mode == "reverse":
            return self.gradient_reverse(input, params, target, metric=metric, warm_up=warm_up, config=config, initstates=initstates,
                                         timing=timing)
        if mode != "forward":
            raise ValueError("mode must be 'forward' or 'reverse'")
        pdict = _as_componentvector(params)["params"]
        names = list(wrt) if wrt is not None else [nm for nm in self.component.param_names]
        if not names:
            raise ValueError("no parameters to differentiate")
        D = max(1, min(int(max_directions), HB_MAX_TANGENTS))
        loss, grads = None, {}
        for g0 in range(0, len(names), D):
            group = tuple(names[g0:g0 + D])
            obj, g, prog, p_mode, htypes = self.objective(input, params, target, metric=metric, warm_up=warm_up, config=config,
                                                          initstates=initstates, n_members=n_members, timing=timing,
                                                          _tangents=group)
            loss = obj if loss is None else loss
            for k, nm in enumerate(group):
                if nm in prog.state_names and nm not in pdict:          # initial value of a state: one per node
                    grads[nm] = g[k].copy() if (self.multinode or n_members is not None) else g[k][0]
                    continue
                modes = {int(p_mode[j]) for j, (pn, _) in enumerate(prog.param_slots) if pn == nm}
                if len(modes) != 1:
                    raise NotImplementedError(f"parameter {nm} is used under different htypes vectors")
                mode = modes.pop()
                if mode == 0:                                   # one value for every node
                    val = g[k].sum()
                    grads[nm] = val if np.ndim(pdict[nm]) == 0 else np.full(np.shape(pdict[nm]), val)
                elif mode == 1:                                 # one value per node / member
                    grads[nm] = g[k].copy()
                else:                                           # gathered through 0-based htypes set (mode - 2)
                    out = np.zeros(np.size(pdict[nm]), dtype=np.float64)
                    np.add.at(out, htypes[mode - 2], g[k])
                    grads[nm] = out
        return loss, grads


    # -- gradients: the reverse-mode adjoint of the stepper (neural-network weights) -------------------
    def gradient_reverse(self, input, params, target, *, metric: str = "SSE", warm_up: int = 1, config=None, initstates=None,
                         timing: Optional[dict] = None):
        """`(loss, grads)` like `gradient`, by ONE backward sweep over the stored state trajectory (templates/adjoint.cu) —
        the reverse-mode counterpart of the reference's `Zygote.gradient` through its ImmutableSolver
        (`/root/reference/README.md:161-187`, `docs/src/tutorials/neuralnetwork_embeding.md`): what training the neural fluxes
        of a hybrid model (M50: 690 weights) needs.  `grads` holds every physical parameter (shape of `params["params"][name]`),
        every state's initial value (per node) and `grads["nns"][chain]` (flat, the layout of `params["nns"][chain]`, summed
        over all nodes).  Metrics that are linear in the squared errors: SSE, MSE, NSE — and SUM, the plain sum of the simulated
        row from `warm_up` on (`target` may be None), which is what the reference's own gradient tests differentiate
        (`test/gradient/run_hybrid_gradient.jl:39-44`: `output[end, :] |> sum`).  Single-node (2D input) and multi-node models."""
        m = metric.lower()
        if m not in ("sse", "mse", "nse", "sum"):
            raise NotImplementedError("the adjoint run covers SSE / MSE / NSE (metrics linear in the squared errors) and SUM")
        if self.route is not None or self.precision != "f64":
            raise NotImplementedError("the adjoint run is for per-node models in Float64")
        cfg = normalize_config(config)
        input = np.asarray(input, dtype=np.float64)
        N, T = self._dims(input, False, None)
        _check_config(cfg, T)
        key = ("adjoint",)
        if key not in self._cache:
            self._cache[key] = CompiledAdjoint(lower_stepper(self.component, objective=True, fast=True, adjoint=True), self.block_threads)
        ck = self._cache[key]
        prog = ck.program
        flat, p_off, p_mode, htypes, _, _, nn = self.prepare(prog, params, N)
        pdict = _as_componentvector(params)["params"]
        init = self._initstates(prog, initstates, N)
        S, NP = len(prog.state_names), len(prog.param_slots)
        tgt = np.zeros(T) if (target is None and m == "sum") else np.asarray(target, dtype=np.float64)
        per_node = 0
        if tgt.ndim == 2:
            if tgt.shape != (N, T):
                raise ValueError("target must be [T] or [N, T]")
            tgt2 = tgt
            tgt, per_node = np.ascontiguousarray(tgt.T), 1
        elif tgt.shape != (T,):
            raise ValueError("target must be [T] or [N, T]")
        else:
            tgt2 = np.broadcast_to(tgt, (N, T))
        w = tgt2[:, warm_up - 1:]
        if m in ("sse", "sum"):
            scale = np.ones(N)
        elif m == "mse":
            scale = np.full(N, 1.0 / w.shape[1])
        else:
            scale = 1.0 / np.sum((w - w.mean(axis=1, keepdims=True)) ** 2, axis=1)
        inp = input if input.flags.f_contiguous else np.asfortranarray(input)
        # forward: the state trajectory [T][N][S]
        fwd = DeviceRun(self, N, T, params, config=cfg, initstates=initstates, rows=list(prog.state_names))
        bufs = {"in": DeviceBuffer.from_host(inp.ravel(order="K")), "traj": DeviceBuffer(max(1, T * N * S) * 8),
                "params": DeviceBuffer.from_host(flat), "tgt": DeviceBuffer.from_host(np.ascontiguousarray(tgt)),
                "scale": DeviceBuffer.from_host(np.ascontiguousarray(scale)), "gp": DeviceBuffer(max(1, NP * N) * 8),
                "gi": DeviceBuffer(max(1, S * N) * 8), "gn": DeviceBuffer(max(1, prog.nn_words) * 8), "loss": DeviceBuffer(N * 8)}
        if htypes is not None:
            bufs["ht"] = DeviceBuffer.from_host(htypes)
        if init is not None:
            bufs["init"] = DeviceBuffer.from_host(init)
        if nn is not None:
            bufs["nn"] = DeviceBuffer.from_host(nn)
        ms_f = fwd.launch(bufs["in"].ptr, bufs["traj"].ptr, None, timed=True) if S else 0.0
        d = AdjointDesc()
        d.struct_size = C.sizeof(AdjointDesc)
        d.n_htype_sets = 0 if htypes is None else htypes.shape[0]
        d.n_nodes, d.n_steps = N, T
        d.input, d.trajectory = bufs["in"].ptr, bufs["traj"].ptr
        d.params, d.n_param_values = bufs["params"].ptr, flat.size
        d.param_offset, d.param_mode = _ptr(p_off), _ptr(p_mode)
        d.htypes = bufs["ht"].ptr if "ht" in bufs else None
        d.initstates = bufs["init"].ptr if "init" in bufs else None
        d.nn_params = bufs["nn"].ptr if "nn" in bufs else None
        d.target, d.target_per_node, d.warm_up = (None if m == "sum" else bufs["tgt"].ptr), per_node, int(warm_up)
        d.min_value = cfg.min_value
        d.scale, d.grad_params, d.grad_init = bufs["scale"].ptr, bufs["gp"].ptr, bufs["gi"].ptr
        d.grad_nn, d.loss = bufs["gn"].ptr, bufs["loss"].ptr
        el = C.c_float(0.0)
        d.elapsed_ms = C.cast(C.byref(el), C.c_void_p)
        check(lib().hb_run_adjoint_device(ck.handle, C.byref(d), None))
        gp = bufs["gp"].to_host((max(1, NP), N))
        gi = bufs["gi"].to_host((max(1, S), N))
        gn = bufs["gn"].to_host((max(1, prog.nn_words),))
        loss = bufs["loss"].to_host((N,))
        for b_ in bufs.values():
            b_.free()
        if m == "nse":
            loss = loss - 1.0                                   # -NSE = -1 + SSE / var(obs): what `objective(metric="NSE")` returns
        if timing is not None:
            timing.update(kernel_ms=el.value, forward_ms=ms_f, kernel=ck)
        grads: Dict[str, object] = {}
        for nm in dict.fromkeys(pn for pn, _ in prog.param_slots):
            out = np.zeros(np.size(pdict[nm]), dtype=np.float64)
            for j, (pn, hs) in enumerate(prog.param_slots):
                if pn != nm:
                    continue
                mode_j = int(p_mode[j])
                if mode_j == 0:
                    out[0] += gp[j].sum()
                elif mode_j == 1:
                    out += gp[j]
                else:
                    np.add.at(out, htypes[mode_j - 2], gp[j])
            grads[nm] = out if np.ndim(pdict[nm]) else float(out[0])
        for si, sn in enumerate(prog.state_names):
            grads[sn] = gi[si].copy()
        if prog.nn_layout:
            grads["nns"] = {name: gn[off:off + n].copy() for name, off, n in prog.nn_layout}
        return loss, grads


def run_component(component: Component, input, params, config=None, *, initstates=None, **kwargs):
    """`component(input, params, config; initstates)` — what `Component.__call__` dispatches to."""
    eng = getattr(component, "_b200_engine", None)
    if eng is None:
        eng = B200Model(component)
        component._b200_engine = eng
    return eng(input, params, config, initstates=initstates, **kwargs)


# ---------------------------------------------------------------------------------------------------
# device-resident runs (benchmarks, multi-GPU shards, chained calls): hb_run_device
# ---------------------------------------------------------------------------------------------------
class DeviceBuffer:
    """A device allocation owned through the C-ABI (no torch needed)."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(lib().hb_malloc_device(C.byref(p), max(self.nbytes, 8)))
        self.ptr = p.value

    @classmethod
    def from_host(cls, a: np.ndarray, stream=None) -> "DeviceBuffer":
        a = np.ascontiguousarray(a)
        b = cls(a.nbytes)
        check(lib().hb_memcpy_h2d(b.ptr, a.ctypes.data_as(C.c_void_p), a.nbytes, stream))
        check(lib().hb_stream_synchronize(stream))
        return b

    def to_host(self, shape, dtype=np.float64, order="C", stream=None) -> np.ndarray:
        out = np.empty(shape, dtype=dtype, order=order)
        check(lib().hb_memcpy_d2h(out.ctypes.data_as(C.c_void_p), self.ptr, out.nbytes, stream))
        check(lib().hb_stream_synchronize(stream))
        return out

    def free(self):
        if getattr(self, "ptr", None):
            lib().hb_free_device(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceRun:
    """Everything but forcing and results staged on the device once; `launch()` is then a pure
    kernel launch on a caller-supplied stream (the unit `bench.py` times as `value`)."""

    def __init__(self, eng: B200Model, N: int, T: int, params, *, config=None, initstates=None,
                 rows: Optional[Sequence[str]] = None, objective: Optional[str] = None, target=None,
                 warm_up: int = 1, ensemble: bool = False, _kmax=None, node_range: Optional[Tuple[int, int]] = None,
                 uh_on_device: bool = False, aux_rows: Sequence[str] = ()):
        """`uh_on_device`: unit-hydrograph ordinates are evaluated in the kernel prologue from the parameters in the device
        buffer (`_kmax` then gives the ordinates to reserve per unit hydrograph) — the host table is only a placeholder."""
        self.eng, self.N, self.T = eng, int(N), int(T)
        self.cfg = normalize_config(config)
        _check_config(self.cfg, T)
        obj = objective is not None
        metric = METRICS[objective.lower()] if obj else 0
        if _kmax is not None:
            kmax = _kmax
        elif eng._has_uh():
            probe = lower_stepper(eng._stepper_component(), rows=rows, objective=obj, fast=eng.fast,
                                  holes=eng._route_rows())
            kmax = eng.prepare(probe, params, N, ensemble=ensemble, node_range=node_range)[5]
        else:
            kmax = []
        self.ck = eng.compiled(rows=rows, objective=obj, metric=metric, shared_forcing=ensemble, uh_kmax=kmax,
                               uh_on_device=uh_on_device, n_nodes=N, aux_rows=aux_rows)
        prog = self.ck.program
        flat, p_off, p_mode, htypes, uhw, kmax_host, nn = eng.prepare(prog, params, N, ensemble=ensemble, node_range=node_range)
        if uh_on_device and uhw:
            uhw = [np.zeros((k, N)) for k in kmax]            # placeholder: the prologue overwrites every ordinate
        self._host = (p_off, p_mode)
        self.bufs = {"params": DeviceBuffer.from_host(flat)}
        self.n_param_values = flat.size
        self.n_sets = 0 if htypes is None else htypes.shape[0]
        if htypes is not None:
            self.bufs["htypes"] = DeviceBuffer.from_host(htypes)
        init = eng._initstates(prog, initstates, N, node_range)
        if init is not None:
            self.bufs["init"] = DeviceBuffer.from_host(init)
        if uhw:
            self.bufs["uhw"] = DeviceBuffer.from_host(np.concatenate(uhw, axis=0))
        if nn is not None:
            self.bufs["nn"] = DeviceBuffer.from_host(nn)
        self.per_node_target = 0
        if obj:
            tgt = np.asarray(target, dtype=np.float64)
            if tgt.ndim == 2:
                tgt, self.per_node_target = np.ascontiguousarray(tgt.T), 1
            self.bufs["target"] = DeviceBuffer.from_host(tgt)
        self.warm_up = int(warm_up)
        self.n_rows = 0 if obj else len(prog.row_names)
        self.n_inputs = len(prog.input_names)
        self.objective = obj

    @property
    def in_bytes(self) -> int:
        return self.T * (1 if self.ck.spec.shared_forcing else self.N) * self.n_inputs * (4 if self.ck.spec.precision else 8)

    @property
    def out_bytes(self) -> int:
        return self.N * 8 if self.objective else self.T * self.N * self.n_rows * (4 if self.ck.spec.precision else 8)

    def launch(self, input_ptr: int, out_ptr: int, stream: Optional[int] = None, final_states_ptr: Optional[int] = None,
               timed: bool = False, aux_ptr: Optional[int] = None, compact_rows: Sequence[int] = ()) -> Optional[float]:
        """`compact_rows`: store only these record rows, `out_ptr` is then `[T][N][len(compact_rows)]` (same kernel)."""
        d = RunDesc()
        d.aux_input = aux_ptr
        d.n_compact_rows = len(compact_rows)
        for k, r in enumerate(compact_rows):
            d.compact_rows[k] = int(r)
        d.struct_size = C.sizeof(RunDesc)
        d.n_htype_sets = self.n_sets
        d.n_nodes, d.n_steps = self.N, self.T
        d.input = input_ptr
        d.params, d.n_param_values = self.bufs["params"].ptr, self.n_param_values
        d.param_offset, d.param_mode = _ptr(self._host[0]), _ptr(self._host[1])
        d.htypes = self.bufs["htypes"].ptr if "htypes" in self.bufs else None
        d.initstates = self.bufs["init"].ptr if "init" in self.bufs else None
        d.uh_weights = self.bufs["uhw"].ptr if "uhw" in self.bufs else None
        d.nn_params = self.bufs["nn"].ptr if "nn" in self.bufs else None
        d.min_value = self.cfg.min_value
        d.warm_up = self.warm_up
        if self.objective:
            d.target, d.target_per_node = self.bufs["target"].ptr, self.per_node_target
            d.objective = out_ptr
        else:
            d.out = out_ptr
        d.final_states = final_states_ptr
        el = C.c_float(0.0)
        if timed:
            d.elapsed_ms = C.cast(C.byref(el), C.c_void_p)
        check(lib().hb_run_device(self.ck.handle, C.byref(d), stream))
        return el.value if timed else None


class RouteStage:
    """Device-resident arguments of the routing stage of a model (`hb_run_route_device`): parameters,
    htypes, upstream CSR, initial state — staged once; `launch()` enqueues the T+1 routing kernels."""

    def __init__(self, eng: B200Model, N: int, T: int, params, cfg: HydroConfig, rec_rows: Sequence[str], init_state,
                 node_range: Optional[Tuple[int, int]] = None, csr=None):
        """`node_range=(n0, n1)` + `csr=(up_ptr, up_idx)`: this process owns nodes `[n0, n1)`; `up_idx`
        addresses local nodes `< N` and ghost slots `>= N` (see `parallel.RoutePartition`)."""
        route = eng.route
        if eng._route_kernel is None:
            eng._route_kernel = CompiledRoute(lower_route(route, fast=eng.fast))
        self.rk = eng._route_kernel
        self.N, self.T, self.cfg = int(N), int(T), cfg
        pdict = _as_componentvector(params)["params"]
        chunks, offs, pos, modes = [], [], 0, []
        h0 = np.asarray(route.htypes, dtype=np.int64) - 1
        n_global = len(h0)
        n0, n1 = node_range if node_range is not None else (0, N)
        if n1 - n0 != N or n1 > n_global or (node_range is None and n_global != N):
            raise ValueError(f"htypes has {n_global} entries but the input has {N} nodes")
        ident = np.array_equal(h0[n0:n1], np.arange(n0, n1))
        h0 = h0[n0:n1]
        for nm, _ in self.rk.program.param_slots:
            if nm not in pdict:
                raise KeyError(f"Parameter {nm} does not exist in params")
            v = np.atleast_1d(np.asarray(pdict[nm], dtype=np.float64)).reshape(-1)
            if h0.size and (h0.min() < 0 or h0.max() >= v.size):
                raise IndexError(f"BoundsError: attempt to access {v.size}-element parameter {nm} at index htypes")
            per_node = v.size == n_global and ident
            modes.append(1 if per_node else 2)
            if per_node:
                v = v[n0:n1]
            offs.append(pos)
            chunks.append(v)
            pos += v.size
        flat = np.ascontiguousarray(np.concatenate(chunks)) if chunks else np.zeros(1)
        self.n_param_values = flat.size
        self.p_off = np.array(offs or [0], dtype=np.int32)
        self.p_mode = np.array(modes or [0], dtype=np.int32)
        up_ptr, up_idx = csr if csr is not None else upstream_csr(route.aggr, N)
        self.bufs = [DeviceBuffer.from_host(flat), DeviceBuffer.from_host(h0.astype(np.int32)),
                     DeviceBuffer.from_host(up_ptr), DeviceBuffer.from_host(up_idx)]
        if init_state is not None:
            init_state = np.asarray(init_state, dtype=np.float64)
            if node_range is not None and init_state.ndim == 1 and init_state.size != N:
                init_state = init_state[n0:n1]
        self.dinit = None if init_state is None else DeviceBuffer.from_host(
            np.ascontiguousarray(np.broadcast_to(init_state, (N,))))
        self.in_row = np.array([list(rec_rows).index(nm) for nm in route.input_names] or [0], dtype=np.int32)
        self.n_rows = len(rec_rows)
        self.s_row = list(rec_rows).index(route.state_names[0])
        self.o_row = list(rec_rows).index(route.output_names[0])

    def launch(self, rec_ptr: Optional[int], stream: Optional[int] = None, timed: bool = False,
               planes: Optional[Tuple[int, int]] = None) -> Optional[float]:
        """`planes = (route_inputs_ptr, route_out_ptr)`: compact `[T][N][n_inputs]` in, `[T][N][2]` out instead of record rows."""
        el = C.c_float(0.0)
        d = RouteDesc()
        if planes is not None:
            d.route_inputs, d.route_out = planes
        d.struct_size = C.sizeof(RouteDesc)
        d.n_htype_sets = 1
        d.n_nodes, d.n_steps = self.N, self.T
        d.records, d.n_rows = rec_ptr, self.n_rows
        d.s_row, d.o_row = self.s_row, self.o_row
        d.in_row = _ptr(self.in_row)
        d.params, d.n_param_values = self.bufs[0].ptr, self.n_param_values
        d.param_offset, d.param_mode = _ptr(self.p_off), _ptr(self.p_mode)
        d.htypes = self.bufs[1].ptr
        d.initstates = self.dinit.ptr if self.dinit is not None else None
        d.up_ptr, d.up_idx = self.bufs[2].ptr, self.bufs[3].ptr
        d.min_value = self.cfg.min_value
        if timed:
            d.elapsed_ms = C.cast(C.byref(el), C.c_void_p)
        check(lib().hb_run_route_device(self.rk.handle, C.byref(d), stream))
        return el.value if timed else None

    def _desc(self, rec_ptr: int) -> RouteDesc:
        d = RouteDesc()
        d.struct_size = C.sizeof(RouteDesc)
        d.n_htype_sets = 1
        d.n_nodes, d.n_steps = self.N, self.T
        d.records, d.n_rows = rec_ptr, self.n_rows
        d.s_row, d.o_row = self.s_row, self.o_row
        d.in_row = _ptr(self.in_row)
        d.params, d.n_param_values = self.bufs[0].ptr, self.n_param_values
        d.param_offset, d.param_mode = _ptr(self.p_off), _ptr(self.p_mode)
        d.htypes = self.bufs[1].ptr
        d.up_ptr, d.up_idx = self.bufs[2].ptr, self.bufs[3].ptr
        d.min_value = self.cfg.min_value
        return d

    def step(self, rec_ptr: int, t: int, s_prev: int, s_next: int, q_cur: int, q_next: int, stream: Optional[int] = None):
        """One routing step on caller-owned ping-pong buffers (`hb_route_step_device`)."""
        d = self._desc(rec_ptr)
        check(lib().hb_route_step_device(self.rk.handle, C.byref(d), C.c_int64(t), s_prev, s_next, q_cur, q_next, stream))

    def free(self):
        for b in self.bufs + ([self.dinit] if self.dinit is not None else []):
            b.free()
        self.bufs, self.dinit = [], None


class RoutedDeviceRun:
    """Device-resident run of a model that ends in a HydroRoute on ANY node list / graph, in three passes that touch every
    byte once:

      1. the per-node stepper storing the route's input rows only -> compact plane `[T][N][RV]` (24 B in, 8 B out per cell-step);
      2. T+1 routing launches on compact planes: route inputs in, `{state, outflow}` out (`hb_route_step`, planes mode);
      3. the per-node stepper again with ALL rows, the route's two rows arriving through its auxiliary input stream — every
         record is written once by a TMA bulk store.

    Round 1 wrote the records first and patched the route's two 8-byte rows into the 104-byte records afterwards: 312 B of
    DRAM traffic per cell-step for the route alone (read-modify-write at sector granularity).  The bucket algebra is now
    evaluated twice (cheap: both stepper passes are HBM-bound) and the whole model moves ~1.6x its algorithmic bytes.
    `legacy=True` keeps the in-place path (used by the per-step sharded run, `parallel.ShardedRoutedRun`)."""

    def __init__(self, eng: B200Model, N: int, T: int, params, *, config=None, initstates=None, legacy: bool = False):
        if eng.route is None or eng.lumped is None:
            raise ValueError("RoutedDeviceRun needs a model with per-node components followed by a HydroRoute")
        self.cfg = normalize_config(config)
        rows = eng.component.var_names
        route = eng.route
        s_name, o_name = route.state_names[0], route.output_names[0]
        init = None if initstates is None else {k: v for k, v in initstates.items() if k != s_name}
        self.legacy = bool(legacy) or any(nm not in rows for nm in route.input_names)
        self.N, self.T, self.n_rows = int(N), int(T), len(rows)
        self.stage = RouteStage(eng, N, T, params, self.cfg, rows, None if initstates is None else initstates.get(s_name))
        if self.legacy:
            self.stepper = DeviceRun(eng, N, T, params, config=self.cfg, initstates=init, rows=rows)
        else:
            self.stepper = DeviceRun(eng, N, T, params, config=self.cfg, initstates=init, rows=rows, aux_rows=[s_name, o_name])
            self.route_in_rows = [list(rows).index(nm) for nm in route.input_names]
            if len(self.route_in_rows) > 4:
                raise NotImplementedError("at most 4 route inputs")
            self.xin = DeviceBuffer(max(1, T * N * len(route.input_names)) * 8)
            self.out2 = DeviceBuffer(max(1, T * N * 2) * 8)
        self.in_bytes, self.out_bytes = self.stepper.in_bytes, T * N * len(rows) * 8

    def launch(self, input_ptr: int, rec_ptr: int, stream: Optional[int] = None, timed: bool = False) -> Optional[float]:
        if self.legacy:
            ms = self.stepper.launch(input_ptr, rec_ptr, stream, timed=timed)
            ms2 = self.stage.launch(rec_ptr, stream, timed=timed)
            return (ms + ms2) if timed else None
        # pass 1 is the SAME kernel as pass 3 (no auxiliary stream yet, only the route's input rows stored): its values are
        # bit-identical to the rows pass 3 writes — a separately compiled "route inputs only" kernel differs in the last bit
        # where the compiler contracts multiplies and adds differently
        ms1 = self.stepper.launch(input_ptr, self.xin.ptr, stream, timed=timed, compact_rows=self.route_in_rows)
        ms2 = self.stage.launch(None, stream, timed=timed, planes=(self.xin.ptr, self.out2.ptr))
        ms3 = self.stepper.launch(input_ptr, rec_ptr, stream, timed=timed, aux_ptr=self.out2.ptr)
        return (ms1 + ms2 + ms3) if timed else None

    def free(self):
        self.stage.free()
        if not self.legacy:
            self.xin.free()
            self.out2.free()


# ---------------------------------------------------------------------------------------------------
# fused dense-grid path (templates/grid.cu): buckets + D8 route in one pass, records written once
# ---------------------------------------------------------------------------------------------------
def dense_grid_shape(aggr, n_nodes: int) -> Optional[Tuple[int, int]]:
    """(rows, cols) when the node list is the complete grid in row-major order, else None."""
    if not isinstance(aggr, GridAggregation):
        return None
    R, Cc = aggr.flwdir.shape
    if R * Cc != n_nodes or len(aggr.positions) != n_nodes:
        return None
    rr, cc = np.divmod(np.arange(n_nodes), Cc)
    if np.array_equal(aggr.positions[:, 0], rr + 1) and np.array_equal(aggr.positions[:, 1], cc + 1):
        return R, Cc
    return None


GRID_VARIANTS = {"tile": 0, "wave": 1}


class CompiledGrid:
    def __init__(self, program: GridProgram, steps_per_chunk: int = 2, block_w: int = 32, block_h: int = 32,
                 variant: str = "tile", stages: int = 0, max_registers: int = 0):
        self.program = program
        sp, rp = program.stepper, program.route
        spec = GridSpec(C.sizeof(GridSpec), len(sp.input_names), len(program.row_names), len(sp.state_names),
                        len(sp.param_slots) + len(rp.param_slots), len(sp.param_slots), len(rp.input_names),
                        program.s_row, program.o_row, steps_per_chunk, block_w, block_h, GRID_VARIANTS[variant], stages,
                        max_registers)
        self.spec = spec
        h = C.c_void_p()
        check(lib().hb_compile_grid(C.byref(spec), program.body.encode(), C.byref(h)))
        self.handle = h

    def source(self) -> str:
        s = C.c_char_p()
        n = C.c_int64()
        check(lib().hb_model_source(self.handle, C.byref(s), C.byref(n)))
        return s.value.decode()

    def info(self) -> Dict[str, int]:
        k = KernelInfo()
        check(lib().hb_model_info(self.handle, C.byref(k)))
        return {f: getattr(k, f) for f, _ in KernelInfo._fields_}

    def __del__(self):
        try:
            if self.handle and _lib is not None:
                _lib.hb_free_model(self.handle)
                self.handle = None
        except Exception:
            pass


class GridDeviceRun:
    """Device-resident run of `[buckets…, HydroRoute]` on a dense D8 grid with a fused kernel.

    `row_range=(g0, g1)`: this process runs grid rows `[g0, g1)` only (a strip of a multi-GPU run, ghost rows
    included — see `parallel.ShardedGridRun`); `N` is then the number of cells of the strip and flow directions
    leaving the strip count as off-grid."""

    def __init__(self, eng: B200Model, N: int, T: int, params, *, config=None, initstates=None,
                 steps_per_chunk: int = 2, block_w: Optional[int] = None, block_h: int = 32, variant: str = "tile", stages: int = 0,
                 max_registers: int = 0, row_range: Optional[Tuple[int, int]] = None):
        if block_w is None:
            block_w = 32 if variant == "tile" else 0       # tile: region width in cells; wave: threads per block (0 = 128)
        if eng.route is None or eng.lumped is None:
            raise ValueError("GridDeviceRun needs per-node components followed by a HydroRoute")
        n_global = len(eng.route.htypes)
        shape = dense_grid_shape(eng.route.aggr, n_global)
        if shape is None:
            raise ValueError("the fused grid kernels need the complete grid in row-major node order")
        g0, g1 = row_range if row_range is not None else (0, shape[0])
        self.rows, self.cols = g1 - g0, shape[1]
        node_range = None if row_range is None else (g0 * self.cols, g1 * self.cols)
        if N != self.rows * self.cols:
            raise ValueError(f"{N} nodes given for a {self.rows} x {self.cols} grid (strip)")
        self.cfg = normalize_config(config)
        _check_config(self.cfg, T)
        rows = eng.component.var_names
        key = ("grid", variant, steps_per_chunk, block_w, block_h, stages, max_registers)
        if key not in eng._cache:
            eng._cache[key] = CompiledGrid(lower_grid(eng.lumped, eng.route, rows, fast=eng.fast), steps_per_chunk, block_w, block_h,
                                           variant, stages, max_registers)
        self.ck = eng._cache[key]
        sp, rp = self.ck.program.stepper, self.ck.program.route
        flat1, off1, mode1, ht1, _, _, _ = eng.prepare(sp, params, N, node_range=node_range)
        dummy_csr = (np.zeros(N + 1, dtype=np.int32), np.zeros(1, dtype=np.int32))     # the grid kernels gather through flwdir
        stage = RouteStage(eng, N, T, params, self.cfg, rows, None if initstates is None else initstates.get(eng.route.state_names[0]),
                           node_range=node_range, csr=dummy_csr)
        flat2 = stage.bufs[0].to_host((stage.n_param_values,))
        n_sets1 = 0 if ht1 is None else ht1.shape[0]
        flat = np.concatenate([flat1, flat2])
        p_off = np.concatenate([off1[:len(sp.param_slots)], stage.p_off[:len(rp.param_slots)] + flat1.size]).astype(np.int32)
        # the route's gather modes address its own htypes set, appended after the stepper's sets
        mode2 = np.where(stage.p_mode[:len(rp.param_slots)] >= 2, stage.p_mode[:len(rp.param_slots)] + n_sets1, stage.p_mode[:len(rp.param_slots)])
        p_mode = np.concatenate([mode1[:len(sp.param_slots)], mode2]).astype(np.int32)
        h_route = np.asarray(eng.route.htypes, dtype=np.int64) - 1
        if node_range is not None:
            h_route = h_route[node_range[0]:node_range[1]]
        hts = ([ht1] if ht1 is not None else []) + [h_route.astype(np.int32)[None, :]]
        htypes = np.ascontiguousarray(np.concatenate(hts, axis=0))
        if p_off.size == 0:
            p_off, p_mode = np.zeros(1, np.int32), np.zeros(1, np.int32)
        self._host = (np.ascontiguousarray(p_off), np.ascontiguousarray(p_mode))
        self.n_sets = htypes.shape[0]
        self.n_param_values = flat.size
        flw = eng.route.aggr.flwdir[g0:g1]
        self.bufs = {"params": DeviceBuffer.from_host(flat), "htypes": DeviceBuffer.from_host(htypes),
                     "flw": DeviceBuffer.from_host(np.where(np.isin(flw, D8_CODES), flw, 0).astype(np.uint8))}
        init = eng._initstates(sp, None if initstates is None else {k: v for k, v in initstates.items()
                                                                     if k != eng.route.state_names[0]}, N, node_range)
        if init is not None:
            self.bufs["init"] = DeviceBuffer.from_host(init)
        if stage.dinit is not None:
            self.bufs["rinit"] = DeviceBuffer.from_host(stage.dinit.to_host((N,)))
        stage.free()
        self.N, self.T, self.n_rows, self.n_states = int(N), int(T), len(rows), len(sp.state_names)
        self.in_bytes, self.out_bytes = T * N * len(sp.input_names) * 8, T * N * len(rows) * 8

    def launch_chunk(self, input_ptr: int, rec_ptr: int, n_steps: int, init_bucket: Optional[int], init_route: Optional[int],
                     final_bucket: Optional[int], final_route: Optional[int], out_first: int = 0, out_count: int = 0,
                     stream: Optional[int] = None, out_window: Optional[Tuple[int, int, int, int]] = None,
                     in_pitched: Optional[Tuple[int, int, int]] = None, out_pitched: Optional[Tuple[int, int, int]] = None):
        """`n_steps` steps from caller-owned state buffers (`[S][N]` / `[N]` device pointers, None = zeros / not wanted).
        Records are written for whole rows `[out_first, out_first + out_count)` (node range, 0 = all) or for the window
        `out_window = (row0, rows, col0, cols)`, as `[T][rows][cols][R]`.  `in_pitched = (t_stride, pitch, col0)` /
        `out_pitched = (t_stride, pitch, base)` (in cells) address the forcing / the records of this sub-grid inside the
        arrays of a wider grid (`hb_grid_desc`).  Asynchronous on `stream`."""
        d = self._desc(input_ptr, rec_ptr, n_steps)
        d.init_bucket, d.init_route, d.final_bucket, d.final_route = init_bucket, init_route, final_bucket, final_route
        if in_pitched is not None:
            d.in_t_stride, d.in_pitch, d.in_col0 = [int(x) for x in in_pitched]
        if out_pitched is not None:
            d.out_t_stride, d.out_pitch, d.out_base = [int(x) for x in out_pitched]
        if out_window is not None:
            d.out_row0, d.out_rows, d.out_col0, d.out_cols = [int(x) for x in out_window]
        elif out_count:
            if out_first % self.cols or out_count % self.cols:
                raise ValueError("an output node range must consist of whole grid rows")
            d.out_row0, d.out_rows = out_first // self.cols, out_count // self.cols
        check(lib().hb_run_grid_device(self.ck.handle, C.byref(d), stream))

    def _desc(self, input_ptr: int, rec_ptr: int, n_steps: int) -> GridDesc:
        d = GridDesc()
        d.struct_size = C.sizeof(GridDesc)
        d.n_htype_sets = self.n_sets
        d.rows, d.cols, d.n_steps = self.rows, self.cols, n_steps
        d.input, d.out = input_ptr, rec_ptr
        d.params, d.n_param_values = self.bufs["params"].ptr, self.n_param_values
        d.param_offset, d.param_mode = _ptr(self._host[0]), _ptr(self._host[1])
        d.htypes = self.bufs["htypes"].ptr
        d.flwdir = self.bufs["flw"].ptr
        d.min_value = self.cfg.min_value
        return d

    def launch(self, input_ptr: int, rec_ptr: int, stream: Optional[int] = None, timed: bool = False) -> Optional[float]:
        d = self._desc(input_ptr, rec_ptr, self.T)
        d.init_bucket = self.bufs["init"].ptr if "init" in self.bufs else None
        d.init_route = self.bufs["rinit"].ptr if "rinit" in self.bufs else None
        el = C.c_float(0.0)
        if timed:
            d.elapsed_ms = C.cast(C.byref(el), C.c_void_p)
        check(lib().hb_run_grid_device(self.ck.handle, C.byref(d), stream))
        return el.value if timed else None

    def check(self, stream: Optional[int] = None) -> int:
        """Synchronise and raise if a warp of the wave kernel gave up waiting; returns the steps per launch used."""
        tc = C.c_int32(0)
        check(lib().hb_grid_status(self.ck.handle, stream, C.byref(tc)))
        return tc.value
