"""Raster ingest: host-side mirrors of the reference's Rasters extension
(`/root/reference/ext/HydroModelsRastersExt/HydroModelsRastersExt.jl`, `netcdf_loader.jl`).

Same names and argument meaning as the reference: `compute_d8_flow_direction(dem)`, `extract_positions_from_raster`,
`extract_positions_from_netcdf`, `load_netcdf_data`, `load_netcdf_data_simple`, `prepare_netcdf_forcing`,
`create_hydroroute_from_dem`.  A "raster" here is a plain `numpy` array `[rows, cols(, time)]` plus its cell sizes
(the reference takes `Rasters.Raster` objects; NetCDF files are read with `scipy.io.netcdf_file`, NetCDF-3 classic).

What runs on the GPU: the D8 steepest-descent stencil (`hb_d8_flow_direction_device`) and the gather of the variable
rasters at the node positions into the stepper's forcing layout `[T][N][V]` (`hb_pack_forcing_device`,
`device_forcing`) — so a gridded run goes NetCDF -> pinned host -> HBM -> packed forcing -> `GridDeviceRun` /
`DeviceRun` without a host-side transpose; `ingest_forcing` does the upload and the gather in one streamed call
(`hb_ingest_forcing`: two chunk slots, host rasters never whole on the device).  Neither has a CPU path here; position lists (one integer pair per
valid cell, computed once per grid) are plain numpy.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np

D8_CODES = (1, 2, 4, 8, 16, 32, 64, 128)     # E, SE, S, SW, W, NW, N, NE  (HydroModelsRastersExt.jl:16-24)


def compute_d8_flow_direction(dem: np.ndarray, cellsize: Tuple[float, float] = (1.0, 1.0)) -> np.ndarray:
    """`compute_d8_flow_direction(dem::Raster)` (`HydroModelsRastersExt.jl:32-109`) on the GPU.

    `dem[rows, cols]` float64 (NaN = NoData), `cellsize = (|step(X)|, |step(Y)|)`.  Returns the integer D8 matrix
    `[rows, cols]` the reference returns (codes 1..128, 0 for NoData cells)."""
    from .engine import DeviceBuffer, check, lib
    dem = np.asarray(dem, dtype=np.float64)
    if dem.ndim != 2:
        raise ValueError("dem must be a [rows, cols] array")
    rows, cols = dem.shape
    d = DeviceBuffer.from_host(np.ascontiguousarray(dem))
    out = DeviceBuffer(rows * cols)
    check(lib().hb_d8_flow_direction_device(d.ptr, rows, cols, cols, 1, float(cellsize[0]), float(cellsize[1]), out.ptr, None))
    codes = out.to_host((rows, cols), dtype=np.uint8)
    d.free()
    out.free()
    return codes.astype(np.int64)


def extract_positions_from_raster(data: np.ndarray, valid_mask: Optional[Callable[[np.ndarray], np.ndarray]] = None) -> List[Tuple[int, int]]:
    """`extract_positions_from_raster` (`HydroModelsRastersExt.jl:124-137`): 1-based `(row, col)` of the valid cells,
    rows outer, columns inner.  `valid_mask` maps the array to booleans (default: not NaN, the numpy stand-in for
    `!ismissing`)."""
    data = np.asarray(data)
    mask = ~np.isnan(data) if valid_mask is None else np.asarray(valid_mask(data), dtype=bool)
    rr, cc = np.nonzero(mask)                       # row-major order = the reference's loop order
    return [(int(r) + 1, int(c) + 1) for r, c in zip(rr, cc)]


def _open_netcdf(file_path: str):
    from scipy.io import netcdf_file
    return netcdf_file(file_path, "r", mmap=False)


def extract_positions_from_netcdf(file_path: str, var_name: str, threshold: float = 0.0) -> List[Tuple[int, int]]:
    """`extract_positions_from_netcdf` (`netcdf_loader.jl:151-177`): cells of the first time slice that are not NaN and
    exceed `threshold`.  Arrays are taken as `[row, col, time]` (the reference indexes `raster[row, col, t]`)."""
    with _open_netcdf(file_path) as ds:
        if var_name not in ds.variables:
            raise KeyError(f"Variable '{var_name}' not found in NetCDF file")
        a = np.array(ds.variables[var_name][:], dtype=np.float64)
    a2 = a[:, :, 0] if a.ndim == 3 else a
    return extract_positions_from_raster(a2, lambda x: ~np.isnan(x) & (x > threshold))


def load_netcdf_data(file_path: str, var_mapping: Mapping[str, str], positions: Sequence[Tuple[int, int]]) -> Dict[str, np.ndarray]:
    """`load_netcdf_data` (`netcdf_loader.jl:29-62`): `{model variable: array [n_timesteps, n_positions]}` from variables
    stored as `[row, col, time]`, sampled at the 1-based `positions`."""
    pos = np.asarray(positions, dtype=np.int64).reshape(-1, 2) - 1
    out: Dict[str, np.ndarray] = {}
    with _open_netcdf(file_path) as ds:
        for model_var, nc_var in var_mapping.items():
            if str(nc_var) not in ds.variables:
                raise KeyError(f"Variable '{nc_var}' not found in NetCDF file")
            a = np.array(ds.variables[str(nc_var)][:], dtype=np.float64)
            out[str(model_var)] = np.ascontiguousarray(a[pos[:, 0], pos[:, 1], :].T)
    return out


def load_netcdf_data_simple(file_path: str, var_mapping: Mapping[str, str]) -> Dict[str, np.ndarray]:
    """`load_netcdf_data_simple` (`netcdf_loader.jl:76-105`): one series per variable; 3-D variables are averaged over
    space."""
    out: Dict[str, np.ndarray] = {}
    with _open_netcdf(file_path) as ds:
        for model_var, nc_var in var_mapping.items():
            if str(nc_var) not in ds.variables:
                raise KeyError(f"Variable '{nc_var}' not found in NetCDF file. Available: {list(ds.variables)}")
            a = np.array(ds.variables[str(nc_var)][:], dtype=np.float64)
            if a.ndim == 1:
                out[str(model_var)] = a
            elif a.ndim == 3:
                out[str(model_var)] = a.mean(axis=(0, 1))
            else:
                raise ValueError(f"Unsupported data dimensions for variable '{nc_var}'")
    return out


def prepare_netcdf_forcing(data: Mapping[str, np.ndarray], var_order: Sequence[str]) -> np.ndarray:
    """`prepare_netcdf_forcing` (`netcdf_loader.jl:119-136`): `[n_timesteps, n_variables, n_positions]` in `var_order`.
    (The model functors take `[variables, nodes, time]`: use `model_input` / `device_forcing` for that.)"""
    first = np.asarray(data[var_order[0]])
    forcing = np.zeros((first.shape[0], len(var_order), first.shape[1]))
    for i, v in enumerate(var_order):
        if v not in data:
            raise KeyError(f"Variable {v} not found in data. Available: {list(data)}")
        forcing[:, i, :] = data[v]
    return forcing


def model_input(data: Mapping[str, np.ndarray], var_order: Sequence[str]) -> np.ndarray:
    """`[variables, nodes, time]` Fortran-ordered (memory `[T][N][V]`) from `load_netcdf_data`'s dictionary."""
    first = np.asarray(data[var_order[0]])
    out = np.empty((len(var_order), first.shape[1], first.shape[0]), order="F")
    for i, v in enumerate(var_order):
        out[i] = np.asarray(data[v]).T
    return out


def create_hydroroute_from_dem(dem: np.ndarray, cellsize: Tuple[float, float] = (1.0, 1.0)):
    """`create_hydroroute_from_dem` (`HydroModelsRastersExt.jl:174-196`): flow directions, valid positions, the D8
    aggregation and all-ones `htypes` (fluxes are the caller's, as in the reference)."""
    from .components import build_aggr_func
    flow_dir = compute_d8_flow_direction(dem, cellsize)
    positions = extract_positions_from_raster(dem)
    return dict(flow_dir=flow_dir, positions=positions, aggr_func=build_aggr_func(flow_dir, positions),
                htypes=np.ones(len(positions), dtype=np.int64))


def device_forcing(rasters: Sequence, positions: Sequence[Tuple[int, int]], n_steps: int, rows: int, cols: int, *,
                   layout: str = "trc", stream: Optional[int] = None):
    """Gather `len(rasters)` DEVICE rasters (pointers or `engine.DeviceBuffer`s, each `n_steps x rows x cols` doubles) at
    the 1-based `positions` into a new device buffer laid out `[T][N][V]` — the forcing `DeviceRun` / `GridDeviceRun`
    take.  `layout`: memory order of a raster, slowest first — "trc" (NetCDF's usual time, row, col) or "tcr" (the
    column-major `[row, col, time]` arrays the reference indexes)."""
    from .engine import DeviceBuffer, check, lib
    if layout not in ("trc", "tcr"):
        raise ValueError("layout must be 'trc' or 'tcr'")
    pos = np.asarray(positions, dtype=np.int64).reshape(-1, 2) - 1
    if pos.size and (pos.min() < 0 or pos[:, 0].max() >= rows or pos[:, 1].max() >= cols):
        raise IndexError("position outside the raster")
    N, V = len(pos), len(rasters)
    ptrs = np.array([r.ptr if hasattr(r, "ptr") else int(r) for r in rasters], dtype=np.uint64)
    d_ptrs = DeviceBuffer.from_host(ptrs)
    # every cell, row-major: the tiled kernel (no position arrays, coalesced on both sides for either memory order)
    dense = V <= 5 and N == rows * cols and np.array_equal(pos[:, 0] * cols + pos[:, 1], np.arange(N))
    d_r = DeviceBuffer.from_host(np.ascontiguousarray(pos[:, 0], dtype=np.int32))
    d_c = DeviceBuffer.from_host(np.ascontiguousarray(pos[:, 1], dtype=np.int32))
    out = DeviceBuffer(n_steps * N * V * 8)
    rs, cs = (cols, 1) if layout == "trc" else (1, rows)
    for t0 in range(0, n_steps, 32768):
        nt = min(32768, n_steps - t0)
        sub = np.array([int(p) + t0 * rows * cols * 8 for p in ptrs], dtype=np.uint64)
        d_sub = d_ptrs if t0 == 0 and nt == n_steps else DeviceBuffer.from_host(sub)
        if dense:
            check(lib().hb_pack_forcing_dense_device(d_sub.ptr, V, nt, rows * cols, rs, cs, rows, cols, out.ptr + t0 * N * V * 8, stream))
        else:
            check(lib().hb_pack_forcing_device(d_sub.ptr, V, nt, rows * cols, rs, cs, d_r.ptr, d_c.ptr, N,
                                               out.ptr + t0 * N * V * 8, stream))
        if d_sub is not d_ptrs:
            check(lib().hb_stream_synchronize(stream))
            d_sub.free()
    check(lib().hb_stream_synchronize(stream))
    for b in (d_ptrs, d_r, d_c):
        b.free()
    return out


def ingest_forcing(rasters: Sequence[np.ndarray], positions: Optional[Sequence[Tuple[int, int]]] = None, *, layout: str = "trc",
                   stream: Optional[int] = None, out=None):
    """HOST rasters -> a device buffer laid out `[T][N][V]`, streamed through the library (`hb_ingest_forcing`): what
    `load_netcdf_data` + `prepare_netcdf_forcing` (`netcdf_loader.jl:29-62, :119-136`) and an upload would do, without ever
    holding the rasters on the device — time chunks are copied (pinned staging by the library's host threads for ordinary
    numpy / memory-mapped arrays) while the previous chunk is being gathered.

    `rasters`: one array per variable, each `[time, rows, cols]` C-ordered ("trc", NetCDF's usual order) or `[rows, cols, time]`
    Fortran-ordered ("tcr", the arrays the reference indexes) — the time axis is the slowest in memory either way.
    `positions`: 1-based `(row, col)` pairs, or None for every cell in row-major order.  `out`: an existing `DeviceBuffer` of
    `T * N * V * 8` bytes to fill (a loop over files reuses one), else a new one is allocated."""
    from .engine import DeviceBuffer, check, lib
    if layout not in ("trc", "tcr"):
        raise ValueError("layout must be 'trc' or 'tcr'")
    arrs = []
    for a in rasters:
        a = np.asarray(a)
        if a.dtype != np.float64 or a.ndim != 3:
            raise TypeError("rasters must be 3-dimensional float64 arrays")
        if layout == "trc" and not a.flags.c_contiguous:
            raise ValueError("layout 'trc' takes C-contiguous [time, rows, cols] arrays")
        if layout == "tcr" and not a.flags.f_contiguous:
            raise ValueError("layout 'tcr' takes Fortran-contiguous [rows, cols, time] arrays")
        arrs.append(a)
    if not arrs or any(a.shape != arrs[0].shape for a in arrs):
        raise ValueError("rasters must be a non-empty list of equally shaped arrays")
    if layout == "trc":
        T, rows, cols = arrs[0].shape
        rs, cs = cols, 1
    else:
        rows, cols, T = arrs[0].shape
        rs, cs = 1, rows
    V = len(arrs)
    ptrs = (C.c_void_p * V)(*[a.ctypes.data for a in arrs])
    bufs = []
    if positions is None:
        N, d_r, d_c = rows * cols, None, None
    else:
        pos = np.asarray(positions, dtype=np.int64).reshape(-1, 2) - 1
        if pos.size and (pos.min() < 0 or pos[:, 0].max() >= rows or pos[:, 1].max() >= cols):
            raise IndexError("position outside the raster")
        N = len(pos)
        d_r = DeviceBuffer.from_host(np.ascontiguousarray(pos[:, 0], dtype=np.int32))
        d_c = DeviceBuffer.from_host(np.ascontiguousarray(pos[:, 1], dtype=np.int32))
        bufs = [d_r, d_c]
    if out is None:
        out = DeviceBuffer(T * N * V * 8)
    elif out.nbytes < T * N * V * 8:
        raise ValueError(f"out holds {out.nbytes} bytes, the forcing needs {T * N * V * 8}")
    try:
        check(lib().hb_ingest_forcing(ptrs, V, T, rows * cols, rs, cs, rows, cols, d_r.ptr if d_r else None, d_c.ptr if d_c else None, N,
                                      out.ptr, stream))
    finally:
        for b in bufs:
            b.free()
    return out
