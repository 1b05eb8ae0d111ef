"""CPU-only: the C-ABI library loads, exports every symbol include/hydrob200.h declares, compiles
kernels without a device, and fails loudly (no CPU fallback) when asked to run without one."""
import ctypes
import os
import re

import numpy as np
import pytest

import hydromodels_jl_b200 as hm
from hydromodels_jl_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "hydrob200.h")).read()
    declared = set(re.findall(r"^\s*(?:const\s+char\s*\*|int)\s*(hb_\w+)\s*\(", hdr, flags=re.M))
    assert {"hb_compile_stepper", "hb_run", "hb_run_device", "hb_last_error", "hb_version"} <= declared
    lib = ctypes.CDLL(engine.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in hydrob200.h but not exported"
    assert engine.lib().hb_version() == 1


def test_struct_sizes_match_header():
    # the library checks struct_size; a mismatch would surface as HB_ERR_INVALID at compile time
    eng = hm.B200Model(hm.models.exphydro())
    ck = eng.compiled()
    assert len(ck.cubin()) > 1000 and "hb_stepper" in ck.source()
    assert "cp.async.bulk" in ck.source()


def test_ctypes_mirrors_have_the_layout_of_the_header(tmp_path):
    """Every field of every struct: offsetof/sizeof from a C compile of include/hydrob200.h against the ctypes mirrors."""
    import subprocess
    pairs = {"hb_stepper_spec": engine.StepperSpec, "hb_run_desc": engine.RunDesc, "hb_kernel_info": engine.KernelInfo,
             "hb_route_spec": engine.RouteSpec, "hb_route_desc": engine.RouteDesc, "hb_grid_spec": engine.GridSpec,
             "hb_grid_desc": engine.GridDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "hydrob200.h"', "int main(void) {"]
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for f, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in pairs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for f, _ in cls._fields_:
            assert int(got[f"{cname}.{f}"]) == getattr(cls, f).offset, f"{cname}.{f}"


def test_nvrtc_error_is_reported():
    from hydromodels_jl_b200.lowering import lower_stepper
    prog = lower_stepper(hm.models.exphydro())
    prog.body = prog.body.replace("hb_out[0] =", "hb_out[0] = undefined_symbol +")
    with pytest.raises(engine.HydroB200Error, match="undefined_symbol"):
        engine.CompiledStepper(prog)


@pytest.mark.skipif(hm.device_count() > 0, reason="needs a box without a GPU")
def test_no_device_means_no_result():
    model = hm.models.exphydro()
    with pytest.raises(hm.NoDeviceError, match="no CPU fallback"):
        model(np.zeros((3, 10)), {"params": dict(f=0.0167, Smax=1709.46, Qmax=18.47, Df=2.674, Tmax=0.17, Tmin=-2.09)})


def test_argument_errors_raise_before_any_device_work():
    model = hm.models.exphydro(htypes=np.arange(1, 9))
    p = {"params": {k: np.full(8, 1.0) for k in ["Tmin", "Tmax", "Df", "Smax", "Qmax", "f"]}}
    with pytest.raises(ValueError, match="only accepts 3D input"):
        model(np.zeros((3, 10)), p)
    with pytest.raises(AssertionError, match="Input variables length mismatch"):
        model(np.zeros((2, 8, 10)), p)
    with pytest.raises(TypeError):
        model(np.zeros((3, 8, 10)), [1.0, 2.0])
    with pytest.raises(NotImplementedError):
        model(np.zeros((3, 8, 10)), p, hm.HydroConfig(solver=hm.ODESolver))
    with pytest.raises(ValueError, match="reads 'b' before it is defined"):
        s = hm.Symbols("a b c", "")
        hm.HydroBucket(fluxes=[hm.HydroFlux("c ~ b + a", symbols=s), hm.HydroFlux("b ~ a", symbols=s)])


def test_host_copy_pool_is_exact_for_any_alignment_and_size():
    """`hb_host_copy`: the parallel streaming-store copy `hb_run` uses to stage pageable arrays (no device involved)."""
    L = engine.lib()
    n = ctypes.c_int(0)
    assert L.hb_host_threads(ctypes.byref(n)) == 0 and n.value >= 1
    rng = np.random.default_rng(0)
    src = rng.integers(0, 256, size=(40 << 20) + 977, dtype=np.uint8)
    for off_s, off_d, size in [(0, 0, 0), (1, 3, 5), (7, 1, 4095), (3, 61, 4096 + 129), (5, 64, (9 << 20) + 13), (0, 32, 40 << 20)]:
        dst = np.full(size + 256, 0xAB, dtype=np.uint8)
        assert L.hb_host_copy(dst[off_d:].ctypes.data, src[off_s:].ctypes.data, size) == 0
        assert np.array_equal(dst[off_d:off_d + size], src[off_s:off_s + size])
        assert np.all(dst[:off_d] == 0xAB) and np.all(dst[off_d + size:] == 0xAB)
