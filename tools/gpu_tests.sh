python -m pytest tests -m gpu -q -x 2>&1 | tail -15
python -c "
import hydromodels_jl_b200 as hm, ctypes
hm.engine.init(0)
v=ctypes.c_double(); print('fp64 dfma probe rc', hm.engine.lib().hb_probe_fp64_tflops(ctypes.byref(v)), 'TFLOP/s', v.value)
v=ctypes.c_double(); print('fp64 dmma probe rc', hm.engine.lib().hb_probe_dmma_tflops(ctypes.byref(v)), 'TFLOP/s', v.value)"
