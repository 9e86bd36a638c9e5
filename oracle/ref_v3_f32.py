"""TEST INFRASTRUCTURE: the ch4/v3 reference built with `using type_calc = float` (all.h:11), oracle/_ref/libref_v3_f32.so - the fp32
oracle SURVEY.md 8c names.  Same wrapper classes as oracle/ref_v3.py (a second instance of that module bound to the fp32 library).

Only tests/ may import this.
"""
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("oracle._ref_v3_f32_impl", os.path.join(_HERE, "ref_v3.py"))
_impl = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_impl)
_impl.LIB_PATH = os.path.join(_HERE, "_ref", "libref_v3_f32.so")
globals().update({k: getattr(_impl, k) for k in dir(_impl) if not k.startswith("__")})
