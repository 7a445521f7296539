"""Import alias: the package directory is named `opencl-raytracing_b200/` (not a valid Python
identifier), so `import raytracing_cuda` loads it under the name of the crate it stands in for
(`raytracing-cuda`, the `--backend cuda` sibling of crates/raytracing-cpu)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "opencl-raytracing_b200")
_spec = importlib.util.spec_from_file_location("raytracing_cuda", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["raytracing_cuda"] = _mod
_spec.loader.exec_module(_mod)
