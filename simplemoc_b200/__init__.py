"""simplemoc_b200 -- B200-native 3D MOC transport sweep behind SimpleMOC's C interface.

The product is the native library `libmoc_b200.so` (CUDA kernels for sm_100a + C-ABI,
sources under csrc/, interface in include/moc_b200.h) and the C driver `SimpleMOC-b200`.
This Python package only binds that library for tests and benchmarks.
"""
from . import api  # noqa: F401
from .api import (DeviceProblem, HostProblem, MocError, default_input, derive,  # noqa: F401
                  device_count, input_from_values, make_grid, read_input_file, small_input)
