"""ocean-bgc_b200 — B200-native drop-in for the Ocean-BGC column hot path.

Host-side Python mirror of the reference's operator interface over the C ABI
(include/bgc_b200.h).  Imported under the module name `ocean_bgc_b200` via
`__graft_entry__.load_package()` (the directory name carries a hyphen).
"""
from . import abi          # noqa: F401
from . import columns      # noqa: F401
from . import sharding     # noqa: F401
from .columns import (BgcColumns, DmsColumns, MacrosColumns, synth_fill, synth_co2_points,  # noqa: F401
                      synth_fill_device)


def __getattr__(name):
    # the CUDA-backed host layer loads the C-ABI shared library on first use, so
    # that `import ocean_bgc_b200` works on a box without the built library
    # (it is an error to *call* anything there: there is no CPU fallback).
    if name == "host":
        import importlib
        return importlib.import_module(".host", __name__)
    raise AttributeError(name)
