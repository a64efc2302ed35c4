"""p_winds_b200 -- B200-native batched forward-model engine for the p-winds
retrieval hot path (hydrogen.ion_fraction -> parker.structure ->
helium.population_fraction -> transit.draw_transit -> transit.radiative_transfer_2d).

The modules `parker`, `hydrogen`, `helium`, `transit`, `lines`, `tools` keep the
reference's function signatures (p_winds v1.4.7); `engine.Engine` is the batched
interface a retrieval uses.  All numerical work runs in hand-written sm_100a CUDA
kernels behind the C ABI declared in include/pwinds_b200.h; there is no CPU
fallback: importing the engine without the built library or without a CUDA device
raises.
"""
from . import lines, tools  # noqa: F401  (pure constants / host I/O helpers)

__all__ = ['parker', 'hydrogen', 'helium', 'oxygen', 'carbon', 'transit', 'retrieval', 'lines', 'tools', 'engine']
__version__ = '0.1.0'


def __getattr__(name):
    # the compute modules import the CUDA library lazily so that `import p_winds_b200`
    # (and the constants) work in a CPU-only build container
    if name in ('parker', 'hydrogen', 'helium', 'oxygen', 'carbon', 'transit', 'retrieval', 'engine', '_lib'):
        import importlib
        return importlib.import_module('.' + name, __name__)
    raise AttributeError(name)
