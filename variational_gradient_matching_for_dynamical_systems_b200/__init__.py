"""B200-native variational gradient matching (VGM) for dynamical systems.

Drop-in modules (same function names and signatures as the reference's ``VGM_modules/``):
``GP_regression``, ``import_odes``, ``rewrite_odes_as_local_linear_combinations``,
``proxies_for_ode_parameters_and_states``, ``kernels``.  All arithmetic runs in hand-written
FP64 CUDA kernels for sm_100a behind the C ABI of ``include/vgm_b200.h`` (``libvgm_b200.so``);
there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
from .codegen import OdeModel, Program, build_host_plan, nvrtc_compile  # noqa: F401


def __getattr__(name):
    if name in ("VGMEngine", "Operators"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
