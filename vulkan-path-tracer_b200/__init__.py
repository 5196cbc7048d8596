"""B200-native wavefront path tracer: the hot path of Zydak/Vulkan-Path-Tracer behind a C-ABI.
The product is `libb200pt.so` (built from csrc/ for sm_100a); `binding` is its ctypes harness."""
from .binding import *  # noqa: F401,F403
from . import binding  # noqa: F401
