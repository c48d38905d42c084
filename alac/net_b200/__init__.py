"""alac.net_b200 -- B200-native (sm_100a) ALAC frame decode path behind the
teekay/ALAC.NET API.  The product is libalacgpu.so (csrc/, C ABI in
include/alacgpu.h); this package holds its build script, the ctypes binding and
the host-side mirrors of the reference's AlacContext / ALACFileReader."""
from ._native import AlacGpuError, LIB_PATH, load  # noqa: F401
from .decoder import BatchDecoder, PinnedBuffer, host_checksum, plan_partition  # noqa: F401
