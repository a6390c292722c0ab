"""pflare_b200 -- B200-native (sm_100a CUDA) AIRG V-cycle apply behind PFLARE's PCAIR / PCPFLAREINV
apply interface.  The product is ``libpflare_b200.so`` (C-ABI: include/pflare_b200.h); this package
is the thin host-side mirror used by tests and bench.py.  No CPU fallback exists."""
from ._capi import PflareB200Error, LIB_PATH, lib  # noqa: F401
from .device import ClusterAIR, DeviceAIR, AFF, AFC, ACF, ACC, INV_AFF, INV_ACC, R, P, COARSE  # noqa: F401
from .pc import PC, PCAIR, PCPFLAREINV  # noqa: F401
from .upload import feed  # noqa: F401
from . import petsc_io  # noqa: F401
