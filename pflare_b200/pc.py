"""Host-side mirror of the reference's PC interface for the apply path (PETSc is absent in this
image, so this plays the role of the PETSc `PC` object in tests and bench.py).

    pc = PC(); pc.setType("air")            # PCSetType(pc, PCAIR)     src/PCAIR.c:3601-3644
    pc.setHierarchy(H)                      # what the reference's PCSetUp built (air_multigrid_data)
    pc.setUp()                              # PCSetUp_AIR_c -> upload hook (src/PCAIR.c:182-201)
    pc.apply(x, y)                          # PCApply_AIR_c (src/PCAIR.c:150-166): y = M^-1 x
    pc.reset(); pc.destroy()                # PCReset_AIR_c / PCDestroy (src/PCAIR.c:134-148)

`setType("pflareinv")` gives PCPFLAREINV (src/PCPFLAREINV.c:618-626): apply = MatMult(mat_inverse).

The hierarchy itself is an *input*: the reference's own setup builds it (BASELINE.json); here it
arrives as a container object exposing what ``air_multigrid_data`` holds and is walked by
``feed`` exactly as the Fortran/C shim of INTEGRATION.md walks the PETSc objects.  Same error
behaviour as PETSc where it matters for the path: apply on a PC that was never set up triggers
setUp (PCApply calls PCSetUp), a wrong-sized vector is an error, and a one-level PCAIR refuses to
run (the reference swaps in PCJACOBI there, src/AIR_MG_Setup.F90:1167-1174 -- outside this path).
"""
import numpy as np

from .device import DeviceAIR, INV_AFF
from ._capi import PflareB200Error
from .upload import feed

PCAIR = "air"
PCPFLAREINV = "pflareinv"


class PC:
    def __init__(self, rank=0, nranks=1, unique_id=None, device=0):
        self._type = None
        self._H = None
        self._dev = None
        self._setup = False
        self._comm = (rank, nranks, unique_id, device)
        self.options = {}

    # PCSetType
    def setType(self, pc_type):
        if pc_type not in (PCAIR, PCPFLAREINV):
            raise ValueError("Unknown PC type %r (PCRegister_PFLARE registers 'air' and 'pflareinv')" % (pc_type,))
        self.reset()
        self._type = pc_type
        return self

    def getType(self):
        return self._type

    def setHierarchy(self, H):
        """Hand over the operators the reference's PCSetUp produced."""
        self.reset()
        self._H = H
        return self

    def setOption(self, key, value):
        self.options[key] = value
        if self._dev is not None:
            self._dev.set_option(key, value)

    # PCSetUp
    def setUp(self):
        if self._type is None:
            raise RuntimeError("PCSetType must be called before PCSetUp")
        if self._H is None:
            raise RuntimeError("PCSetUp: no operators set")
        if self._setup:
            return self
        H = self._H
        rank, nranks, uid, device = self._comm
        if self._type == PCAIR and H.no_levels < 2:
            raise PflareB200Error(6, "PCAIR with a single level: the reference falls back to PCJACOBI "
                                     "(src/AIR_MG_Setup.F90:1167-1174); not part of the accelerated path")
        self._dev = DeviceAIR(H.no_levels, rank=rank, nranks=nranks, unique_id=uid, device=device)
        for k, v in self.options.items():
            self._dev.set_option(k, v)
        feed_fn = getattr(H, "feed", None)
        if feed_fn is not None:
            feed_fn(self._dev)
        else:
            feed(H, self._dev)
        if hasattr(H, "local_rows"):
            self._n = H.local_rows()
        else:      # rows of level 1 (a container read from disk carries no separate copy of the system matrix)
            self._n = H.levels[0].n if H.levels else H.coarse_matrix.shape[0]
        self._setup = True
        return self

    # PCApply
    def apply(self, x, y=None):
        if not self._setup:
            self.setUp()
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.ndim != 1 or x.size != self._n:
            raise ValueError("PCApply: vector has %d local entries, the PC has %d rows" % (x.size, self._n))
        if self._type == PCAIR:
            out = self._dev.apply(x)
        else:
            out = self._dev.inv_apply(1, INV_AFF, x)
        if y is not None:
            y[...] = out
            return y
        return out

    def device(self):
        if not self._setup:
            self.setUp()
        return self._dev

    # PCView: the complexity numbers of print_stats (src/AIR_MG_Stats.F90:256-416) that the apply determines
    def view(self):
        d = self.device()
        s = d.stats()
        H = self._H
        lines = ["PC Object: type %s" % self._type, "  levels: %d" % H.no_levels]
        if hasattr(H, "sizes"):
            lines.append("  rows per level: %s" % (H.sizes(),))
        lines.append("  nnz traversed per V-cycle: %d" % int(s["nnz_per_cycle"]))
        lines.append("  algorithmic HBM bytes per V-cycle: %d" % int(s["algorithmic_bytes"]))
        lines.append("  kernel launches per V-cycle: %d (tail levels %d)" % (int(s["kernel_launches"]), int(s["tail_levels"])))
        return "\n".join(lines)

    # PCReset / PCDestroy
    def reset(self):
        if self._dev is not None:
            self._dev.close()
        self._dev = None
        self._setup = False

    def destroy(self):
        self.reset()
        self._H = None
        self._type = None
