"""hiergen -- synthetic AIR hierarchy generator (INPUT GENERATOR, not the product).

The reference (PFLARE) builds the multigrid hierarchy in its own Fortran/PETSc
``PCSetUp`` (``src/AIR_MG_Setup.F90:44-1231``) and only the *apply* is
re-implemented in this repository.  Neither gfortran nor PETSc exist in the
build container or on the GPU box, so tests and ``bench.py`` need a stand-in
that manufactures hierarchies with the same structure (CF splitting, A_ff,
A_fc, approximate inverses, R=[Z I], P=[W;I], coarse operators).  This package
is that stand-in: a numpy/scipy (+ small OpenMP C++ helper) restatement of the
reference's setup.  It is not bit-reproducible against a gfortran build
(the reference seeds the compiler RNG) and is never
on the measured path: it only produces the operators that both the CPU oracle
and the CUDA path consume.
"""
from .problems import adv_1d, adv_diff_fd, dg_upwind_surrogate, read_petsc_binary, parilu_factors  # noqa: F401
from .setup import AirOptions, Hierarchy, Level, Inverse, build_hierarchy, build_pflareinv  # noqa: F401
from pflare_b200.upload import feed  # noqa: F401  (the upload-hook walker lives with the product's boundary: oracle and CUDA path are fed by the same code)
from .partition import partition, split_ownership, scatter_vector, LocalHierarchy  # noqa: F401
