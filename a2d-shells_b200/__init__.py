"""a2d-shells_b200 — B200-native assembly of MITC4 director-shell residual, tangent and
geometric stiffness into 6x6 BCSR matrices (the one hot path of a2d-shells / TACS).

The product is the CUDA library `lib/liba2ds_b200.so` behind the C ABI of
`include/a2ds.h`.  This package is only the thin Python face of that ABI (ctypes)
used by the tests and the benchmark driver, plus generators for the synthetic
structured meshes the benchmark runs on.  There is no CPU path: importing works
anywhere, creating an `Assembler` needs the library and a B200.
"""
from . import meshes  # noqa: F401
from .capi import (Assembler, Mesh, Partition, partition_rcb, halo_from_distribute, A2dsError, host_pattern, host_color_elements, host_color_elements_hashed, lib_path, load_library, SCATTER_ATOMIC,  # noqa: F401
                   SCATTER_COLORED, SCATTER_ATOMIC_COLOR_ORDER, STIFFNESS_MATRIX, GEOMETRIC_STIFFNESS_MATRIX, MASS_MATRIX,
                   QUAD4_SHELL, QUAD4_NONLINEAR_SHELL, TRANSFORM_NATURAL, TRANSFORM_REF_AXIS)
from .constitutive import iso_shell_tables, iso_mass_moments  # noqa: F401
