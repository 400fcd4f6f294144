"""sbp_b200 -- B200-native hot path of SummationByPartsOperatorsExtra.jl.

Batched application of assembled SBP derivative operators, fused linear-advection RHS with SAT boundary
terms and P-weighted reductions, executed by hand-written CUDA (sm_100a) kernels in `libsbp_b200.so`
behind a C ABI (`include/sbp_b200.h`).  This Python package is the host-side mirror of the reference's
Julia operator interface (the image has no Julia); `julia/SBPB200.jl` is the `ccall` shim a Julia user
loads instead.  There is no CPU fallback: importing works without a GPU (construction is host-side numpy),
but creating a `Context` or applying an operator requires the CUDA library and a device.

The directory name contains a dot, so import it through the repo-root shim:  `import sbp_b200`.
"""
from . import _lib
from ._lib import (ArgumentError, CudaError, DimensionMismatch, NcclError, SbpError, UnsupportedError, build, load)
from .device import Context, DeviceBatch, RawBuffer
from .operators import (DerivativeOperator, GaussLegendre, GaussRadauLeft, GaussRadauRight, LobattoLegendre,
                        MatrixDerivativeOperator, Matrix, MultidimensionalMatrixDerivativeOperator, NodalBasis,
                        OperatorTable, PolynomialBasesDerivativeOperator, SubcellOperator, abs2, couple_subcell,
                        find_corners, grid, identity, integrate, interpolation_matrix, legendre_derivative_operator,
                        mass_matrix, mass_matrix_boundary, mattsson_nordstrom_2004, mul_,
                        neighborhood_sparsity_pattern, polynomialbases_derivative_operator, power,
                        scale_by_inverse_mass_matrix_, scale_by_mass_matrix_, tensor_product_operator_2D,
                        upload_csr, upload_dense, upload_sparse, upload_table)
from .conservation_laws import (AnalysisCallback, LowStorage3SpTableau,
                                MultidimensionalLinearAdvectionNonperiodicSemidiscretization,
                                ODEProblem, VariableLinearAdvectionNonperiodicSemidiscretization,
                                analyze_quantities, quantities, semidiscretize, solve_lowstorage_3sp,
                                solve_ssprk53, step_schedule, tstops)
from .container import load_operator, save_operator
from .objective import ConstructionObjective, get_nsigma, skew_index_map
from .sharding import shard_range

__version__ = "0.1.0"
