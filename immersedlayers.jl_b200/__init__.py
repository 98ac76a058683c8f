"""ilm-b200: B200-native immersed-layer operator hot path.

Host-side mirror of the ImmersedLayers.jl operator API on `BasicILMCache`
(regularize!/interpolate!, staggered stencils, LGF inverse Laplacian, Schur
builders, surface-point solve) over the C ABI of libilm_b200.so
(include/ilm_b200.h).  Import as `import ilm_b200` (see ../ilm_b200.py).
"""
from . import _lib, bodies, lgf  # noqa: F401
from . import timemarching  # noqa: F401,E402
from . import forcing  # noqa: F401,E402
from . import tools  # noqa: F401,E402
from . import helmholtz  # noqa: F401,E402
from .helmholtz import (  # noqa: F401,E402
    ScalarPotentialCache, VectorFieldCache, VectorPotentialCache, curlv_masked_from_masked_curlv, divv_masked_from_masked_divv,
    masked_curlv_from_curlv_masked, masked_divv_from_divv_masked, potentials_from_masked_fields, scalarpotential_from_divv,
    scalarpotential_from_masked_divv, scalarpotential_uniformvecfield, vecfield_from_scalarpotential, vecfield_from_vectorpotential,
    vecfield_helmholtz, vecfield_uniformvecfield, vectorpotential_from_curlv, vectorpotential_from_masked_curlv,
    vectorpotential_uniformvecfield)
from ._lib import DimensionMismatch, IlmError, MethodError  # noqa: F401
from .forcing import (  # noqa: F401,E402
    AreaForcingModel, AreaRegionCache, ForcingModelAndRegion, LineForcingModel, LineRegionCache, PointForcingModel,
    PointRegionCache, RigidTransform, SpatialGaussian, apply_forcing)
from .api import (  # noqa: F401
    EdgeGradient, SurfaceVectorCache, TensorData, create_CL2invCT, create_RTHR, create_RTHR_direct, convolve, stokes_flow, convective_derivative, w_cross_v, create_RTLinvR_direct, neumann_poisson, create_CLinvCT_scalar, create_GLinvD_symm,
    normal_dot_interpolate, normal_interpolate_symm, regularize_normal_dot, regularize_normal_symm,
    surface_divergence_symm, surface_grad_symm,
    XEdges, YEdges, Dual, Edges, GridScaling, IndexScaling, LU, Nodes, PhysicalGrid, Primal, ScalarData,
    SurfaceScalarCache, VectorData, complementary_mask, create_CLinvCT, create_GLinvD,
    create_GLinvD_cross, create_RTLinvR, create_nRTRn, create_surface_filter, curl,
    dirichlet_poisson, dirichlet_solve, create_schur_sharded, divergence, grad, interpolate, inverse_laplacian, laplacian, mask,
    matvec_pow, normal_cross_interpolate, normal_interpolate, regularize, regularize_normal,
    regularize_normal_cross, surface_curl, surface_curl_cross, surface_divergence,
    surface_divergence_cross, surface_grad, surface_grad_cross)
