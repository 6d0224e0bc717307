"""patchmixturekriging_b200 -- B200-native (sm_100a) local-GP fit + mixture-query hot path of
RoyCCWang/PatchMixtureKriging behind the reference's own function surface
(reference src/PatchMixtureKriging.jl:54-71).  Python host mirror over the C ABI in include/pmk.h;
julia/PatchMixtureKrigingB200.jl is the same layer written for the reference's own language."""
from .kernels import (BrownianBridge10, BrownianBridge1eps, BrownianBridge1ϵ, BrownianBridge20, BrownianBridge2eps,
                      BrownianBridge2ϵ, GaussianKernel1DType, RationalQuadraticKernelType, Spline12KernelType,
                      Spline32KernelType, Spline34KernelType)
from .partition import (BSPTree, fetchhyperplanes, findpartition, gethyperplane, organizetrainingsets,
                        organizetrainingsets_device, setuppartition, setuppartition_device)
from ._lib import Handle, PMKError, PosDefException, LIB_PATH
from .mixturegp import (MixtureGPDebugType, MixtureGPType, fitmixtureGP_, loadmixtureGP, querymixtureGP, querymixtureGP_,
                        savemixtureGP)
from .rkhs import RKHSProblemType, constructkernelmatrix, evalkernel, evalquery, fitRKHS_, query_, setupGPquery

__all__ = [
    "BrownianBridge10", "BrownianBridge20", "BrownianBridge1ϵ", "BrownianBridge2ϵ", "BrownianBridge1eps", "BrownianBridge2eps",
    "GaussianKernel1DType", "Spline34KernelType", "Spline12KernelType", "Spline32KernelType", "RationalQuadraticKernelType",
    "BSPTree", "setuppartition", "setuppartition_device", "organizetrainingsets", "organizetrainingsets_device", "fetchhyperplanes", "findpartition", "gethyperplane",
    "MixtureGPType", "MixtureGPDebugType", "fitmixtureGP_", "querymixtureGP", "querymixtureGP_", "savemixtureGP", "loadmixtureGP",
    "RKHSProblemType", "fitRKHS_", "query_", "constructkernelmatrix", "evalkernel", "evalquery", "setupGPquery",
    "Handle", "PMKError", "PosDefException", "LIB_PATH",
]
