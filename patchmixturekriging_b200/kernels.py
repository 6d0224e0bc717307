"""Kernel parameter types, mirroring reference src/misc/declarations.jl:25-100 (same names, same
single scalar field).  Evaluation happens on the GPU (csrc/pmk_common.cuh); these are descriptors."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class _Kernel:
    kernel_id = -1
    stationary = False

    @property
    def params(self) -> np.ndarray:
        return np.array([float(getattr(self, self.__dataclass_fields__.__iter__().__next__()))], dtype=np.float64)


@dataclass(frozen=True)
class GaussianKernel1DType(_Kernel):      # declarations.jl:65-67, kernel.jl:350-357
    ϵ_sq: float
    kernel_id = 0
    stationary = True


@dataclass(frozen=True)
class Spline34KernelType(_Kernel):        # declarations.jl:29-31, kernel.jl:299-313
    a: float
    kernel_id = 1
    stationary = True


@dataclass(frozen=True)
class BrownianBridge10(_Kernel):          # declarations.jl:77-79, kernel.jl:156-158
    a: float = 1.0
    kernel_id = 2


@dataclass(frozen=True)
class BrownianBridge20(_Kernel):          # declarations.jl:81-83, kernel.jl:218-225
    a: float = 1.0
    kernel_id = 3


@dataclass(frozen=True)
class BrownianBridge1ϵ(_Kernel):          # declarations.jl:94-96, kernel.jl:168-174
    ϵ: float
    kernel_id = 4


@dataclass(frozen=True)
class BrownianBridge2ϵ(_Kernel):          # declarations.jl:98-100, kernel.jl:176-193
    ϵ: float
    kernel_id = 5


@dataclass(frozen=True)
class Spline12KernelType(_Kernel):        # kernel.jl:316-330
    a: float
    kernel_id = 6
    stationary = True


@dataclass(frozen=True)
class Spline32KernelType(_Kernel):        # declarations.jl:25-27, kernel.jl:333-347
    a: float
    kernel_id = 7
    stationary = True


@dataclass(frozen=True)
class RationalQuadraticKernelType(_Kernel):   # declarations.jl:33-35, kernel.jl:360-366
    a: float
    kernel_id = 8
    stationary = True


BrownianBridge1eps = BrownianBridge1ϵ
BrownianBridge2eps = BrownianBridge2ϵ
