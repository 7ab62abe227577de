"""cloudy_b200 — B200-native batched evaluation of Cloudy.jl's coalescence / sedimentation tendencies.

Host-side mirror of the reference's Julia API for the hot path (names and argument meaning follow
CliMA/Cloudy.jl v0.6.0); the arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of
``libcloudy_b200.so`` (include/cloudy_b200.h).  There is no CPU fallback."""
from . import _lib
from ._lib import (EXPONENTIAL, GAMMA, LOGNORMAL, MONODISPERSE, MODEL_BOX, MODEL_RAINSHAFT, CloudyError)
from .context import Context, default_context
from .helpers import (get_dist_moment_ind, get_dist_moments_ind_range, get_moments_normalizing_factors, rflatten)
from .distributions import (PrimitiveParticleDistribution, ExponentialPrimitiveParticleDistribution,
                            GammaPrimitiveParticleDistribution, LognormalPrimitiveParticleDistribution,
                            MonodispersePrimitiveParticleDistribution, moment, get_moments, density, nparams,
                            update_dist_from_moments, moment_source_helper, integrate_SimpsonEvenFast,
                            compute_threshold, compute_thresholds, normed_density)
from .kernel_tensors import (CoalescenceTensor, get_normalized_kernel_tensor, check_symmetry, polyfit,
                             ConstantKernelFunction, LinearKernelFunction, HydrodynamicKernelFunction,
                             LongKernelFunction, get_normalized_kernel_func)
from .coalescence import (CoalescenceData, get_coal_ints, AnalyticalCoalStyle, NumericalCoalStyle, FixedThreshold,
                          MovingThreshold, log_grid, build_config)
from .sedimentation import get_sedimentation_flux, get_cond_evap, get_standard_N_q
from .ensemble import (ParcelEnsemble, CoalescenceModel, ModelParameters, make_box_model_rhs, make_rainshaft_rhs)
