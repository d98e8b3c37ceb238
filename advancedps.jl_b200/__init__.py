"""B200-native SMC / particle-MCMC inner loop behind the AdvancedPS.jl sampler surface.

Layout: ``csrc/`` holds the sm_100a CUDA kernels and the C-ABI shim (``libaps_b200.so``,
declared in ``include/aps_b200.h``); the Python modules are the host-side mirror of the
reference's sampler interface (src/smc.jl, src/container.jl, src/resampling.jl) over that ABI.
"""
from . import _abi, models  # noqa: F401
from ._abi import (  # noqa: F401
    RESAMPLE_MULTINOMIAL, RESAMPLE_RESIDUAL, RESAMPLE_STRATIFIED, RESAMPLE_SYSTEMATIC,
    SAMPLER_PG, SAMPLER_PGAS, SAMPLER_SMC,
)
from .sampler import (  # noqa: F401,E402
    DEFAULT_RESAMPLER, PG, PGAS, SMC, ApsError, DeviceParticleContainer, ParticleContainer, PGSample, PGState, ResampleWithESSThreshold,
    SMCSample, Trace, TracedSSM, effectiveSampleSize, getweight, getweights, increase_logweight_, logZ, randcat,
    resample_multinomial, resample_propagate_, resample_residual, resample_stratified, resample_systematic,
    reset_logweights_, reweight_, sample, step, sweep_,
)
