"""gadfit_b200 -- B200-native Levenberg-Marquardt inner loop of GADfit's global
multi-dataset fit, behind the reference's own solver interface.

Host side only: ``LMsolver`` mirrors ``gadfit::LMsolver``; all arithmetic is in
``libgadfit_b200.so`` (hand-written sm_100a CUDA, built by
``__graft_entry__.build()``).  There is no CPU fallback.
"""
from .lm_solver import (  # noqa: F401
    CudaError, GADfitException, GadfOptions, InsufficientIntegrationWorkspace,
    InvalidArgument, LateAddDatasetCall, LMsolver, Loss, NcclError,
    NegativeDegreesOfFreedom, NoFittingParameters, NoGlobalParameters,
    NotAvailable, SetParInvalidIndex, SolverSettings, UninitializedParameter,
    UnknownModel, UnknownOperation, comm_unique_id, io, load_models,
    load_test_models, measure_fp64_peak, models, partition,
)
from .user_models import define_model, model_source  # noqa: F401,E402
