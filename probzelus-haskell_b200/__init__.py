"""probzelus-haskell_b200: B200-native streaming particle-filter step of ProbZelus-haskell.

Host-side mirror of the reference's inference interface (haskell/src/Inference.hs:101-114,
haskell/src/Util/ZStream.hs:16,65-69) over the C ABI of include/pzgpu.h.  The directory
name carries a hyphen (it is the repository's name); load it with
`__graft_entry__.load_package()` or importlib, which registers it as the module
`probzelus_haskell_b200`.

There is no CPU fallback: importing works without a GPU (so the ABI can be inspected), but
every call that computes raises PzError when libpzgpu.so or a CUDA device is missing.
"""
from .infer import (  # noqa: F401
    PzError, ModelDesc, StepSummary, Posterior, ZStream, lib, lib_path, exported_symbols,
    rw1d, stt, mtt_track, mtt, zparticles, zdsparticles, infer, run_stream,
    resample_systematic, resample_multinomial, comm_unique_id, zparticles_sharded,
    PZ_RESAMPLE_SYSTEMATIC, PZ_RESAMPLE_MULTINOMIAL_REF, json_line_tracks, json_line_generic1d, json_line_mcam,
    walker, beta_bernoulli, zparticles_multi, importance_resampling, dist_sample, dist_score, WalkerParams,
    PZ_MODEL_RW1D, PZ_MODEL_LINGAUSS, PZ_MODEL_MTT, PZ_MODEL_WALKER, PZ_MODEL_BETA_BERNOULLI, PZ_MODEL_MULTICAM, multicam,
)
