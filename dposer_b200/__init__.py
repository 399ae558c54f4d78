"""dposer_b200 -- B200-native (sm_100a) implementation of DPoser's data-parallel hot path.

Host code mirrors the reference's call surface (module names follow the reference:
``model.ScoreModelFC``, ``sde_lib.subVPSDE``, ``utils.get_score_fn``,
``sampling.get_sampling_fn``, ``body_model.BodyModel`` ...); the arithmetic runs in
hand-written CUDA kernels behind the C ABI in ``include/dposer_b200.h``.
There is no CPU fallback.
"""
__version__ = '0.1.0'
