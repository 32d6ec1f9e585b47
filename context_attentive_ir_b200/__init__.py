"""B200-native scoring hot path of wasiahmad/context_attentive_ir."""
from .rankers import ARCI, ARCII, CDSSM, DRMM, DSSM, DUET, ESM, MatchTensor  # noqa: F401
from .multitask import CARS, MNSRF, M_MATCH_TENSOR  # noqa: F401

__all__ = ['ARCI', 'ARCII', 'DSSM', 'CDSSM', 'ESM', 'MatchTensor', 'DRMM', 'DUET', 'CARS', 'MNSRF', 'M_MATCH_TENSOR']
