"""hwer_b200 -- B200-native serving hot path of the Hybrid-Weighted-Embedding-Recommender (`hwer`).

Importable as `hwer_b200` (see /hwer_b200.py at the repo root; the directory name carries a hyphen).
The public names mirror `hwer/__init__.py`'s for the part of the reference this package replaces.
"""
from .recommendation_base import Edge, MultiKNN, Node, RecommendationBase  # noqa: F401
from .recommenders import ContentRecommendation, GcnNCF  # noqa: F401
from .utils import NodeNotFoundException, unit_length, unit_length_violations  # noqa: F401
from . import utils  # noqa: F401
from . import ops, sharded, table_io, validation  # noqa: F401

__version__ = "0.1.0"
