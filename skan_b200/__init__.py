"""skan_b200 -- B200-native (sm_100a) implementation of jni/skan's skeleton-to-graph hot path.

Drop-in for the reference's ``skan.Skeleton`` / ``skan.summarize`` / ``skan.csr`` API on that
path; see DESIGN.md for the scope and INTEGRATION.md for the C-ABI boundary.
"""
from . import csr  # noqa: F401
from .csr import PathGraph, Skeleton, sholl_analysis, skeleton_to_csgraph, summarize  # noqa: F401
from . import image_stats  # noqa: F401,E402

__version__ = '0.1.0'
__all__ = ['Skeleton', 'summarize', 'sholl_analysis', 'PathGraph', 'skeleton_to_csgraph', 'csr', 'image_stats']
