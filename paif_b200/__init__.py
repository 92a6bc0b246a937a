"""paif_b200 — B200-native (sm_100a) implementation of PAIF's fusion hot path.

``Network_Fusion_Searched`` is a drop-in for the reference class of the same name
(core/model_fusion_auto.py:599-640); ``install()`` rebinds it inside the reference's module so
``test_original.py`` / ``robust_test.py`` pick it up unchanged.
"""
from .genotypes import Genotype, fusion_at
from .fusion import Network_Fusion_Searched, Network_Fusion_Searched_showfeatures

__all__ = ["Network_Fusion_Searched", "Network_Fusion_Searched_showfeatures", "Genotype", "fusion_at", "install"]


def install(module=None):
    """Rebind ``core.model_fusion_auto.Network_Fusion_Searched`` to the B200 drop-in.

    Must run before the reference scripts execute ``from core.model_fusion_auto import ...``
    (test_original.py:716 constructs by imported name; core/model_fusion_auto.py:1039 by module
    global).  Returns the patched module."""
    if module is None:
        import core.model_fusion_auto as module
    module.Network_Fusion_Searched = Network_Fusion_Searched
    if hasattr(module, "Network_Fusion_Searched_showfeatures"):
        module.Network_Fusion_Searched_showfeatures = Network_Fusion_Searched_showfeatures
    return module
