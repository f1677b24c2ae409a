"""sngnn_b200 -- B200-native (sm_100a) implementation of SNGNN's similarity-navigated aggregation path.

Drop-in surface (same names / ctor signatures as R: models/models.py:35-334):
    SNGNN, SNGNN_Plus, SNGNN_Plus_Plus, SNConv, SNConv_plus, SNConv_plus_plus
plus `simknn` (all-pairs similarity-kNN builder) and `toolbox` (Sim-GFA metric names).
All compute goes through the C-ABI library `libsng.so` (include/sng.h); there is no CPU fallback.
"""
__version__ = "0.1.0"

_LAZY = {
    "SNGNN": "models", "SNGNN_Plus": "models", "SNGNN_Plus_Plus": "models",
    "SNConv": "models", "SNConv_plus": "models", "SNConv_plus_plus": "models",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        return getattr(importlib.import_module(f"{__name__}.{_LAZY[name]}"), name)
    raise AttributeError(name)
