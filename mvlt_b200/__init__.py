"""mvlt_b200: B200-native (sm_100a) implementation of the MVLT / PVLT data-parallel hot path."""
__version__ = "0.1.0"
