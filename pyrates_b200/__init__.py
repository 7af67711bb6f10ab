"""pyrates_b200 — B200-native (sm_100a) numerical-integration backend for PyRates."""
__version__ = "0.1.0"
