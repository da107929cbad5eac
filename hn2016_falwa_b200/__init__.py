"""hn2016_falwa_b200 — B200-native (sm_100a, fp64) implementation of falwa's QGField hot path."""
__version__ = "0.1.0"
