"""clive2_b200 — B200-native bidirectional path tracing hot path behind Clive2's Python surface."""
__version__ = "0.1.0"
