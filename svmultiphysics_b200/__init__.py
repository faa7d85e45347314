"""svmultiphysics_b200 — B200-native element assembly + FSILS solve behind a C ABI (include/svb200.h)."""
__version__ = "0.1.0"
