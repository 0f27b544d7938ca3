"""csbsr_b200 -- B200-native (sm_100a) implementation of the CSBSR degrade -> blind-SR -> segment -> AIU/AHD95 path."""
__version__ = "0.1.0"
