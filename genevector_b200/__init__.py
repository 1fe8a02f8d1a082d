"""genevector_b200 -- B200-native all-pairs gene co-expression engine (GeneVector mi_backend)."""
__version__ = "0.1.0"
