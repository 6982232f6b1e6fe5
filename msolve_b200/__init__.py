"""msolve_b200 -- B200-native GF(p) linear algebra for msolve/neogb's F4."""
__version__ = "0.1.0"
