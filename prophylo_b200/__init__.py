"""prophylo_b200 -- B200-native PPP scan (the one data-parallel hot path of malaybasu/ProPhylo)."""
__version__ = "0.1.0"
