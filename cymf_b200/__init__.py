"""cymf_b200 -- B200-native factor-update hot path of CyMF (placeholder; filled in below)."""
