"""surround360_b200 -- B200-native stereo-panorama render path (Surround360 drop-in)."""
