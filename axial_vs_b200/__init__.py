"""axial_vs_b200 -- B200-native axial-trajectory attention (drop-in for Axial-VS / MaXTron's hot path)."""
__version__ = "0.1.0"
