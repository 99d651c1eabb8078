"""B200-native Marigold-DC guided denoising loop (drop-in for tier4/depth_completion's hot path)."""
__version__ = "0.1.0"
