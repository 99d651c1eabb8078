"""B200-native Marigold-DC guided denoising loop (drop-in for tier4/depth_completion's hot path)."""
__version__ = "0.1.0"


def __getattr__(name):  # lazy: importing the package must not require torch / CUDA
    if name == "MarigoldDepthCompletionPipeline":
        from .pipeline import MarigoldDepthCompletionPipeline

        return MarigoldDepthCompletionPipeline
    raise AttributeError(name)
