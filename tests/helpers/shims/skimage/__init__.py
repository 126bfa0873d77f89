from . import io, transform   # dataset/*.py: `from skimage import io, transform` (unused)
