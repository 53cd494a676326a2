__version__ = "0.1.2"  # tracks the reference package version (rasterizer/version.py:1)
