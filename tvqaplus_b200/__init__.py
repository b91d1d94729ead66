"""B200-native STAGE forward hot path (drop-in for jayleicn/TVQAplus ``model/``)."""
__version__ = "0.1.0"
