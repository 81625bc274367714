"""Drop-in service classes: same names, signatures, return types and error
behaviour as the reference's app/services modules for the image hot path."""
from .color_extractor import ColorExtractor, ColorPalette  # noqa: F401
from .font_detector import FontDetector  # noqa: F401
from .image_processor import ImageProcessor  # noqa: F401
from .image_transformer import ImageTransformer  # noqa: F401
from .models import ColorFeatures  # noqa: F401
from .shadow_analyzer import ShadowAnalyzer  # noqa: F401
from .shape_analyzer import ShapeAnalyzer  # noqa: F401
from .text_extractor import TextExtractor  # noqa: F401
from .utils import PreprocessingMode, validate_and_preprocess_image  # noqa: F401
