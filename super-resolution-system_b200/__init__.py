"""super-resolution-system_b200 -- B200-native (sm_100a) drop-in for the local hot path of
squellww/super-resolution-system: overlap tiling, Laplacian-pyramid tile fusion, PSNR / Gaussian SSIM.

The modules keep the reference's names, so putting this directory on sys.path makes
`from tiling_module import TilingModule`, `from blending_module import BlendingModule, TileInfo`,
`from quality_assessment_module import QualityAssessmentModule` and
`from main import PipelineConfig, SuperResolutionPipeline` resolve to the GPU implementation.
All compute goes through libsrfuse.so (csrc/, C ABI in include/srfuse.h); there is no CPU fallback.
"""
from . import _native, build, distributed  # noqa: F401
from .tiling_module import TilingModule, Tile, TileMetadata, PaddingMode, TileStatus  # noqa: F401
from .blending_module import BlendingModule, TileInfo, WeightType, FusionMethod  # noqa: F401
from .quality_assessment_module import QualityAssessmentModule, QualityThresholds, ScaleConfig  # noqa: F401
from .config import TilingConfig, BlendingConfig, QualityAssessmentConfig  # noqa: F401
from .main import PipelineConfig, PipelineResult, SuperResolutionPipeline  # noqa: F401

__version__ = "0.1.0"
