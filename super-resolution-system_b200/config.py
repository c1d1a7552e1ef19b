"""Configuration dataclasses with the reference's field names (/root/reference/config.py:33-55 TilingConfig,
134-150 BlendingConfig, 180-201 QualityAssessmentConfig).  The reference file itself does not parse (encoding
damage, SyntaxError at line 85) and nothing imports it; the recoverable fields are re-declared here."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List


@dataclass
class TilingConfig:
    block_size: int = 2048
    output_block_size: int = 4096
    overlap_ratio: float = 0.2
    min_overlap_ratio: float = 0.1
    max_overlap_ratio: float = 0.3
    padding_mode: str = "mirror"  # mirror | replicate | reflect | constant
    enable_content_aware: bool = True
    face_protection_distance: float = 0.5
    cache_dir: str = "./cache"
    enable_l1_cache: bool = True
    enable_l2_cache: bool = True
    l1_cache_size: int = 50

    def tiling_kwargs(self) -> dict:
        """keyword arguments for TilingModule(...)"""
        return dict(block_size=self.block_size, overlap_ratio=self.overlap_ratio, padding_mode=self.padding_mode,
                    output_scale=self.output_block_size / self.block_size, l1_cache_size=self.l1_cache_size,
                    enable_content_aware=self.enable_content_aware)


@dataclass
class BlendingConfig:
    method: str = "laplacian"  # laplacian | poisson | weighted
    num_pyramid_levels: int = 6
    weight_function: str = "cosine"  # linear | cosine | sigmoid
    seam_detection_threshold: float = 0.95
    enable_color_correction: bool = True
    poisson_mode: str = "NORMAL"


@dataclass
class QualityAssessmentConfig:
    device: str = "cpu"
    psnr_threshold: float = 35.0
    ssim_threshold: float = 0.98
    lpips_threshold: float = 0.05
    scale_factors: List[float] = field(default_factory=lambda: [0.1, 0.2, 0.4])
    scale_names: Dict[float, str] = field(default_factory=lambda: {0.1: "structure_color", 0.2: "mid_frequency",
                                                                   0.4: "high_frequency"})
