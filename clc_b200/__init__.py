"""clc_b200 -- B200-native (sm_100a) implementation of the CLC conditional-latent hot path:
reference matching (Pearson correlation + top-k gather), CLM fusion, and the ChARM entropy
stage (GaussianConditional / EntropyBottleneck / bpp), forward and backward, behind the
reference's own nn.Module / operator API.  See DESIGN.md and include/clc_b200.h."""
from ._lib import lib as _load_library  # noqa: F401  (loading is lazy; ops fail loudly if missing)
from .entropy_models import EntropyBottleneck, GaussianConditional
from .loss import RateDistortionLoss, compute_bpp, compute_psnr
from .clm import CLM, DeformableAlignment, SimpleCLM, clm_fuse
from .matching import (L2_or_pearson_corr, SI_Finder_at_Decoder_Feature_Domain, SI_Wraper,
                       create_gaussian_masks, match_and_gather, match_topk, topk_rows)

__all__ = [
    "EntropyBottleneck", "GaussianConditional", "RateDistortionLoss", "compute_bpp", "compute_psnr",
    "SimpleCLM", "CLM", "DeformableAlignment", "clm_fuse", "L2_or_pearson_corr", "SI_Finder_at_Decoder_Feature_Domain", "SI_Wraper",
    "create_gaussian_masks", "match_and_gather", "match_topk", "topk_rows",
]
from . import ans, ops, retrieval  # noqa: E402,F401  (range coder behind compress()/decompress(); raw op wrappers)
