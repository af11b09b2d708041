"""bfa_b200 -- importable name of the package that lives in `bournemouth-forced-aligner_b200/`
(the directory name required by the repo layout is not a valid Python identifier)."""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "bournemouth-forced-aligner_b200")
__path__.append(_impl)

from ._cabi import BfaError, BfaParams, BfaShape, default_params  # noqa: E402,F401
from .aligner import (AlignmentUtils, BatchResult, ViterbiDecoder, _calculate_confidences,  # noqa: E402,F401
                      align_host, extend_soft_boundaries_func, _calculate_confidences_batch)

from .postprocess import (CoverageError, align_words, analyze_alignment_coverage, convert_to_ms,  # noqa: E402,F401
                          ensure_target_coverage, post_process_segment, stamps_to_ms)
from .pipeline import PhonemeTimestampAligner, log_softmax_rows, stitch_log_softmax  # noqa: E402,F401

__all__ = ["PhonemeTimestampAligner", "stitch_log_softmax", "log_softmax_rows", "AlignmentUtils", "ViterbiDecoder", "_calculate_confidences", "align_host", "BatchResult", "convert_to_ms", "stamps_to_ms", "extend_soft_boundaries_func", "_calculate_confidences_batch", "ensure_target_coverage", "CoverageError", "align_words", "analyze_alignment_coverage", "post_process_segment",
           "BfaError", "BfaParams", "BfaShape", "default_params"]
