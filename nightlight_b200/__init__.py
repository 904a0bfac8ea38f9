"""nightlight_b200 -- B200 (sm_100a) implementation of mlnoga/nightlight's stacking hot path.

The product is ``libnightlight_cuda.so`` (hand-written CUDA behind the C ABI of
``include/nightlight_cuda.h``).  This package is the thin ctypes binding plus host-side mirrors of
the reference's operators for this path -- ``OpStack`` (internal/ops/stack/stack.go:66-227),
``project`` (internal/fits/project.go:26-76) and ``find_stars`` (internal/star/findstars.go:59-100)
-- with the same field names, argument meaning and error behaviour, so the parity tests read like
calls into the reference.  There is no CPU fallback: importing works anywhere, but every compute
call needs the built library and a CUDA device and raises otherwise.
"""
from .binding import (  # noqa: F401
    NightlightError, load_library, library_path, Context, StackJob, Star, STAR_DTYPE,
    ST_MEDIAN, ST_MEAN, ST_SIGMA, ST_WINSOR_SIGMA, ST_MAD_SIGMA, ST_LINEAR_FIT, ST_AUTO,
    W_NONE, W_EXPOSURE, W_INVERSE_NOISE, W_INVERSE_HFR, DECLARED_SYMBOLS, NUMERICS_AMD64, NUMERICS_PUREGO,
)
from .ops import (OpStack, OpStackBatches, project, transform_invert, find_bright_pixels, find_stars, get_weights,  # noqa: F401
                  find_sigmas_and_stack, estimate_noise, project_scaled, fits_decode, fits_encode, partition,
                  median_filter3x3, stats, bad_pixel_map, op_bad_pixel)

__version__ = "0.1.0"
