"""B200-native relative-entropy coding engine for the MIRACLE samplers of
gergely-flamich/compression-without-quantization (drop-in for code/coded_importance_sampler.py and
code/coded_greedy_sampler.py).  See DESIGN.md / INTEGRATION.md at the repository root."""
from ._lib import CwqError, LIB_PATH, load as load_library  # noqa: F401
from .binary_io import (elias_delta_code, elias_delta_decode, from_bit_string, read_bin_code, to_bit_string,  # noqa: F401
                        write_bin_code)
from .coded_greedy_sampler import (code_greedy_sample, code_grouped_greedy_sample,  # noqa: F401
                                   code_grouped_greedy_sample_batch, decode_greedy_sample,
                                   decode_grouped_greedy_sample, decode_grouped_greedy_sample_batch)
from .coded_importance_sampler import (code_grouped_importance_sample, code_grouped_importance_sample_batch,  # noqa: F401
                                       code_importance_sample, decode_grouped_importance_sample,
                                       decode_grouped_importance_sample_batch, decode_importance_sample)
from .coding import ArithmeticCoder  # noqa: F401
from .misc import stateless_normal_sample  # noqa: F401
from .pln_codec import build_empirical_dists, code_image_latents, decode_image_latents, latent_permutations  # noqa: F401

__version__ = "0.1.0"
