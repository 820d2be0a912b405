"""research_image_retrieval_b200 — the B200-native retrieval hot path of Mak-GIBA/research_image_retrieval.

pool -> L2 -> whiten -> L2 -> Q·Xᵀ -> top-k -> alpha-QE -> revisited mAP, as PyTorch host code over the C ABI of
librir.so (include/rir.h), which launches hand-written sm_100a kernels.  No Triton, no dispatcher, no CPU fallback.

The module names mirror the reference's (`src/benchmark/utils/{evaluate,helpfunc}.py`, `networks/`), so a caller
switches with an import change — see INTEGRATION.md.
"""
from ._lib import LIB_PATH, RirError, load  # noqa: F401
from .evaluate import (compute_ap, compute_map, compute_map_and_print, compute_map_full, gnd_positions,  # noqa: F401
                       revisited_map, revisited_map_full)
from .helpfunc import extract_database, extract_vectors, extract_vectors_device, scale_mean_l2  # noqa: F401
from .pooling import (DescriptorHead, G2Pooling, GeMPooling, MACPooling, clear_prepared_whitening, gem,  # noqa: F401
                      gem_l2_whiten, gem_pool, l2n, mac_pool, prepare_whitening, spoc, spoc_pool, ultron_gem_pooling,
                      whiten)
from .search import (Database, HostQueryPipeline, ShardedDatabase, alpha_query_expansion, merge_topk, pack_descriptors, rank,  # noqa: F401
                     rerank_topk, search_with_aqe, shard_bounds, sim_topk)

from .formats import DescriptorStore, RoxfordAndRparis, gnd_to_csr  # noqa: F401
from .whitening import ConvDimReduction, pca_covariance, pcawhitenlearn_shrinkage  # noqa: F401

__version__ = "0.1.0"
