from .rrng import RRNG  # noqa: F401
from . import _native  # noqa: F401
from ._native import Context, FlexgseaError, NonFiniteScoreError, LIB_PATH  # noqa: F401
from .api import (flexgsea, flexgsea_s2n, flexgsea_lm, flexgsea_weighted_ks, flexgsea_maxmean,  # noqa: F401
                  flexgsea_mean, flexgsea_calc_sig, flexgsea_calc_sig_simple, read_gmt,
                  filter_gene_sets, gene_sets_to_csr, ModelMatrix, EList, flexgsea_limma, get_context,
                  get_group, sharded_sig)
