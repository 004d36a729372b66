"""auncel_b200: B200-native error-bounded IVF-Flat query path."""
from .index import (Error_sys, FaissException, IndexIVFFlat, METRIC_INNER_PRODUCT, METRIC_L2,  # noqa: F401
                    merge_tables)
