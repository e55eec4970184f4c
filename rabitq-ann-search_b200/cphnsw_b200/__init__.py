"""cphnsw_b200 -- B200-native query path of CP-HNSW behind the reference's CPIndex API."""
from .index import CPIndex, set_host_module  # noqa: F401

__all__ = ["CPIndex", "set_host_module"]
