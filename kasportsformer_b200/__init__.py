"""kasportsformer_b200 -- B200-native (sm_100a) KASportsFormer inference forward + MPJPE reduction.

Public surface mirrors the reference's: `KASportsFormer`, `load_model`, `yaml_config_reader`,
`total_parameters_count`; plus `evaluate` (GPU evaluation epilogue), `clipstore` (packed clip shards + pinned
double-buffered feeder), `serving` (variable-length video front end) and `synthetic` helpers.
"""
from .model import KASportsFormer, load_model, yaml_config_reader, total_parameters_count, AttrDict

__all__ = ["KASportsFormer", "load_model", "yaml_config_reader", "total_parameters_count", "AttrDict"]
