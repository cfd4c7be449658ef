"""Hot-path-relevant pipeline parameters (defaults of the reference's ``PipelineParams``,
litegs/arguments.py:69-77).  Any object with these attributes is accepted by ``litegs_b200.render``."""
from dataclasses import dataclass


@dataclass
class PipelineParams:
    cluster_size: int = 128
    tile_size: tuple = (8, 16)
    sparse_grad: bool = True
    enable_transmitance: bool = False
    enable_depth: bool = False
    input_color_type: str = "sh"
