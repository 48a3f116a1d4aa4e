"""B200-native (sm_100a) implementation of DF-Prior's per-ray volume-rendering hot path.

Public surface mirrors the reference's seam for this path:

* ``Renderer``                  -- src/utils/Renderer.py (render_batch_ray, render_img, eval_points, eval_points_tsdf)
* ``DF`` / ``get_model``        -- src/conv_onet/models/decoder.py, src/conv_onet/config.py
* ``Mesher``                    -- the dense point query of src/utils/Mesher.py (eval_points, get_grid_uniform, query loop)
* ``functional``                -- the same operations as plain functions
* ``dist``                      -- ray / voxel-block sharding across GPUs
* ``mapper_ops``                -- the Mapper-side rows next to the path: masked-grid Adam, frustum mask, ray batch
                                   assembly, loss (src/Mapper.py:90-158, 351-388, 408-482)
* ``io``                        -- the reference's on-disk formats: checkpoints, TSDF prior files, pretrained decoders

All arithmetic runs in libdfprior_b200.so (hand-written CUDA behind a C ABI, include/dfprior_b200.h).
There is no CPU fallback: importing is cheap, but any call raises if the library is not built.
"""
from . import _lib, common, functional  # noqa: F401
from .decoder import DF, MLP, DenseLayer, GaussianFourierFeatureTransform, get_model, mlp_tsdf  # noqa: F401
from .mesher import Mesher  # noqa: F401
from .renderer import Renderer  # noqa: F401
from . import dist  # noqa: F401
from . import mapper_ops  # noqa: F401
from . import io  # noqa: F401
from . import graphs  # noqa: F401
from .mapper_ops import MaskedGridAdam  # noqa: F401

__all__ = ["Renderer", "Mesher", "dist", "graphs", "mapper_ops", "MaskedGridAdam", "DF", "MLP", "mlp_tsdf", "DenseLayer", "GaussianFourierFeatureTransform", "get_model",
           "functional", "common"]
