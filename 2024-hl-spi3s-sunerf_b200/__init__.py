"""sunerf_b200 - B200-native (sm_100a) implementation of SuNeRF's ray-render hot path.

Drop-in surface (same names, constructor arguments, forward contracts and state_dict keys as the reference):

    sunerf/model/model.py            -> sunerf_b200.model      NeRF, NeRF_DT, EmissionModel, PositionalEncoding, Sine
    sunerf/model/stellar_model.py    -> sunerf_b200.model      SimpleStar
    sunerf/train/sampling.py         -> sunerf_b200.sampling   StratifiedSampler, SphericalSampler, HierarchicalSampler
    sunerf/rendering/*.py            -> sunerf_b200.rendering  SuNeRFRendering, EmissionRadiativeTransfer,
                                                               DensityTemperatureRadiativeTransfer
    sunerf/train/scaling.py          -> sunerf_b200.trainer    ImageAsinhScaling
    Lightning training_step + optim  -> sunerf_b200.trainer    RayTrainer (ray-sharded, one NCCL all-reduce/step)
    sunerf/evaluation/loader.py      -> sunerf_b200.image_render  ObserverRenderer.render_observer_image (rays on device)
    (whole forward / backward chain as one C call each: sunerf_b200.fused.FusedRender over snf_render_fused_{fwd,bwd})

The directory is named after the upstream repo (`2024-hl-spi3s-sunerf_b200`, not a valid Python identifier);
`import sunerf_b200` resolves to it through the small alias package at the repo root.
"""
from . import _lib
from ._lib import SnfError, build
from . import ops
from .model import NeRF, NeRF_DT, EmissionModel, PositionalEncoding, Sine, SimpleStar
from .sampling import StratifiedSampler, SphericalSampler, HierarchicalSampler
from .rendering import SuNeRFRendering, EmissionRadiativeTransfer, DensityTemperatureRadiativeTransfer
from .trainer import RayTrainer, ImageAsinhScaling
from . import rays, parallel, image_render, checkpoint, ray_store
from .ray_store import RayStore
from .image_render import ObserverRenderer
from .fused import FusedRender
from .parallel import ReplicatedRendering

__all__ = ['SnfError', 'build', 'ops', 'NeRF', 'NeRF_DT', 'EmissionModel', 'PositionalEncoding', 'Sine', 'SimpleStar',
           'StratifiedSampler', 'SphericalSampler', 'HierarchicalSampler', 'SuNeRFRendering', 'EmissionRadiativeTransfer',
           'DensityTemperatureRadiativeTransfer', 'RayTrainer', 'ImageAsinhScaling', 'ObserverRenderer', 'RayStore', 'FusedRender',
           'ReplicatedRendering']
