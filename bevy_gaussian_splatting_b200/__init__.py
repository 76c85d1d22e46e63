"""bevy_gaussian_splatting_b200 -- B200-native forward splat path behind the reference's plugin surface.

Product = `csrc/` (hand-written sm_100a kernels + the C ABI of include/bgs.h, built to libbgs.so).
The Python modules here are the host-side mirror of the reference interface for this path
(same names as mosure/bevy_gaussian_splatting: src/lib.rs:7-29 re-exports) used by tests and bench.
"""
from .abi import BgsError  # noqa: F401
from .camera import GaussianCamera, View, headless_view, orbit_view, perspective_view  # noqa: F401
from .gaussian import (PlanarGaussian3d, random_gaussians_3d, random_gaussians_3d_seeded,  # noqa: F401
                       SH_COEFF_COUNT)
from .io import load_cloud, parse_ply_3d  # noqa: F401
from .gcloud import decode_gcloud, encode_gcloud, read_gcloud, write_gcloud  # noqa: F401
from .plugin import CloudTransform, GaussianSplattingPlugin, PlanarGaussian3dHandle  # noqa: F401
from .settings import (CloudSettings, DrawMode, GaussianColorSpace, GaussianMode, RadixSortDepthBits,  # noqa: F401
                       RasterizeMode, ShaderDefines, SortMode)
