"""smallvcm_b200 — host-side Python face of libsvcm.so, the B200-native replacement for
SmallVCM's VertexCM::RunIteration hot path.

The product is the C-ABI CUDA library (include/svcm.h, smallvcm_b200/csrc/); the reference-side
C++ glue is smallvcm_b200/host/vertexcm_gpu.hxx.  This package is the thin ctypes layer the
tests and bench.py use: it mirrors the reference's renderer interface
(AbstractRenderer / VertexCM, reference src/renderer.hxx:33-70, src/vertexcm.hxx:208-214)
and never computes anything itself — without the compiled library it raises.
"""
from . import abi
from .renderer import (Scene, VertexCM, SvcmError, load_library, library_path, hashgrid_build,
                       intersect_batch, occluded_batch, bsdf_evaluate_batch, bsdf_sample_batch,
                       light_batch, camera_ray_batch,
                       rng_stream, rng_stream_device, image_convert_batch, image_header, comm_unique_id, SCENE_NAMES)

__all__ = ["abi", "Scene", "VertexCM", "SvcmError", "load_library", "library_path",
           "hashgrid_build", "intersect_batch", "occluded_batch", "bsdf_evaluate_batch",
           "bsdf_sample_batch", "light_batch", "camera_ray_batch", "rng_stream", "rng_stream_device", "image_convert_batch", "image_header", "comm_unique_id", "SCENE_NAMES"]
