"""prometheus_b200 — B200-native (sm_100a) GS-decoder rendering hot path of
houstonwhy/Prometheus: unproject -> EWA project -> tile bin + radix sort ->
alpha composite (fwd + bwd), behind the reference's own operator interface.

Public surface (mirrors the reference, see DESIGN.md / INTEGRATION.md):
  rasterization(...)            gsplat.rasterization contract
  GaussianConverter, GaussianRenderer   module-level mirrors
  install_as_gsplat()           plug the operator under the unmodified reference
  DecodeRenderPipeline          streaming host->device->host decode+render (copies overlap compute)
  GraphedDecodeRender           the same launch set captured once per shape and replayed as a CUDA graph
  TrajectoryRenderer, sample_from_dense_cameras, export_video   novel-view video along a camera path
  render_loss_terms             mse / disparity / validity mask / entropy term of the stage-1 loss in one pass each way
"""
from .rasterization import rasterization  # noqa: F401
from .renderer import GaussianConverter, GaussianRenderer, install_as_gsplat  # noqa: F401
from .pipeline import DecodeRenderPipeline  # noqa: F401
from .graphed import GraphedDecodeRender  # noqa: F401
from .losses import render_loss_terms  # noqa: F401
from .trajectory import (TrajectoryRenderer, export_ply_for_gaussians, export_video,  # noqa: F401
                         sample_from_dense_cameras)

__all__ = ["rasterization", "GaussianConverter", "GaussianRenderer", "install_as_gsplat", "DecodeRenderPipeline", "GraphedDecodeRender", "render_loss_terms",
           "TrajectoryRenderer", "sample_from_dense_cameras", "export_video", "export_ply_for_gaussians"]
