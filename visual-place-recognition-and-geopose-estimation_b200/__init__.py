"""B200-native (sm_100a) SALAD aggregation head, geopose heads/losses and descriptor kNN: the one
data-parallel hot path of anshium/visual-place-recognition-and-geopose-estimation, behind the
reference's own module interface.  Importable as ``vpr_b200`` (see /vpr_b200.py at the repo root;
the directory name carries the reference's hyphenated name)."""
from ._lib import LIB_PATH, lib  # noqa: F401
from .heads import GeoposeHeads, sincos_to_degrees  # noqa: F401
from .knn import knn_topk, knn_topk_sharded, merge_topk  # noqa: F401
from .parallel import (DataParallelStep, GraphedFinetuneStep, allreduce_grads, bind_to_gpu_numa_node, decorrelate_dropout,  # noqa: F401
                       init_distributed, overlap_grad_allreduce, shard_range)
from .metrics import (LatLonScaler, latlon_metrics, maae, recall_at_k, write_angle_predictions_csv,  # noqa: F401
                      write_latlon_predictions_csv)
from .model import (DINOv2RegressionModel, VPRModel, load_reference_checkpoint, load_upstream_salad_weights,  # noqa: F401
                    save_reference_checkpoint)
from .preprocess import GpuPreprocess  # noqa: F401
from .seam import BackboneSeam  # noqa: F401
from .salad import GraphedSaladStep, SaladHead, dropout_keep_mask, sinkhorn, sinkhorn_backward  # noqa: F401

__all__ = ["SaladHead", "GraphedSaladStep", "GeoposeHeads", "knn_topk", "knn_topk_sharded", "merge_topk", "sinkhorn",
           "sinkhorn_backward", "dropout_keep_mask", "DataParallelStep", "GraphedFinetuneStep", "allreduce_grads", "overlap_grad_allreduce", "init_distributed", "bind_to_gpu_numa_node",
           "shard_range", "sincos_to_degrees", "BackboneSeam", "GpuPreprocess", "VPRModel", "DINOv2RegressionModel", "load_reference_checkpoint",
           "load_upstream_salad_weights", "save_reference_checkpoint", "LatLonScaler", "latlon_metrics", "maae", "recall_at_k",
           "write_latlon_predictions_csv", "write_angle_predictions_csv", "decorrelate_dropout", "lib", "LIB_PATH"]
