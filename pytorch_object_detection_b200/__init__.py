"""B200-native FCOS post-processing / target assignment / loss hot path (sm_100a CUDA behind a C ABI).

Drop-in names of the reference's ``model/modules/head.py`` and ``model/loss.py``:

    from pytorch_object_detection_b200 import FCOSHead, ClipBoxes, FCOSGenTargets, FCOSLoss

Importing the package loads ``libb200det.so`` (building it with nvcc when absent); there is no
CPU or other-backend fallback.
"""
from . import _lib

_lib.load()      # fail loudly at import when the CUDA library is missing and cannot be built

from . import ops  # noqa: E402  (registers the torch.library ops of the b200det namespace)
from .data import DeviceCollate, collate_images, pack_gt  # noqa: E402
from .eval import coco_results, eval_ap_2d, eval_ap_batched, sort_by_score  # noqa: E402
from .head import ClipBoxes, FCOSGenTargets, FCOSHead  # noqa: E402
from .loss import (FCOSLoss, FCOSTargetLoss, compute_cls_loss, compute_cnt_loss, compute_reg_loss,  # noqa: E402
                   focal_loss_from_logits, giou_loss, iou_loss)

__all__ = ["FCOSHead", "ClipBoxes", "FCOSGenTargets", "FCOSLoss", "FCOSTargetLoss", "compute_cls_loss", "compute_cnt_loss",
           "compute_reg_loss", "iou_loss", "giou_loss", "focal_loss_from_logits", "DeviceCollate", "collate_images", "pack_gt", "eval_ap_2d", "eval_ap_batched", "coco_results",
           "sort_by_score", "ops"]
