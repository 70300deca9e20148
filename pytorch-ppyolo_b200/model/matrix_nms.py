"""Matrix-NMS entry points (reference ``model/matrix_nms.py``: ``jaccard`` :33-47, ``matrix_nms`` :102-151).

``matrix_nms`` keeps the reference signature and return convention (``[M,6]`` rows
``[label, score, x0, y0, x1, y1]`` sorted by decayed score, or one row of ``-1`` when nothing
survives) but runs the threshold -> top-k -> IoU/decay -> top-k pipeline in the batched CUDA
kernels of ``csrc/nms.cu``; ties in score are broken by (box, class) index, i.e. like a stable sort.
"""
from ppyolo_b200 import ops


def jaccard(box_a, box_b):
    """Pairwise IoU of two xyxy box sets, [A,4] x [B,4] -> [A,B]."""
    return ops.pairwise_iou(box_a, box_b)


def matrix_nms(bboxes, scores, score_threshold, post_threshold, nms_top_k, keep_top_k, use_gaussian=False,
               gaussian_sigma=2.):
    """Single-image Matrix-NMS: ``bboxes`` [B,4], ``scores`` [B,C]."""
    return ops.matrix_nms_batched(bboxes[None], scores[None], score_threshold=score_threshold,
                                  post_threshold=post_threshold, nms_top_k=nms_top_k, keep_top_k=keep_top_k,
                                  use_gaussian=use_gaussian, gaussian_sigma=gaussian_sigma)[0]
