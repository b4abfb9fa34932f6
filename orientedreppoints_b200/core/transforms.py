"""Mirror of rbbox2result (mmdet/core/bbox/transforms.py:356-375)."""
import numpy as np


def rbbox2result(bboxes, labels, num_classes):
    """-> list of (num_classes-1) float32 arrays [k_c, 27]; empty detections give zeros((0, 9)) like the reference."""
    if bboxes.shape[0] == 0:
        return [np.zeros((0, 9), dtype=np.float32) for _ in range(num_classes - 1)]
    bboxes = bboxes.cpu().numpy()
    labels = labels.cpu().numpy()
    return [bboxes[labels == i, :] for i in range(num_classes - 1)]
