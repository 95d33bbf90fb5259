"""heatmaps_to_keypoints restated from detectron2 v0.5 structures/keypoints.py [d2-memory] (the source is not
vendored under /root/reference; call site: centermask/modeling/centermask/keypoint_head.py:113).

Per ROI: the K heatmaps [K, S, S] are resized with bicubic interpolation (align_corners=False) to the ROI's own
(ceil(h), ceil(w)) pixels; the arg-max pixel of each map gives the keypoint location (pixel centre, scaled by
size / ceil(size), offset by the box corner); column 2 is the logit there, column 3 the score
exp(logit - max) / sum over the S x S pool-resolution map of exp(map - max)."""
import torch
import torch.nn.functional as F


@torch.no_grad()
def heatmaps_to_keypoints(maps, rois):
    offset_x = rois[:, 0]
    offset_y = rois[:, 1]
    widths = (rois[:, 2] - rois[:, 0]).clamp(min=1)
    heights = (rois[:, 3] - rois[:, 1]).clamp(min=1)
    widths_ceil = widths.ceil()
    heights_ceil = heights.ceil()
    num_rois, num_keypoints = maps.shape[:2]
    xy_preds = maps.new_zeros(rois.shape[0], num_keypoints, 4)
    width_corrections = widths / widths_ceil
    height_corrections = heights / heights_ceil
    keypoints_idx = torch.arange(num_keypoints, device=maps.device)
    for i in range(num_rois):
        outsize = (int(heights_ceil[i]), int(widths_ceil[i]))
        roi_map = F.interpolate(maps[[i]], size=outsize, mode="bicubic", align_corners=False).squeeze(0)
        max_score, _ = roi_map.view(num_keypoints, -1).max(1)
        max_score = max_score.view(num_keypoints, 1, 1)
        tmp_full_resolution = (roi_map - max_score).exp_()
        tmp_pool_resolution = (maps[i] - max_score).exp_()
        roi_map_scores = tmp_full_resolution / tmp_pool_resolution.sum((1, 2), keepdim=True)
        w = roi_map.shape[2]
        pos = roi_map.view(num_keypoints, -1).argmax(1)
        x_int = pos % w
        y_int = (pos - x_int) // w
        x = (x_int.float() + 0.5) * width_corrections[i]
        y = (y_int.float() + 0.5) * height_corrections[i]
        xy_preds[i, :, 0] = x + offset_x[i]
        xy_preds[i, :, 1] = y + offset_y[i]
        xy_preds[i, :, 2] = roi_map[keypoints_idx, y_int, x_int]
        xy_preds[i, :, 3] = roi_map_scores[keypoints_idx, y_int, x_int]
    return xy_preds
