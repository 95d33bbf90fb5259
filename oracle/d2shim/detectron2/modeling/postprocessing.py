"""detector_postprocess restated from detectron2 v0.5 modeling/postprocessing.py."""
import torch
from ..structures import Instances, ROIMasks


def detector_postprocess(results, output_height, output_width, mask_threshold=0.5):
    new_size = (output_height, output_width)
    scale_x, scale_y = (output_width / results.image_size[1], output_height / results.image_size[0])
    results = Instances(new_size, **results.get_fields())
    output_boxes = results.pred_boxes
    output_boxes.scale(scale_x, scale_y)
    output_boxes.clip(results.image_size)
    results = results[output_boxes.nonempty()]
    if results.has("pred_masks"):
        roi_masks = ROIMasks(results.pred_masks[:, 0, :, :])
        results.pred_masks = roi_masks.to_bitmasks(
            results.pred_boxes, output_height, output_width, mask_threshold).tensor
    if results.has("pred_keypoints"):
        results.pred_keypoints[:, :, 0] *= scale_x
        results.pred_keypoints[:, :, 1] *= scale_y
    return results
