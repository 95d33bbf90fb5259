import torch
import torch.nn.functional as F


class ImageList(object):
    """padded batch + per-image (h, w) (detectron2.structures.ImageList semantics)."""

    def __init__(self, tensor, image_sizes):
        self.tensor = tensor
        self.image_sizes = image_sizes

    def __len__(self):
        return len(self.image_sizes)

    def __getitem__(self, idx):
        size = self.image_sizes[idx]
        return self.tensor[idx, ..., : size[0], : size[1]]

    @property
    def device(self):
        return self.tensor.device

    @staticmethod
    def from_tensors(tensors, size_divisibility=0, pad_value=0.0):
        assert len(tensors) > 0
        image_sizes = [(im.shape[-2], im.shape[-1]) for im in tensors]
        max_h = max(s[0] for s in image_sizes)
        max_w = max(s[1] for s in image_sizes)
        if size_divisibility > 1:
            stride = size_divisibility
            max_h = (max_h + (stride - 1)) // stride * stride
            max_w = (max_w + (stride - 1)) // stride * stride
        if len(tensors) == 1:
            h, w = image_sizes[0]
            batched = F.pad(tensors[0], [0, max_w - w, 0, max_h - h], value=pad_value).unsqueeze_(0)
        else:
            batch_shape = [len(tensors)] + list(tensors[0].shape[:-2]) + [max_h, max_w]
            batched = tensors[0].new_full(batch_shape, pad_value)
            for img, pad_img in zip(tensors, batched):
                pad_img[..., : img.shape[-2], : img.shape[-1]].copy_(img)
        return ImageList(batched.contiguous(), image_sizes)
