"""paddle.vision.transforms. Parity: python/paddle/vision/transforms/{transforms,functional}.py.
Images are numpy HWC arrays, PIL images or CHW tensors."""
from __future__ import annotations

import math
import numbers
import random

import numpy as np
import torch

from ...tensor import Tensor

try:
    from PIL import Image
except Exception:  # pragma: no cover
    Image = None


def _is_pil(img):
    return Image is not None and isinstance(img, Image.Image)


def _to_np(img):
    if _is_pil(img):
        return np.asarray(img)
    if isinstance(img, torch.Tensor):
        a = img.detach().cpu().as_subclass(torch.Tensor).numpy()
        return a.transpose(1, 2, 0) if a.ndim == 3 else a
    return np.asarray(img)


def _like(out, ref):
    if _is_pil(ref):
        return Image.fromarray(out.astype(np.uint8) if out.dtype != np.uint8 else out)
    if isinstance(ref, torch.Tensor):
        a = out.transpose(2, 0, 1) if out.ndim == 3 else out
        return torch.from_numpy(np.ascontiguousarray(a)).as_subclass(Tensor)
    return out


# ------------------------------------------------------------------------------------------------ functional
def to_tensor(pic, data_format="CHW"):
    a = _to_np(pic)
    if a.ndim == 2:
        a = a[:, :, None]
    t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype == torch.uint8:
        t = t.float() / 255.0
    else:
        t = t.float()
    if data_format == "CHW":
        t = t.permute(2, 0, 1)
    return t.contiguous().as_subclass(Tensor)


def _interp(a, size, interpolation="bilinear"):
    t = torch.from_numpy(np.ascontiguousarray(a)).float()
    chw = t.permute(2, 0, 1)[None] if t.dim() == 3 else t[None, None]
    mode = {"nearest": "nearest", "bilinear": "bilinear", "bicubic": "bicubic", "area": "area", "lanczos": "bicubic", "box": "area", "hamming": "bilinear"}[interpolation]
    kw = {"align_corners": False} if mode in ("bilinear", "bicubic") else {}
    out = torch.nn.functional.interpolate(chw, size=size, mode=mode, antialias=mode in ("bilinear", "bicubic"), **kw)[0]
    out = out.permute(1, 2, 0) if t.dim() == 3 else out[0]
    out = out.numpy()
    if a.dtype == np.uint8:
        out = np.clip(np.round(out), 0, 255).astype(np.uint8)
    return out


def resize(img, size, interpolation="bilinear"):
    a = _to_np(img)
    h, w = a.shape[:2]
    if isinstance(size, int):
        if (w <= h and w == size) or (h <= w and h == size):
            return img
        if w < h:
            ow, oh = size, int(size * h / w)
        else:
            oh, ow = size, int(size * w / h)
    else:
        oh, ow = size
    return _like(_interp(a, (oh, ow), interpolation), img)


def crop(img, top, left, height, width):
    a = _to_np(img)
    return _like(a[top:top + height, left:left + width], img)


def center_crop(img, output_size):
    if isinstance(output_size, numbers.Number):
        output_size = (int(output_size), int(output_size))
    a = _to_np(img)
    h, w = a.shape[:2]
    th, tw = output_size
    return crop(img, int(round((h - th) / 2.0)), int(round((w - tw) / 2.0)), th, tw)


def hflip(img):
    return _like(_to_np(img)[:, ::-1].copy(), img)


def vflip(img):
    return _like(_to_np(img)[::-1].copy(), img)


def pad(img, padding, fill=0, padding_mode="constant"):
    if isinstance(padding, int):
        l = r = t = b = padding
    elif len(padding) == 2:
        l, t = padding
        r, b = padding
    else:
        l, t, r, b = padding
    a = _to_np(img)
    pw = ((t, b), (l, r)) + (((0, 0),) if a.ndim == 3 else ())
    if padding_mode == "constant":
        out = np.pad(a, pw, mode="constant", constant_values=fill)
    else:
        out = np.pad(a, pw, mode={"edge": "edge", "reflect": "reflect", "symmetric": "symmetric"}[padding_mode])
    return _like(out, img)


def normalize(img, mean, std, data_format="CHW", to_rgb=False):
    if isinstance(img, torch.Tensor):
        t = img.as_subclass(torch.Tensor).float()
        shape = (-1, 1, 1) if data_format == "CHW" else (1, 1, -1)
        m = torch.as_tensor(mean, dtype=t.dtype).reshape(shape)
        s = torch.as_tensor(std, dtype=t.dtype).reshape(shape)
        return ((t - m) / s).as_subclass(Tensor)
    a = _to_np(img).astype(np.float32)
    if to_rgb:
        a = a[..., ::-1]
    shape = (-1, 1, 1) if data_format == "CHW" else (1, 1, -1)
    return (a - np.asarray(mean, np.float32).reshape(shape)) / np.asarray(std, np.float32).reshape(shape)


def to_grayscale(img, num_output_channels=1):
    a = _to_np(img).astype(np.float32)
    g = (0.299 * a[..., 0] + 0.587 * a[..., 1] + 0.114 * a[..., 2])
    g = np.clip(g, 0, 255).astype(np.uint8)[..., None]
    if num_output_channels == 3:
        g = np.repeat(g, 3, -1)
    return _like(g if num_output_channels == 3 else g, img) if not _is_pil(img) else Image.fromarray(g[..., 0] if num_output_channels == 1 else g)


def _blend(a, b, f):
    return np.clip(a.astype(np.float32) * f + b.astype(np.float32) * (1 - f), 0, 255).astype(np.uint8)


def adjust_brightness(img, brightness_factor):
    a = _to_np(img)
    return _like(_blend(a, np.zeros_like(a), brightness_factor), img)


def adjust_contrast(img, contrast_factor):
    a = _to_np(img)
    mean = (0.299 * a[..., 0] + 0.587 * a[..., 1] + 0.114 * a[..., 2]).mean() if a.ndim == 3 and a.shape[-1] == 3 else a.mean()
    return _like(_blend(a, np.full_like(a, int(round(mean))), contrast_factor), img)


def adjust_saturation(img, saturation_factor):
    a = _to_np(img)
    g = (0.299 * a[..., 0] + 0.587 * a[..., 1] + 0.114 * a[..., 2])[..., None]
    return _like(_blend(a, np.repeat(g, 3, -1).astype(a.dtype), saturation_factor), img)


def adjust_hue(img, hue_factor):
    if not -0.5 <= hue_factor <= 0.5:
        raise ValueError("hue_factor is not in [-0.5, 0.5].")
    a = _to_np(img).astype(np.float32) / 255.0
    r, g, b = a[..., 0], a[..., 1], a[..., 2]
    mx, mn = a.max(-1), a.min(-1)
    d = mx - mn
    h = np.zeros_like(mx)
    m = d > 0
    rc, gc, bc = (mx - r) / np.where(m, d, 1), (mx - g) / np.where(m, d, 1), (mx - b) / np.where(m, d, 1)
    h = np.where(mx == r, bc - gc, np.where(mx == g, 2.0 + rc - bc, 4.0 + gc - rc))
    h = (h / 6.0) % 1.0
    h = np.where(m, h, 0.0)
    s = np.where(mx > 0, d / np.where(mx > 0, mx, 1), 0)
    v = mx
    h = (h + hue_factor) % 1.0
    i = np.floor(h * 6.0)
    f = h * 6.0 - i
    p, q, t = v * (1 - s), v * (1 - s * f), v * (1 - s * (1 - f))
    i = i.astype(np.int32) % 6
    out = np.stack([np.choose(i, [v, q, p, p, t, v]), np.choose(i, [t, v, v, q, p, p]), np.choose(i, [p, p, t, v, v, q])], -1)
    return _like(np.clip(out * 255.0 + 0.5, 0, 255).astype(np.uint8), img)


def _affine_grid_sample(a, matrix, out_hw, interpolation="nearest", fill=0):
    """matrix maps output pixel coords -> input pixel coords (2x3)."""
    h, w = a.shape[:2]
    oh, ow = out_hw
    ys, xs = np.meshgrid(np.arange(oh, dtype=np.float32), np.arange(ow, dtype=np.float32), indexing="ij")
    sx = matrix[0][0] * xs + matrix[0][1] * ys + matrix[0][2]
    sy = matrix[1][0] * xs + matrix[1][1] * ys + matrix[1][2]
    t = torch.from_numpy(np.ascontiguousarray(a)).float()
    chw = t.permute(2, 0, 1)[None] if t.dim() == 3 else t[None, None]
    gx = torch.from_numpy((sx + 0.5) / w * 2 - 1)
    gy = torch.from_numpy((sy + 0.5) / h * 2 - 1)
    grid = torch.stack([gx, gy], -1)[None]
    out = torch.nn.functional.grid_sample(chw - fill, grid, mode="bilinear" if interpolation == "bilinear" else "nearest", padding_mode="zeros", align_corners=False) + fill
    out = out[0].permute(1, 2, 0).numpy() if t.dim() == 3 else out[0, 0].numpy()
    return np.clip(np.round(out), 0, 255).astype(np.uint8) if a.dtype == np.uint8 else out


def rotate(img, angle, interpolation="nearest", expand=False, center=None, fill=0):
    a = _to_np(img)
    h, w = a.shape[:2]
    cx, cy = ((w - 1) / 2.0, (h - 1) / 2.0) if center is None else center
    th = math.radians(angle)
    c, s = math.cos(th), math.sin(th)
    oh, ow = h, w
    if expand:
        ow = int(math.ceil(round(abs(w * c) + abs(h * s), 6)))
        oh = int(math.ceil(round(abs(w * s) + abs(h * c), 6)))
    ocx, ocy = (ow - 1) / 2.0, (oh - 1) / 2.0
    m = [[c, -s, cx - c * ocx + s * ocy], [s, c, cy - s * ocx - c * ocy]]
    return _like(_affine_grid_sample(a, m, (oh, ow), interpolation, fill if isinstance(fill, numbers.Number) else 0), img)


def affine(img, angle, translate, scale, shear, interpolation="nearest", fill=0, center=None):
    a = _to_np(img)
    h, w = a.shape[:2]
    cx, cy = ((w - 1) / 2.0, (h - 1) / 2.0) if center is None else center
    rot = math.radians(angle)
    sx, sy = [math.radians(s) for s in (shear if isinstance(shear, (list, tuple)) else (shear, 0.0))]
    # forward matrix M = T * C * R * S * Sc * C^-1 ; we need its inverse
    a_ = math.cos(rot - sy) / math.cos(sy)
    b_ = -math.cos(rot - sy) * math.tan(sx) / math.cos(sy) - math.sin(rot)
    c_ = math.sin(rot - sy) / math.cos(sy)
    d_ = -math.sin(rot - sy) * math.tan(sx) / math.cos(sy) + math.cos(rot)
    fwd = np.array([[scale * a_, scale * b_, 0], [scale * c_, scale * d_, 0], [0, 0, 1]], dtype=np.float64)
    T1 = np.array([[1, 0, cx + translate[0]], [0, 1, cy + translate[1]], [0, 0, 1]], dtype=np.float64)
    T2 = np.array([[1, 0, -cx], [0, 1, -cy], [0, 0, 1]], dtype=np.float64)
    inv = np.linalg.inv(T1 @ fwd @ T2)
    return _like(_affine_grid_sample(a, inv[:2].tolist(), (h, w), interpolation, fill if isinstance(fill, numbers.Number) else 0), img)


def perspective(img, startpoints, endpoints, interpolation="nearest", fill=0):
    a = _to_np(img)
    h, w = a.shape[:2]
    A, B = [], []
    for (x, y), (u, v) in zip(endpoints, startpoints):
        A += [[x, y, 1, 0, 0, 0, -u * x, -u * y], [0, 0, 0, x, y, 1, -v * x, -v * y]]
        B += [u, v]
    coef = np.linalg.lstsq(np.asarray(A, np.float64), np.asarray(B, np.float64), rcond=None)[0]
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    den = coef[6] * xs + coef[7] * ys + 1
    sx = (coef[0] * xs + coef[1] * ys + coef[2]) / den
    sy = (coef[3] * xs + coef[4] * ys + coef[5]) / den
    t = torch.from_numpy(np.ascontiguousarray(a)).float()
    chw = t.permute(2, 0, 1)[None] if t.dim() == 3 else t[None, None]
    grid = torch.stack([torch.from_numpy((sx + 0.5) / w * 2 - 1), torch.from_numpy((sy + 0.5) / h * 2 - 1)], -1)[None].float()
    out = torch.nn.functional.grid_sample(chw, grid, mode="bilinear" if interpolation == "bilinear" else "nearest", padding_mode="zeros", align_corners=False)
    out = out[0].permute(1, 2, 0).numpy() if t.dim() == 3 else out[0, 0].numpy()
    return _like(np.clip(np.round(out), 0, 255).astype(np.uint8) if a.dtype == np.uint8 else out, img)


def erase(img, i, j, h, w, v, inplace=False):
    if isinstance(img, torch.Tensor):
        out = img if inplace else img.clone()
        out[..., i:i + h, j:j + w] = torch.as_tensor(v, dtype=out.dtype)
        return out
    a = _to_np(img).copy()
    a[i:i + h, j:j + w] = v
    return _like(a, img)


# ------------------------------------------------------------------------------------------------ transforms
class BaseTransform:
    def __init__(self, keys=None):
        self.keys = keys

    def _apply_image(self, img):
        raise NotImplementedError

    def __call__(self, inputs):
        if isinstance(inputs, tuple) and self.keys:
            return tuple(self._apply_image(x) if k == "image" else x for x, k in zip(inputs, self.keys))
        return self._apply_image(inputs)


class Compose:
    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, data):
        for f in self.transforms:
            data = f(data)
        return data


class ToTensor(BaseTransform):
    def __init__(self, data_format="CHW", keys=None):
        super().__init__(keys)
        self.data_format = data_format

    def _apply_image(self, img):
        return to_tensor(img, self.data_format)


class Resize(BaseTransform):
    def __init__(self, size, interpolation="bilinear", keys=None):
        super().__init__(keys)
        self.size, self.interpolation = size, interpolation

    def _apply_image(self, img):
        return resize(img, self.size, self.interpolation)


class RandomResizedCrop(BaseTransform):
    def __init__(self, size, scale=(0.08, 1.0), ratio=(3.0 / 4, 4.0 / 3), interpolation="bilinear", keys=None):
        super().__init__(keys)
        self.size = (size, size) if isinstance(size, int) else tuple(size)
        self.scale, self.ratio, self.interpolation = scale, ratio, interpolation

    def _apply_image(self, img):
        a = _to_np(img)
        h, w = a.shape[:2]
        area = h * w
        for _ in range(10):
            ta = random.uniform(*self.scale) * area
            ar = math.exp(random.uniform(math.log(self.ratio[0]), math.log(self.ratio[1])))
            cw, ch = int(round(math.sqrt(ta * ar))), int(round(math.sqrt(ta / ar)))
            if 0 < cw <= w and 0 < ch <= h:
                i, j = random.randint(0, h - ch), random.randint(0, w - cw)
                return resize(crop(img, i, j, ch, cw), self.size, self.interpolation)
        return resize(center_crop(img, min(h, w)), self.size, self.interpolation)


class CenterCrop(BaseTransform):
    def __init__(self, size, keys=None):
        super().__init__(keys)
        self.size = size

    def _apply_image(self, img):
        return center_crop(img, self.size)


class RandomCrop(BaseTransform):
    def __init__(self, size, padding=None, pad_if_needed=False, fill=0, padding_mode="constant", keys=None):
        super().__init__(keys)
        self.size = (size, size) if isinstance(size, int) else tuple(size)
        self.padding, self.pad_if_needed, self.fill, self.padding_mode = padding, pad_if_needed, fill, padding_mode

    def _apply_image(self, img):
        if self.padding is not None:
            img = pad(img, self.padding, self.fill, self.padding_mode)
        a = _to_np(img)
        h, w = a.shape[:2]
        th, tw = self.size
        if self.pad_if_needed and (w < tw or h < th):
            img = pad(img, (max(0, tw - w), max(0, th - h)), self.fill, self.padding_mode)
            a = _to_np(img)
            h, w = a.shape[:2]
        i, j = random.randint(0, h - th), random.randint(0, w - tw)
        return crop(img, i, j, th, tw)


class RandomHorizontalFlip(BaseTransform):
    def __init__(self, prob=0.5, keys=None):
        super().__init__(keys)
        self.prob = prob

    def _apply_image(self, img):
        return hflip(img) if random.random() < self.prob else img


class RandomVerticalFlip(BaseTransform):
    def __init__(self, prob=0.5, keys=None):
        super().__init__(keys)
        self.prob = prob

    def _apply_image(self, img):
        return vflip(img) if random.random() < self.prob else img


class Normalize(BaseTransform):
    def __init__(self, mean=0.0, std=1.0, data_format="CHW", to_rgb=False, keys=None):
        super().__init__(keys)
        self.mean = [mean] * 3 if isinstance(mean, numbers.Number) else mean
        self.std = [std] * 3 if isinstance(std, numbers.Number) else std
        self.data_format, self.to_rgb = data_format, to_rgb

    def _apply_image(self, img):
        return normalize(img, self.mean, self.std, self.data_format, self.to_rgb)


class Transpose(BaseTransform):
    def __init__(self, order=(2, 0, 1), keys=None):
        super().__init__(keys)
        self.order = order

    def _apply_image(self, img):
        a = _to_np(img)
        if a.ndim == 2:
            a = a[..., None]
        return a.transpose(self.order)


class Pad(BaseTransform):
    def __init__(self, padding, fill=0, padding_mode="constant", keys=None):
        super().__init__(keys)
        self.padding, self.fill, self.padding_mode = padding, fill, padding_mode

    def _apply_image(self, img):
        return pad(img, self.padding, self.fill, self.padding_mode)


class Grayscale(BaseTransform):
    def __init__(self, num_output_channels=1, keys=None):
        super().__init__(keys)
        self.n = num_output_channels

    def _apply_image(self, img):
        return to_grayscale(img, self.n)


def _rand_factor(v):
    return random.uniform(max(0, 1 - v), 1 + v)


class BrightnessTransform(BaseTransform):
    def __init__(self, value, keys=None):
        super().__init__(keys)
        self.value = value

    def _apply_image(self, img):
        return adjust_brightness(img, _rand_factor(self.value)) if self.value else img


class ContrastTransform(BaseTransform):
    def __init__(self, value, keys=None):
        super().__init__(keys)
        self.value = value

    def _apply_image(self, img):
        return adjust_contrast(img, _rand_factor(self.value)) if self.value else img


class SaturationTransform(BaseTransform):
    def __init__(self, value, keys=None):
        super().__init__(keys)
        self.value = value

    def _apply_image(self, img):
        return adjust_saturation(img, _rand_factor(self.value)) if self.value else img


class HueTransform(BaseTransform):
    def __init__(self, value, keys=None):
        super().__init__(keys)
        self.value = value

    def _apply_image(self, img):
        return adjust_hue(img, random.uniform(-self.value, self.value)) if self.value else img


class ColorJitter(BaseTransform):
    def __init__(self, brightness=0, contrast=0, saturation=0, hue=0, keys=None):
        super().__init__(keys)
        self.ts = [BrightnessTransform(brightness), ContrastTransform(contrast), SaturationTransform(saturation), HueTransform(hue)]

    def _apply_image(self, img):
        order = list(range(4))
        random.shuffle(order)
        for i in order:
            img = self.ts[i]._apply_image(img)
        return img


class RandomRotation(BaseTransform):
    def __init__(self, degrees, interpolation="nearest", expand=False, center=None, fill=0, keys=None):
        super().__init__(keys)
        self.degrees = (-degrees, degrees) if isinstance(degrees, numbers.Number) else tuple(degrees)
        self.interpolation, self.expand, self.center, self.fill = interpolation, expand, center, fill

    def _apply_image(self, img):
        return rotate(img, random.uniform(*self.degrees), self.interpolation, self.expand, self.center, self.fill)


class RandomAffine(BaseTransform):
    def __init__(self, degrees, translate=None, scale=None, shear=None, interpolation="nearest", fill=0, center=None, keys=None):
        super().__init__(keys)
        self.degrees = (-degrees, degrees) if isinstance(degrees, numbers.Number) else tuple(degrees)
        self.translate, self.scale, self.shear = translate, scale, shear
        self.interpolation, self.fill, self.center = interpolation, fill, center

    def _apply_image(self, img):
        a = _to_np(img)
        h, w = a.shape[:2]
        angle = random.uniform(*self.degrees)
        tr = (0, 0) if self.translate is None else (round(random.uniform(-self.translate[0] * w, self.translate[0] * w)), round(random.uniform(-self.translate[1] * h, self.translate[1] * h)))
        sc = 1.0 if self.scale is None else random.uniform(*self.scale)
        sh = (0.0, 0.0)
        if self.shear is not None:
            s = self.shear if isinstance(self.shear, (list, tuple)) else (-self.shear, self.shear)
            sh = (random.uniform(s[0], s[1]), random.uniform(s[2], s[3]) if len(s) == 4 else 0.0)
        return affine(img, angle, tr, sc, sh, self.interpolation, self.fill, self.center)


class RandomPerspective(BaseTransform):
    def __init__(self, prob=0.5, distortion_scale=0.5, interpolation="nearest", fill=0, keys=None):
        super().__init__(keys)
        self.prob, self.d, self.interpolation, self.fill = prob, distortion_scale, interpolation, fill

    def _apply_image(self, img):
        if random.random() >= self.prob:
            return img
        a = _to_np(img)
        h, w = a.shape[:2]
        hw, hh = int(self.d * w / 2), int(self.d * h / 2)
        tl = (random.randint(0, hw), random.randint(0, hh))
        tr = (w - 1 - random.randint(0, hw), random.randint(0, hh))
        br = (w - 1 - random.randint(0, hw), h - 1 - random.randint(0, hh))
        bl = (random.randint(0, hw), h - 1 - random.randint(0, hh))
        return perspective(img, [(0, 0), (w - 1, 0), (w - 1, h - 1), (0, h - 1)], [tl, tr, br, bl], self.interpolation, self.fill)


class RandomErasing(BaseTransform):
    def __init__(self, prob=0.5, scale=(0.02, 0.33), ratio=(0.3, 3.3), value=0, inplace=False, keys=None):
        super().__init__(keys)
        self.prob, self.scale, self.ratio, self.value, self.inplace = prob, scale, ratio, value, inplace

    def _apply_image(self, img):
        if random.random() >= self.prob:
            return img
        shape = img.shape if isinstance(img, torch.Tensor) else _to_np(img).shape
        h, w = (shape[-2], shape[-1]) if isinstance(img, torch.Tensor) else shape[:2]
        area = h * w
        for _ in range(10):
            ea = random.uniform(*self.scale) * area
            ar = math.exp(random.uniform(math.log(self.ratio[0]), math.log(self.ratio[1])))
            eh, ew = int(round(math.sqrt(ea * ar))), int(round(math.sqrt(ea / ar)))
            if eh < h and ew < w:
                i, j = random.randint(0, h - eh), random.randint(0, w - ew)
                v = np.random.normal(size=(eh, ew) if not isinstance(img, torch.Tensor) else ()).astype(np.float32) if self.value == "random" else self.value
                return erase(img, i, j, eh, ew, v, self.inplace)
        return img
