"""Stand-ins for the reference's TransInfo / CommonTransforms / Config objects (utils/tranform.py:19,126-171,
configs/__init__.py:22-44) with the default validation settings (configs/decode_cfg.yaml, trans_cfg.json)."""
from collections import namedtuple

import numpy as np

TransInfo = namedtuple('TransInfo', ['img_path', 'img_size'])


class _Configer:
    def get(self, *key):
        if key == ('val_trans', 'trans_seq'):
            return []
        raise KeyError(key)


class IdentityTransforms:
    """detransform_pixel of CommonTransforms with val_trans.trans_seq = [] : (y,x) -> (x,y)."""
    configer = _Configer()

    def detransform_pixel(self, pixels, info):
        return pixels.reshape(-1, 2)[:, ::-1]


class _ResizeConfiger:
    def __init__(self, scale):
        self.scale = scale

    def get(self, *key):
        if key == ('val_trans', 'trans_seq'):
            return ['resize']
        if key == ('val_trans', 'resize'):
            return {'target_size': self.scale}
        raise KeyError(key)


class ResizeTransforms:
    """detransform_pixel of CommonTransforms with val_trans.trans_seq = ['resize'], resize.target_size = scale
    (utils/tranform.py:157-171): flip to (x,y), then the inverse affine map to the original image size."""

    def __init__(self, scale):
        self.scale = scale
        self.configer = _ResizeConfiger(scale)

    def detransform_pixel(self, pixels, info):
        from oracle.ref_decode import detransform_pixel
        return detransform_pixel(np.asarray(pixels), info.img_size, self.scale)


class DecodeCfg:
    def __init__(self, **kw):
        self.cls_th, self.iou_th, self.kp_th = 0.3, 0.2, 20000
        self.obj_pixel_th, self.wh_delta, self.alpha_ratio, self.draw_flag = 2, 0.1, 2, False
        self.__dict__.update(kw)


def unpack_bits(bits: np.ndarray, W: int) -> np.ndarray:
    """[..., H, Ww] 32-bit words -> [..., H, W] uint8"""
    b = np.ascontiguousarray(bits).view(np.uint8).reshape(bits.shape[:-1] + (-1,))
    return np.unpackbits(b, axis=-1, bitorder="little")[..., :W]
