"""Results writer behind the decode path: the on-disk formats of the reference's
utils/eval_util.py — `{epoch}_dets.json` / `{epoch}_infos.json` (:65-70) and, per image, `<basename>pred.txt`
plus one `results/<basename>_<class>_<k>.png` mask per detection (:100-125, the cityscapesscripts instance-level
layout).  The masks come from the device rasteriser (`image.polys_to_masks`, one launch per image) instead of
one `poly_to_mask` / cv2.fillPoly call per detection; the files are byte-identical to the reference's.

The model loop (`eval_outputs` :35-63) and the cityscapesscripts scorer (:126-127) stay with the caller: they are
outside the decode path."""
from __future__ import annotations

import json
import os

import numpy as np

from . import image


class NpEncoder(json.JSONEncoder):
    """numpy scalars / arrays -> JSON numbers / lists (reference :23-32)."""

    def default(self, obj):
        if isinstance(obj, np.integer):
            return int(obj)
        if isinstance(obj, np.floating):
            return float(obj)
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        return super().default(obj)


def save_dets(dets_list, info_list, output_dir, epoch):
    """`{epoch}_dets.json`, `{epoch}_infos.json` (reference :65-70); returns the two paths."""
    dets_path = os.path.join(output_dir, "{}_dets.json".format(epoch))
    infos_path = os.path.join(output_dir, "{}_infos.json".format(epoch))
    with open(dets_path, "w") as f:
        f.write(json.dumps(dets_list, cls=NpEncoder))
    with open(infos_path, "w") as f:
        f.write(json.dumps(info_list, cls=NpEncoder))
    return dets_path, infos_path


def load_dets(output_dir, epoch):
    """Inverse of save_dets (reference :74-75)."""
    with open(os.path.join(output_dir, "{}_dets.json".format(epoch))) as f:
        dets_list = json.load(f)
    with open(os.path.join(output_dir, "{}_infos.json".format(epoch))) as f:
        info_list = json.load(f)
    return dets_list, info_list


def write_results(dets_list, info_list, output_dir, label_names, label_ids, logger=None):
    """Per image: `<basename>pred.txt` with one `<png> <label id> <score>` line per detection, classes in label
    order, and the 0/255 mask PNGs under `results/` (reference :100-125).

    dets_list[i]: [(cls, conf, centre, polygon [K,2] (x,y)), ...] as returned by decode_output or loaded back from
    `{epoch}_dets.json`; info_list[i] = (img_path, img_size (H, W))."""
    import cv2
    results_dir = os.path.join(output_dir, "results")
    if not os.path.exists(results_dir):
        os.mkdir(results_dir)
    for i, dets in enumerate(dets_list):
        im_name, img_size = info_list[i][0], info_list[i][1]
        basename = os.path.splitext(os.path.basename(im_name))[0]
        if logger is not None and i % 10 == 0:
            logger.write("i: {}: {}".format(i, basename))
        masks = image.polys_to_masks([np.array(d[3]) for d in dets], img_size=tuple(img_size)) if len(dets) else []
        with open(os.path.join(output_dir, basename + "pred.txt"), "w") as fid_txt:
            for j in range(len(label_names)):
                for k in range(len(dets)):
                    center_cls, center_conf = dets[k][0], dets[k][1]
                    if center_cls != j:
                        continue
                    pngname = os.path.join("results", basename + "_" + label_names[j] + "_{}.png".format(k))
                    fid_txt.write("{} {} {}\n".format(pngname, label_ids[j], float(center_conf)))
                    cv2.imwrite(os.path.join(output_dir, pngname), (masks[k] * 255).astype(np.uint8))
