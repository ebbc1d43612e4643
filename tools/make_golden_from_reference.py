#!/usr/bin/env python
"""Generate golden fixtures by running the REFERENCE's own Python code.

Runs only in the authoring container (needs /root/reference and cv2).  It imports
``/root/reference/lib/utils/image.py`` unmodified - with stub modules for the two imports
that cannot be satisfied offline (``coviar_py2`` needs FFmpeg, ``bbox.bbox_transform`` is
a Cython build) and that the called functions never touch - and records
``transform_mv_res`` / ``resize`` outputs for seeded inputs.

Outputs: tests/golden/ref_transform_mv_res.npz  (inputs + reference outputs)
The GPU box has no /root/reference; tests read only the committed .npz.

NOTE on version drift: the reference pins opencv-python 3.2.0.6 (README.md:31); this
container has cv2 4.13.  The stride-16 stage is exact arithmetic on these inputs and
cannot drift; the im_scale stage is float32 bilinear and is recorded as produced here.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "tests", "golden", "ref_transform_mv_res.npz")


def load_reference_image_module():
    stub_coviar = types.ModuleType("coviar_py2")
    stub_bbox = types.ModuleType("bbox")
    stub_bt = types.ModuleType("bbox.bbox_transform")
    stub_bt.clip_boxes = lambda boxes, shape: boxes
    stub_bbox.bbox_transform = stub_bt
    sys.modules.setdefault("coviar_py2", stub_coviar)
    sys.modules.setdefault("bbox", stub_bbox)
    sys.modules.setdefault("bbox.bbox_transform", stub_bt)
    spec = importlib.util.spec_from_file_location("ref_image", os.path.join(REF, "lib/utils/image.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference_image_module()
    rng = np.random.default_rng(20261017)
    cases = {}
    # (h, w, im_scale): multiples of 16, ragged, up- and down-scaling, canonical aspect
    specs = [(96, 160, 1.0), (75, 131, 1.0), (90, 160, 0.78125), (72, 96, 1.25),
             (48, 80, 2.0), (150, 250, 1.0), (117, 203, 600.0 / 117.0 / 4.0)]
    means = np.array([0.0, 0.0, 0.0])
    for i, (h, w, s) in enumerate(specs):
        blk = rng.integers(-48, 49, size=(-(-h // 16), -(-w // 16), 2)).astype(np.int32)
        blk[rng.random(blk.shape[:2]) < 0.4] = 0
        mv = np.repeat(np.repeat(blk, 16, 0), 16, 1)[:h, :w].astype(np.float32)
        if i % 2 == 1:      # per-pixel (non macroblock-constant) field too
            mv = rng.integers(-48, 49, size=(h, w, 2)).astype(np.float32)
        res = rng.integers(-64, 65, size=(h, w, 3)).astype(np.float32)
        mv_t, res_t = ref.transform_mv_res(mv.copy(), res.copy(), s, means, 1.0)
        cases["mv_in_%d" % i] = mv
        cases["res_in_%d" % i] = res
        cases["scale_%d" % i] = np.float64(s)
        cases["mv_out_%d" % i] = np.asarray(mv_t)          # float64, as the reference returns
        cases["res_out_%d" % i] = np.asarray(res_t)
    # non-zero pixel means / scale exercise the aliasing loop of image.py:217-218
    mv = rng.integers(-16, 17, size=(64, 64, 2)).astype(np.float32)
    res = rng.integers(-64, 65, size=(64, 64, 3)).astype(np.float32)
    pm = np.array([103.06, 115.90, 123.15])
    mv_t, res_t = ref.transform_mv_res(mv.copy(), res.copy(), 1.0, pm, 0.5)
    cases.update(mv_in_m=mv, res_in_m=res, means_m=pm, pscale_m=np.float64(0.5),
                 mv_out_m=np.asarray(mv_t), res_out_m=np.asarray(res_t))
    # resize(): the im_scale rule, on image shapes only (image content irrelevant)
    shapes = [(720, 1280), (480, 640), (360, 480), (1080, 1920), (600, 1000), (500, 333)]
    scales = []
    for (h, w) in shapes:
        _, sc = ref.resize(np.zeros((h, w, 3), np.float32), 600, 1000, stride=0)
        scales.append(sc)
    cases["resize_shapes"] = np.array(shapes)
    cases["resize_scales"] = np.array(scales, dtype=np.float64)
    cases["n_cases"] = np.int64(len(specs))
    np.savez_compressed(OUT, **cases)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
