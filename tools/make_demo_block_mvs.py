#!/usr/bin/env python
"""Coherent (real-video) macroblock motion for the bank-conflict measurements of DESIGN.md section 3: block-match the
144 demo frames under /root/reference/demo/ILSVRC2015_val_00007010/ (the only video the reference ships) and store one
integer vector per 16x16 macroblock, in coviar's accumulated convention (mv = dst - src, the displacement back to the
GOP's key frame; lib/utils/image.py:52-54 negates it).  GOP = 12 frames (config.py KEY_FRAME_INTERVAL): frame 12k is the
key frame, frames 12k+1..12k+11 are matched against it directly.  Frames are resized to 1000x600 (the network scale of
SCALES = (600,1000)), so the field is 38x63 macroblocks like the synthetic one.  Two-level search, SAD on grey levels (+-12 at half resolution, then +-1 at full resolution).

Runs in the authoring container only (the GPU box has no /root/reference); output: tests/golden/demo_block_mvs.npz.
"""
import glob
import os

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/demo/ILSVRC2015_val_00007010"
W, H, B = 1000, 600, 16


def block_sad(a, b, blk):
    d = np.abs(a.astype(np.int16) - b.astype(np.int16)).astype(np.int32)
    hh, ww = d.shape[0] // blk * blk, d.shape[1] // blk * blk
    return d[:hh, :ww].reshape(hh // blk, blk, ww // blk, blk).sum(axis=(1, 3))


def shift(img, dx, dy):
    """img sampled at (x+dx, y+dy), edge-replicated."""
    h, w = img.shape
    ys = np.clip(np.arange(h) + dy, 0, h - 1)
    xs = np.clip(np.arange(w) + dx, 0, w - 1)
    return img[ys][:, xs]


def match(cur, key):
    ph, pw = -(-H // B) * B, -(-W // B) * B
    cur = np.pad(cur, ((0, ph - H), (0, pw - W)), mode="edge")
    key = np.pad(key, ((0, ph - H), (0, pw - W)), mode="edge")
    c2, k2 = cv2.resize(cur, (pw // 2, ph // 2), interpolation=cv2.INTER_AREA), cv2.resize(key, (pw // 2, ph // 2), interpolation=cv2.INTER_AREA)
    nby, nbx = ph // B, pw // B
    best = np.full((nby, nbx), 1 << 30, np.int64)
    vec = np.zeros((nby, nbx, 2), np.int32)
    R = 10
    for dy in range(-R, R + 1):
        for dx in range(-R, R + 1):
            s = block_sad(c2, shift(k2, dx, dy), B // 2) + (abs(dx) + abs(dy))      # tiny bias towards zero motion (flat areas)
            m = s < best
            best[m] = s[m]
            vec[m] = (2 * dx, 2 * dy)
    # +-1 refinement at full resolution: one gather of the key frame per candidate (every block shifted by its own vector)
    out = vec.copy()
    best = np.full((nby, nbx), 1 << 30, np.int64)
    yy, xx = np.mgrid[0:ph, 0:pw]
    vpx = np.repeat(np.repeat(vec, B, axis=0), B, axis=1)
    for ddy in (-1, 0, 1):
        for ddx in (-1, 0, 1):
            ky = np.clip(yy + vpx[..., 1] + ddy, 0, ph - 1)
            kx = np.clip(xx + vpx[..., 0] + ddx, 0, pw - 1)
            s = block_sad(cur, key[ky, kx], B)
            m = s < best
            best[m] = s[m]
            out[m] = vec[m] + (ddx, ddy)
    return -out                     # src = dst + (dx,dy)  ->  coviar's mv = dst - src


def main():
    files = sorted(glob.glob(os.path.join(SRC, "*.JPEG")))
    frames = [cv2.resize(cv2.cvtColor(cv2.imread(f), cv2.COLOR_BGR2GRAY), (W, H), interpolation=cv2.INTER_AREA) for f in files]
    mvs = []
    for k in range(0, len(frames) - 11, 12):
        for t in range(1, 12):
            mvs.append(match(frames[k + t], frames[k]))
    mvs = np.stack(mvs)[:, : -(-H // B), : -(-W // B)]
    assert np.abs(mvs).max() < 128
    zero = float((np.abs(mvs).sum(-1) == 0).mean())
    same_as_left = float((mvs[:, :, 1:] == mvs[:, :, :-1]).all(-1).mean())
    dst = os.path.join(ROOT, "tests", "golden", "demo_block_mvs.npz")
    np.savez_compressed(dst, block_mv=mvs.astype(np.int8), gop_pos=np.tile(np.arange(1, 12), len(mvs) // 11).astype(np.int8))
    print(dst, os.path.getsize(dst), "bytes;", mvs.shape, "zero blocks %.3f, equal to left neighbour %.3f, |mv| p50/p90/max = %s"
          % (zero, same_as_left, np.percentile(np.abs(mvs), [50, 90, 100]).tolist()))


if __name__ == "__main__":
    main()
