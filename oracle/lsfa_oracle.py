"""CPU oracle for LSFA's non-key-frame feature-propagation + aggregation path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product path (``lsfa_b200``) never does and
fails loudly when its CUDA library is missing.

It is a NumPy restatement of the arithmetic the reference graph performs for this
path.  Every function cites the reference lines it follows (paths relative to
``/root/reference``; ``SYM`` = ``dff_rfcn/symbols/resnet_v1_101_flownet_rfcn.py``).

Pinning status (see DESIGN.md "Oracle"):

* MV / residual preparation (``transform_mv_res``, ``resize``; rows a2-a6 of
  SURVEY.md section 8a): **pinned** against the reference's own Python code
  imported in the authoring container (``tools/make_golden_from_reference.py``
  -> ``tests/golden/ref_transform_mv_res_*.npz``) and against ``cv2.resize``.
* ``GridGenerator(transform_type='warp')``, ``BilinearSampler``,
  ``L2Normalization``, ``softmax`` live in Apache MXNet @75a9e187d, which is NOT
  vendored in the reference tree and cannot be built offline: **parity unpinned**
  for those rows.  The formulas below restate MXNet's published CPU operators
  (src/operator/grid_generator-inl.h, src/operator/bilinear_sampler.cc,
  src/operator/l2_normalization-inl.h, src/operator/nn/softmax-inl.h) in their
  exact op order and precision, and are cross-checked against an independent
  implementation (``torch.nn.functional.grid_sample(align_corners=True)``).
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32
F64 = np.float64

RCNN_STRIDE = 16          # lib/utils/image.py:202 (rcnn_stride default)
L2NORM_EPS = 1e-10        # MXNet L2Normalization default eps (SYM:112-113 pass none)

POOL_CENTRE2X2 = 0        # what cv2.resize(fx=1/16, INTER_LINEAR) really computes
POOL_AVG16 = 1            # literal 16x16 block average (north_star wording)

W_NONE, W_ADD, W_MEAN, W_LOGITS, W_COSINE = 0, 1, 2, 3, 4


# --------------------------------------------------------------------------------------
# a1  raw MV fetch: sign and horizontal flip          lib/utils/image.py:52-60
# --------------------------------------------------------------------------------------
def mv_sign_flip(mv_coviar, flipped=False):
    """``motion_vector = -coviar.load(...)`` (image.py:53-54) and the h-flip of
    image.py:56-60: reverse x, negate channel 0.  (h,w,2) int32 -> (h,w,2) f32."""
    mv = -np.asarray(mv_coviar).astype(F32)
    if flipped:
        mv = mv[:, ::-1].copy()
        mv[:, :, 0] = -mv[:, :, 0]
    return mv


def im_scale_for(h, w, target_size=600, max_size=1000):
    """Scale chosen by ``resize`` (image.py:266-286), SCALES=(600,1000) config.py:25."""
    im_size_min = min(h, w)
    im_size_max = max(h, w)
    im_scale = float(target_size) / float(im_size_min)
    if np.round(im_scale * im_size_max) > max_size:
        im_scale = float(max_size) / float(im_size_max)
    return im_scale


# --------------------------------------------------------------------------------------
# upstream: coviar's accumulated MV field   coviar_data_loader.c:71-139,318-328
# --------------------------------------------------------------------------------------
def coviar_accumulate(mvs, counts, height, width):
    """Pure-Python literal loop (small cases only): per P-frame, every vector with dst != src, in
    list order, copies accu_old[src] to accu[dst] for the pixels of its w x h block whose dst
    AND src are inside the frame; accu starts as the identity; mv = (x,y) - accu."""
    ident = np.stack(np.meshgrid(np.arange(width), np.arange(height)), -1).astype(np.int32)  # [y,x] = (x,y)
    old = ident.copy()
    new = ident.copy()
    for t in range(mvs.shape[0]):
        for i in range(min(int(counts[t]), mvs.shape[1])):
            w, h, sx, sy, dx, dy = (int(v) for v in mvs[t, i])
            if dx - sx == 0 and dy - sy == 0:
                continue
            for xs in range(int(-1 * w / 2), int(w / 2)):          # C division truncates toward zero
                for ys in range(int(-1 * h / 2), int(h / 2)):
                    pdx, pdy, psx, psy = dx + xs, dy + ys, sx + xs, sy + ys
                    if 0 <= pdy < height and 0 <= pdx < width and 0 <= psy < height and 0 <= psx < width:
                        new[pdy, pdx] = old[psy, psx]
        old = new.copy()
    return ident - new


def coviar_residual(iframe, cur, mv):
    """coviar_data_loader.c:141-175 (accumulate case): cur - iframe[(x,y) - mv], int32."""
    h, w, _ = cur.shape
    ident = np.stack(np.meshgrid(np.arange(w), np.arange(h)), -1)
    src = ident - mv
    return cur.astype(np.int32) - iframe[src[..., 1], src[..., 0]].astype(np.int32)


def synth_mv_lists(rng, T, height, width, block=16, max_disp=24, p_move=0.6, extra=4):
    """Per P-frame motion-vector lists as FFmpeg exports them for MPEG-4: one vector per 16x16
    macroblock (dst = block centre), a share of them static, plus a few 8x8 vectors that overlap
    earlier ones and some that point outside the frame.  Returns mvs (T,M,6), counts (T,)."""
    bx, by = -(-width // block), -(-height // block)
    M = bx * by + extra
    mvs = np.zeros((T, M, 6), np.int32)
    counts = np.zeros(T, np.int32)
    for t in range(T):
        k = 0
        for j in range(by):
            for i in range(bx):
                dx, dy = i * block + block // 2, j * block + block // 2
                if rng.random() < p_move:
                    ox, oy = (int(v) for v in rng.integers(-max_disp, max_disp + 1, 2))
                else:
                    ox = oy = 0
                mvs[t, k] = (block, block, dx + ox, dy + oy, dx, dy)
                k += 1
        for _ in range(extra):
            dx, dy = int(rng.integers(0, width)), int(rng.integers(0, height))
            ox, oy = (int(v) for v in rng.integers(-max_disp, max_disp + 1, 2))
            mvs[t, k] = (8, 8, dx + ox, dy + oy, dx, dy)
            k += 1
        counts[t] = k
    return mvs, counts


# --------------------------------------------------------------------------------------
# a2  stage-1 resize by im_scale                      lib/utils/image.py:204-205
# --------------------------------------------------------------------------------------
def _cv_round(v):
    """cvRound: round-half-to-even (lrint) like OpenCV's dsize computation."""
    return int(np.rint(v))


def linear_resize_coeffs(dn, sn, scale):
    """Per-destination source index and weight of cv2.resize INTER_LINEAR (float32
    path): ``fx = (float)((d+0.5)*inv_scale - 0.5); sx = floor(fx); fx -= sx`` with the
    edge rules of cv::resize (sx<0 -> (0,0); sx>=sn-1 -> (sn-1,0))."""
    inv = 1.0 / scale
    d = np.arange(dn)
    f = ((d + 0.5) * inv - 0.5).astype(F32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(F32)).astype(F32)
    lo = s < 0
    f[lo] = 0
    s[lo] = 0
    hi = s >= sn - 1
    f[hi] = 0
    s[hi] = sn - 1
    return s, f


def resize_linear_f32(src, scale):
    """NumPy transcription of ``cv2.resize(src.astype(f32), None, None, fx=scale,
    fy=scale, INTER_LINEAR)`` (image.py:204-205).  Horizontal pass then vertical pass,
    float32, no FMA.  scale == 1 is the identity (cv2 copies)."""
    src = np.asarray(src, dtype=F32)
    if scale == 1.0:
        return src.copy()
    h, w = src.shape[:2]
    dh, dw = _cv_round(h * scale), _cv_round(w * scale)
    xi, xa = linear_resize_coeffs(dw, w, scale)
    yi, ya = linear_resize_coeffs(dh, h, scale)
    xi1 = np.minimum(xi + 1, w - 1)
    yi1 = np.minimum(yi + 1, h - 1)
    a0 = (F32(1) - xa)[None, :, None]
    a1 = xa[None, :, None]
    hr = src[:, xi, :] * a0 + src[:, xi1, :] * a1
    b0 = (F32(1) - ya)[:, None, None]
    b1 = ya[:, None, None]
    return (hr[yi] * b0 + hr[yi1] * b1).astype(F32)


# --------------------------------------------------------------------------------------
# a3-a6  pad-16, colour/mean aliasing, stride-16 reduction, rescale, layout
#        lib/utils/image.py:207-228
# --------------------------------------------------------------------------------------
def pad_to_stride(x, stride=RCNN_STRIDE):
    """image.py:207-215: zero-pad bottom/right to multiples of 16 into a float64 array."""
    h, w, c = x.shape
    ph = int(np.ceil(h / float(stride)) * stride)
    pw = int(np.ceil(w / float(stride)) * stride)
    out = np.zeros((ph, pw, c), dtype=F64)
    out[:h, :w] = x
    return out


def res_colour_mean_inplace(padded_res, pixel_means=(0.0, 0.0, 0.0), pixel_scale=1.0):
    """image.py:217-218, reproduced WITH its in-place aliasing: the loop reads channels
    it has already overwritten, so with zero means the result is (ch2, ch1, ch2)."""
    for i in range(3):
        padded_res[:, :, i] = (padded_res[:, :, 2 - i] - pixel_means[2 - i]) * pixel_scale
    return padded_res


def pool_stride16(padded, mode=POOL_CENTRE2X2, stride=RCNN_STRIDE):
    """image.py:220-222: ``cv2.resize(padded, fx=fy=1/16, INTER_LINEAR)`` on float64.
    Source coordinate of destination d is 16d+7.5, so the result is the mean of the 2x2
    centre pixels of every 16x16 block, evaluated horizontal-pass-first:
    ``((p[7,7]+p[7,8]) + (p[8,7]+p[8,8])) * 0.25`` (bit-exact vs cv2, see tests).
    ``POOL_AVG16`` is the literal block mean (row-major summation), offered because the
    north-star text says "average-pooled"; it is NOT what the reference computes."""
    ph, pw, c = padded.shape
    assert ph % stride == 0 and pw % stride == 0
    if mode == POOL_CENTRE2X2:
        lo, hi = stride // 2 - 1, stride // 2
        a = padded[lo::stride, lo::stride]
        b = padded[lo::stride, hi::stride]
        cc = padded[hi::stride, lo::stride]
        d = padded[hi::stride, hi::stride]
        return ((a + b) + (cc + d)) * 0.25
    elif mode == POOL_AVG16:
        blk = padded.reshape(ph // stride, stride, pw // stride, stride, c)
        acc = np.zeros((ph // stride, pw // stride, c), dtype=F64)
        for r in range(stride):           # fixed row-major order (matches the CUDA kernel)
            for q in range(stride):
                acc = acc + blk[:, r, :, q, :]
        return acc * (1.0 / (stride * stride))
    raise ValueError("unknown pool mode %r" % (mode,))


def transform_mv_res(motion_vector, res_diff, im_scale, pixel_means=(0.0, 0.0, 0.0),
                     pixel_scale=1.0, mode=POOL_CENTRE2X2, use_cv2=False):
    """Whole of image.py:202-228.  (h,w,2),(h,w,3) -> f64 (1,2,H,W),(1,3,H,W); the cast to
    float32 happens later in ``mx.nd.array`` (core/loader.py:140) - see ``to_f32``."""
    if use_cv2:
        import cv2
        mv = cv2.resize(np.asarray(motion_vector, dtype=F32), None, None, fx=im_scale,
                        fy=im_scale, interpolation=cv2.INTER_LINEAR)
        rs = cv2.resize(np.asarray(res_diff, dtype=F32), None, None, fx=im_scale,
                        fy=im_scale, interpolation=cv2.INTER_LINEAR)
    else:
        mv = resize_linear_f32(motion_vector, im_scale)
        rs = resize_linear_f32(res_diff, im_scale)
    pmv = pad_to_stride(mv)
    prs = res_colour_mean_inplace(pad_to_stride(rs), pixel_means, pixel_scale)
    rmv = pool_stride16(pmv, mode)
    rrs = pool_stride16(prs, mode)
    scale = im_scale * (1.0 / RCNN_STRIDE)
    rmv = rmv * scale
    th, tw, _ = rrs.shape
    return (rmv.transpose(2, 0, 1).reshape(1, 2, th, tw),
            rrs.transpose(2, 0, 1).reshape(1, 3, th, tw))


def to_f32(x):
    """``mx.nd.array`` default dtype cast (core/loader.py:140)."""
    return np.asarray(x).astype(F32)


def mv_pool(mv_batch, im_scale=1.0, mode=POOL_CENTRE2X2):
    """Batched a3+a5+a6 on already stage-1-resized MVs: (N,h,w,2) int32|f32 ->
    flow (N,2,H,W) f32 in feature cells.  ch0 = x-flow, ch1 = y-flow."""
    outs = []
    scale = im_scale * (1.0 / RCNN_STRIDE)
    for mv in mv_batch:
        p = pool_stride16(pad_to_stride(np.asarray(mv, dtype=F32)), mode) * scale
        outs.append(p.transpose(2, 0, 1))
    return to_f32(np.stack(outs))


def res_pool(res_batch, pixel_means=(0.0, 0.0, 0.0), pixel_scale=1.0, mode=POOL_CENTRE2X2):
    """Batched a3+a4+a5 for the residual: (N,h,w,3) -> (N,3,H,W) f32."""
    outs = []
    for rs in res_batch:
        p = res_colour_mean_inplace(pad_to_stride(np.asarray(rs, dtype=F32)),
                                    pixel_means, pixel_scale)
        outs.append(pool_stride16(p, mode).transpose(2, 0, 1))
    return to_f32(np.stack(outs))


# --------------------------------------------------------------------------------------
# a7  mx.sym.GridGenerator(transform_type='warp')     call-sites SYM:306,320,468,571,678
#     MXNet src/operator/grid_generator-inl.h (kWarp branch), all float32
# --------------------------------------------------------------------------------------
def grid_dst_xy(H, W):
    """MXNet's ``grid_dst``: x = i - int(i/W)*W, y = int(i/W), generated in float32."""
    i = np.arange(H * W, dtype=F32)
    q = (i / F32(W)).astype(np.int32).astype(F32)
    x = i - q * F32(W)
    return x.reshape(H, W), q.reshape(H, W)


def grid_generator_warp(flow):
    """grid[n,0] = (flow[n,0] + x) / ((W-1)/2) - 1 ; grid[n,1] = (flow[n,1] + y) /
    ((H-1)/2) - 1.  Op order add, div, sub; IEEE float32; no FMA."""
    flow = np.asarray(flow, dtype=F32)
    N, two, H, W = flow.shape
    assert two == 2
    x, y = grid_dst_xy(H, W)
    half_w = F32((F64(F32(W)) - 1.0) / 2.0)
    half_h = F32((F64(F32(H)) - 1.0) / 2.0)
    grid = np.empty_like(flow)
    with np.errstate(divide="ignore", invalid="ignore"):
        grid[:, 0] = (flow[:, 0] + x) / half_w - F32(1)
        grid[:, 1] = (flow[:, 1] + y) / half_h - F32(1)
    return grid


# --------------------------------------------------------------------------------------
# a8  mx.sym.BilinearSampler(data, grid)              call-sites SYM:307,321,469,572,679
#     MXNet src/operator/bilinear_sampler.cc BilinearSamplerForward (CPU)
# --------------------------------------------------------------------------------------
def sampler_coords(grid, Hi, Wi):
    """De-normalisation + floor + top-left weights, exactly as the CPU operator:
    ``real = (g + 1) * (dim - 1) / 2`` in float32; ``w = 1.0 - (real - floor)`` evaluated
    in double (the literal is a double) and stored to float32.
    Returns x0,y0 (int32) and wx,wy (float32) of shape (N,Ho,Wo)."""
    grid = np.asarray(grid, dtype=F32)
    gx, gy = grid[:, 0], grid[:, 1]
    x_real = (gx + F32(1)) * F32(Wi - 1) / F32(2)
    y_real = (gy + F32(1)) * F32(Hi - 1) / F32(2)
    fx, fy = np.floor(x_real), np.floor(y_real)
    # static_cast<int> of an out-of-range float is UB in C++; clamp far outside the
    # image so every tap is still "not between" (same observable result: zeros).
    big = F32(2 ** 24)
    x0 = np.clip(np.nan_to_num(fx, nan=-big, posinf=big, neginf=-big), -big, big).astype(np.int32)
    y0 = np.clip(np.nan_to_num(fy, nan=-big, posinf=big, neginf=-big), -big, big).astype(np.int32)
    with np.errstate(invalid="ignore"):
        wx = (1.0 - (x_real - x0.astype(F32)).astype(F64)).astype(F32)
        wy = (1.0 - (y_real - y0.astype(F32)).astype(F64)).astype(F32)
    return x0, y0, wx, wy


def bilinear_sampler(data, grid):
    """out[n,c,h,w] = v00*wy*wx + v01*wy*(1.0-wx) + v10*(1.0-wy)*wx + v11*(1.0-wy)*(1.0-wx)
    with taps outside [0,Wi-1]x[0,Hi-1] contributing 0.  C++ promotion rules are kept:
    the first product is float32, every term containing ``1.0 - w`` is double, the sum is
    double, the store rounds to float32."""
    data = np.asarray(data, dtype=F32)
    N, C, Hi, Wi = data.shape
    No, two, Ho, Wo = grid.shape
    assert No == N and two == 2
    x0, y0, wx, wy = sampler_coords(grid, Hi, Wi)
    out = np.empty((N, C, Ho, Wo), dtype=F32)

    def tap(n, yy, xx):
        ok = (xx >= 0) & (xx <= Wi - 1) & (yy >= 0) & (yy <= Hi - 1)
        v = data[n][:, np.clip(yy, 0, Hi - 1), np.clip(xx, 0, Wi - 1)]
        return np.where(ok[None], v, F32(0))

    for n in range(N):
        v00 = tap(n, y0[n], x0[n])
        v01 = tap(n, y0[n], x0[n] + 1)
        v10 = tap(n, y0[n] + 1, x0[n])
        v11 = tap(n, y0[n] + 1, x0[n] + 1)
        wxf, wyf = wx[n][None], wy[n][None]
        omx = 1.0 - wxf.astype(F64)
        omy = 1.0 - wyf.astype(F64)
        t0 = ((v00 * wyf) * wxf).astype(F64)                    # float32 product chain
        t1 = (v01 * wyf).astype(F64) * omx
        t2 = (v10.astype(F64) * omy) * wxf.astype(F64)
        t3 = (v11.astype(F64) * omy) * omx
        out[n] = (((t0 + t1) + t2) + t3).astype(F32)
    return out


def warp(key, flow):
    """GridGenerator(warp) followed by BilinearSampler (SYM:571-572)."""
    return bilinear_sampler(key, grid_generator_warp(flow))


# --------------------------------------------------------------------------------------
# backward of a7 / a8 (needed by get_train_symbol SYM:305-307,319-321)
#     MXNet src/operator/bilinear_sampler.cc BilinearSamplerBackward, grid_generator-inl.h Backward
#     (kWarp branch).  Not vendored in the reference tree: parity unpinned, like the forward rows;
#     cross-checked against torch autograd of grid_sample(align_corners=True) in the tests.
# --------------------------------------------------------------------------------------
def bilinear_sampler_backward(data, grid, out_grad):
    """Returns (grad_data, grad_grid) for req = kWriteTo.

    Per output pixel p with top-left index (x0,y0) and top-left weights (wx,wy):
      grad_data[c, tap] += og[c,p] * {wy*wx, wy*(1-wx), (1-wy)*wx, (1-wy)*(1-wx)}   (taps inside the plane only)
      gy -= og[c,p] * (v01 - v11 + (v00 - v01 - v10 + v11) * wx)
      gx -= og[c,p] * (v10 - v11 + (v00 - v01 - v10 + v11) * wy)      (v = 0 outside the plane)
      grad_grid[1,p] = gy * (Hi-1)/2 ; grad_grid[0,p] = gx * (Wi-1)/2
    Sums are taken in float64 here (MXNet accumulates sequentially in float32; the C port
    ``lsfa_ref_bilinear_sampler_backward`` keeps that order) - the GPU gate is a tolerance."""
    data = np.asarray(data, dtype=F32)
    og = np.asarray(out_grad, dtype=F32)
    N, C, Hi, Wi = data.shape
    _, _, Ho, Wo = og.shape
    x0, y0, wx, wy = sampler_coords(grid, Hi, Wi)
    gdata = np.zeros((N, C, Hi * Wi), dtype=F64)
    ggrid = np.zeros((N, 2, Ho, Wo), dtype=F64)
    for n in range(N):
        wxd, wyd = wx[n].astype(F64).ravel(), wy[n].astype(F64).ravel()
        ogn = og[n].reshape(C, -1).astype(F64)
        vals = []
        for dy, dx, w in ((0, 0, wyd * wxd), (0, 1, wyd * (1.0 - wxd)), (1, 0, (1.0 - wyd) * wxd),
                          (1, 1, (1.0 - wyd) * (1.0 - wxd))):
            yy, xx = (y0[n] + dy).ravel(), (x0[n] + dx).ravel()
            ok = (xx >= 0) & (xx <= Wi - 1) & (yy >= 0) & (yy <= Hi - 1)
            q = np.clip(yy, 0, Hi - 1).astype(np.int64) * Wi + np.clip(xx, 0, Wi - 1)
            idx = np.nonzero(ok)[0]
            for c in range(C):
                np.add.at(gdata[n, c], q[idx], ogn[c, idx] * w[idx])
            v = np.where(ok[None], data[n].reshape(C, -1)[:, q], F32(0)).astype(F64)
            vals.append(v)
        v00, v01, v10, v11 = vals
        t = v00 - v01 - v10 + v11
        gy = -np.sum(ogn * (v01 - v11 + t * wxd[None]), axis=0)
        gx = -np.sum(ogn * (v10 - v11 + t * wyd[None]), axis=0)
        ggrid[n, 1] = (gy * (Hi - 1) / 2).reshape(Ho, Wo)
        ggrid[n, 0] = (gx * (Wi - 1) / 2).reshape(Ho, Wo)
    return gdata.reshape(N, C, Hi, Wi).astype(F32), ggrid.astype(F32)


def grid_generator_warp_backward(grad_grid):
    """gdata = grad / [(W-1)/2, (H-1)/2] broadcast over (N,2,H,W), float32."""
    g = np.asarray(grad_grid, dtype=F32)
    N, two, H, W = g.shape
    half_w = F32((F64(F32(W)) - 1.0) / 2.0)
    half_h = F32((F64(F32(H)) - 1.0) / 2.0)
    out = np.empty_like(g)
    with np.errstate(divide="ignore", invalid="ignore"):
        out[:, 0] = g[:, 0] / half_w
        out[:, 1] = g[:, 1] / half_h
    return out


def warp_backward(key, flow, out_grad):
    """Backward of SYM:306-307 / 571-572: (grad_key, grad_flow)."""
    gk, gg = bilinear_sampler_backward(key, grid_generator_warp(flow), out_grad)
    return gk, grid_generator_warp_backward(gg)


# --------------------------------------------------------------------------------------
# a9  scale multiply (SYM:308,470,680); a10 res_diff_ada + add (SYM:57-67,575-576)
# --------------------------------------------------------------------------------------
def scale_mul(warp_feat, scale_map):
    return np.asarray(warp_feat, F32) * np.asarray(scale_map, F32)


def rnet_conv0(res, weight, bias):
    """Shipped ``res_diff_ada``: one 1x1 convolution 3->C with bias (SYM:66).
    res (N,3,H,W), weight (C,3), bias (C,) -> (N,C,H,W) float32 (k = 0,1,2 in order)."""
    res = np.asarray(res, F32)
    weight = np.asarray(weight, F32).reshape(-1, 3)
    acc = np.zeros((res.shape[0], weight.shape[0]) + res.shape[2:], dtype=F32)
    for k in range(3):
        acc = acc + weight[None, :, k, None, None] * res[:, k:k + 1]
    return acc + np.asarray(bias, F32)[None, :, None, None]


# --------------------------------------------------------------------------------------
# a12/a13/a14 weights                                 SYM:94-148, 315, 476
# --------------------------------------------------------------------------------------
def softmax_pair(l_first, l_second):
    """``mx.sym.softmax(axis=0)`` over two stacked maps: exp(x - max) / sum(exp(x - max)),
    float32 (MXNet nn/softmax-inl.h)."""
    a, b = np.asarray(l_first, F32), np.asarray(l_second, F32)
    m = np.maximum(a, b)
    ea, eb = np.exp(a - m).astype(F32), np.exp(b - m).astype(F32)
    s = ea + eb
    return ea / s, eb / s


def l2norm_channel(e, eps=L2NORM_EPS):
    """``L2Normalization(mode='channel')``: e / sqrt(sum_c e^2 + eps)  (SYM:112-113)."""
    e = np.asarray(e, F32)
    nrm = np.sqrt(np.sum(e * e, axis=1, keepdims=True, dtype=F32) + F32(eps)).astype(F32)
    return e / nrm


def cosine_weight(embed_a, embed_b):
    """``compute_weight`` SYM:111-116: sum_c l2n(a) * l2n(b), keepdims -> (N,1,H,W)."""
    return np.sum(l2norm_channel(embed_a) * l2norm_channel(embed_b), axis=1, keepdims=True,
                  dtype=F32)


def aggregate_logits(warp_feat, cur_feat, logit_warp, logit_cur):
    """Tail of ``Nq_net`` SYM:104-108: softmax over the two sources, tiled over C,
    ``w1*warp + w2*conv``.  logits are (N,1,H,W) or (N,H,W)."""
    lw = np.asarray(logit_warp, F32).reshape(warp_feat.shape[0], 1, *warp_feat.shape[2:])
    lc = np.asarray(logit_cur, F32).reshape(lw.shape)
    w1, w2 = softmax_pair(lw, lc)
    return w1 * np.asarray(warp_feat, F32) + w2 * np.asarray(cur_feat, F32)


def aggregate_cosine(warp_feat, cur_feat, emb_warp, emb_cur):
    """``Fgfa_net`` SYM:132-148 given the two embedding tensors (embed net = dense convs,
    out of scope): l1 = cos(emb_warp, emb_cur), l2 = cos(emb_cur, emb_cur)."""
    l1 = cosine_weight(emb_warp, emb_cur)
    l2 = cosine_weight(emb_cur, emb_cur)
    return aggregate_logits(warp_feat, cur_feat, l1, l2)


def aggregate_mean(warp_feat, cur_feat):
    """SYM:315,476: 0.5 * (warp + conv_feat)."""
    return F32(0.5) * (np.asarray(warp_feat, F32) + np.asarray(cur_feat, F32))


def choose_feat(conv_feat, conv_feat_prop, eq_flag):
    """``ChooseFeat`` operator_py/choose_feat.py:23-31 batched: flag==1 keeps the raw
    current feature, else the propagated/aggregated one."""
    flag = np.asarray(eq_flag).astype(bool).reshape(-1, 1, 1, 1)
    return np.where(flag, conv_feat, conv_feat_prop)


def tile_as(data_content, n):
    """operator_py/tile_as.py:16-19."""
    return np.tile(data_content, (n, 1, 1, 1))


# --------------------------------------------------------------------------------------
# The fused op family (SURVEY.md section 8a "op family to build")
# --------------------------------------------------------------------------------------
def warp_scale_aggregate(key, flow, cur=None, scale_map=None, res=None, rnet_w=None,
                         rnet_b=None, weight_mode=W_NONE, logits=None, emb_warp=None,
                         emb_cur=None, bypass=None, key_index=None):
    """src0 = bilinear(key[key_index], grid_warp(flow)) [* scale_map] [+ rnet(res)];
    out = blend(src0, cur) by weight_mode; frames with bypass!=0 return cur."""
    key = np.asarray(key, F32)
    if key_index is not None:
        key = key[np.asarray(key_index)]
    src0 = warp(key, flow)
    if scale_map is not None:
        src0 = scale_mul(src0, scale_map)
    if res is not None:
        src0 = src0 + rnet_conv0(res, rnet_w, rnet_b)
    if weight_mode == W_NONE:
        out = src0
    elif weight_mode == W_ADD:                       # fuse_small_net 'add' SYM:236
        out = np.asarray(cur, F32) + src0
    elif weight_mode == W_MEAN:
        out = aggregate_mean(src0, cur)
    elif weight_mode == W_LOGITS:
        logits = np.asarray(logits, F32)
        out = aggregate_logits(src0, cur, logits[:, 0], logits[:, 1])
    elif weight_mode == W_COSINE:
        out = aggregate_cosine(src0, cur, emb_warp, emb_cur)
    else:
        raise ValueError(weight_mode)
    if bypass is not None and weight_mode != W_NONE:
        out = choose_feat(np.asarray(cur, F32), out, bypass)
    return out.astype(F32)


# --------------------------------------------------------------------------------------
# Reference graphs restated as compositions
# --------------------------------------------------------------------------------------
def cur_frame_path(feat_key, motion_vector, res_diff, rnet_w, rnet_b, small_net_feat):
    """``get_cur_test_symbol`` SYM:570-586 as shipped (yaml:49-60): warp(MV) +
    rnet_conv0(res) then ``fuse_small_net`` 'add' with the current-frame feature."""
    return warp_scale_aggregate(feat_key, motion_vector, cur=small_net_feat, res=res_diff,
                                rnet_w=rnet_w, rnet_b=rnet_b, weight_mode=W_ADD)


def key_frame_path_nq(feat_key_old, flow, scale_map, conv_feat, nq_logits, is_first):
    """``get_key_test_symbol`` SYM:467-477 with Nq weights (shipped)."""
    return warp_scale_aggregate(feat_key_old, flow, cur=conv_feat, scale_map=scale_map,
                                weight_mode=W_LOGITS, logits=nq_logits, bypass=is_first)


def key_frame_path_fgfa(feat_key_old, flow, scale_map, conv_feat, emb_warp, emb_cur, is_first):
    """Same with the cosine-embedding aggregator (SYM:473-474)."""
    return warp_scale_aggregate(feat_key_old, flow, cur=conv_feat, scale_map=scale_map,
                                weight_mode=W_COSINE, emb_warp=emb_warp, emb_cur=emb_cur,
                                bypass=is_first)


def batch_path(conv_feat_key, flow, scale_map):
    """``get_batch_test_symbol`` SYM:675-680: key feature tiled over the batch, warped by
    per-frame flow, times scale map."""
    n = flow.shape[0]
    return warp_scale_aggregate(conv_feat_key, flow, scale_map=scale_map,
                                key_index=np.zeros(n, dtype=np.int64))


# --------------------------------------------------------------------------------------
# The dense convolutions around the path (library GEMMs in the product; here only so the exact
# two-phase key-frame graphs can be checked end to end at small sizes)
# --------------------------------------------------------------------------------------
def conv2d(x, w, b, pad=0):
    """mx.sym.Convolution, stride 1: x (N,Ci,H,W), w (Co,Ci,k,k), b (Co,) -> (N,Co,H,W) (float64 accumulate)."""
    x = np.asarray(x, F64)
    w = np.asarray(w, F64)
    k = w.shape[2]
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    N, Ci, Hp, Wp = xp.shape
    H, W = Hp - k + 1, Wp - k + 1
    out = np.zeros((N, w.shape[0], H, W), F64)
    for dy in range(k):
        for dx in range(k):
            out += np.einsum("nchw,oc->nohw", xp[:, :, dy:dy + H, dx:dx + W], w[:, :, dy, dx])
    return (out + np.asarray(b, F64)[None, :, None, None]).astype(F32)


def embed_net(x, w1, b1, w2, b2, w3, b3):
    """get_embednet SYM:118-130."""
    x = np.maximum(conv2d(x, w1, b1), 0)
    x = np.maximum(conv2d(x, w2, b2, pad=1), 0)
    return conv2d(x, w3, b3)


def nq_net(x, w1, b1, w2, b2, w3, b3):
    """Nq_net convs SYM:97-101."""
    x = np.maximum(conv2d(x, w1, b1, pad=1), 0)
    x = np.maximum(conv2d(x, w2, b2), 0)
    return conv2d(x, w3, b3)


def key_frame_fgfa_full(feat_key_old, flow, scale_map, conv_feat, embed_params, is_first=None):
    """get_key_test_symbol, Fgfa branch, convolutions included (SYM:468-470,473-474,132-148,477)."""
    wp = scale_mul(warp(feat_key_old, flow), scale_map)
    n = conv_feat.shape[0]
    emb = embed_net(np.concatenate([conv_feat, wp], 0), *embed_params)
    out = aggregate_cosine(wp, conv_feat, emb[n:], emb[:n])
    return choose_feat(np.asarray(conv_feat, F32), out, is_first) if is_first is not None else out


def key_frame_nq_full(feat_key_old, flow, scale_map, conv_feat, nq_params, is_first=None):
    """get_key_test_symbol, Nq branch (shipped), convolutions included (SYM:468-472,94-109,477)."""
    wp = scale_mul(warp(feat_key_old, flow), scale_map)
    n = conv_feat.shape[0]
    q = nq_net(np.concatenate([wp, conv_feat], 0), *nq_params)
    out = aggregate_logits(wp, conv_feat, q[:n], q[n:])
    return choose_feat(np.asarray(conv_feat, F32), out, is_first) if is_first is not None else out


def warp_scale_aggregate_backward(out_grad, key, flow, cur=None, scale_map=None, res=None, rnet_w=None, rnet_b=None,
                                  weight_mode=W_NONE, logits=None, bypass=None):
    """Gradients of warp_scale_aggregate (modes none/add/mean/logits) w.r.t. key, flow, scale_map, cur, logits, res,
    rnet_w, rnet_b from d/d(out) - float64 chain rule on top of the a7/a8 backward restatement (get_train_symbol
    SYM:306-338 differentiates exactly these operators).  Returns a dict (entries only for inputs that exist)."""
    g = np.asarray(out_grad, F64)
    N, C, H, W = g.shape
    wp = np.asarray(warp(key, flow), F64)
    sc = np.ones_like(wp) if scale_map is None else np.asarray(scale_map, F64)
    src0 = wp * sc
    if res is not None:
        rw = np.asarray(rnet_w, F64).reshape(C, 3)
        src0 = src0 + np.einsum("cj,njhw->nchw", rw, np.asarray(res, F64)) + np.asarray(rnet_b, F64).reshape(1, C, 1, 1)
    ww = np.ones((N, 1, H, W))
    wc = np.zeros((N, 1, H, W))
    if weight_mode == W_ADD:
        wc[:] = 1.0
    elif weight_mode == W_MEAN:
        ww[:] = 0.5
        wc[:] = 0.5
    elif weight_mode == W_LOGITS:
        a, b = softmax_pair(logits[:, 0:1], logits[:, 1:2])
        ww, wc = np.asarray(a, F64), np.asarray(b, F64)
    live = np.ones((N, 1, 1, 1)) if bypass is None else (np.asarray(bypass).reshape(N, 1, 1, 1) == 0).astype(F64)
    out = {}
    gs0 = ww * g * live                                   # d/d(src0)
    if cur is not None and weight_mode != W_NONE:
        out["cur"] = (wc * g * live + g * (1.0 - live)).astype(F32)
    if scale_map is not None:
        out["scale"] = (gs0 * wp).astype(F32)
    gk, gf = warp_backward(key, flow, (gs0 * sc).astype(F32))
    out["key"], out["flow"] = gk, gf
    if weight_mode == W_LOGITS:
        t1 = np.sum(g * src0, axis=1, keepdims=True)
        t2 = np.sum(g * np.asarray(cur, F64), axis=1, keepdims=True)
        d = ww * wc * (t1 - t2) * live
        out["logits"] = np.concatenate([d, -d], axis=1).astype(F32)
    if res is not None:
        out["res"] = np.einsum("cj,nchw->njhw", rw, gs0).astype(F32)
        out["rnet_w"] = np.einsum("nchw,njhw->cj", gs0, np.asarray(res, F64)).astype(F32)
        out["rnet_b"] = gs0.sum(axis=(0, 2, 3)).astype(F32)
    return out


# --------------------------------------------------------------------------------------
# The graph switches around the non-key step that the shipped yaml does not select (SYM:57-67, 209-272, 326-328),
# restated so that lsfa_b200.graphs.cur_frame_step can be checked for every one of them.  float64 convolutions.
# --------------------------------------------------------------------------------------
def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-np.asarray(x, F64)))


def res_diff_ada(res_diff, convs):
    """SYM:57-67 without BatchNorm: `convs` = [(w,b) 3x3 + ReLU] * rnet_num_conv + [(w,b) 1x1 -> 1024]."""
    x = np.asarray(res_diff, F32)
    for (w, b) in convs[:-1]:
        x = np.maximum(conv2d(x, w, b, pad=1), 0)
    w, b = convs[-1]
    return conv2d(x, w, b)


def fuse_small_net(warp_feat, cur_small, params, fuse_type):
    """SYM:229-272 (no BatchNorm, no cur_scale): cur_small = the small net's feature (N,256,H,W)."""
    wp = np.asarray(warp_feat, F32)
    if fuse_type == "add":                                   # SYM:230-236 (shipped)
        return conv2d(cur_small, *params["fuse_reduce_add"], pad=1) + wp
    if fuse_type == "addv2":                                 # SYM:237-244
        c = np.maximum(conv2d(cur_small, *params["fuse_reduce_add_conv1"], pad=1), 0)
        return conv2d(c, *params["fuse_reduce_add_conv2"]) + wp
    if fuse_type in ("concat", "concatv1"):                  # SYM:245-262
        c1 = conv2d(cur_small, *params["fuse_reduce_c1"], pad=1)
        c2 = conv2d(wp, *params["fuse_reduce_c2"], pad=1)
        cat = conv2d(np.concatenate([c2, c1], axis=1), *params["fuse_reduce"], pad=1)
        if fuse_type == "concat":
            return cat
        cat = np.maximum(cat, 0)
        sfeat = cat.mean(axis=(2, 3), keepdims=True, dtype=F64).astype(F32)
        sfeat = np.maximum(conv2d(sfeat, *params["s_feat_conv1"]), 0)
        sfeat = _sigmoid(conv2d(sfeat, *params["s_feat_conv2"])).astype(F32)
        return cat * sfeat + cat
    if fuse_type == "concatv2":                              # SYM:263-272
        c = conv2d(cur_small, *params["fuse_reduce_c1"], pad=1)
        cat = np.concatenate([wp, c], axis=1)
        sfeat = cat.mean(axis=(2, 3), keepdims=True, dtype=F64).astype(F32)
        sfeat = np.maximum(conv2d(sfeat, *params["s_feat_conv1"]), 0)
        sfeat = _sigmoid(conv2d(sfeat, *params["s_feat_conv2"])).astype(F32)
        return c * sfeat + wp
    raise ValueError(fuse_type)


def cur_frame_step(feat_key, flow, res_diff, rnet_convs, cur_small, small_params, fuse_type="add", res_fuse="add",
                   fuse_downsample=None):
    """get_cur_test_symbol's tail (SYM:570-586) with every switch: warp(key, MV); res_diff_ada; fuse_type 'add' | 'concat'
    (SYM:326-328: 1x1 conv over Concat_1(warp, res_diff)); fuse_small_net."""
    wp = warp(feat_key, flow)
    rd = res_diff_ada(res_diff, rnet_convs)
    if res_fuse == "add":
        wp = wp + rd
    else:
        wp = conv2d(np.concatenate([wp, rd], axis=1), *fuse_downsample)
    return fuse_small_net(wp, cur_small, small_params, fuse_type)


# --------------------------------------------------------------------------------------
# Stated-tolerance bf16 variant of the two networks (the tensor-core kernels of SURVEY 8f rank 2):
# same graphs as embed_net / nq_net above with the product's rounding points made explicit -
# inputs, convolution weights and the two hidden activations of the embedding net are rounded to
# bf16; every accumulation, the biases, the Nq tail and compute_weight stay float32/float64.
# --------------------------------------------------------------------------------------
def embed_cosine_logits_bf16(x, w1, b1, w2, b2, w3, b3):
    """x = Concat_0(conv_feat, warp_feat) (2N,C,H,W) -> logits (N,2,H,W): [n,0] = <e_warp^, e_cur^>, [n,1] = <e_cur^, e_cur^>
    (get_embednet SYM:118-130, compute_weight SYM:111-116, Fgfa_net SYM:133-139)."""
    r = bf16_round
    n = x.shape[0] // 2
    h1 = r(np.maximum(conv2d(r(x), r(w1), b1), 0))
    h2 = r(np.maximum(conv2d(h1, r(w2), b2, pad=1), 0))
    e = conv2d(h2, r(w3), b3)
    l1 = cosine_weight(e[n:], e[:n])
    l2 = cosine_weight(e[:n], e[:n])
    return np.concatenate([l1, l2], axis=1).astype(F32)


def nq_logits_bf16(x, w1, b1, w2, b2, w3, b3):
    """x = Concat_0(warp_feat, conv_feat) (2N,C,H,W) -> logits (N,2,H,W): [n,0] warp, [n,1] current (Nq_net SYM:95-101)."""
    r = bf16_round
    n = x.shape[0] // 2
    q = np.maximum(conv2d(r(x), r(w1), b1, pad=1), 0)
    q = np.maximum(conv2d(q, w2, b2), 0)
    q = conv2d(q, w3, b3)
    return np.concatenate([q[:n], q[n:]], axis=1).astype(F32)


# --------------------------------------------------------------------------------------
# Algorithmic bytes - used by bench.py and DESIGN.md
# --------------------------------------------------------------------------------------
def algorithmic_bytes_per_frame(C, H, W, feat_bytes=4, variant="V2", E=2048):
    F = C * H * W * feat_bytes
    HW = H * W
    if variant == "V0":
        return 2 * F + 8 * HW
    if variant == "V0p":
        return 2 * F + 32 * HW
    if variant == "V1":
        return 3 * F + 32 * HW + 12 * HW + 16384
    if variant == "V2":
        return 4 * F + 32 * HW + 8 * HW
    if variant == "V3":
        return 4 * F + 32 * HW + 8 * HW + 2 * E * HW * feat_bytes
    raise ValueError(variant)


# --------------------------------------------------------------------------------------
# Synthetic inputs
# --------------------------------------------------------------------------------------
def synth_raw_mv(rng, n, h=600, w=1000, max_px=32, zero_frac=0.5):
    """int32 (n,h,w,2), constant per 16x16 macroblock, ``zero_frac`` of blocks static."""
    bh, bw = -(-h // 16), -(-w // 16)
    blk = rng.integers(-max_px, max_px + 1, size=(n, bh, bw, 2), dtype=np.int32)
    blk[rng.random((n, bh, bw)) < zero_frac] = 0
    full = np.repeat(np.repeat(blk, 16, axis=1), 16, axis=2)
    return np.ascontiguousarray(full[:, :h, :w])


def synth_features(rng, shape):
    """post-ReLU-like: max(N(0,1),0) float32."""
    return np.maximum(rng.standard_normal(shape, dtype=F32), F32(0))


def synth_scale_map(rng, shape):
    return (F32(1) + F32(0.1) * rng.standard_normal(shape, dtype=F32)).astype(F32)


def bf16_round(x):
    """Round-to-nearest-even float32 -> bfloat16 -> float32 (for the bf16 parity gate)."""
    u = np.asarray(x, F32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(F32).reshape(np.shape(x))
