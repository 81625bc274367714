"""Plain-numpy restatements of every arithmetic op on the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The formulas are the published
behaviour of OpenCV 4.x (`opencv-python-headless` 4.13.0.92 here) for the
exact calls the reference makes; they were pinned empirically bit-for-bit
against that binary and are re-asserted against it in
tests/test_oracle_vs_cv2.py.  Citations are relative to /root/reference.

Everything here is slow, single-purpose and deliberately simple: integer ops
in int32/int64, float ops forced to float32 step by step where OpenCV rounds
step by step.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32

# --------------------------------------------------------------------------
# A.1  cv2.cvtColor(BGR2GRAY)   shape_analyzer.pyc src L18, shadow_analyzer.pyc
#      src L8, app/services/analyze/text_extractor.py:27, font_detector.py:28
# --------------------------------------------------------------------------
GRAY_B, GRAY_G, GRAY_R, GRAY_SHIFT = 3735, 19235, 9798, 15


def bgr2gray(bgr: np.ndarray) -> np.ndarray:
    """Y = (3735*B + 19235*G + 9798*R + 16384) >> 15  (Q15 fixed point)."""
    b = bgr[..., 0].astype(np.int32)
    g = bgr[..., 1].astype(np.int32)
    r = bgr[..., 2].astype(np.int32)
    return ((GRAY_B * b + GRAY_G * g + GRAY_R * r + (1 << (GRAY_SHIFT - 1))) >> GRAY_SHIFT).astype(np.uint8)


def bgr2rgb(bgr: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(BGR2RGB): pure channel swap (color_extractor.py:151)."""
    return np.ascontiguousarray(bgr[..., ::-1])


# --------------------------------------------------------------------------
# A.2  cv2.GaussianBlur(src, (5,5), 0) on u8   shape_analyzer.pyc src L21,
#      shadow_analyzer.pyc src L9, image_transformer.pyc src L103 (3-channel)
# --------------------------------------------------------------------------
def gaussian_blur5(src: np.ndarray) -> np.ndarray:
    """Separable [1,4,6,4,1]x[1,4,6,4,1], BORDER_REFLECT_101, (sum+128)>>8."""
    if src.ndim == 3:
        return np.stack([gaussian_blur5(src[..., c]) for c in range(src.shape[2])], axis=-1)
    h, w = src.shape
    p = src.astype(np.int32)
    if h == 1 and w == 1:
        return src.copy()
    py = _reflect101_pad(p, 2, axis=0)
    px = _reflect101_pad(py, 2, axis=1)
    k = (1, 4, 6, 4, 1)
    hs = sum(k[i] * px[:, i:i + w] for i in range(5))
    vs = sum(k[i] * hs[i:i + h, :] for i in range(5))
    return ((vs + 128) >> 8).astype(np.uint8)


def _reflect101_idx(i: np.ndarray, n: int) -> np.ndarray:
    """OpenCV borderInterpolate(BORDER_REFLECT_101) for any offset / any n."""
    if n == 1:
        return np.zeros_like(i)
    period = 2 * (n - 1)
    i = np.mod(i, period)
    return np.where(i >= n, period - i, i)


def _reflect101_pad(a: np.ndarray, r: int, axis: int) -> np.ndarray:
    n = a.shape[axis]
    idx = _reflect101_idx(np.arange(-r, n + r), n)
    return np.take(a, idx, axis=axis)


# --------------------------------------------------------------------------
# A.3  cv2.Canny(img, 50, 150)  (aperture 3, L1 gradient)  shape_analyzer.pyc L24
# --------------------------------------------------------------------------
TG22 = 13573  # tan(22.5 deg) in Q15


def sobel3(g: np.ndarray):
    """3x3 Sobel dx, dy as int32 with BORDER_REPLICATE."""
    p = np.pad(g.astype(np.int32), 1, mode="edge")
    h, w = g.shape
    s = lambda dy, dx: p[1 + dy:1 + dy + h, 1 + dx:1 + dx + w]  # noqa: E731
    dx = (s(-1, 1) + 2 * s(0, 1) + s(1, 1)) - (s(-1, -1) + 2 * s(0, -1) + s(1, -1))
    dy = (s(1, -1) + 2 * s(1, 0) + s(1, 1)) - (s(-1, -1) + 2 * s(-1, 0) + s(-1, 1))
    return dx, dy


def canny_nms(g: np.ndarray, low: int, high: int):
    """Returns (weak, strong) boolean maps: weak = survives NMS with mag>low
    (strong pixels included), strong = weak & mag>high."""
    dx, dy = sobel3(g)
    mag = np.abs(dx) + np.abs(dy)
    h, w = g.shape
    mp = np.pad(mag, 1, mode="constant")  # zero magnitude outside the image
    m = lambda oy, ox: mp[1 + oy:1 + oy + h, 1 + ox:1 + ox + w]  # noqa: E731
    x = np.abs(dx).astype(np.int64)
    y = np.abs(dy).astype(np.int64) << 15
    tg22x = x * TG22
    tg67x = tg22x + (x << 16)
    horiz = y < tg22x
    vert = (~horiz) & (y > tg67x)
    diag = ~(horiz | vert)
    keep_h = (mag > m(0, -1)) & (mag >= m(0, 1))
    keep_v = (mag > m(-1, 0)) & (mag >= m(1, 0))
    s_neg = (dx ^ dy) < 0
    keep_d = np.where(s_neg, (mag > m(-1, 1)) & (mag > m(1, -1)), (mag > m(-1, -1)) & (mag > m(1, 1)))
    weak = (mag > low) & ((horiz & keep_h) | (vert & keep_v) | (diag & keep_d))
    strong = weak & (mag > high)
    return weak, strong


def hysteresis(weak: np.ndarray, strong: np.ndarray) -> np.ndarray:
    """8-connected components of `weak` that contain a `strong` pixel."""
    from scipy import ndimage

    lab, n = ndimage.label(weak, structure=np.ones((3, 3)))
    keep = np.zeros(n + 1, bool)
    keep[np.unique(lab[strong])] = True
    keep[0] = False
    return keep[lab]


def hysteresis_iterative(weak: np.ndarray, strong: np.ndarray) -> np.ndarray:
    """Same result by fixed-point propagation (no scipy); used to cross-check."""
    e = strong.copy()
    while True:
        p = np.pad(e, 1)
        grown = np.zeros_like(e)
        for dy in (0, 1, 2):
            for dx in (0, 1, 2):
                grown |= p[dy:dy + e.shape[0], dx:dx + e.shape[1]]
        new = weak & grown
        if np.array_equal(new, e):
            return e
        e = new


def canny(g: np.ndarray, low: int = 50, high: int = 150) -> np.ndarray:
    weak, strong = canny_nms(g, low, high)
    return (hysteresis(weak, strong) * 255).astype(np.uint8)


# --------------------------------------------------------------------------
# A.4  cv2.dilate(edges, ones((3,3)), iterations=1)   shape_analyzer.pyc L27-28
# --------------------------------------------------------------------------
def dilate3(src: np.ndarray) -> np.ndarray:
    """3x3 max; neighbours outside the image are ignored."""
    h, w = src.shape
    p = np.pad(src, 1, mode="constant")
    out = np.zeros_like(src)
    for dy in (0, 1, 2):
        for dx in (0, 1, 2):
            out = np.maximum(out, p[dy:dy + h, dx:dx + w])
    return out


# --------------------------------------------------------------------------
# A.5  cv2.adaptiveThreshold(src,255,GAUSSIAN_C,BINARY_INV,11,2)
#      shadow_analyzer.pyc src L17-18, app/services/analyze/font_detector.py:31-35
# --------------------------------------------------------------------------
def gaussian_kernel_f32(ksize: int = 11) -> np.ndarray:
    """float32(cv2.getGaussianKernel(ksize, 0)): sigma = 0.3*((k-1)*0.5-1)+0.8,
    doubles exp(-x^2/(2 sigma^2)) normalised by their sum, then cast to f32."""
    sigma = 0.3 * ((ksize - 1) * 0.5 - 1) + 0.8
    scale2x = -0.5 / (sigma * sigma)
    xs = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(scale2x * xs * xs)
    k = k / k.sum()
    return k.astype(f32)


def _fma(a: np.ndarray, b, c: np.ndarray) -> np.ndarray:
    # float32 fused multiply-add: exact product and sum in float64 (24-bit x
    # 24-bit products are exact in 53 bits; the sum is then rounded once).
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(f32)


def _mad(a: np.ndarray, b, c: np.ndarray) -> np.ndarray:
    # separately rounded float32 multiply and add
    return ((a.astype(f32) * f32(b)).astype(f32) + c.astype(f32)).astype(f32)


def gauss11_f32(src_u8: np.ndarray) -> np.ndarray:
    """11x11 sigma=2 Gaussian mean in float32 with the op order of the installed OpenCV binary
    (cv2 4.13.0.92: what adaptiveThreshold's internal GaussianBlur(CV_32F, BORDER_REPLICATE) returns, compared
    bit for bit on widths 2..49, 101, 203, 330, 1001, 1366, 1444, 1921, heights >= 2, 1 and 8 threads):
      row pass    = left-to-right chain  s = k0*x[-5];  s = op(x[i-5], k[i], s), i = 1..10
      column pass = symmetric-pair chain v = k5*r[0];   v = op(r[+i] + r[-i], k[5+i], v), i = 1..5
    with op = fused multiply-add in the vectorised columns and separately rounded multiply + add in the tail:
      row pass:    columns < n4 fused (8-wide, then one 4-wide vector); the last w % 4 columns run the scalar loop,
                   whose compiled code is mul + add for taps 1..8 and fused for taps 9 and 10;
      column pass: columns < n8 fused; the last w % 8 columns mul + add,
    n8 = w - w % 8, n4 = n8 + 4 if w - n8 >= 4 else n8.  Widths that are a multiple of 8 (1920, 3840) have no tail.
    (Images with a single row or column take another OpenCV code path and are not covered.)"""
    k = gaussian_kernel_f32(11)
    h, w = src_u8.shape
    n8 = w - w % 8
    n4 = n8 + 4 if w - n8 >= 4 else n8
    p = np.pad(src_u8.astype(f32), 5, mode="edge")
    s = (k[0] * p[:, 0:w]).astype(f32)
    for i in range(1, 11):
        x = p[:, i:i + w]
        t = _fma(x, k[i], s)
        if n4 < w and i <= 8:
            t[:, n4:] = _mad(x[:, n4:], k[i], s[:, n4:])
        s = t
    v = (k[5] * s[5:5 + h]).astype(f32)
    for i in range(1, 6):
        pr = (s[5 + i:5 + i + h] + s[5 - i:5 - i + h]).astype(f32)
        t = _fma(pr, k[5 + i], v)
        if n8 < w:
            t[:, n8:] = _mad(pr[:, n8:], k[5 + i], v[:, n8:])
        v = t
    return v


def adaptive_threshold(src_u8: np.ndarray, c: int = 2) -> np.ndarray:
    mean = np.rint(gauss11_f32(src_u8)).clip(0, 255).astype(np.uint8)
    return np.where(src_u8.astype(np.int32) - mean.astype(np.int32) <= -c, 255, 0).astype(np.uint8)


# --------------------------------------------------------------------------
# A.6  cv2.threshold(gray,0,255,BINARY+OTSU)   text_extractor.py:40, then
#      `if np.mean(binary) > 127: bitwise_not`   text_extractor.py:43-44
# --------------------------------------------------------------------------
def otsu_threshold(gray: np.ndarray) -> int:
    hist = np.bincount(gray.ravel(), minlength=256).astype(np.int64)
    return otsu_from_hist(hist, gray.size)


def otsu_from_hist(hist: np.ndarray, n: int) -> int:
    scale = 1.0 / n
    mu = 0.0
    for i in range(256):
        mu += i * float(hist[i])
    mu *= scale
    mu1 = 0.0
    q1 = 0.0
    max_sigma = 0.0
    max_val = 0
    eps = float(np.finfo(np.float32).eps)  # FLT_EPSILON (probed: 1 px of 16.7M is skipped, 3 px are not)
    for i in range(256):
        p_i = float(hist[i]) * scale
        mu1 *= q1
        q1 += p_i
        q2 = 1.0 - q1
        if min(q1, q2) < eps or max(q1, q2) > 1.0 - eps:
            continue
        mu1 = (mu1 + i * p_i) / q1
        mu2 = (mu - q1 * mu1) / q2
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2)
        if sigma > max_sigma:
            max_sigma = sigma
            max_val = i
    return max_val


def otsu_binary(gray: np.ndarray):
    t = otsu_threshold(gray)
    return t, np.where(gray > t, 255, 0).astype(np.uint8)


def text_mask(bgr: np.ndarray) -> np.ndarray:
    """TextExtractor.preprocess_image for images that skip the tiny-image
    INTER_CUBIC upscale (text_extractor.py:15-46 with h>=30 and w>=100)."""
    gray = bgr2gray(bgr) if bgr.ndim == 3 else bgr
    _, binary = otsu_binary(gray)
    if np.mean(binary) > 127:
        binary = 255 - binary
    return binary


# --------------------------------------------------------------------------
# A.7  cv2.resize(INTER_AREA), downscale   utils.py:127, image_processor.py:112,
#      image_transformer.pyc src L53, L170-173
# --------------------------------------------------------------------------
def area_tab(ssize: int, dsize: int, scale: float):
    """OpenCV computeResizeAreaTab: list of (dst index, src index, f32 weight)."""
    tab = []
    for d in range(dsize):
        fsx1 = d * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1 = math.ceil(fsx1)
        sx2 = math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((d, sx1 - 1, f32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((d, sx, f32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((d, sx2, f32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def resize_area(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """INTER_AREA down-scale (dw<=sw and dh<=sh), all three OpenCV code paths."""
    sh, sw = src.shape[:2]
    if (dw, dh) == (sw, sh):
        return src.copy()
    scale_x = 1.0 / (dw / sw)
    scale_y = 1.0 / (dh / sh)
    isx, isy = int(round(scale_x)), int(round(scale_y))
    is_int = abs(scale_x - isx) < np.finfo(np.float64).eps and abs(scale_y - isy) < np.finfo(np.float64).eps
    if is_int:
        s = src.astype(np.int64).reshape((dh, isy, dw, isx) + src.shape[2:]).sum(axis=(1, 3))
        if isx == 2 and isy == 2:
            return ((s + 2) >> 2).astype(np.uint8)
        area = isx * isy
        # OpenCV: saturate_cast<uchar>(float(sum) * float(1/area)) -> rint
        v = (s.astype(f32) * f32(1.0 / area)).astype(f32)
        return np.rint(v).clip(0, 255).astype(np.uint8)
    xt = area_tab(sw, dw, scale_x)
    yt = area_tab(sh, dh, scale_y)
    s = src.astype(f32)
    buf = np.zeros((sh, dw) + src.shape[2:], f32)
    for (d, sx, a) in xt:
        buf[:, d] = (buf[:, d] + (s[:, sx] * a).astype(f32)).astype(f32)
    out = np.zeros((dh, dw) + src.shape[2:], f32)
    first = set()
    for (d, sy, b) in yt:
        t = (b * buf[sy]).astype(f32)
        if d not in first:
            out[d] = t
            first.add(d)
        else:
            out[d] = (out[d] + t).astype(f32)
    return np.rint(out).clip(0, 255).astype(np.uint8)


def linear_tab(ssize: int, dsize: int, clamp_weights: bool):
    """Per-axis source indices and 11-bit fixed-point weights of cv2's INTER_LINEAR u8 path.
    OpenCV clamps the fraction to 0 at the borders for the x axis only; for y it keeps the
    fraction and clips the two row indices instead (they may then be the same row)."""
    scale = float(ssize) / dsize
    idx = np.zeros(dsize, np.int64)
    wts = np.zeros((dsize, 2), np.int32)
    for d in range(dsize):
        fx = f32((d + 0.5) * scale - 0.5)
        sx = int(np.floor(fx))
        fx = f32(fx - f32(sx))
        if clamp_weights:
            if sx < 0:
                fx, sx = f32(0), 0
            if sx >= ssize - 1:
                fx, sx = f32(0), ssize - 1
        idx[d] = sx
        wts[d, 0] = int(np.rint(f32(f32(1.0) - fx) * f32(2048)))
        wts[d, 1] = int(np.rint(fx * f32(2048)))
    return idx, wts


def resize_linear(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR) on u8 -- the `performance`
    preprocessing mode, app/services/analyze/utils.py:136-143.  Fixed point: horizontal
    H = S[x0]*a0 + S[x1]*a1 (weights scaled by 2048), vertical
    (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2; exact 2x2 ratios take INTER_AREA."""
    sh, sw = src.shape[:2]
    if sw == 2 * dw and sh == 2 * dh:
        return resize_area(src, dw, dh)
    xi, xa = linear_tab(sw, dw, True)
    yi, ya = linear_tab(sh, dh, False)
    S = src.astype(np.int32)
    if S.ndim == 2:
        S = S[:, :, None]
    x1 = np.minimum(xi + 1, sw - 1)
    H = S[:, xi, :] * xa[:, 0][None, :, None] + S[:, x1, :] * xa[:, 1][None, :, None]
    y0, y1 = np.clip(yi, 0, sh - 1), np.clip(yi + 1, 0, sh - 1)
    b0, b1 = ya[:, 0][:, None, None], ya[:, 1][:, None, None]
    out = (((b0 * (H[y0] >> 4)) >> 16) + ((b1 * (H[y1] >> 4)) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out if src.ndim == 3 else out[:, :, 0]


def lanczos4_coeffs(x) -> np.ndarray:
    """cv::interpolateLanczos4: 8 float32 coefficients for the fraction x (float32)."""
    import math

    s45 = 0.70710678118654752440084436210485
    cs = ((1, 0), (-s45, -s45), (0, 1), (s45, -s45), (-1, 0), (s45, s45), (0, -1), (-s45, s45))
    x = f32(x)
    coeffs = np.zeros(8, f32)
    total = f32(0)
    y0 = -float(f32(x + f32(3))) * math.pi * 0.25
    s0, c0 = math.sin(y0), math.cos(y0)
    for i in range(8):
        y0_ = f32(f32(x + f32(3)) - f32(i))
        if abs(float(y0_)) >= 1e-6:
            y = -float(y0_) * math.pi * 0.25
            coeffs[i] = f32((cs[i][0] * s0 + cs[i][1] * c0) / (y * y))
        else:
            coeffs[i] = f32(1e30)
        total = f32(total + coeffs[i])
    total = f32(f32(1.0) / total)
    return (coeffs * total).astype(f32)


def lanczos4_tab(ssize: int, dsize: int):
    scale = float(ssize) / dsize
    first = np.zeros(dsize, np.int64)
    wts = np.zeros((dsize, 8), np.int64)
    for d in range(dsize):
        fx = f32((d + 0.5) * scale - 0.5)
        sx = int(np.floor(fx))
        fx = f32(fx - f32(sx))
        first[d] = sx - 3
        wts[d] = np.clip(np.rint(lanczos4_coeffs(fx) * f32(2048)), -32768, 32767).astype(np.int64)
    return first, wts


def resize_lanczos4(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LANCZOS4) on u8 -- the `high_quality`
    preprocessing mode, app/services/analyze/utils.py:128-135.  8 taps per axis with weights
    scaled by 2048 (short), indices clipped into the image, dst = sat_u8((sum + 2^21) >> 22)."""
    sh, sw = src.shape[:2]
    xf, xa = lanczos4_tab(sw, dw)
    yf, ya = lanczos4_tab(sh, dh)
    S = src.astype(np.int64)
    if S.ndim == 2:
        S = S[:, :, None]
    H = np.zeros((sh, dw, S.shape[2]), np.int64)
    for k in range(8):
        H += S[:, np.clip(xf + k, 0, sw - 1), :] * xa[:, k][None, :, None]
    out = np.zeros((dh, dw, S.shape[2]), np.int64)
    for k in range(8):
        out += H[np.clip(yf + k, 0, sh - 1)] * ya[:, k][:, None, None]
    out = np.clip((out + (1 << 21)) >> 22, 0, 255).astype(np.uint8)
    return out if src.ndim == 3 else out[:, :, 0]


def performance_resize_shape(h: int, w: int, max_dim: int = 1000):
    """utils.py:136-143: LINEAR to max-dim 1000 -> (new_w, new_h) or None."""
    if max(h, w) <= max_dim:
        return None
    scale = max_dim / max(h, w)
    return int(w * scale), int(h * scale)


def auto_resize_shape(h: int, w: int, max_dim: int = 2000):
    """`auto` preprocessing output size, utils.py:120-126 (Python float math)."""
    if max(h, w) <= max_dim:
        return h, w
    scale = max_dim / max(h, w)
    return int(h * scale), int(w * scale)


def fit_resize_shape(h: int, w: int, max_width: int = 1920, max_height: int = 1080):
    """ImageProcessor.resize_image output size, image_processor.py:106-110."""
    s = min(max_width / w, max_height / h, 1)
    return int(h * s), int(w * s)


# --------------------------------------------------------------------------
# A.9  cv2.convertScaleAbs(x, alpha, beta=0)   image_transformer.pyc src L139-142
# --------------------------------------------------------------------------
def convert_scale_abs(src: np.ndarray, alpha: float, beta: float = 0.0) -> np.ndarray:
    v = (src.astype(f32) * f32(alpha)).astype(f32)
    if beta != 0.0:
        v = (v + f32(beta)).astype(f32)
    return np.rint(np.abs(v)).clip(0, 255).astype(np.uint8)


def adjust_brightness_contrast(img: np.ndarray, brightness: float = 1.0, contrast: float = 1.0) -> np.ndarray:
    return convert_scale_abs(convert_scale_abs(img, brightness), contrast)


# --------------------------------------------------------------------------
# a4/a5  noise + np.unique(axis=0)   color_extractor.py:220-225, :177
# --------------------------------------------------------------------------
def apply_noise(rgb_pixels: np.ndarray, noise_i8: np.ndarray) -> np.ndarray:
    """clip(int32(pixels) + int8 noise, 0, 255).astype(u8)  (color_extractor.py:225)."""
    return np.clip(rgb_pixels.astype(np.int32) + noise_i8.astype(np.int32), 0, 255).astype(np.uint8)


def make_noise(shape, seed: int) -> np.ndarray:
    """The reference's noise tensor under np.random.seed(seed) (color_extractor.py:224)."""
    st = np.random.get_state()
    np.random.seed(seed)
    n = np.random.normal(0, 0.5, shape).astype(np.int8)
    np.random.set_state(st)
    return n


def rgb_keys(px_rgb: np.ndarray) -> np.ndarray:
    p = px_rgb.reshape(-1, 3).astype(np.uint32)
    return (p[:, 0] << 16) | (p[:, 1] << 8) | p[:, 2]


def unique_colors(px_rgb: np.ndarray) -> np.ndarray:
    """np.unique(pixels, axis=0) restated as a 2^24-bit bitmap + ordered
    compaction; key = R<<16|G<<8|B ascending == lexicographic (R,G,B)."""
    seen = np.zeros(1 << 24, bool)
    seen[rgb_keys(px_rgb)] = True
    uk = np.flatnonzero(seen).astype(np.uint32)
    return np.stack([uk >> 16, (uk >> 8) & 255, uk & 255], 1).astype(np.uint8)


def unique_colors_counts(px_rgb: np.ndarray):
    cnt = np.bincount(rgb_keys(px_rgb), minlength=1 << 24)
    uk = np.flatnonzero(cnt).astype(np.uint32)
    return np.stack([uk >> 16, (uk >> 8) & 255, uk & 255], 1).astype(np.uint8), cnt[uk].astype(np.int64)


# --------------------------------------------------------------------------
# A.8  cv2.kmeans   color_extractor.py:189-196
# --------------------------------------------------------------------------
class CvRNG:
    """cv::RNG (multiply-with-carry)."""
    COEFF = 4164903690

    def __init__(self, seed: int):
        self.state = (seed & 0xFFFFFFFFFFFFFFFF) or 0xFFFFFFFF

    def next(self) -> int:
        self.state = ((self.state & 0xFFFFFFFF) * self.COEFF + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uint(self) -> int:
        return self.next()

    def double(self) -> float:
        t = self.next()
        return float((t << 32) | self.next()) * 5.4210108624275221700372640043497e-20


def l2sqr(data: np.ndarray, c: np.ndarray) -> np.ndarray:
    """f32 squared distance, d = ((0 + t0^2) + t1^2) + t2^2, separate mul/add."""
    d = np.zeros(len(data), f32)
    for j in range(data.shape[1]):
        t = (data[:, j] - c[j]).astype(f32)
        d = (d + (t * t).astype(f32)).astype(f32)
    return d


def assign(data: np.ndarray, centers: np.ndarray):
    """Nearest centre, f32 distances, strict '<' so the lowest index wins ties."""
    d = np.stack([l2sqr(data, centers[k]) for k in range(len(centers))], axis=1)
    return d.argmin(1).astype(np.int32), d


def _pp_pick_sequential(dist64: np.ndarray, p: float) -> int:
    n = len(dist64)
    for ci in range(n - 1):
        p -= dist64[ci]
        if p <= 0:
            return ci
    return n - 1


def _pp_pick_prefix(dist64: np.ndarray, p: float) -> int:
    """Same index via an exact prefix sum.  dist values are exact integers
    (integer-valued colours), so partial sums are exact in float64 and the
    sequential `p -= d` never rounds before the sign test decides."""
    n = len(dist64)
    pre = np.cumsum(dist64[: n - 1])
    hit = np.flatnonzero(pre >= p)
    return int(hit[0]) if len(hit) else n - 1


def centers_pp(data: np.ndarray, k: int, rng: CvRNG, trials: int = 3, sequential: bool = False):
    """cv::generateCentersPP (kmeans++ with 3 trials per centre)."""
    n = len(data)
    pick = _pp_pick_sequential if sequential else _pp_pick_prefix
    cidx = [rng.uint() % n]
    dist = l2sqr(data, data[cidx[0]])
    sum0 = float(np.cumsum(dist.astype(np.float64))[-1])
    for _ in range(1, k):
        best = np.inf
        bestc = -1
        bestd = None
        for _ in range(trials):
            p = rng.double() * sum0
            ci = pick(dist.astype(np.float64), p)
            td = np.minimum(l2sqr(data, data[ci]), dist)
            s = float(np.cumsum(td.astype(np.float64))[-1])
            if s < best:
                best, bestc, bestd = s, ci, td
        cidx.append(bestc)
        sum0 = best
        dist = bestd
    return data[cidx].copy(), cidx


def _sums_f32_sequential(data: np.ndarray, labels: np.ndarray, k: int):
    c = np.zeros((k, data.shape[1]), f32)
    cnt = np.bincount(labels, minlength=k).astype(np.int64)
    for j in range(k):
        s = data[labels == j]
        if len(s):
            c[j] = np.cumsum(s, axis=0, dtype=f32)[-1]
    return c, cnt


def center_shift(centers: np.ndarray, old: np.ndarray) -> float:
    """max_k sum_j t^2 with t = float32(c - old) widened to double (OpenCV:
    `double t = center[j] - old_center[j]` on float operands)."""
    t = (centers.astype(f32) - old.astype(f32)).astype(f32).astype(np.float64)
    return float((t * t).sum(1).max())


def centers_update_cv(data: np.ndarray, labels: np.ndarray, k: int, old: np.ndarray | None = None):
    """One cv2.kmeans centre update INCLUDING the empty-cluster repair:
    f32 sequential sums in index order, repair, then sum * f32(1/count).
    Returns (centers f32, labels possibly modified by the repair).  As in
    OpenCV, a repair overwrites the donor cluster's row of `old` (in place)
    with its provisional mean, which the following shift test then sees."""
    labels = labels.copy()
    c, cnt = _sums_f32_sequential(data, labels, k)
    for j in range(k):
        if cnt[j] != 0:
            continue
        max_k = 0
        for k1 in range(1, k):
            if cnt[max_k] < cnt[k1]:
                max_k = k1
        base = (c[max_k] * f32(f32(1.0) / f32(cnt[max_k]))).astype(f32)
        if old is not None:
            old[max_k] = base
        members = np.flatnonzero(labels == max_k)
        d = l2sqr(data[members], base)
        far = int(members[len(d) - 1 - int(np.argmax(d[::-1]))])  # last max wins ('<=')
        cnt[max_k] -= 1
        cnt[j] += 1
        labels[far] = j
        c[max_k] = (c[max_k] - data[far]).astype(f32)
        c[j] = (c[j] + data[far]).astype(f32)
    centers = (c * (f32(1.0) / cnt.astype(f32))[:, None]).astype(f32)
    return centers, labels


def centers_update_exact(data_int: np.ndarray, labels: np.ndarray, k: int, weights: np.ndarray | None = None,
                         old: np.ndarray | None = None):
    """The pinned exact-sum rule of SURVEY.md A.8: integer sums,
    c = float32(double(sum) / double(count)).  Empty-cluster repair follows
    cv2's rule (biggest cluster, farthest member, last max wins) and moves ONE
    point (one unit of weight)."""
    labels = labels.copy()
    w = np.ones(len(data_int), np.int64) if weights is None else weights.astype(np.int64)
    d64 = data_int.astype(np.int64)
    sums = np.zeros((k, d64.shape[1]), np.int64)
    for j in range(d64.shape[1]):
        # float64 bincount is exact here: every partial sum is an integer < 2^53
        sums[:, j] = np.bincount(labels, weights=(d64[:, j] * w).astype(np.float64), minlength=k).astype(np.int64)
    cnt = np.bincount(labels, weights=w.astype(np.float64), minlength=k).astype(np.int64)
    moved = []
    for j in range(k):
        if cnt[j] != 0:
            continue
        max_k = 0
        for k1 in range(1, k):
            if cnt[max_k] < cnt[k1]:
                max_k = k1
        base = (sums[max_k].astype(np.float64) / float(cnt[max_k])).astype(f32)
        if old is not None:
            old[max_k] = base
        members = np.flatnonzero(labels == max_k)
        d = l2sqr(data_int[members].astype(f32), base)
        far = int(members[len(d) - 1 - int(np.argmax(d[::-1]))])
        cnt[max_k] -= 1
        cnt[j] += 1
        sums[max_k] -= d64[far]
        sums[j] += d64[far]
        if weights is None:
            labels[far] = j
        moved.append((far, j))
    centers = (sums.astype(np.float64) / cnt.astype(np.float64)[:, None]).astype(f32)
    return centers, labels, sums, cnt


def cv_max_count(max_iter: int) -> int:
    """cv::kmeans clamps criteria.maxCount to [2, 100] (`std::min(std::max(criteria.maxCount, 2), 100)`):
    the reference's 200 (color_extractor.py:190) is therefore 100 iterations.  Verified against the installed
    cv2: maxCount 100, 101 and 200 give identical results on a slow-converging list, 99 differs
    (tests/test_oracle_vs_cv2.py)."""
    return min(max(int(max_iter), 2), 100)


def lloyd_cv(data: np.ndarray, init_centers: np.ndarray, max_iter: int = 200, eps: float = 0.2):
    """cv2.kmeans(data,K,labels0,crit,1,KMEANS_USE_INITIAL_LABELS) with
    labels0 = assign(init_centers): plain Lloyd from given centres, f32 sums.
    Returns (centers, labels, iters, compactness)."""
    max_iter = cv_max_count(max_iter)
    k = len(init_centers)
    eps2 = eps * eps
    labels, _ = assign(data, init_centers)
    centers = np.zeros_like(init_centers, dtype=f32)
    it = 0
    while True:
        old = centers.copy()
        centers, labels = centers_update_cv(data, labels, k, old)
        shift = np.inf if it == 0 else center_shift(centers, old)
        it += 1
        if it == max(max_iter, 2) or shift <= eps2:
            _, d = assign(data, centers)
            comp = float(d[np.arange(len(data)), labels].astype(np.float64).sum())
            return centers, labels, it, comp
        labels, _ = assign(data, centers)


def lloyd_exact(data_u8: np.ndarray, init_centers: np.ndarray, max_iter: int = 200, eps: float = 0.2,
                weights: np.ndarray | None = None):
    """Per-pixel mode oracle: same loop as `lloyd_cv` but with the exact-sum
    centre rule (order-independent, so identical across any sharding).
    `weights` lets the caller pass unique colours + pixel counts instead of
    the raw pixel list (mathematically identical).  Returns
    (centers f32, labels, iters, sums int64 (K,3), counts int64 (K,))."""
    k = len(init_centers)
    eps2 = eps * eps
    df = data_u8.astype(f32)
    labels, _ = assign(df, init_centers.astype(f32))
    centers = np.zeros((k, data_u8.shape[1]), f32)
    it = 0
    while True:
        old = centers.copy()
        centers, labels, sums, cnt = centers_update_exact(data_u8, labels, k, weights, old)
        shift = np.inf if it == 0 else center_shift(centers, old)
        it += 1
        if it == max(max_iter, 2) or shift <= eps2:
            return centers, labels, it, sums, cnt
        labels, _ = assign(df, centers)


def cv_kmeans(data: np.ndarray, k: int, rng: CvRNG, attempts: int = 10, max_iter: int = 200, eps: float = 0.2):
    """Full cv2.kmeans(data, K, None, (EPS+MAX_ITER, max_iter, eps), attempts,
    KMEANS_PP_CENTERS).  Returns (compactness, labels, centers)."""
    max_iter = cv_max_count(max_iter)
    eps2 = eps * eps
    best = (np.inf, None, None)
    for _ in range(attempts):
        centers, _ = centers_pp(data, k, rng)
        it = 1  # iteration 0 generated the centres (shift = DBL_MAX, never last)
        labels, _ = assign(data, centers)
        while True:
            old = centers.copy()
            centers, labels = centers_update_cv(data, labels, k, old)
            shift = center_shift(centers, old)
            it += 1
            if it == max(max_iter, 2) or shift <= eps2:
                _, d = assign(data, centers)
                comp = float(d[np.arange(len(data)), labels].astype(np.float64).sum())
                break
            labels, _ = assign(data, centers)
        if comp < best[0]:
            best = (comp, labels.copy(), centers.copy())
    return best


# --------------------------------------------------------------------------
# a7  palette tail   color_extractor.py:231-284  (kept literal: argsort(-counts)
#     with numpy's default sort so tie order is whatever numpy gives)
# --------------------------------------------------------------------------
def is_light_color(rgb) -> bool:
    r, g, b = [x / 255.0 for x in rgb]
    return 0.2126 * r + 0.7152 * g + 0.0722 * b > 0.6


def palette_tail(centers_u8: np.ndarray, labels: np.ndarray) -> dict:
    centers = centers_u8
    if len(centers) > 1:
        counts = np.bincount(labels, minlength=len(centers))
        order = np.argsort(-counts)
        centers = centers[order]
    hexes = ["#%02x%02x%02x" % (int(c[0]), int(c[1]), int(c[2])) for c in centers]
    hexes = [c for c in hexes if c.lower() not in ("#ffffff", "#000000")]
    if not hexes:
        bg = "#000000" if is_light_color((255, 255, 255)) else "#FFFFFF"
        return {"primary": bg, "background": bg, "accent": [bg] * 3}
    primary = hexes[0]
    accent = [c for c in hexes if c != primary][:3]
    while len(accent) < 3:
        accent.append(accent[-1] if accent else primary)
    rgb = tuple(int(primary[i:i + 2], 16) for i in (1, 3, 5))
    bg = "#FFFFFF" if not is_light_color(rgb) else "#000000"
    return {"primary": primary, "background": bg, "accent": accent[:3]}


def dominant_colors(px_rgb_noised: np.ndarray, n_colors: int, rng: CvRNG):
    """ColorExtractor._get_dominant_colors (color_extractor.py:173-201)."""
    uniq = unique_colors(px_rgb_noised)
    k = min(n_colors, len(uniq))
    if k <= 1:
        return uniq, np.zeros(len(uniq), np.int64)
    _, labels, centers = cv_kmeans(uniq.astype(f32), k, rng)
    return centers.astype(np.uint8), labels


def extract_colors(bgr: np.ndarray, n_colors: int, noise_i8: np.ndarray, cv_seed: int) -> dict:
    """ColorExtractor.extract_colors (color_extractor.py:203-284) for a
    (H,W,3) u8 BGR ndarray with H>4, given the noise tensor the reference
    would draw and the cv2 RNG seed it would run kmeans++ under."""
    px = apply_noise(bgr2rgb(bgr).reshape(-1, 3), noise_i8.reshape(-1, 3))
    centers, labels = dominant_colors(px, n_colors, CvRNG(cv_seed))
    return palette_tail(centers, labels)


# --------------------------------------------------------------------------
# service-level pipelines
# --------------------------------------------------------------------------
def shape_mask(bgr: np.ndarray, low: int = 50, high: int = 150) -> np.ndarray:
    """ShapeAnalyzer.preprocess_image (shape_analyzer.pyc src L6-30)."""
    return dilate3(canny(gaussian_blur5(bgr2gray(bgr)), low, high))


def shadow_parts(bgr: np.ndarray):
    """ShadowAnalyzer.preprocess_image + analyze_shadow_level
    (shadow_analyzer.pyc src L5-31).  Returns (blurred, mask, sum, count, level)."""
    blurred = gaussian_blur5(bgr2gray(bgr))
    mask = adaptive_threshold(blurred, 2)
    sel = blurred[mask == 255]
    s, n = int(sel.astype(np.int64).sum()), int(sel.size)
    return blurred, mask, s, n, shadow_level(s, n)


def shadow_level(s: int, n: int) -> str:
    if n == 0:
        return "Low"
    avg_darkness = 255 - s / n
    if avg_darkness < 30:
        return "Low"
    if avg_darkness < 60:
        return "Moderate"
    return "High"


def font_mask(bgr: np.ndarray) -> np.ndarray:
    """FontDetector.preprocess_image (font_detector.py:16-37)."""
    return adaptive_threshold(bgr2gray(bgr), 2)
