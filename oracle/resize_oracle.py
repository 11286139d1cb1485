"""numpy restatement of the six ``skimage.transform.resize(..., order=1, ...)`` calls of
``pix2pose_model/recognition.py`` (:82, :103, :121, :134, :144, :146).

TEST INFRASTRUCTURE (see oracle/__init__.py).  scikit-image is not vendored by the reference, is
not pinned in its requirements.txt and is absent from this image, so *parity is unpinned* here; the
semantics below are the documented behaviour of skimage 0.14-0.18 ``resize`` -> ``warp`` ->
``_warp_fast`` with ``anti_aliasing`` off (SURVEY.md §8c pin):

* output pixel (r, c) samples the input at ``(r + 0.5) * in/out - 0.5`` (pixel centres),
* bilinear weights from ``floor`` / ``ceil`` neighbours,
* ``mode='constant'``: neighbours outside the image read ``cval``;
  ``mode='reflect'``: mirrored without repeating the edge (numpy.pad 'reflect'),
* ``clip=True`` (default): the result is clipped to the input's [min, max]; output pixels exactly
  equal to ``cval`` are kept when ``cval`` lies outside that range,
* bool input is converted to 0.0 / 1.0; channels are warped independently, the clip range is
  taken over the whole array.
"""
import numpy as np


def _reflect(idx, dim):
    """skimage ``coord_map`` mode 'R' for the only out-of-range indices bilinear resizing can touch."""
    if dim == 1:
        return np.zeros_like(idx)
    cmax = dim - 1
    idx = np.abs(idx)                      # -k -> k
    period = 2 * cmax
    idx = idx % period
    return np.where(idx > cmax, period - idx, idx)


def _axis(n_in, n_out, mode):
    src = (np.arange(n_out, dtype=np.float64) + 0.5) * (float(n_in) / n_out) - 0.5
    lo = np.floor(src).astype(np.int64)
    hi = np.ceil(src).astype(np.int64)
    w = src - lo
    if mode == "reflect":
        return _reflect(lo, n_in), _reflect(hi, n_in), w, None, None
    return np.clip(lo, 0, n_in - 1), np.clip(hi, 0, n_in - 1), w, (lo < 0) | (lo >= n_in), (hi < 0) | (hi >= n_in)


def resize(image, output_shape, order=1, mode="reflect", cval=0.0, clip=True):
    assert order == 1 and mode in ("reflect", "constant")
    img = np.asarray(image)
    img = img.astype(np.float64)
    squeeze = img.ndim == 2
    if squeeze:
        img = img[:, :, None]
    H, W, C = img.shape
    oh, ow = int(output_shape[0]), int(output_shape[1])
    rlo, rhi, rw, rlo_o, rhi_o = _axis(H, oh, mode)
    clo, chi, cw, clo_o, chi_o = _axis(W, ow, mode)

    def px(ri, ci, ro, co):
        v = img[ri[:, None], ci[None, :], :]
        if mode == "constant":
            out = ro[:, None] | co[None, :]
            v = np.where(out[:, :, None], cval, v)
        return v

    cw3, rw3 = cw[None, :, None], rw[:, None, None]
    top = (1 - cw3) * px(rlo, clo, rlo_o, clo_o) + cw3 * px(rlo, chi, rlo_o, chi_o)
    bot = (1 - cw3) * px(rhi, clo, rhi_o, clo_o) + cw3 * px(rhi, chi, rhi_o, chi_o)
    out = (1 - rw3) * top + rw3 * bot
    if clip:
        mn, mx = img.min(), img.max()
        keep = mode == "constant" and not (mn <= cval <= mx)
        if keep:
            cmask = out == cval
        out = np.clip(out, mn, mx)
        if keep:
            out[cmask] = cval
    return out[:, :, 0] if squeeze else out
