"""Restatement of ``skimage.registration.phase_cross_correlation`` (CPU oracle).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

scikit-image is a third-party dependency of the reference that is neither
vendored under ``/root/reference`` nor pinned (``install_requirements.sh:54``
names it, ``:65-67`` runs ``pip install -U`` with no version) and is not
installed in this image.  The 3-tuple unpack at the reference's call sites
(``stitcher_process.py:683,706``; ``stitcher.py:510,523``;
``zarr_stitcher.py:351``) and the default ``normalization="phase"`` imply
scikit-image >= 0.19 semantics.  This file restates the published algorithm of
``skimage/registration/_phase_cross_correlation.py`` (Guizar-Sicairos, Thurman,
Fienup, "Efficient subpixel image registration algorithms", Opt. Lett. 33, 2008)
for the only configuration the reference uses: ``space="real"``, no masks,
``disambiguate=False``, ``normalization="phase"``, any ``upsample_factor``.

It runs on ``scipy.fft`` in complex128 -- the FFT backend scikit-image itself
uses -- so a faithful restatement is expected to agree with a real install to
the last bit of the returned shift.
"""
from __future__ import annotations

import numpy as np
import scipy.fft as sfft


def upsampled_dft(data, upsampled_region_size, upsample_factor=1, axis_offsets=None):
    """Matrix-multiply DFT over a small upsampled window (skimage ``_upsampled_dft``).

    For every axis, *last axis first*, the array is contracted with
    ``exp(-2j*pi * (arange(size) - offset)[:, None] * fftfreq(n, upsample_factor))``;
    ``tensordot`` moves the new axis to the front each time, so after all axes
    the result is indexed in the original axis order.
    """
    data = np.asarray(data)
    if not hasattr(upsampled_region_size, "__iter__"):
        upsampled_region_size = [upsampled_region_size] * data.ndim
    elif len(upsampled_region_size) != data.ndim:
        raise ValueError("shape of upsampled region sizes must be equal to input data's number of dimensions.")
    if axis_offsets is None:
        axis_offsets = [0] * data.ndim
    elif len(axis_offsets) != data.ndim:
        raise ValueError("number of axis offsets must be equal to input data's number of dimensions.")

    im2pi = 1j * 2 * np.pi
    dim_properties = list(zip(data.shape, upsampled_region_size, axis_offsets))
    for n_items, ups_size, ax_offset in dim_properties[::-1]:
        kernel = (np.arange(ups_size) - ax_offset)[:, None] * sfft.fftfreq(n_items, upsample_factor)
        kernel = np.exp(-im2pi * kernel)
        kernel = kernel.astype(data.dtype, copy=False)
        data = np.tensordot(kernel, data, axes=(1, -1))
    return data


def _compute_phasediff(cross_correlation_max):
    return np.arctan2(cross_correlation_max.imag, cross_correlation_max.real)


def _compute_error(cross_correlation_max, src_amp, target_amp):
    amp = src_amp * target_amp
    if amp == 0:
        return np.nan
    error = 1.0 - cross_correlation_max * cross_correlation_max.conj() / amp
    return np.sqrt(np.abs(error))


def phase_cross_correlation(reference_image, moving_image, *, upsample_factor=1,
                            normalization="phase", return_details=False):
    """Sub-pixel shift that registers ``moving_image`` onto ``reference_image``.

    Returns ``(shift[float64, ndim], error, phasediff)`` like scikit-image; the
    reference ignores the last two (``stitcher_process.py:683-685``).  With
    ``return_details`` a dict of the intermediate integer indices (what the CUDA
    kernels emit) is appended as a fourth element.
    """
    reference_image = np.asarray(reference_image)
    moving_image = np.asarray(moving_image)
    if reference_image.shape != moving_image.shape:
        raise ValueError("images must be same shape")

    # real-space inputs: forward transforms (integer input -> float64 -> complex128)
    src_freq = sfft.fftn(reference_image)
    target_freq = sfft.fftn(moving_image)

    shape = src_freq.shape
    image_product = src_freq * target_freq.conj()
    if normalization == "phase":
        eps = np.finfo(image_product.real.dtype).eps
        image_product /= np.maximum(np.abs(image_product), 100 * eps)
    elif normalization is not None:
        raise ValueError("normalization must be either phase or None")
    cross_correlation = sfft.ifftn(image_product)

    # whole-pixel peak: first maximum of |cc| in C order
    coarse = np.unravel_index(np.argmax(np.abs(cross_correlation)), cross_correlation.shape)
    midpoint = np.array([np.fix(axis_size / 2) for axis_size in shape])

    float_dtype = image_product.real.dtype
    shift = np.stack(coarse).astype(float_dtype, copy=False)
    shift[shift > midpoint] -= np.array(shape)[shift > midpoint]

    fine = None
    if upsample_factor == 1:
        src_amp = np.sum(np.real(src_freq * src_freq.conj())) / src_freq.size
        target_amp = np.sum(np.real(target_freq * target_freq.conj())) / target_freq.size
        CCmax = cross_correlation[coarse]
    else:
        upsample_factor = np.array(upsample_factor, dtype=float_dtype)
        shift = np.round(shift * upsample_factor) / upsample_factor
        upsampled_region_size = np.ceil(upsample_factor * 1.5)
        dftshift = np.fix(upsampled_region_size / 2.0)
        sample_region_offset = dftshift - shift * upsample_factor
        cross_correlation = upsampled_dft(image_product.conj(), upsampled_region_size,
                                          upsample_factor, sample_region_offset).conj()
        fine = np.unravel_index(np.argmax(np.abs(cross_correlation)), cross_correlation.shape)
        CCmax = cross_correlation[fine]
        maxima = np.stack(fine).astype(float_dtype, copy=False)
        maxima -= dftshift
        shift += maxima / upsample_factor
        src_amp = np.sum(np.real(src_freq * src_freq.conj()))
        target_amp = np.sum(np.real(target_freq * target_freq.conj()))

    # an axis of length 1 carries no shift information
    for dim in range(src_freq.ndim):
        if shape[dim] == 1:
            shift[dim] = 0

    error = _compute_error(CCmax, src_amp, target_amp)
    phasediff = _compute_phasediff(CCmax)
    if return_details:
        details = {
            "coarse": tuple(int(i) for i in coarse),
            "fine": None if fine is None else tuple(int(i) for i in fine),
            "ccmax": complex(CCmax),
        }
        return shift, error, phasediff, details
    return shift, error, phasediff


def shift_from_indices(coarse, fine, shape, upsample_factor):
    """Rebuild the float64 shift from the integer peak indices.

    This is the host-side half of the CUDA path (the kernels return integer
    indices; the float64 arithmetic below is skimage's, term by term), kept here
    so tests can check ``shift_from_indices(details) == shift`` bit for bit.
    """
    shape = tuple(int(s) for s in shape)
    midpoint = np.array([np.fix(n / 2) for n in shape])
    shift = np.array(coarse, dtype=np.float64)
    shift[shift > midpoint] -= np.array(shape)[shift > midpoint]
    if upsample_factor != 1:
        uf = np.array(upsample_factor, dtype=np.float64)
        shift = np.round(shift * uf) / uf
        dftshift = np.fix(np.ceil(uf * 1.5) / 2.0)
        maxima = np.array(fine, dtype=np.float64) - dftshift
        shift += maxima / uf
    for dim, n in enumerate(shape):
        if n == 1:
            shift[dim] = 0
    return shift
