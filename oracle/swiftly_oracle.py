"""
CPU ORACLE for the SwiFTly facet<->subgrid hot path.  TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the reference algorithm
(ska-sdp-distributed-fourier-transform, package ``ska_sdp_exec_swiftly``).  It
exists so that the CUDA path can be checked against something that runs on any
host, including the GPU box where ``/root/reference`` does not exist.

Rules (see DESIGN.md "Oracle"):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import this module;
  * the product package never imports it and has no CPU fallback.

Pinning: ``tests/golden/make_golden.py`` imports the real reference from
``/root/reference`` (numpy backend) and stores its outputs as fixtures under
``tests/golden/``; ``tests/test_oracle.py`` checks this module against those
fixtures bit-for-bit (same numpy.fft calls on the same data => identical
rounding) and against the reference's own analytic known-answer tests (direct
DFT of point sources, reference ``tests/test_core.py``).  Parity is therefore
PINNED.

Every function cites the reference lines it restates.  Paths are relative to
``/root/reference/src/ska_sdp_exec_swiftly/``.

Conventions: all transforms are "centred" (zero frequency / zero pixel at
index n//2).  Data movement (pad / extract / roll) is expressed here as
modular index maps instead of the reference's chains of numpy.roll /
numpy.pad / slicing; the arithmetic (which products are formed, which
numpy.fft routine is called on which data) is unchanged, so results are
bit-identical to the reference.
"""

from __future__ import annotations

import numpy
import scipy.special

__all__ = [
    "OracleCore",
    "centred_fft",
    "centred_ifft",
    "pad_mid",
    "extract_mid",
    "coordinates",
    "facet_from_sources",
    "subgrid_from_sources",
    "full_cover_offsets",
    "cover_mask",
    "forward_reference_order",
    "backward_reference_order",
]


# --------------------------------------------------------------------------
# helpers (fourier_transform/fourier_algorithm.py)
# --------------------------------------------------------------------------
def _axis_index(ndim, axis, idx):
    """Index tuple selecting ``idx`` along ``axis`` and everything elsewhere."""
    sel = [slice(None)] * ndim
    sel[axis] = idx
    return tuple(sel)


def _along(vec, ndim, axis):
    """Reshape 1-D ``vec`` so it broadcasts along ``axis`` of an ndim array.

    Restates ``broadcast`` (fourier_algorithm.py:38-50).
    """
    shape = [1] * ndim
    shape[axis] = len(vec)
    return numpy.reshape(vec, shape)


def pad_mid(a, n, axis):
    """Zero-pad ``a`` to length ``n`` along ``axis`` keeping index n0//2 at n//2.

    Restates ``pad_mid`` (fourier_algorithm.py:53-73): left padding is
    ``n//2 - n0//2``.
    """
    n0 = a.shape[axis]
    if n0 == n:
        return a
    shape = list(a.shape)
    shape[axis] = n
    out = numpy.zeros(shape, dtype=a.dtype)
    lo = n // 2 - n0 // 2
    out[_axis_index(a.ndim, axis, slice(lo, lo + n0))] = a
    return out


def extract_mid(a, n, axis):
    """Cut the central ``n`` samples along ``axis`` (inverse of pad_mid).

    Restates ``extract_mid`` (fourier_algorithm.py:76-93): the cut starts at
    ``shape//2 - n//2`` for both parities of ``n``.
    """
    if n > a.shape[axis]:
        raise AssertionError("extract_mid: n larger than array")
    lo = a.shape[axis] // 2 - n // 2
    return a[_axis_index(a.ndim, axis, slice(lo, lo + n))]


def centred_fft(a, axis):
    """Image -> grid centred DFT (fourier_algorithm.py:96-107)."""
    return numpy.fft.fftshift(
        numpy.fft.fft(numpy.fft.ifftshift(a, axis), axis=axis), axis
    )


def centred_ifft(a, axis):
    """Grid -> image centred inverse DFT (fourier_algorithm.py:110-122)."""
    return numpy.fft.fftshift(
        numpy.fft.ifft(numpy.fft.ifftshift(a, axis), axis=axis), axis
    )


def coordinates(n):
    """Coordinates in [-0.5, 0.5) with 0 at n//2 (fourier_algorithm.py:125-138)."""
    half = n // 2
    return (numpy.arange(n) - half) / n


def _cyclic_window(length, start, count):
    """Indices ``(start + i) mod length`` for ``i < count``."""
    return numpy.mod(start + numpy.arange(count), length)


# --------------------------------------------------------------------------
# ground truth generators (fourier_algorithm.py:218-315)
# --------------------------------------------------------------------------
def facet_from_sources(sources, image_size, facet_size, facet_offsets, masks=None):
    """Facet pixels for a list of ``(intensity, *coords)`` point sources.

    Restates ``make_facet_from_sources`` (fourier_algorithm.py:218-264).
    """
    dims = len(facet_offsets)
    facet = numpy.zeros(dims * [facet_size], dtype=complex)
    origin = numpy.array(facet_offsets, dtype=int) - facet_size // 2
    for intensity, *coord in sources:
        rel = numpy.mod(numpy.array(coord) - origin, image_size)
        if numpy.any(rel >= facet_size):
            continue
        facet[tuple(rel)] += intensity
    for axis, mask in enumerate(masks or []):
        if mask is not None:
            facet *= _along(numpy.asarray(mask), dims, axis)
    return facet


def subgrid_from_sources(
    sources, image_size, subgrid_size, subgrid_offsets, masks=None
):
    """Direct DFT of point sources on a subgrid (the analytic truth).

    Restates ``make_subgrid_from_sources`` (fourier_algorithm.py:267-315):
    ``sum_s I_s / N^d * exp(2 pi i <uv, x_s> / N)`` with uv running over
    ``off - size//2 ... off + (size+1)//2 - 1`` per axis.
    """
    dims = len(subgrid_offsets)
    axes = [
        numpy.arange(off - subgrid_size // 2, off + (subgrid_size + 1) // 2)
        for off in subgrid_offsets
    ]
    out = numpy.zeros(dims * [subgrid_size], dtype=complex)
    for intensity, *coords in sources:
        phase = numpy.zeros(dims * [subgrid_size])
        for axis, (uv, x) in enumerate(zip(axes, coords)):
            phase = phase + _along(uv * x, dims, axis)
        out += (intensity / image_size**dims) * numpy.exp(
            (2j * numpy.pi / image_size) * phase
        )
    for axis, mask in enumerate(masks or []):
        if mask is not None:
            out *= _along(numpy.asarray(mask), dims, axis)
    return out


# --------------------------------------------------------------------------
# covers / masks (api_helper.py:213-253)
# --------------------------------------------------------------------------
def full_cover_offsets(N, chunk):
    """Offsets ``chunk * arange(ceil(N / chunk))`` (api_helper.py:221)."""
    return chunk * numpy.arange(int(numpy.ceil(N / chunk)))


def cover_mask(N, chunk, index):
    """0/1 mask of chunk ``index`` of a full cover (api_helper.py:221-240,243-253).

    Borders sit half way between neighbouring offsets; the mask is all ones
    when ``chunk`` divides ``N``.
    """
    offs = full_cover_offsets(N, chunk)
    border = (offs + numpy.hstack([offs[1:], [N + offs[0]]])) // 2
    left = (border[index - 1] - offs[index] + chunk // 2) % N
    right = border[index] - offs[index] + chunk // 2
    mask = numpy.zeros((chunk,))
    mask[slice(left, right)] = 1
    return mask


# --------------------------------------------------------------------------
# the eight primitives (fourier_transform/core.py)
# --------------------------------------------------------------------------
class OracleCore:
    """numpy restatement of ``SwiftlyCore`` (core.py:20-484)."""

    def __init__(self, W, N, xM_size, yN_size):
        self.W = W
        self.N = N
        self.xM_size = xM_size
        self.yN_size = yN_size
        # core.py:55-74
        if N % yN_size != 0:
            raise ValueError(f"Image size {N} not divisible by facet size {yN_size}!")
        if N % xM_size != 0:
            raise ValueError(
                f"Image size {N} not divisible by subgrid size {xM_size}!"
            )
        if (xM_size * yN_size) % N != 0:
            raise ValueError(
                f"Contribution size not integer with image size {N}, "
                f"subgrid size {xM_size} and facet size {yN_size}!"
            )
        self.xM_yN_size = xM_size * yN_size // N  # core.py:48
        pswf = self.pswf_window(W, yN_size)
        # core.py:104-108
        self._Fb = 1 / pswf[1:]
        # core.py:110-117
        step = int(N / xM_size)
        self._Fn = pswf[(yN_size // 2) % step :: step]

    # core.py:76-92
    @property
    def subgrid_off_step(self):
        return self.N // self.yN_size

    @property
    def facet_off_step(self):
        return self.N // self.xM_size

    @staticmethod
    def pswf_window(W, yN_size):
        """PSWF samples at facet resolution (core.py:119-150).

        ``pro_ang1(0, 0, pi W / 2, 2 x)`` for x = coordinates(yN); evaluated
        in chunks of 500 like the reference (scipy segfault work-around), the
        first sample (x = -1, NaN) is zeroed.
        """
        pswf = numpy.empty(yN_size, dtype=float)
        x2 = 2 * coordinates(yN_size)
        for lo in range(1, yN_size, 500):
            pswf[lo : lo + 500] = scipy.special.pro_ang1(
                0, 0, numpy.pi * W / 2, x2[lo : lo + 500]
            )[0]
        pswf[0] = 0
        return pswf

    # -- shared -----------------------------------------------------------
    @staticmethod
    def _emit(result, out, accumulate=False):
        """core.py:152-186 (``_copy_to_out``)."""
        if out is None:
            return result
        if out.shape != result.shape:
            raise ValueError(f"Output shape is {out.shape}, expected {result.shape}!")
        if accumulate:
            out[:] += result
        else:
            out[:] = result
        return out

    def _fb_window(self, size):
        """``extract_mid(Fb, size)`` (core.py:213-215, 474-476)."""
        return extract_mid(self._Fb, size, 0)

    # -- facet -> subgrid ---------------------------------------------------
    def prepare_facet(self, facet, facet_off, axis, out=None):
        """core.py:189-222: Fb-weight, pad to yN, roll by facet_off, centred iFFT."""
        yN = self.yN_size
        fs = facet.shape[axis]
        weighted = facet * _along(self._fb_window(fs), facet.ndim, axis)
        shape = list(facet.shape)
        shape[axis] = yN
        line = numpy.zeros(shape, dtype=weighted.dtype)
        dest = _cyclic_window(yN, yN // 2 - fs // 2 + facet_off, fs)
        line[_axis_index(facet.ndim, axis, dest)] = weighted
        return self._emit(centred_ifft(line, axis), out)

    def _facet_window(self, subgrid_off):
        """Source indices of the m-sample window used by a subgrid.

        core.py:243-252: roll by -s, extract_mid(m), roll by +s with
        ``s = subgrid_off * yN // N``.
        """
        yN, m = self.yN_size, self.xM_yN_size
        s = subgrid_off * yN // self.N
        t = numpy.arange(m)
        return numpy.mod(yN // 2 - m // 2 + numpy.mod(t - s, m) + s, yN)

    def extract_from_facet(self, prep_facet, subgrid_off, axis, out=None):
        """core.py:224-253 (pure gather)."""
        src = self._facet_window(subgrid_off)
        result = prep_facet[_axis_index(prep_facet.ndim, axis, src)]
        return self._emit(result, out)

    def add_to_subgrid(self, facet_contrib, facet_off, axis, out=None):
        """core.py:255-285: centred FFT_m, roll by -sf, times Fn, pad to xM, roll by +sf."""
        xM, m = self.xM_size, self.xM_yN_size
        sf = facet_off * xM // self.N
        spectrum = centred_fft(facet_contrib, axis)
        u = numpy.arange(m)
        rolled = spectrum[_axis_index(spectrum.ndim, axis, numpy.mod(u + sf, m))]
        weighted = _along(self._Fn, spectrum.ndim, axis) * rolled
        shape = list(facet_contrib.shape)
        shape[axis] = xM
        result = numpy.zeros(shape, dtype=weighted.dtype)
        dest = numpy.mod(xM // 2 - m // 2 + u + sf, xM)
        result[_axis_index(spectrum.ndim, axis, dest)] = weighted
        return self._emit(result, out, accumulate=True)

    def finish_subgrid(self, summed_contribs, subgrid_off, subgrid_size, out=None):
        """core.py:287-325: per axis centred iFFT_xM, roll by -off, extract_mid(size)."""
        dims = summed_contribs.ndim
        if not isinstance(subgrid_off, list):
            if dims != 1:
                raise ValueError("Subgrid offset must be given for every dimension!")
            subgrid_off = [subgrid_off]
        xM = self.xM_size
        tmp = summed_contribs
        for axis in range(dims):
            img = centred_ifft(tmp, axis)
            src = _cyclic_window(
                xM, xM // 2 - subgrid_size // 2 + subgrid_off[axis], subgrid_size
            )
            tmp = img[_axis_index(dims, axis, src)]
        return self._emit(tmp, out)

    # -- subgrid -> facet ---------------------------------------------------
    def prepare_subgrid(self, subgrid, subgrid_off, out=None):
        """core.py:328-368: per axis pad to xM, roll by +off, centred FFT."""
        dims = subgrid.ndim
        if dims == 1 and not isinstance(subgrid_off, tuple):
            subgrid_off = (subgrid_off,)
        if len(subgrid_off) != dims:
            raise ValueError("Dimensionality mismatch between subgrid and offsets!")
        xM = self.xM_size
        tmp = subgrid
        for axis in range(dims):
            size = tmp.shape[axis]
            shape = list(tmp.shape)
            shape[axis] = xM
            line = numpy.zeros(shape, dtype=tmp.dtype)
            dest = _cyclic_window(xM, xM // 2 - size // 2 + subgrid_off[axis], size)
            line[_axis_index(dims, axis, dest)] = tmp
            tmp = centred_fft(line, axis)
        return self._emit(tmp, out)

    def extract_from_subgrid(self, FSi, facet_off, axis, out=None):
        """core.py:370-406: roll by -sf, extract_mid(m), times Fn, roll by +sf, centred iFFT_m."""
        xM, m = self.xM_size, self.xM_yN_size
        sf = facet_off * xM // self.N
        u = numpy.arange(m)
        window = FSi[_axis_index(FSi.ndim, axis, numpy.mod(xM // 2 - m // 2 + u + sf, xM))]
        weighted = _along(self._Fn, FSi.ndim, axis) * window
        rolled = numpy.empty_like(weighted)
        rolled[_axis_index(FSi.ndim, axis, numpy.mod(u + sf, m))] = weighted
        return self._emit(centred_ifft(rolled, axis), out)

    def add_to_facet(self, subgrid_contrib, subgrid_off, axis, out=None):
        """core.py:408-449 (pure scatter-add, transpose of extract_from_facet)."""
        yN = self.yN_size
        shape = list(subgrid_contrib.shape)
        shape[axis] = yN
        result = numpy.zeros(shape, dtype=subgrid_contrib.dtype)
        dest = self._facet_window(subgrid_off)
        result[_axis_index(subgrid_contrib.ndim, axis, dest)] = subgrid_contrib
        return self._emit(result, out, accumulate=True)

    def finish_facet(self, MiNjSi_sum, facet_off, facet_size, axis, out=None):
        """core.py:452-484: centred FFT_yN, roll by -facet_off, extract_mid, times Fb."""
        yN = self.yN_size
        img = centred_fft(MiNjSi_sum, axis)
        src = _cyclic_window(yN, yN // 2 - facet_size // 2 + facet_off, facet_size)
        cut = img[_axis_index(img.ndim, axis, src)]
        result = _along(self._fb_window(facet_size), img.ndim, axis) * cut
        return self._emit(result, out)


# --------------------------------------------------------------------------
# task bodies in the reference's order (api_helper.py / api.py), serial
# --------------------------------------------------------------------------
def forward_reference_order(core, facets, facet_offs, subgrid_offs, subgrid_size,
                            subgrid_masks=None, keep=None):
    """Serial facet->subgrid transform in the reference's task order.

    ``facets``: list of 2-D arrays; ``facet_offs``: list of (off0, off1);
    ``subgrid_offs``: list of (off0, off1); returns list of finished subgrids.

    Restates SwiftlyForward (api.py:238-324) with the task bodies
    ``extract_column`` (api_helper.py:200-210) and ``sum_and_finish_subgrid``
    (api_helper.py:73-112).  The reference iterates a python ``set`` of facet
    off1 values; here the grouping order is sorted (addition order only).

    ``keep``: optional dict that receives intermediates (BF_F, NMBF_BF,
    contributions) of the *last* subgrid for stage-by-stage parity checks.
    """
    BF_F = [core.prepare_facet(f, off0, axis=0) for f, (off0, _) in zip(facets, facet_offs)]
    column_cache = {}
    results = []
    for isg, (sg0, sg1) in enumerate(subgrid_offs):
        if sg0 not in column_cache:
            column_cache.clear()  # lru_forward = 1
            column_cache[sg0] = [
                core.prepare_facet(
                    core.extract_from_facet(bf, sg0, axis=0), off1, axis=1
                )
                for bf, (_, off1) in zip(BF_F, facet_offs)
            ]
        NMBF_BF = column_cache[sg0]
        contribs = [core.extract_from_facet(nb, sg1, axis=1) for nb in NMBF_BF]
        acc = None
        for off1 in sorted({o1 for _, o1 in facet_offs}):
            col = None
            for c, (off0, o1) in zip(contribs, facet_offs):
                if o1 != off1:
                    continue
                col = core.add_to_subgrid(c, off0, axis=0, out=col)
            acc = core.add_to_subgrid(col, off1, axis=1, out=acc)
        sg = core.finish_subgrid(acc, [sg0, sg1], subgrid_size)
        if subgrid_masks is not None:
            m0, m1 = subgrid_masks[isg]
            if m0 is not None:
                sg = sg * numpy.asarray(m0)[:, None]
            if m1 is not None:
                sg = sg * numpy.asarray(m1)[None, :]
        results.append(sg)
        if keep is not None:
            keep["BF_F"] = BF_F
            keep["NMBF_BF"] = NMBF_BF
            keep["contribs"] = contribs
            keep["acc"] = acc
    return results


def backward_reference_order(core, subgrids, subgrid_offs, facet_offs, facet_size,
                             facet_masks=None):
    """Serial subgrid->facet transform in the reference's task order.

    Restates SwiftlyBackward (api.py:347-463) with ``prepare_and_split_subgrid``
    (api_helper.py:115-139), ``accumulate_column`` (:142-152),
    ``accumulate_facet`` (:155-179) and ``finish_facet`` (:182-197) for
    ``lru_backward = 1`` (a column is folded when the subgrid off0 changes).
    """
    nf = len(facet_offs)
    MNAF_BMNAF = [None] * nf
    current_off0 = None
    NAF_MNAF = [None] * nf

    def fold(sg_off0):
        for j, (_, off1) in enumerate(facet_offs):
            part = core.finish_facet(NAF_MNAF[j], off1, facet_size, axis=1)
            if facet_masks is not None and facet_masks[j][1] is not None:
                part = part * numpy.asarray(facet_masks[j][1])[None, :]
            MNAF_BMNAF[j] = core.add_to_facet(part, sg_off0, axis=0, out=MNAF_BMNAF[j])

    for sg, (sg0, sg1) in zip(subgrids, subgrid_offs):
        if current_off0 is not None and sg0 != current_off0:
            fold(current_off0)
            NAF_MNAF = [None] * nf
        current_off0 = sg0
        prepared = core.prepare_subgrid(sg, (sg0, sg1))
        by_off0 = {}
        for j, (off0, off1) in enumerate(facet_offs):
            if off0 not in by_off0:
                by_off0[off0] = core.extract_from_subgrid(prepared, off0, axis=0)
            c = core.extract_from_subgrid(by_off0[off0], off1, axis=1)
            NAF_MNAF[j] = core.add_to_facet(c, sg1, axis=1, out=NAF_MNAF[j])
    if current_off0 is not None:
        fold(current_off0)
    out = []
    for j, (off0, _) in enumerate(facet_offs):
        f = core.finish_facet(MNAF_BMNAF[j], off0, facet_size, axis=0)
        if facet_masks is not None and facet_masks[j][0] is not None:
            f = f * numpy.asarray(facet_masks[j][0])[:, None]
        out.append(f)
    return out
