"""
``ska_sdp_func``-shaped adapter: class ``Swiftly(N, yN_size, xM_size, W)`` with the ten entry
points the reference's ``SwiftlyCoreFunc`` calls on
``ska_sdp_func.fourier_transforms.swiftly.Swiftly`` (reference ``fourier_transform/core.py``
:508-510 construction, :684-929 calls): every method takes 2-D complex128 numpy arrays --
possibly transposed strided views, exactly what ``_auto_broadcast_create`` passes for
``axis=0`` (core.py:605-617) -- transforms along the LAST axis and writes into the given output
array.  Everything runs in the CUDA kernels behind ``libswiftly_b200.so``.

Putting this class in the place of the native library,

    import ska_sdp_func.fourier_transforms.swiftly as native
    native.Swiftly = ska_sdp_distributed_fourier_transform_b200.sdp_func_compat.Swiftly

lets an UNMODIFIED reference (``SwiftlyConfig(backend="ska_sdp_func")``, ``SwiftlyCoreFunc``)
run on the GPU; ``tests/test_emu_reference_dropin.py`` does exactly that with the reference's
own unit tests.
"""

import ctypes

import numpy

from . import _lib
from .pswf import window_tables


class Swiftly:
    """Drop-in for ``ska_sdp_func.fourier_transforms.swiftly.Swiftly`` (note the argument order)."""

    def __init__(self, N, yN_size, xM_size, W, device=0):
        self._lib = _lib.load()
        self.N, self.yN_size, self.xM_size, self.W = N, yN_size, xM_size, W
        Fb, Fn = window_tables(W, N, xM_size, yN_size)
        self._plan = ctypes.c_void_p()
        rc = self._lib.swiftly_b200_create(
            float(W), int(N), int(xM_size), int(yN_size),
            Fb.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
            Fn.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), int(device),
            ctypes.byref(self._plan))
        _lib.check(self._lib, rc)

    def __del__(self):
        plan = getattr(self, "_plan", None)
        if plan is not None and plan.value:
            self._lib.swiftly_b200_destroy(plan)
            self._plan = None

    # ------------------------------------------------------------------ plumbing
    @staticmethod
    def _lines(a):
        """Lines = rows of the 2-D (possibly strided) complex128 array, along the last axis."""
        if a.ndim != 2 or a.dtype != numpy.complex128:
            raise ValueError("expected a 2-D complex128 array")
        if any(s % a.itemsize for s in a.strides):
            raise ValueError("array strides must be multiples of the item size")
        s0, s1 = (s // a.itemsize for s in a.strides)
        return _lib.Lines(a.ctypes.data, a.shape[0], a.shape[1], s0, s1, _lib.HOST)

    def _call(self, name, in_arr, out_arr, offset, *extra):
        in_arr = numpy.asarray(in_arr)
        if in_arr.dtype != numpy.complex128:
            in_arr = in_arr.astype(numpy.complex128)
        din, dout = self._lines(in_arr), self._lines(out_arr)
        rc = getattr(self._lib, name)(self._plan, ctypes.byref(din), ctypes.byref(dout),
                                      int(offset), *extra, ctypes.c_void_p(0))
        _lib.check(self._lib, rc)

    # ------------------------------------------------------------------ facet -> subgrid
    def prepare_facet(self, facet, prep_facet_out, facet_offset):
        self._call("swiftly_b200_prepare_facet", facet, prep_facet_out, facet_offset)

    def extract_from_facet(self, prep_facet, contribution_out, subgrid_offset):
        self._call("swiftly_b200_extract_from_facet", prep_facet, contribution_out, subgrid_offset)

    def add_to_subgrid(self, contribution, subgrid_image_inout, facet_offset):
        self._call("swiftly_b200_add_to_subgrid", contribution, subgrid_image_inout, facet_offset)

    def add_to_subgrid_2d(self, contribution, subgrid_image_inout, facet_offset0, facet_offset1):
        tmp = numpy.zeros((self.xM_size, contribution.shape[1]), dtype=numpy.complex128)
        self._call("swiftly_b200_add_to_subgrid", contribution.T, tmp.T, facet_offset0)
        self._call("swiftly_b200_add_to_subgrid", tmp, subgrid_image_inout, facet_offset1)

    def finish_subgrid(self, subgrid_image, subgrid_out, subgrid_offset):
        self._call("swiftly_b200_finish_subgrid", subgrid_image, subgrid_out, subgrid_offset,
                   ctypes.c_void_p(0))

    # ------------------------------------------------------------------ subgrid -> facet
    def prepare_subgrid_inplace(self, subgrid_inout, subgrid_offset):
        # already padded to xM by the caller (core.py:833-840): roll + centred FFT in place
        self._call("swiftly_b200_prepare_subgrid", subgrid_inout, subgrid_inout, subgrid_offset)

    def prepare_subgrid_inplace_2d(self, subgrid_inout, subgrid_offset0, subgrid_offset1):
        self._call("swiftly_b200_prepare_subgrid", subgrid_inout.T, subgrid_inout.T, subgrid_offset0)
        self._call("swiftly_b200_prepare_subgrid", subgrid_inout, subgrid_inout, subgrid_offset1)

    def extract_from_subgrid(self, subgrid_image, contribution_out, facet_offset):
        self._call("swiftly_b200_extract_from_subgrid", subgrid_image, contribution_out,
                   facet_offset)

    def add_to_facet(self, contribution, prep_facet_inout, subgrid_offset):
        self._call("swiftly_b200_add_to_facet", contribution, prep_facet_inout, subgrid_offset)

    def finish_facet(self, prep_facet, facet_out, facet_offset):
        self._call("swiftly_b200_finish_facet", prep_facet, facet_out, facet_offset,
                   ctypes.c_void_p(0))
