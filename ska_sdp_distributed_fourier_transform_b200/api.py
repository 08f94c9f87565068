"""
Application programming interface of the B200 SwiFTly transform.

Same surface as the reference's ``api.py`` (``FacetConfig`` :39-70,
``SubgridConfig`` :73-104, ``SwiftlyConfig`` :107-214, ``SwiftlyForward`` :217-324,
``SwiftlyBackward`` :327-463, ``make_full_subgrid_cover`` / ``make_full_facet_cover``
:593-612) -- constructor arguments, method names and properties are kept so a
driver written for ``ska_sdp_exec_swiftly`` only changes its import.  The
mechanism is different: instead of building a Dask task graph that is shipped to
CPU workers, every call enqueues hand-written CUDA kernels on the current CUDA
stream of one GPU; "tasks" are light handles around device tensors
(``.result()`` / ``.compute()`` copy to the host, ``.tensor`` stays on the GPU).
``lru_forward`` / ``lru_backward`` keep their meaning (number of subgrid columns
whose intermediates stay resident); ``queue_size`` bounds the number of results
in flight.  ``dask_client`` / ``client`` are accepted and ignored.
"""

import collections
import logging

import numpy

from .api_helper import (
    accumulate_column,
    accumulate_facet,
    extract_column,
    finish_facet,
    make_full_cover_config,
    make_mask_from_slice,
    prepare_and_split_subgrid,
    sum_and_finish_subgrid,
)
from .core import SwiftlyCoreB200

try:
    import torch
except ImportError:  # pragma: no cover
    torch = None

__all__ = [
    "FacetConfig",
    "SubgridConfig",
    "SwiftlyConfig",
    "SwiftlyForward",
    "SwiftlyBackward",
    "make_full_facet_cover",
    "make_full_subgrid_cover",
]

log = logging.getLogger("fourier-logger")


class _ChunkConfig:
    """Offset, size and (lazily materialised) masks of a facet or subgrid."""

    def __init__(self, off0, off1, size, mask0=None, mask1=None):
        self.off0 = off0
        self.off1 = off1
        self.size = size
        self._mask0 = mask0
        self._mask1 = mask1

    @staticmethod
    def _materialise(mask):
        # a mask is either an array or ``[[slices], size]``
        if isinstance(mask, list):
            return make_mask_from_slice(mask[0], mask[1])
        return mask

    @property
    def mask0(self):
        """Mask along axis 0 (vertical)."""
        return self._materialise(self._mask0)

    @property
    def mask1(self):
        """Mask along axis 1 (horizontal)."""
        return self._materialise(self._mask1)

    def __repr__(self):
        return f"{self.__class__.__name__}(off0={self.off0}, off1={self.off1}, size={self.size})"


class FacetConfig(_ChunkConfig):
    """Facet configuration (offsets of the facet mid-point, size, masks)."""


class SubgridConfig(_ChunkConfig):
    """Subgrid configuration (offsets of the subgrid mid-point, size, masks)."""


class SwiftlyConfig:
    """SwiFTly configuration: sizes, window parameter and the processing core.

    :param W: PSWF parameter
    :param fov: field of view (kept for compatibility, unused like in the reference)
    :param N: image size
    :param yB_size: facet size
    :param yN_size: padded facet size
    :param xA_size: subgrid size
    :param xM_size: padded subgrid size
    :param dask_client: accepted for compatibility, ignored
    :param backend: ``"b200"`` (alias ``"cuda"``); the reference's backend names
        (``"numpy"``, ``"ska_sdp_func"``) are accepted and run on the same CUDA core
        (with a log warning) so that a reference driver only changes its import; any
        other name raises ``ValueError`` like the reference (api.py:137-143)
    :param device: CUDA device index (default: current device)
    :param core: optionally a ready-made object with the eight-primitive interface
    """

    # pylint: disable=too-many-arguments,too-many-instance-attributes
    def __init__(self, W, fov, N, yB_size, yN_size, xA_size, xM_size, dask_client=None,
                 backend="b200", device=None, core=None, **_other_args):
        self._W = W
        self._fov = fov
        self._N = N
        self._yB_size = yB_size
        self._yN_size = yN_size
        self._xA_size = xA_size
        self._xM_size = xM_size
        self.dask_client = dask_client
        if core is not None:
            self._core = core
        elif backend in ("b200", "cuda", "numpy", "ska_sdp_func"):
            if backend in ("numpy", "ska_sdp_func"):
                # a driver written for the reference passes its CPU backend names
                # (api.py:137-143); there is no CPU implementation here -- the same
                # primitives run on the GPU core
                log.warning("backend=%r requested: this package has no CPU backend, "
                            "using the B200 CUDA core", backend)
            self._core = SwiftlyCoreB200(W, N, xM_size, yN_size, device=device)
        else:
            raise ValueError(f"Unknown SwiFTly backend: {backend}")
        # the reference hands out a dask.delayed handle to the scattered core; here the
        # core itself plays that role (its methods run immediately on the GPU)
        self.core_task = self._core

    @property
    def core(self):
        """The processing core (eight SwiFTly primitives on the GPU)."""
        return self._core

    @property
    def image_size(self):
        """Size of the entire (virtual) image in pixels."""
        return self._N

    @property
    def max_facet_size(self):
        """Maximum size of a facet in pixels."""
        return self._yB_size

    @property
    def max_subgrid_size(self):
        """Maximum size of a subgrid in pixels."""
        return self._xA_size

    @property
    def pswf_parameter(self):
        """Parameter of the window function."""
        return self._W

    @property
    def internal_facet_size(self):
        """Padded facet size used internally."""
        return self._yN_size

    @property
    def internal_subgrid_size(self):
        """Padded subgrid size used internally."""
        return self._xM_size

    @property
    def facet_off_step(self):
        """All facet offsets must be divisible by this."""
        return self._core.facet_off_step

    @property
    def subgrid_off_step(self):
        """All subgrid offsets must be divisible by this."""
        return self._core.subgrid_off_step


def make_full_subgrid_cover(swiftlyconfig):
    """Subgrid configs covering the whole grid."""
    return make_full_cover_config(
        swiftlyconfig.image_size, swiftlyconfig.max_subgrid_size, SubgridConfig
    )


def make_full_facet_cover(swiftlyconfig):
    """Facet configs covering the whole image."""
    return make_full_cover_config(
        swiftlyconfig.image_size, swiftlyconfig.max_facet_size, FacetConfig
    )


# ---------------------------------------------------------------------- task handles
class DeviceTask:
    """Result handle: a device tensor plus the CUDA event that marks it complete.

    Stands in for the reference's dask futures / delayed objects.
    """

    def __init__(self, tensor):
        self.tensor = tensor
        self._event = None
        if torch is not None and hasattr(tensor, "is_cuda") and tensor.is_cuda:
            self._event = torch.cuda.Event()
            self._event.record(torch.cuda.current_stream(tensor.device))

    def done(self):
        """True once the GPU has produced the result."""
        return self._event is None or self._event.query()

    def wait(self):
        """Block until the result is complete."""
        if self._event is not None:
            self._event.synchronize()
        return self

    def result(self):
        """The result as a host numpy array."""
        self.wait()
        t = self.tensor
        return t.detach().cpu().numpy() if hasattr(t, "detach") else numpy.asarray(t)

    compute = result


def _resolve(data):
    """Turn whatever a caller passes as facet / subgrid data into an array or tensor."""
    if isinstance(data, DeviceTask):
        return data.tensor
    for attr in ("compute", "result"):
        if hasattr(data, attr) and not hasattr(data, "shape"):
            return getattr(data, attr)()
    if callable(data) and not hasattr(data, "shape"):
        return data()
    return data


class _TaskQueue:
    """Bound on the number of unfinished results (``queue_size`` of the reference)."""

    def __init__(self, max_task):
        self.max_task = max(1, int(max_task))
        self.pending = collections.deque()

    def process(self, tasks):
        for task in tasks:
            self.pending.append(task)
        while len(self.pending) > self.max_task:
            self.pending.popleft().wait()

    def wait_all_done(self):
        while self.pending:
            self.pending.popleft().wait()


class _LRU:
    """Least-recently-used cache keyed by subgrid column offset."""

    def __init__(self, size):
        self.size = max(1, int(size))
        self.data = collections.OrderedDict()

    def get(self, key):
        if key not in self.data:
            return None
        self.data.move_to_end(key)
        return self.data[key]

    def set(self, key, value):
        """Insert; returns the evicted ``(key, value)`` or ``(None, None)``."""
        self.data[key] = value
        self.data.move_to_end(key)
        if len(self.data) <= self.size:
            return None, None
        return self.data.popitem(last=False)

    def pop_all(self):
        while self.data:
            yield self.data.popitem(last=False)


def _device_of(core):
    dev = getattr(core, "tensor_device", None)
    if dev is not None:
        return dev
    return torch.device("cuda", core.device)


def _to_device(data, device, dtype=None):
    """complex128 tensor on ``device`` from a numpy array / tensor."""
    dtype = dtype or torch.complex128
    if isinstance(data, torch.Tensor):
        return data.to(device=device, dtype=dtype, non_blocking=True)
    arr = numpy.asarray(data)
    if dtype == torch.complex128 and arr.dtype != numpy.complex128:
        arr = arr.astype(numpy.complex128)
    return torch.from_numpy(numpy.ascontiguousarray(arr)).to(device, non_blocking=True)


def _upload_iter(datas, device):
    """Yield device tensors for ``datas`` in order.

    Host sources are uploaded on a side stream one item ahead of the consumer, so that the
    H2D copy of facet ``j + 1`` overlaps the kernels working on facet ``j`` (pinned host
    memory makes the copy asynchronous; pageable memory still works, without overlap).
    """
    datas = list(datas)
    use_side = device.type == "cuda"
    side = torch.cuda.Stream(device) if use_side else None

    def upload(data):
        data = _resolve(data)
        if isinstance(data, torch.Tensor) and data.device == device:
            return _to_device(data, device), None
        if not use_side:
            return _to_device(data, device), None
        with torch.cuda.stream(side):
            t = _to_device(data, device)
            ev = torch.cuda.Event()
            ev.record(side)
        return t, ev

    nxt = upload(datas[0]) if datas else None
    for j in range(len(datas)):
        cur = nxt
        nxt = upload(datas[j + 1]) if j + 1 < len(datas) else None
        t, ev = cur
        if ev is not None:
            torch.cuda.current_stream(device).wait_event(ev)
            t.record_stream(torch.cuda.current_stream(device))
        yield t


def _device_mask(mask, device):
    """float64 device mask, or None when the mask is absent or all ones."""
    if mask is None:
        return None
    arr = numpy.asarray(mask, dtype=float)
    if arr.all() and (arr == 1).all():
        return None
    return torch.from_numpy(numpy.ascontiguousarray(arr)).to(device)


# ---------------------------------------------------------------------- forward
class SwiftlyForward:
    """Facet -> subgrid streaming transform on one GPU.

    :param swiftly_config: ``SwiftlyConfig``
    :param facet_tasks: list of ``(FacetConfig, data)``; ``data`` is a numpy array, a
        (CUDA) tensor, or something with ``.compute()`` / ``.result()``
    :param lru_forward: number of subgrid columns whose prepared facet columns
        (``NMBF_BF``) stay resident
    :param queue_size: maximum number of unfinished subgrid results
    :param client: accepted for compatibility, ignored
    :param bf_f_buffers: optional preallocated ``(yN, size)`` device tensors that
        receive the axis-0 prepared facets (they may reuse storage of facets that
        were consumed earlier in the list; facets are processed in list order)
    """

    # pylint: disable=too-many-arguments,too-many-instance-attributes
    def __init__(self, swiftly_config, facet_tasks, lru_forward=1, queue_size=20, client=None,
                 bf_f_buffers=None):
        self.config = swiftly_config
        self.facet_tasks = list(facet_tasks)
        self.core = swiftly_config.core
        self.device = _device_of(self.core)
        self.BF_Fs_persist = None
        self._bf_f_buffers = bf_f_buffers
        self.task_queue = _TaskQueue(queue_size)
        self.lru = _LRU(lru_forward)
        self._client = client
        # facets that share off0 form a "facet row": their contributions are combined
        # along axis 1 first, the rows then along axis 0
        rows = collections.OrderedDict()
        for idx, (cfg, _) in enumerate(self.facet_tasks):
            rows.setdefault(cfg.off0, []).append(idx)
        self._rows = list(rows.items())
        self._strips = None
        self._prep0 = None
        self._prep1 = None
        self._masks = {}
        self._fused = bool(getattr(self.core, "fused_forward_supported", lambda: False)())

    # -- stage 1: prepare every facet along axis 0 (once) --------------------------------
    def _get_BF_Fs(self):
        if self.BF_Fs_persist is None:
            out = []
            uploads = _upload_iter([data for _, data in self.facet_tasks], self.device)
            for idx, ((cfg, _), facet) in enumerate(zip(self.facet_tasks, uploads)):
                buf = None if self._bf_f_buffers is None else self._bf_f_buffers[idx]
                if self._fused:
                    # pre-windowed along axis 1 (the K2 kernel then skips its Fb multiply)
                    out.append(self.core.prepare_facet(facet, cfg.off0, axis=0, out=buf,
                                                       window_lines=True))
                else:
                    out.append(self.core.prepare_facet(facet, cfg.off0, axis=0, out=buf))
                del facet
            # facets are dead from here on: drop the references held by the task list
            self.facet_tasks = [(cfg, None) for cfg, _ in self.facet_tasks]
            self.BF_Fs_persist = out
            # (stage 1's scratch -- bounded at 2 GiB -- stays cached in the plan: freeing and
            # re-allocating it per transform cost 100-350 ms of cudaMalloc per step, measured.
            # ``core.release_scratch()`` gives it back explicitly.)
        return self.BF_Fs_persist

    # -- stage 2: per subgrid column ------------------------------------------------------
    def get_NMBF_BFs_off0(self, off0, BF_Fs):
        """Prepared facet columns for subgrid column ``off0`` (LRU cached)."""
        cached = self.lru.get(off0)
        if cached is None:
            reuse = None
            if len(self.lru.data) >= self.lru.size:
                # recycle the buffers of the column that is about to be evicted
                _, reuse = self.lru.data.popitem(last=False)
            if self._fused:
                cached = self.core.extract_columns(
                    BF_Fs, off0, [cfg.off1 for cfg, _ in self.facet_tasks], outs=reuse,
                    prewindowed=True)
            else:
                cached = [extract_column(self.core, BF_F, off0, cfg.off1)
                          for (cfg, _), BF_F in zip(self.facet_tasks, BF_Fs)]
            self.lru.set(off0, cached)
        return cached

    # -- stage 3: per subgrid ---------------------------------------------------------------
    def _gen_subgrid(self, subgrid_config, NMBF_BFs):
        core = self.core
        sg = subgrid_config
        if not self._fused:
            contribs = [core.extract_from_facet(nb, sg.off1, axis=1) for nb in NMBF_BFs]
            return sum_and_finish_subgrid(
                core, contribs, [cfg for cfg, _ in self.facet_tasks], sg
            )
        m = core.xM_yN_size
        shape = (len(self._rows), m, sg.size)
        if self._strips is None or tuple(self._strips.shape) != shape:
            # stored TRANSPOSED -- (row, xA, m), contribution index contiguous -- so that the
            # axis-0 kernel reads unit-stride lines; the axis-1 kernel's finished lines are
            # scattered into this layout by the TMA engine (bulk tensor stores)
            self._strips = torch.empty((shape[0], shape[2], shape[1]), dtype=torch.complex128,
                                       device=self.device).transpose(1, 2)
            self._prep0 = None
        mask0 = self._mask(sg, 0)
        mask1 = self._mask(sg, 1)
        # the argument blocks are built once per subgrid column (axis 1) / once per transform
        # (axis 0) and reused: per subgrid only offsets, masks and the output pointer change
        key = id(NMBF_BFs)
        if self._prep1 is None or self._prep1[0] != key or self._prep1[1] != sg.size:
            groups = [[(NMBF_BFs[j], self.facet_tasks[j][0].off1) for j in members]
                      for _, members in self._rows]
            self._prep1 = (key, sg.size, core.prepare_sum_finish(
                groups, 1, m, sg.size, (self._strips.stride(1), self._strips.stride(2))), NMBF_BFs)
        nrows = len(self._rows)
        self._prep1[2].launch([sg.off1] * nrows, [mask1] * nrows, out=self._strips,
                              out_group_stride=self._strips.stride(0))
        out = torch.empty((sg.size, sg.size), dtype=torch.complex128, device=self.device)
        if self._prep0 is None or self._prep0[0] != sg.size:
            sources = [[(self._strips[r], off0) for r, (off0, _) in enumerate(self._rows)]]
            self._prep0 = (sg.size, core.prepare_sum_finish(
                sources, 0, sg.size, sg.size, (out.stride(1), out.stride(0))))
        self._prep0[1].launch([sg.off0], [mask0], out=out)
        return out

    def _mask(self, sg, axis):
        """Device mask of a subgrid along ``axis`` (None when absent or all ones), cached by
        (axis, offset, size): a cover has only a few distinct masks per axis."""
        off = sg.off0 if axis == 0 else sg.off1
        key = (axis, off, sg.size)
        if key not in self._masks:
            self._masks[key] = _device_mask(sg.mask0 if axis == 0 else sg.mask1, self.device)
        return self._masks[key]

    def get_subgrid_task(self, subgrid_config):
        """Enqueue the computation of one subgrid and return its handle."""
        BF_Fs = self._get_BF_Fs()
        NMBF_BFs = self.get_NMBF_BFs_off0(subgrid_config.off0, BF_Fs)
        task = DeviceTask(self._gen_subgrid(subgrid_config, NMBF_BFs))
        self.task_queue.process([task])
        return task


# ---------------------------------------------------------------------- backward
class SwiftlyBackward:
    """Subgrid -> facet streaming accumulation on one GPU.

    :param swiftly_config: ``SwiftlyConfig``
    :param facets_config_list: facets to produce
    :param lru_backward: number of subgrid columns whose facet column accumulators
        (``NAF_MNAF``) stay resident before they are folded into the facets
    :param queue_size: maximum number of unfinished results
    :param client: accepted for compatibility, ignored
    """

    # pylint: disable=too-many-arguments
    def __init__(self, swiftly_config, facets_config_list, lru_backward=1, queue_size=20,
                 client=None):
        self.config = swiftly_config
        self.core = swiftly_config.core
        self.device = _device_of(self.core)
        self.facets_config_list = list(facets_config_list)
        self.MNAF_BMNAFs_persist = [None for _ in self.facets_config_list]
        self.task_queue = _TaskQueue(queue_size)
        self.lru = _LRU(lru_backward)
        self._client = client
        # fused kernels: one launch per subgrid for all facets (extract axis 1 + accumulate
        # column), one launch per finished column for all facets (finish axis 1 + mask + add
        # axis 0); otherwise the reference's task bodies run primitive by primitive
        self._fused = bool(hasattr(self.core, "subgrid_to_facets")
                           and getattr(self.core, "fused_backward_supported", lambda: False)())
        self._masks1 = None

    def add_new_subgrid_task(self, subgrid_config, new_subgrid_task):
        """Fold one subgrid into the facet accumulators."""
        off0, off1 = subgrid_config.off0, subgrid_config.off1
        subgrid = _to_device(_resolve(new_subgrid_task), self.device)
        if self._fused:
            done = self._add_subgrid_fused(subgrid, off0, off1)
        else:
            pieces = prepare_and_split_subgrid(
                self.core, subgrid, [off0, off1], self.facets_config_list
            )
            done = self.update_off0_NAF_MNAFs(off0, off1, pieces)
        self.task_queue.process(done)
        return done

    def _add_subgrid_fused(self, subgrid, off0, off1):
        core = self.core
        prepared = core.prepare_subgrid(subgrid, (off0, off1))
        blocks = {}
        for cfg in self.facets_config_list:
            if cfg.off0 not in blocks:
                blocks[cfg.off0] = core.extract_from_subgrid(prepared, cfg.off0, axis=0)
        column = self.lru.get(off0)
        if column is None:
            reuse = None
            if len(self.lru.data) >= self.lru.size:
                # fold the column that is about to be evicted NOW and recycle its buffers:
                # allocating the new column first would hold two columns (2 x 16 GiB at
                # N=65536) next to the 128 GiB of facet accumulators
                old_off0, reuse = self.lru.data.popitem(last=False)
                self.update_MNAF_BMNAFs(old_off0, reuse)
            if reuse is not None:
                column = reuse
                for acc in column:
                    acc.zero_()
            else:
                shape = (core.xM_yN_size, core.yN_size)
                column = [torch.zeros(shape, dtype=torch.complex128, device=self.device)
                          for _ in self.facets_config_list]
        core.subgrid_to_facets(
            [blocks[cfg.off0] for cfg in self.facets_config_list], column,
            [cfg.off1 for cfg in self.facets_config_list], off1)
        tasks = [DeviceTask(column[-1])] if column else []
        old_off0, old_column = self.lru.set(off0, column)
        if old_off0 is not None:
            self.update_MNAF_BMNAFs(old_off0, old_column)
        return tasks

    def update_off0_NAF_MNAFs(self, off0, off1, new_NAF_NAF_tasks):
        """Accumulate along axis 1 into the column accumulators of column ``off0``."""
        column = self.lru.get(off0)
        if column is None:
            column = [None for _ in self.facets_config_list]
        column = [
            accumulate_column(self.core, piece, acc, off1)
            for piece, acc in zip(new_NAF_NAF_tasks, column)
        ]
        tasks = [DeviceTask(column[-1])] if column else []
        old_off0, old_column = self.lru.set(off0, column)
        if old_off0 is not None:
            self.update_MNAF_BMNAFs(old_off0, old_column)
        return tasks

    def update_MNAF_BMNAFs(self, off0, new_NAF_MNAFs):
        """Finish a subgrid column along axis 1 and fold it into the facets (axis 0)."""
        if self._fused:
            core = self.core
            for j, cfg in enumerate(self.facets_config_list):
                if self.MNAF_BMNAFs_persist[j] is None:
                    self.MNAF_BMNAFs_persist[j] = torch.zeros(
                        (core.yN_size, cfg.size), dtype=torch.complex128, device=self.device)
            if self._masks1 is None:
                self._masks1 = [_device_mask(cfg.mask1, self.device)
                                for cfg in self.facets_config_list]
            core.fold_column(new_NAF_MNAFs, self.MNAF_BMNAFs_persist,
                             [cfg.off1 for cfg in self.facets_config_list], self._masks1, off0)
            return self.MNAF_BMNAFs_persist
        self.MNAF_BMNAFs_persist = [
            accumulate_facet(self.core, col, acc, cfg, off0)
            for cfg, col, acc in zip(
                self.facets_config_list, new_NAF_MNAFs, self.MNAF_BMNAFs_persist
            )
        ]
        return self.MNAF_BMNAFs_persist

    def finish(self):
        """Flush all pending columns and finish the facets; returns result handles."""
        for old_off0, old_column in self.lru.pop_all():
            self.update_MNAF_BMNAFs(old_off0, old_column)
        tasks = []
        for j, cfg in enumerate(self.facets_config_list):
            # release every accumulator as soon as its facet is finished: (yN, size) goes,
            # (size, size) stays -- at N=65536 on one GPU the 128 GiB of accumulators and the
            # 64 GiB of finished facets never coexist
            acc = self.MNAF_BMNAFs_persist[j]
            self.MNAF_BMNAFs_persist[j] = None
            facet = finish_facet(self.core, acc, cfg)
            del acc
            tasks.append(DeviceTask(facet))
        self.task_queue.process(tasks)
        self.task_queue.wait_all_done()
        return tasks
