/*
 * swiftly_b200.h -- C ABI of the B200-native SwiFTly facet<->subgrid hot path.
 *
 * This is the drop-in boundary: the entry points below are what the reference's
 * native-backend adapter binds.  In the reference
 * (ska-sdp-distributed-fourier-transform, src/ska_sdp_exec_swiftly/
 * fourier_transform/core.py) the adapter class `SwiftlyCoreFunc` (core.py:487-929)
 * forwards each of the eight SwiFTly primitives to a method of the native object
 * `ska_sdp_func.fourier_transforms.swiftly.Swiftly(N, yN_size, xM_size, W)`
 * (core.py:508-510).  Those native methods always transform along the LAST axis
 * of a 2-D array and receive axis-0 work as strided transposed views
 * (core.py:577-630).  The functions here take the same information in plain C:
 * a batch of 1-D "lines" described by a base pointer, a line count, a line
 * length, and line/element strides -- so a C-ordered 2-D array along axis 1, the
 * same array along axis 0 (transposed view) and 1-D arrays are all one call.
 *
 * All samples are complex128 (interleaved re, im doubles).  Offsets are in image
 * / grid pixels exactly as in the reference (any integer, taken modulo).
 * Functions return 0 on success, a negative SWIFTLY_B200_E* code otherwise;
 * swiftly_b200_last_error() gives the message of the calling thread's last
 * failure.  A handle is immutable after creation: concurrent calls on
 * different streams are safe.  There is NO CPU implementation behind this ABI:
 * if no CUDA device is usable every call fails.
 */
#ifndef SWIFTLY_B200_H
#define SWIFTLY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWIFTLY_B200_OK 0
#define SWIFTLY_B200_EINVAL (-1)      /* bad argument / shape (Python: ValueError)     */
#define SWIFTLY_B200_ECUDA (-2)       /* CUDA runtime failure (Python: RuntimeError)   */
#define SWIFTLY_B200_EUNSUPPORTED (-3) /* FFT size not supported by this build          */

#define SWIFTLY_B200_DEVICE 0 /* `data` is a device pointer (cudaMalloc / torch)  */
#define SWIFTLY_B200_HOST 1   /* `data` is a host pointer; the library stages it  */

typedef struct swiftly_b200 swiftly_b200; /* opaque plan: tables live on one device */

/* A batch of 1-D lines of complex128 samples.
 * sample (line l, index i) lives at data[(l * line_stride + i * elem_stride)] (complex elements). */
typedef struct swiftly_b200_lines {
    void* data;
    int64_t n_lines;
    int64_t size;
    int64_t line_stride;
    int64_t elem_stride;
    int32_t location; /* SWIFTLY_B200_DEVICE or SWIFTLY_B200_HOST */
} swiftly_b200_lines;

/* Plan creation.  Replaces `Swiftly(N, yN_size, xM_size, W)` (core.py:508-510) and
 * `SwiftlyCore.__init__` / `check_params` (core.py:39-74).  Fb (yN_size-1 doubles,
 * core.py:104-108) and Fn (xM_size*yN_size/N doubles, core.py:110-117) are the
 * PSWF-derived window tables computed by the caller with the reference's scipy
 * formula (core.py:119-150); they are copied to the device.
 * Parameter violations (N % yN, N % xM, xM*yN % N) return SWIFTLY_B200_EINVAL. */
int swiftly_b200_create(double W, int64_t N, int64_t xM_size, int64_t yN_size,
                        const double* Fb, const double* Fn, int device, swiftly_b200** plan);
void swiftly_b200_destroy(swiftly_b200* plan);
const char* swiftly_b200_last_error(void);
/* Free the plan's per-stream scratch buffers (up to 2 GiB after prepare_facet along the strided
 * axis at N = 65536); they are re-created on demand.  Synchronises the owning streams. */
void swiftly_b200_release_scratch(swiftly_b200* plan);
/* Library identification: "swiftly_b200 <version> cuda sm_100a" (or "... EMULATED" for
 * the test-only host build, which the product never loads). */
const char* swiftly_b200_build_info(void);

int64_t swiftly_b200_contribution_size(const swiftly_b200* plan); /* xM_yN_size, core.py:48 */

/* ---- facet -> subgrid ------------------------------------------------------------ */
/* SwiftlyCore.prepare_facet (core.py:189-222) / Swiftly.prepare_facet (core.py:684-692).
 * in: n_lines x facet_size, out: n_lines x yN_size (overwritten). */
int swiftly_b200_prepare_facet(const swiftly_b200* plan, const swiftly_b200_lines* in,
                               const swiftly_b200_lines* out, int64_t facet_off, void* stream);
/* SwiftlyCore.extract_from_facet (core.py:224-253) / Swiftly.extract_from_facet (:713-721).
 * in: n_lines x yN_size, out: n_lines x xM_yN_size (overwritten). */
int swiftly_b200_extract_from_facet(const swiftly_b200* plan, const swiftly_b200_lines* in,
                                    const swiftly_b200_lines* out, int64_t subgrid_off,
                                    void* stream);
/* SwiftlyCore.add_to_subgrid (core.py:255-285) / Swiftly.add_to_subgrid (:742-750).
 * in: n_lines x xM_yN_size, out: n_lines x xM_size, ACCUMULATED into (out +=). */
int swiftly_b200_add_to_subgrid(const swiftly_b200* plan, const swiftly_b200_lines* in,
                                const swiftly_b200_lines* out, int64_t facet_off, void* stream);
/* One axis of SwiftlyCore.finish_subgrid (core.py:287-325) / Swiftly.finish_subgrid
 * (core.py:795-812, called once per axis).  in: n_lines x xM_size, out: n_lines x
 * subgrid_size (overwritten).  mask: optional subgrid_size doubles on the same
 * location as `out` data (0/1 mask of api_helper.py:107-111 folded into the store) or NULL. */
int swiftly_b200_finish_subgrid(const swiftly_b200* plan, const swiftly_b200_lines* in,
                                const swiftly_b200_lines* out, int64_t subgrid_off,
                                const double* mask, void* stream);

/* ---- subgrid -> facet ------------------------------------------------------------ */
/* One axis of SwiftlyCore.prepare_subgrid (core.py:328-368) / Swiftly.prepare_subgrid_inplace
 * (core.py:837-853).  in: n_lines x subgrid_size, out: n_lines x xM_size (overwritten). */
int swiftly_b200_prepare_subgrid(const swiftly_b200* plan, const swiftly_b200_lines* in,
                                 const swiftly_b200_lines* out, int64_t subgrid_off, void* stream);
/* SwiftlyCore.extract_from_subgrid (core.py:370-406) / Swiftly.extract_from_subgrid (:866-876).
 * in: n_lines x xM_size, out: n_lines x xM_yN_size (overwritten). */
int swiftly_b200_extract_from_subgrid(const swiftly_b200* plan, const swiftly_b200_lines* in,
                                      const swiftly_b200_lines* out, int64_t facet_off,
                                      void* stream);
/* SwiftlyCore.add_to_facet (core.py:408-449) / Swiftly.add_to_facet (:890-900).
 * in: n_lines x xM_yN_size, out: n_lines x yN_size, ACCUMULATED into (out +=). */
int swiftly_b200_add_to_facet(const swiftly_b200* plan, const swiftly_b200_lines* in,
                              const swiftly_b200_lines* out, int64_t subgrid_off, void* stream);
/* SwiftlyCore.finish_facet (core.py:452-484) / Swiftly.finish_facet (:916-926).
 * in: n_lines x yN_size, out: n_lines x facet_size (overwritten).  mask as in finish_subgrid
 * (api_helper.py:175-176, 195-196) or NULL. */
int swiftly_b200_finish_facet(const swiftly_b200* plan, const swiftly_b200_lines* in,
                              const swiftly_b200_lines* out, int64_t facet_off,
                              const double* mask, void* stream);

/* ---- fused forward path (device memory only) ---------------------------------------- */
/* The reference's `extract_column` task (api_helper.py:200-210) in one kernel:
 * extract_from_facet(BF_F, subgrid_off0, axis=0) followed by prepare_facet(., facet_off1,
 * axis=1).  bf_f: yN_size lines (rows of the axis-0 prepared facet) of facet_size samples;
 * out: xM_yN_size lines of yN_size samples (overwritten). */
int swiftly_b200_extract_column(const swiftly_b200* plan, const swiftly_b200_lines* bf_f,
                                const swiftly_b200_lines* out, int64_t subgrid_off0,
                                int64_t facet_off1, void* stream);

/* One input of swiftly_b200_sum_finish_axis: `n_lines` lines (n_lines of the output) of
 * `size` samples.  size == yN_size: lines of a prepared facet, the contribution window for
 * `subgrid_off` is extracted on the fly (extract_from_facet, core.py:224-253);
 * size == xM_yN_size: lines that already are contributions. */
typedef struct swiftly_b200_source {
    const void* data; /* device pointer */
    int64_t line_stride;
    int64_t elem_stride;
    int64_t size;
    int64_t facet_off; /* facet offset along the transformed axis */
} swiftly_b200_source;

/* One axis of the reference's `sum_and_finish_subgrid` task (api_helper.py:73-112) in one
 * kernel: for every line, sum_g add_to_subgrid(extract(source_g), facet_off_g) is built in
 * shared memory and finished (finish_subgrid along this axis, mask folded in).  out:
 * n_lines x subgrid_size (overwritten).  Returns SWIFTLY_B200_EUNSUPPORTED when the
 * (xM_yN_size, xM_size) pair has no fused instantiation (callers then use the primitives). */
int swiftly_b200_sum_finish_axis(const swiftly_b200* plan, const swiftly_b200_source* sources,
                                 int n_sources, const swiftly_b200_lines* out,
                                 int64_t subgrid_off, const double* mask, void* stream);
/* The same for several independent source groups in ONE launch (e.g. all facet rows of a
 * subgrid): group g takes sources [sum(group_sizes[:g]), ... + group_sizes[g]) and writes
 * out->data + g * out_group_stride (complex elements); `out` describes one group's lines. */
int swiftly_b200_sum_finish_axis_grouped(const swiftly_b200* plan,
                                         const swiftly_b200_source* sources,
                                         const int32_t* group_sizes, int n_groups,
                                         const swiftly_b200_lines* out, int64_t out_group_stride,
                                         int64_t subgrid_off, const double* mask, void* stream);
/* As above, but the groups may belong to different subgrids (a batch of the multi-GPU
 * driver): subgrid_offs[g] and masks[g] (masks or masks[g] may be NULL) per group
 * (any number of groups: larger jobs are cut into several launches). */
int swiftly_b200_sum_finish_axis_batched(const swiftly_b200* plan,
                                         const swiftly_b200_source* sources,
                                         const int32_t* group_sizes, int n_groups,
                                         const swiftly_b200_lines* out, int64_t out_group_stride,
                                         const int64_t* subgrid_offs, const double* const* masks,
                                         void* stream);
/* As swiftly_b200_sum_finish_axis_batched, but every group writes to its OWN buffer:
 * out_ptrs[g] is the base address of group g's output, `out` gives the (common) shape and
 * strides.  The sharded driver passes the owners' peer-mapped receive buffers, so the strips
 * of a subgrid travel over NVLink as the kernel's epilogue (TMA bulk tensor stores) instead
 * of through a separate collective (replaces the Dask transfer of contributions,
 * reference api.py:263-277). */
int swiftly_b200_sum_finish_axis_scattered(const swiftly_b200* plan,
                                           const swiftly_b200_source* sources,
                                           const int32_t* group_sizes, int n_groups,
                                           const swiftly_b200_lines* out, void* const* out_ptrs,
                                           const int64_t* subgrid_offs,
                                           const double* const* masks, void* stream);
/* Ordering between the ranks of the sharded transform (peer_sync.cu): `flags_dev_table` is a
 * DEVICE array of n_peers device pointers -- this rank's mappings of every rank's flag array
 * (n_peers int64 each, peer-mapped / symmetric memory).  signal stores `value` into entry
 * my_rank of every rank's array (system-scope release, after everything queued before on the
 * stream); wait spins (system-scope acquire) until all n_peers entries of `my_flags` are
 * >= value, or sets *status (device int) to 1 + peer after timeout_s. */
int swiftly_b200_peer_signal(const swiftly_b200* plan, void* const* flags_dev_table, int n_peers,
                             int my_rank, int64_t value, void* stream);
int swiftly_b200_peer_wait(const swiftly_b200* plan, const void* my_flags, int n_peers,
                           int64_t value, double timeout_s, void* status, void* stream);
/* swiftly_b200_extract_column for n_facets (<= 64) facets in ONE launch: bf_f[f] / out[f]
 * as in swiftly_b200_extract_column (contiguous rows), facet_off1[f] per facet. */
int swiftly_b200_extract_columns(const swiftly_b200* plan, int n_facets,
                                 const swiftly_b200_lines* bf_f, const swiftly_b200_lines* out,
                                 int64_t subgrid_off0, const int64_t* facet_off1, void* stream);
/* The pair the fused forward driver uses between stage 1 and K2: prepare_facet whose output
 * LINES are pre-multiplied by the Fb window of the other axis (line l by Fb_c[l], the factor
 * prepare_facet(axis 1) would apply to sample l), and extract_columns that takes such rows and
 * skips its own window multiply.  Together they equal prepare_facet + extract_columns
 * (core.py:189-222 applied along axis 0, then api_helper.py:200-210). */
int swiftly_b200_prepare_facet_windowed(const swiftly_b200* plan, const swiftly_b200_lines* in,
                                        const swiftly_b200_lines* out, int64_t facet_off,
                                        void* stream);
int swiftly_b200_extract_columns_windowed(const swiftly_b200* plan, int n_facets,
                                          const swiftly_b200_lines* bf_f,
                                          const swiftly_b200_lines* out, int64_t subgrid_off0,
                                          const int64_t* facet_off1, void* stream);
/* ---- fused backward path (device memory only) --------------------------------------- */
/* One subgrid into the column accumulators of n_facets (<= 64) facets in ONE launch: per
 * facet `extract_from_subgrid(block, facet_off1, axis=1)` followed by `accumulate_column` =
 * `add_to_facet(., subgrid_off1, axis=1, out=acc)` (api_helper.py:115-152).  blocks[f]: the
 * (xM_yN_size x xM_size) result of extract_from_subgrid(axis 0) for the facet's off0;
 * accs[f]: (xM_yN_size x yN_size) accumulator NAF_MNAF, ACCUMULATED into. */
int swiftly_b200_subgrid_to_facets(const swiftly_b200* plan, int n_facets,
                                   const swiftly_b200_lines* blocks,
                                   const swiftly_b200_lines* accs, const int64_t* facet_off1,
                                   int64_t subgrid_off1, void* stream);
/* Fold a finished subgrid column into n_facets facet accumulators in ONE launch: per facet
 * `finish_facet(acc, facet_off1, size, axis=1)`, optional mask1, `add_to_facet(., subgrid_off0,
 * axis=0, out=facet_acc)` (api_helper.py:155-179).  facet_accs[f]: (yN_size x facet_size)
 * accumulator MNAF_BMNAF, ACCUMULATED into; mask1 or mask1[f] may be NULL. */
int swiftly_b200_fold_column(const swiftly_b200* plan, int n_facets,
                             const swiftly_b200_lines* accs,
                             const swiftly_b200_lines* facet_accs, const int64_t* facet_off1,
                             const double* const* mask1, int64_t subgrid_off0, void* stream);

/* xM_size / xM_yN_size if the fused kernel exists for this plan, else 0. */
int swiftly_b200_sum_finish_axis_supported(const swiftly_b200* plan);

#ifdef __cplusplus
}
#endif
#endif /* SWIFTLY_B200_H */
