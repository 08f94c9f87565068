// SwiFTly B200 -- C ABI of the fused forward-path entry points (device memory only).
#include <vector>

#include "capi_util.h"

using namespace swiftly;

// Fused extract_from_facet(axis 0) + prepare_facet(axis 1): the reference's
// `extract_column` task (api_helper.py:200-210).
extern "C" int swiftly_b200_extract_column(const swiftly_b200* h, const swiftly_b200_lines* bf_f,
                                           const swiftly_b200_lines* out, int64_t subgrid_off0,
                                           int64_t facet_off1, void* stream) {
    if (!h || !bf_f || !out) return einval("extract_column: NULL argument");
    if (bf_f->location != SWIFTLY_B200_DEVICE || out->location != SWIFTLY_B200_DEVICE)
        return einval("extract_column: device arrays only");
    const int64_t yN = h->yN, m = h->m, fs = bf_f->size;
    if (bf_f->n_lines != yN)
        return einval("extract_column: prepared facet must have yN_size lines, has " +
                      std::to_string(bf_f->n_lines));
    if (out->n_lines != m || out->size != yN)
        return einval("extract_column: output must be xM_yN_size lines of yN_size samples");
    if (fs > yN - 1) return einval("extract_column: facet size must be at most yN_size - 1");
    SW_CUDA(cudaSetDevice(h->device), "cudaSetDevice");
    const int64_t sc = floordiv(subgrid_off0 * yN, h->N);
    PrepareFacetOp op;
    op.g.in = (const cplx*)bf_f->data;
    op.g.out = (cplx*)out->data;
    op.g.in_ls = bf_f->line_stride;
    op.g.in_es = bf_f->elem_stride;
    op.g.out_ls = out->line_stride;
    op.g.out_es = out->elem_stride;
    op.g.n_lines = m;
    op.n = (int)yN;
    op.fs = (int)fs;
    op.fb = h->d_Fb + ((yN - 1) / 2 - fs / 2);
    op.shift_in = (int)pmod(fs / 2 - facet_off1, yN);
    op.scale = 1.0 / (double)yN;
    op.rm_m = (int)m;
    op.rm_s_m = (int)pmod(sc, m);
    op.rm_base = (int)pmod(yN / 2 - m / 2 + sc, yN);
    op.rm_mod = (int)yN;
    return run_prepare_facet(h, op, false, (cudaStream_t)stream);
}

extern "C" int swiftly_b200_sum_finish_axis_supported(const swiftly_b200* h) {
    return h ? subgrid_axis_conc((int)h->m, (int)h->xM) : 0;
}

// Fused extract_from_facet + add_to_subgrid (summed over sources) + finish_subgrid along
// one axis (SubgridAxisKernel, kernels.cuh), for several independent source groups in one
// launch: group g = sources [first_g, first_g + group_sizes[g]) -> out + g * out_group_stride.
static int sum_finish_groups(const swiftly_b200* h, const swiftly_b200_source* sources,
                             const int32_t* group_sizes, int n_groups,
                             const swiftly_b200_lines* out, int64_t out_group_stride,
                             const int64_t* subgrid_offs, const double* const* masks,
                             void* stream) {
    if (!h || !sources || !out || !group_sizes) return einval("sum_finish_axis: NULL argument");
    if (n_groups > SW_MAX_GROUPS)
        return einval("sum_finish_axis: at most " + std::to_string(SW_MAX_GROUPS) + " groups");
    if (out->location != SWIFTLY_B200_DEVICE) return einval("sum_finish_axis: device arrays only");
    const int64_t yN = h->yN, xM = h->xM, m = h->m;
    const int conc = subgrid_axis_conc((int)m, (int)xM);
    if (!conc) {
        set_error("sum_finish_axis: no fused kernel for m=" + std::to_string(m) +
                  ", xM=" + std::to_string(xM));
        return SWIFTLY_B200_EUNSUPPORTED;
    }
    if (n_groups < 1) return einval("sum_finish_axis: need at least one group");
    const int64_t sz = out->size;
    if (sz > xM) return einval("sum_finish_axis: subgrid size exceeds padded subgrid size");
    SW_CUDA(cudaSetDevice(h->device), "cudaSetDevice");

    // per group: schedule sources into rounds of `conc` with pairwise disjoint windows
    std::vector<std::vector<std::vector<int>>> rounds((size_t)n_groups);
    std::vector<int> pos_of;
    int first = 0;
    size_t max_rounds = 1;
    for (int g = 0; g < n_groups; ++g) {
        if (group_sizes[g] < 0) return einval("sum_finish_axis: negative group size");
        for (int i = first; i < first + group_sizes[g]; ++i) {
            const swiftly_b200_source& sr = sources[i];
            if (!sr.data) return einval("sum_finish_axis: NULL source pointer");
            if (sr.size != yN && sr.size != m)
                return einval("sum_finish_axis: source line length must be yN_size or xM_yN_size");
            const int64_t sf = floordiv(sr.facet_off * xM, h->N);
            pos_of.push_back((int)pmod(xM / 2 - m / 2 + sf, xM));
            bool placed = false;
            for (auto& r : rounds[g]) {
                if ((int)r.size() >= conc) continue;
                bool clash = false;
                for (int j : r) {
                    int64_t d1 = pmod(pos_of[j] - pos_of[i], xM);
                    int64_t d2 = pmod(pos_of[i] - pos_of[j], xM);
                    if (d1 < m || d2 < m) {
                        clash = true;
                        break;
                    }
                }
                if (!clash) {
                    r.push_back(i);
                    placed = true;
                    break;
                }
            }
            if (!placed) rounds[g].push_back(std::vector<int>(1, i));
        }
        first += group_sizes[g];
        if (rounds[g].size() > max_rounds) max_rounds = rounds[g].size();
    }
    const int slots = (int)max_rounds * conc;
    if ((int64_t)slots * n_groups > SW_MAX_SOURCES)
        return einval("sum_finish_axis: too many sources for one launch (" +
                      std::to_string(first) + " in " + std::to_string(n_groups) + " groups)");
    SubgridAxisArgs a;
    for (int i = 0; i < SW_MAX_SOURCES; ++i) {
        a.src[i].base = nullptr;
        a.src[i].ls = a.src[i].es = 0;
        a.src[i].wbase = a.src[i].s_m = a.src[i].sf_m = a.src[i].pos_base = 0;
        a.src[i].wmod = 1;
    }
    for (int g = 0; g < n_groups; ++g) {
        const int64_t sc = floordiv(subgrid_offs[g] * yN, h->N);
        for (size_t r = 0; r < rounds[g].size(); ++r) {
            for (size_t c = 0; c < rounds[g][r].size(); ++c) {
                const int i = rounds[g][r][c];
                const swiftly_b200_source& sr = sources[i];
                SgSource& d = a.src[(size_t)g * slots + r * conc + c];
                d.base = (const cplx*)sr.data;
                d.ls = sr.line_stride;
                d.es = sr.elem_stride;
                if (sr.size == yN && yN != m) {  // window of a prepared facet line
                    d.wbase = (int)pmod(yN / 2 - m / 2 + sc, yN);
                    d.s_m = (int)pmod(sc, m);
                    d.wmod = (int)yN;
                } else {  // already a contribution
                    d.wbase = 0;
                    d.s_m = 0;
                    d.wmod = (int)m;
                }
                const int64_t sf = floordiv(sr.facet_off * xM, h->N);
                d.sf_m = (int)pmod(sf, m);
                d.pos_base = pos_of[i];
            }
        }
    }
    // does the first round of every group tile the accumulator (conc disjoint windows of m
    // samples with conc * m == xM)?  Then it may store instead of accumulate.
    a.first_round_tiles = ((int64_t)conc * m == xM) ? 1 : 0;
    for (int g = 0; g < n_groups && a.first_round_tiles; ++g)
        if (rounds[g].empty() || (int)rounds[g][0].size() != conc) a.first_round_tiles = 0;
    a.n_slots = slots;
    a.n_groups = n_groups;
    a.n_lines = out->n_lines;
    a.out = (cplx*)out->data;
    a.out_ls = out->line_stride;
    a.out_es = out->elem_stride;
    a.out_gs = out_group_stride;
    a.sz = (int)sz;
    for (int g = 0; g < SW_MAX_GROUPS; ++g) {
        const int gg = g < n_groups ? g : 0;
        a.start[g] = (int)pmod(xM / 2 - sz / 2 + subgrid_offs[gg], xM);
        a.mask[g] = masks ? masks[gg] : nullptr;
    }
    return run_subgrid_axis(h, a, (cudaStream_t)stream);
}

extern "C" int swiftly_b200_sum_finish_axis_grouped(const swiftly_b200* h,
                                                    const swiftly_b200_source* sources,
                                                    const int32_t* group_sizes, int n_groups,
                                                    const swiftly_b200_lines* out,
                                                    int64_t out_group_stride,
                                                    int64_t subgrid_off, const double* mask,
                                                    void* stream) {
    if (n_groups < 1 || n_groups > SW_MAX_GROUPS)
        return einval("sum_finish_axis: between 1 and " + std::to_string(SW_MAX_GROUPS) + " groups");
    int64_t offs[SW_MAX_GROUPS];
    const double* masks[SW_MAX_GROUPS];
    for (int g = 0; g < n_groups; ++g) {
        offs[g] = subgrid_off;
        masks[g] = mask;
    }
    return sum_finish_groups(h, sources, group_sizes, n_groups, out, out_group_stride, offs, masks,
                             stream);
}

// Groups that belong to DIFFERENT subgrids (a batch of the multi-GPU driver): per-group
// subgrid offset and mask.
extern "C" int swiftly_b200_sum_finish_axis_batched(const swiftly_b200* h,
                                                    const swiftly_b200_source* sources,
                                                    const int32_t* group_sizes, int n_groups,
                                                    const swiftly_b200_lines* out,
                                                    int64_t out_group_stride,
                                                    const int64_t* subgrid_offs,
                                                    const double* const* masks, void* stream) {
    if (!subgrid_offs) return einval("sum_finish_axis: NULL subgrid offsets");
    return sum_finish_groups(h, sources, group_sizes, n_groups, out, out_group_stride,
                             subgrid_offs, masks, stream);
}

extern "C" int swiftly_b200_sum_finish_axis(const swiftly_b200* h,
                                            const swiftly_b200_source* sources, int n_sources,
                                            const swiftly_b200_lines* out, int64_t subgrid_off,
                                            const double* mask, void* stream) {
    if (n_sources < 0) return einval("sum_finish_axis: negative source count");
    int32_t one = n_sources;
    return swiftly_b200_sum_finish_axis_grouped(h, sources, &one, 1, out, 0, subgrid_off, mask,
                                                stream);
}

// extract_column for several facets in one launch (same subgrid_off0; per-facet off1).
extern "C" int swiftly_b200_extract_columns(const swiftly_b200* h, int n_facets,
                                            const swiftly_b200_lines* bf_f,
                                            const swiftly_b200_lines* out,
                                            int64_t subgrid_off0, const int64_t* facet_off1,
                                            void* stream) {
    if (!h || !bf_f || !out || !facet_off1) return einval("extract_columns: NULL argument");
    if (n_facets < 1 || n_facets > SW_MAX_COLUMN_FACETS)
        return einval("extract_columns: between 1 and " + std::to_string(SW_MAX_COLUMN_FACETS) +
                      " facets per call");
    const int64_t yN = h->yN, m = h->m;
    ExtractColumnsOp op;
    for (int f = 0; f < n_facets; ++f) {
        const swiftly_b200_lines& i = bf_f[f];
        const swiftly_b200_lines& o = out[f];
        if (i.location != SWIFTLY_B200_DEVICE || o.location != SWIFTLY_B200_DEVICE)
            return einval("extract_columns: device arrays only");
        if (i.n_lines != yN || i.elem_stride != 1 || o.elem_stride != 1)
            return einval("extract_columns: prepared facets must be yN_size contiguous rows");
        if (o.n_lines != m || o.size != yN)
            return einval("extract_columns: output must be xM_yN_size lines of yN_size samples");
        if (i.size > yN - 1) return einval("extract_columns: facet size must be at most yN_size - 1");
        ColumnFacet& F = op.fac[f];
        F.in = (const cplx*)i.data;
        F.out = (cplx*)o.data;
        F.in_ls = i.line_stride;
        F.out_ls = o.line_stride;
        F.fs = (int)i.size;
        F.shift_in = (int)pmod(i.size / 2 - facet_off1[f], yN);
        F.fb_off = (int)((yN - 1) / 2 - i.size / 2);
        F.pad_ = 0;
    }
    SW_CUDA(cudaSetDevice(h->device), "cudaSetDevice");
    const int64_t sc = floordiv(subgrid_off0 * yN, h->N);
    op.g.in = nullptr;
    op.g.out = nullptr;
    op.g.in_ls = op.g.in_es = op.g.out_ls = op.g.out_es = 0;
    op.g.n_lines = (int64_t)n_facets * m;
    op.fb = h->d_Fb;
    op.n = (int)yN;
    op.lines_per = (int)m;
    op.scale = 1.0 / (double)yN;
    op.rm_s_m = (int)pmod(sc, m);
    op.rm_base = (int)pmod(yN / 2 - m / 2 + sc, yN);
    return run_extract_columns(h, op, false, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------- fused backward path
// One subgrid -> column accumulators of all facets (SubgridToFacetsOp, kernels.cuh).
extern "C" int swiftly_b200_subgrid_to_facets(const swiftly_b200* h, int n_facets,
                                              const swiftly_b200_lines* blocks,
                                              const swiftly_b200_lines* accs,
                                              const int64_t* facet_off1, int64_t subgrid_off1,
                                              void* stream) {
    if (!h || !blocks || !accs || !facet_off1) return einval("subgrid_to_facets: NULL argument");
    if (n_facets < 1 || n_facets > SW_MAX_COLUMN_FACETS)
        return einval("subgrid_to_facets: between 1 and " +
                      std::to_string(SW_MAX_COLUMN_FACETS) + " facets per call");
    const int64_t yN = h->yN, xM = h->xM, m = h->m;
    SubgridToFacetsOp op;
    for (int f = 0; f < n_facets; ++f) {
        const swiftly_b200_lines& i = blocks[f];
        const swiftly_b200_lines& o = accs[f];
        if (i.location != SWIFTLY_B200_DEVICE || o.location != SWIFTLY_B200_DEVICE)
            return einval("subgrid_to_facets: device arrays only");
        if (i.n_lines != m || i.size != xM || i.elem_stride != 1)
            return einval("subgrid_to_facets: blocks must be xM_yN_size contiguous lines of xM_size");
        if (o.n_lines != m || o.size != yN || o.elem_stride != 1)
            return einval("subgrid_to_facets: accumulators must be xM_yN_size lines of yN_size");
        const int64_t sf = floordiv(facet_off1[f] * xM, h->N);
        BackFacet& F = op.fac[f];
        F.in = (const cplx*)i.data;
        F.out = (cplx*)o.data;
        F.in_ls = i.line_stride;
        F.out_ls = o.line_stride;
        F.sf_m = (int)pmod(sf, m);
        F.base_x = (int)pmod(xM / 2 - m / 2 + sf, xM);
    }
    SW_CUDA(cudaSetDevice(h->device), "cudaSetDevice");
    const int64_t sc = floordiv(subgrid_off1 * yN, h->N);
    op.g.in = nullptr;
    op.g.out = nullptr;
    op.g.in_ls = op.g.in_es = op.g.out_ls = op.g.out_es = 0;
    op.g.n_lines = (int64_t)n_facets * m;
    op.fn = h->d_Fn;
    op.m = (int)m;
    op.xM = (int)xM;
    op.yN = (int)yN;
    op.lines_per = (int)m;
    op.s_m = (int)pmod(sc, m);
    op.base_y = (int)pmod(yN / 2 - m / 2 + sc, yN);
    op.scale = 1.0 / (double)m;
    return run_subgrid_to_facets(h, op, false, (cudaStream_t)stream);
}

// Fold one finished subgrid column into all facet accumulators (FoldColumnOp, kernels.cuh).
extern "C" int swiftly_b200_fold_column(const swiftly_b200* h, int n_facets,
                                        const swiftly_b200_lines* accs,
                                        const swiftly_b200_lines* facet_accs,
                                        const int64_t* facet_off1, const double* const* mask1,
                                        int64_t subgrid_off0, void* stream) {
    if (!h || !accs || !facet_accs || !facet_off1) return einval("fold_column: NULL argument");
    if (n_facets < 1 || n_facets > SW_MAX_COLUMN_FACETS)
        return einval("fold_column: between 1 and " + std::to_string(SW_MAX_COLUMN_FACETS) +
                      " facets per call");
    const int64_t yN = h->yN, m = h->m;
    FoldColumnOp op;
    for (int f = 0; f < n_facets; ++f) {
        const swiftly_b200_lines& i = accs[f];
        const swiftly_b200_lines& o = facet_accs[f];
        if (i.location != SWIFTLY_B200_DEVICE || o.location != SWIFTLY_B200_DEVICE)
            return einval("fold_column: device arrays only");
        if (i.n_lines != m || i.size != yN || i.elem_stride != 1)
            return einval("fold_column: column accumulators must be xM_yN_size lines of yN_size");
        if (o.n_lines != yN || o.elem_stride != 1 || o.size > yN - 1)
            return einval("fold_column: facet accumulators must be yN_size lines of facet size");
        FoldFacet& F = op.fac[f];
        F.in = (const cplx*)i.data;
        F.out = (cplx*)o.data;
        F.mask = mask1 ? mask1[f] : nullptr;
        F.in_ls = i.line_stride;
        F.out_ls = o.line_stride;
        F.fs = (int)o.size;
        F.start1 = (int)pmod(yN / 2 - o.size / 2 + facet_off1[f], yN);
        F.fb_off = (int)((yN - 1) / 2 - o.size / 2);
        F.pad_ = 0;
    }
    SW_CUDA(cudaSetDevice(h->device), "cudaSetDevice");
    const int64_t sc = floordiv(subgrid_off0 * yN, h->N);
    op.g.in = nullptr;
    op.g.out = nullptr;
    op.g.in_ls = op.g.in_es = op.g.out_ls = op.g.out_es = 0;
    op.g.n_lines = (int64_t)n_facets * m;
    op.fb = h->d_Fb;
    op.n = (int)yN;
    op.lines_per = (int)m;
    op.s0_m = (int)pmod(sc, m);
    op.base0 = (int)pmod(yN / 2 - m / 2 + sc, yN);
    return run_fold_column(h, op, false, (cudaStream_t)stream);
}
