// gvl_host_variants.cuh -- HOST-buffer entries of the `variants` / `variant-windows` outputs (included at the end of
// gvl_host.cu: it uses that file's upload helpers).  Argument lists follow the reference's #[pyfunction]s
// (src/ffi/mod.rs:255-630, 2808); the reference returns freshly allocated arrays, here the results stay in device scratch
// and the caller pulls each one with gvl_variants_fetch once it knows the sizes.
#pragma once

namespace gvl {

constexpr size_t VS_IN = 30;   // scratch slots: inputs 30..39, intermediates 40..43, results 44..51
constexpr size_t VS_TMP = 40;
constexpr size_t VS_RES = 44;

// one i64 from the device (the total behind an offsets array)
static int read_i64(gvl_ctx *ctx, const int64_t *dev, int64_t *out) {
    GVL_CUDA(cudaMemcpyAsync(out, dev, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->own_stream));
    GVL_CUDA(cudaStreamSynchronize(ctx->own_stream));
    return GVL_OK;
}

static int d2h(gvl_ctx *ctx, void *host, const void *dev, int64_t bytes) {
    if (bytes <= 0) return GVL_OK;
    GVL_CUDA(cudaMemcpyAsync(host, dev, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->own_stream));
    GVL_CUDA(cudaStreamSynchronize(ctx->own_stream));
    return GVL_OK;
}

static int result(gvl_ctx *ctx, size_t j, int64_t bytes, void **dev) {
    int rc;
    if ((rc = scratch(ctx, VS_RES + j, bytes + 16, dev))) return rc;
    if (ctx->var_results.size() <= j) ctx->var_results.resize(j + 1, {nullptr, 0});
    ctx->var_results[j] = {*dev, bytes};
    return GVL_OK;
}

}  // namespace gvl

extern "C" {

int gvl_variants_fetch(gvl_ctx *ctx, int which, void *host_out, int64_t bytes) {
    if (!ctx || which < 0 || (size_t)which >= ctx->var_results.size() || !ctx->var_results[which].first)
        return fail(GVL_ERR_STATE, "gvl_variants_fetch: no result %d", which);
    if (bytes != ctx->var_results[which].second)
        return fail(GVL_ERR_ARG, "gvl_variants_fetch: result %d holds %lld bytes, caller asked for %lld", which,
                    (long long)ctx->var_results[which].second, (long long)bytes);
    if (bytes && !host_out) return fail(GVL_ERR_ARG, "gvl_variants_fetch: host_out is NULL");
    GVL_CUDA(cudaSetDevice(ctx->device));
    return d2h(ctx, host_out, ctx->var_results[which].first, bytes);
}

int gvl_gather_rows(gvl_ctx *ctx, const int64_t *geno_offset_idx, int64_t n_rows, const int64_t *geno_offsets, int64_t n_geno,
                    const void *data, int64_t n_data, int64_t *out_offsets, int64_t *total) {
    if (!ctx || !out_offsets || !total || n_rows < 0 || n_geno < 0 || n_data < 0 || (n_rows && (!geno_offset_idx || !geno_offsets)))
        return fail(GVL_ERR_ARG, "gvl_gather_rows: NULL argument or bad size");
    GVL_CUDA(cudaSetDevice(ctx->device));
    ctx->var_results.clear();
    int rc;
    const void *d_go, *d_data;
    if ((rc = static_dev(ctx, geno_offsets, sizeof(int64_t) * 2 * n_geno, VS_IN, &d_go))) return rc;
    if ((rc = static_dev(ctx, data, 4 * n_data, VS_IN + 1, &d_data))) return rc;
    Packer pk;
    size_t i_goi = pk.add(geno_offset_idx, sizeof(int64_t) * n_rows);
    if ((rc = pk.upload(ctx, VS_IN + 2))) return rc;
    void *d_off, *d_out;
    if ((rc = scratch(ctx, VS_TMP, sizeof(int64_t) * (n_rows + 1), &d_off))) return rc;
    const int64_t *starts = (const int64_t *)d_go, *stops = starts ? starts + n_geno : nullptr;
    if ((rc = gvl_dev_gather_rows_offsets(ctx, pk.ptr<int64_t>(i_goi), n_rows, starts, stops, (int64_t *)d_off, ctx->own_stream)))
        return rc;
    if ((rc = d2h(ctx, out_offsets, d_off, sizeof(int64_t) * (n_rows + 1)))) return rc;
    *total = out_offsets[n_rows];
    if (*total < 0) return fail(GVL_ERR_ARG, "gvl_gather_rows: negative row length");
    if ((rc = result(ctx, 0, 4 * *total, &d_out))) return rc;
    return gvl_dev_gather_rows(ctx, pk.ptr<int64_t>(i_goi), n_rows, starts, d_data, (const int64_t *)d_off, *total, d_out,
                               ctx->own_stream);
}

int gvl_gather_alleles(gvl_ctx *ctx, const int32_t *v_idxs, int64_t n, const uint8_t *allele_bytes, int64_t n_bytes,
                       const int64_t *allele_offsets, int64_t n_table, int64_t *seq_offsets, int64_t *total) {
    if (!ctx || !seq_offsets || !total || n < 0 || n_bytes < 0 || n_table < 0 || (n && (!v_idxs || !allele_offsets)))
        return fail(GVL_ERR_ARG, "gvl_gather_alleles: NULL argument or bad size");
    GVL_CUDA(cudaSetDevice(ctx->device));
    ctx->var_results.clear();
    int rc;
    const void *d_ab, *d_ao;
    if ((rc = static_dev(ctx, allele_bytes, n_bytes, VS_IN, &d_ab))) return rc;
    if ((rc = static_dev(ctx, allele_offsets, sizeof(int64_t) * (n_table + 1), VS_IN + 1, &d_ao))) return rc;
    Packer pk;
    size_t i_v = pk.add(v_idxs, sizeof(int32_t) * n);
    if ((rc = pk.upload(ctx, VS_IN + 2))) return rc;
    void *d_off, *d_out;
    if ((rc = scratch(ctx, VS_TMP, sizeof(int64_t) * (n + 1), &d_off))) return rc;
    if ((rc = gvl_dev_gather_alleles_offsets(ctx, pk.ptr<int32_t>(i_v), n, (const int64_t *)d_ao, (int64_t *)d_off,
                                             ctx->own_stream)))
        return rc;
    if ((rc = d2h(ctx, seq_offsets, d_off, sizeof(int64_t) * (n + 1)))) return rc;
    *total = seq_offsets[n];
    if ((rc = result(ctx, 0, *total, &d_out))) return rc;
    return gvl_dev_gather_alleles(ctx, pk.ptr<int32_t>(i_v), n, (const uint8_t *)d_ab, (const int64_t *)d_ao,
                                  (const int64_t *)d_off, *total, nullptr, 1, d_out, ctx->own_stream);
}

int gvl_rc_alleles(gvl_ctx *ctx, uint8_t *byte_data, int64_t n_bytes, const int64_t *seq_offsets, int64_t n_alleles,
                   const int64_t *var_offsets, int64_t n_rows, const uint8_t *to_rc_row) {
    if (!ctx || n_bytes < 0 || n_alleles < 0 || n_rows < 0 || !seq_offsets || !var_offsets || (n_rows && !to_rc_row) ||
        (n_bytes && !byte_data))
        return fail(GVL_ERR_ARG, "gvl_rc_alleles: NULL argument or bad size");
    if (seq_offsets[0] != 0 || seq_offsets[n_alleles] > n_bytes)
        return fail(GVL_ERR_ARG, "gvl_rc_alleles: seq_offsets must start at 0 and end inside byte_data");
    GVL_CUDA(cudaSetDevice(ctx->device));
    int rc;
    Packer pk;
    size_t i_b = pk.add(byte_data, n_bytes);
    size_t i_so = pk.add(seq_offsets, sizeof(int64_t) * (n_alleles + 1));
    size_t i_vo = pk.add(var_offsets, sizeof(int64_t) * (n_rows + 1));
    size_t i_rc = pk.add(to_rc_row, n_rows);
    if ((rc = pk.upload(ctx, VS_IN))) return rc;
    uint8_t *d_b = const_cast<uint8_t *>(pk.ptr<uint8_t>(i_b));
    if ((rc = gvl_dev_rc_alleles(ctx, d_b, pk.ptr<int64_t>(i_so), n_alleles, pk.ptr<int64_t>(i_vo), n_rows, pk.ptr<uint8_t>(i_rc),
                                 seq_offsets[n_alleles], ctx->own_stream)))
        return rc;
    return d2h(ctx, byte_data, d_b, n_bytes);
}

int gvl_compact_keep(gvl_ctx *ctx, const void *values, int64_t n, const int64_t *row_offsets, int64_t n_rows,
                     const uint8_t *keep, int64_t *new_offsets, int64_t *total) {
    if (!ctx || !row_offsets || !new_offsets || !total || n < 0 || n_rows < 0 || (n && (!values || !keep)))
        return fail(GVL_ERR_ARG, "gvl_compact_keep: NULL argument or bad size");
    GVL_CUDA(cudaSetDevice(ctx->device));
    ctx->var_results.clear();
    int rc;
    Packer pk;
    size_t i_v = pk.add(values, 4 * n);
    size_t i_ro = pk.add(row_offsets, sizeof(int64_t) * (n_rows + 1));
    size_t i_k = pk.add(keep, n);
    if ((rc = pk.upload(ctx, VS_IN))) return rc;
    void *d_pos, *d_no, *d_out;
    if ((rc = scratch(ctx, VS_TMP, sizeof(int64_t) * (n + 1), &d_pos))) return rc;
    if ((rc = scratch(ctx, VS_TMP + 1, sizeof(int64_t) * (n_rows + 1), &d_no))) return rc;
    if ((rc = gvl_dev_compact_keep_offsets(ctx, pk.ptr<uint8_t>(i_k), n, pk.ptr<int64_t>(i_ro), n_rows, (int64_t *)d_pos,
                                           (int64_t *)d_no, ctx->own_stream)))
        return rc;
    if ((rc = read_i64(ctx, (const int64_t *)d_pos + n, total))) return rc;
    if ((rc = d2h(ctx, new_offsets, d_no, sizeof(int64_t) * (n_rows + 1)))) return rc;
    if ((rc = result(ctx, 0, 4 * *total, &d_out))) return rc;
    return gvl_dev_compact_keep(ctx, pk.ptr<uint8_t>(i_v), pk.ptr<uint8_t>(i_k), n, (const int64_t *)d_pos, d_out, ctx->own_stream);
}

int gvl_fill_empty_fixed(gvl_ctx *ctx, const void *data, int64_t n_data, const int64_t *offsets, int64_t n_rows, int64_t inner,
                         uint32_t fill_bits, int64_t *new_offsets, int64_t *new_total) {
    if (!ctx || !offsets || !new_offsets || !new_total || n_data < 0 || n_rows < 0 || inner < 0 || (n_data && !data))
        return fail(GVL_ERR_ARG, "gvl_fill_empty_fixed: NULL argument or bad size");
    GVL_CUDA(cudaSetDevice(ctx->device));
    ctx->var_results.clear();
    int rc;
    Packer pk;
    size_t i_d = pk.add(data, 4 * n_data);
    size_t i_o = pk.add(offsets, sizeof(int64_t) * (n_rows + 1));
    if ((rc = pk.upload(ctx, VS_IN))) return rc;
    void *d_no, *d_out;
    if ((rc = scratch(ctx, VS_TMP, sizeof(int64_t) * (n_rows + 1), &d_no))) return rc;
    if ((rc = gvl_dev_fill_empty_offsets(ctx, pk.ptr<int64_t>(i_o), n_rows, (int64_t *)d_no, ctx->own_stream))) return rc;
    if ((rc = d2h(ctx, new_offsets, d_no, sizeof(int64_t) * (n_rows + 1)))) return rc;
    *new_total = new_offsets[n_rows];
    if ((rc = result(ctx, 0, 4 * *new_total * inner, &d_out))) return rc;
    return gvl_dev_fill_empty_fixed(ctx, pk.ptr<uint8_t>(i_d), pk.ptr<int64_t>(i_o), n_rows, (const int64_t *)d_no, *new_total,
                                    inner, fill_bits, d_out, ctx->own_stream);
}

int gvl_fill_empty_seq(gvl_ctx *ctx, const void *data, int itemsize, int64_t n_data, const int64_t *var_offsets, int64_t n_rows,
                       const int64_t *seq_offsets, int64_t n_vars, const void *dummy, int64_t dummy_len,
                       int64_t *new_var_offsets, int64_t *n_new_vars, int64_t *total) {
    if (!ctx || !var_offsets || !seq_offsets || !new_var_offsets || !n_new_vars || !total || n_data < 0 || n_rows < 0 ||
        n_vars < 0 || dummy_len < 0 || (n_data && !data) || (dummy_len && !dummy))
        return fail(GVL_ERR_ARG, "gvl_fill_empty_seq: NULL argument or bad size");
    if (itemsize != 1 && itemsize != 4) return fail(GVL_ERR_ARG, "gvl_fill_empty_seq: items of 1 or 4 bytes");
    GVL_CUDA(cudaSetDevice(ctx->device));
    ctx->var_results.clear();
    int rc;
    Packer pk;
    size_t i_d = pk.add(data, itemsize * n_data);
    size_t i_vo = pk.add(var_offsets, sizeof(int64_t) * (n_rows + 1));
    size_t i_so = pk.add(seq_offsets, sizeof(int64_t) * (n_vars + 1));
    size_t i_du = pk.add(dummy, itemsize * dummy_len);
    if ((rc = pk.upload(ctx, VS_IN))) return rc;
    void *d_nv, *d_src, *d_ns, *d_out;
    if ((rc = scratch(ctx, VS_TMP, sizeof(int64_t) * (n_rows + 1), &d_nv))) return rc;
    if ((rc = gvl_dev_fill_empty_offsets(ctx, pk.ptr<int64_t>(i_vo), n_rows, (int64_t *)d_nv, ctx->own_stream))) return rc;
    if ((rc = d2h(ctx, new_var_offsets, d_nv, sizeof(int64_t) * (n_rows + 1)))) return rc;
    *n_new_vars = new_var_offsets[n_rows];
    if ((rc = scratch(ctx, VS_TMP + 1, sizeof(int64_t) * (*n_new_vars + 1), &d_src))) return rc;
    if ((rc = result(ctx, 1, sizeof(int64_t) * (*n_new_vars + 1), &d_ns))) return rc;
    if ((rc = gvl_dev_fill_empty_seq_offsets(ctx, pk.ptr<int64_t>(i_vo), n_rows, pk.ptr<int64_t>(i_so), dummy_len,
                                             (const int64_t *)d_nv, *n_new_vars, (int64_t *)d_src, (int64_t *)d_ns,
                                             ctx->own_stream)))
        return rc;
    if ((rc = read_i64(ctx, (const int64_t *)d_ns + *n_new_vars, total))) return rc;
    if ((rc = result(ctx, 0, itemsize * *total, &d_out))) return rc;
    return gvl_dev_fill_empty_seq(ctx, pk.ptr<uint8_t>(i_d), itemsize, pk.ptr<int64_t>(i_so), pk.ptr<uint8_t>(i_du),
                                  (const int64_t *)d_src, (const int64_t *)d_ns, *n_new_vars, *total, d_out, ctx->own_stream);
}

int gvl_assemble_variant_buffers(gvl_ctx *ctx, int64_t mode, const int32_t *v_idxs, int64_t n, const uint8_t *alt_global,
                                 const int64_t *alt_off_global, const uint8_t *ref_global, const int64_t *ref_off_global,
                                 int64_t n_variants, int want_ref_bytes, int want_flank, int64_t ref_mode, int64_t alt_mode,
                                 int64_t flank_len, const void *lut, int tok_bytes, const int32_t *v_contigs,
                                 const int32_t *v_starts, const int32_t *ilens, const uint8_t *reference,
                                 const int64_t *ref_offsets, int64_t n_contigs, uint8_t pad_char, int32_t *n_fields,
                                 int32_t *field_kind, int64_t *field_items, int32_t *field_tok) {
    if (!ctx || !n_fields || !field_kind || !field_items || !field_tok || n < 0 || n_variants < 0 || n_contigs < 0 || flank_len < 0 ||
        !alt_off_global || !v_starts || !ilens || !ref_offsets || (n && !v_idxs))
        return fail(GVL_ERR_ARG, "gvl_assemble_variant_buffers: NULL argument or bad size");
    if (mode != 0 && mode != 1) return fail(GVL_ERR_ARG, "gvl_assemble_variant_buffers: mode must be 0 (variants) or 1 (windows)");
    const bool tokens = mode == 1 || want_flank;
    if (tokens && (!lut || (tok_bytes != 1 && tok_bytes != 4)))  // windows.rs:187 / src/ffi/mod.rs:514: the reference panics
        return fail(GVL_ERR_ARG, "gvl_assemble_variant_buffers: tokens requested but no token LUT (1- or 4-byte tokens)");
    if (mode == 1 && ref_mode == 2 && (!ref_global || !ref_off_global))  // windows.rs:272-273
        return fail(GVL_ERR_ARG, "gvl_assemble_variant_buffers: bare ref allele needs the REF byte table");
    GVL_CUDA(cudaSetDevice(ctx->device));
    ctx->var_results.clear();
    int rc;
    const void *d;
    gvl_sparse_tables t;
    memset(&t, 0, sizeof(t));
    t.n_variants = n_variants;
    t.n_contigs = n_contigs;
    if ((rc = static_dev(ctx, alt_off_global, sizeof(int64_t) * (n_variants + 1), VS_IN, &d))) return rc;
    t.alt_offsets = (const int64_t *)d;
    if ((rc = static_dev(ctx, alt_global, alt_off_global[n_variants], VS_IN + 1, &d))) return rc;
    t.alt_alleles = (const uint8_t *)d;
    if ((rc = static_dev(ctx, v_starts, sizeof(int32_t) * n_variants, VS_IN + 2, &d))) return rc;
    t.v_starts = (const int32_t *)d;
    if ((rc = static_dev(ctx, ilens, sizeof(int32_t) * n_variants, VS_IN + 3, &d))) return rc;
    t.ilens = (const int32_t *)d;
    if ((rc = static_dev(ctx, ref_offsets, sizeof(int64_t) * (n_contigs + 1), VS_IN + 4, &d))) return rc;
    t.ref_offsets = (const int64_t *)d;
    if ((rc = static_dev(ctx, reference, ref_offsets[n_contigs], VS_IN + 5, &d))) return rc;
    t.ref = (const uint8_t *)d;
    const uint8_t *d_rg = nullptr;
    const int64_t *d_ro = nullptr;
    if (ref_global && ref_off_global) {
        if ((rc = static_dev(ctx, ref_off_global, sizeof(int64_t) * (n_variants + 1), VS_IN + 6, &d))) return rc;
        d_ro = (const int64_t *)d;
        if ((rc = static_dev(ctx, ref_global, ref_off_global[n_variants], VS_IN + 7, &d))) return rc;
        d_rg = (const uint8_t *)d;
    }
    Packer pk;
    size_t i_v = pk.add(v_idxs, sizeof(int32_t) * n);
    size_t i_c = pk.add(v_contigs, sizeof(int32_t) * n);
    size_t i_l = pk.add(tokens ? lut : nullptr, 256 * (int64_t)tok_bytes);
    if ((rc = pk.upload(ctx, VS_IN + 8))) return rc;
    const int32_t *dv = pk.ptr<int32_t>(i_v), *dc = pk.ptr<int32_t>(i_c);
    const void *dl = pk.ptr<uint8_t>(i_l);
    cudaStream_t st = ctx->own_stream;
    int nf = 0;
    // one field: offsets (result 2 nf + 1), total, data (result 2 nf)
    auto alleles = [&](int kind, const uint8_t *bytes, const int64_t *aoff, const void *l) -> int {
        void *d_off, *d_out;
        int64_t tot = 0;
        int r;
        if ((r = result(ctx, 2 * nf + 1, sizeof(int64_t) * (n + 1), &d_off))) return r;
        if ((r = gvl_dev_gather_alleles_offsets(ctx, dv, n, aoff, (int64_t *)d_off, st))) return r;
        if ((r = read_i64(ctx, (const int64_t *)d_off + n, &tot))) return r;
        const int tb = l ? tok_bytes : 1;
        if ((r = result(ctx, 2 * nf, tot * tb, &d_out))) return r;
        if ((r = gvl_dev_gather_alleles(ctx, dv, n, bytes, aoff, (const int64_t *)d_off, tot, l, tb, d_out, st))) return r;
        field_kind[nf] = kind, field_items[nf] = tot, field_tok[nf] = tb;
        nf++;
        return GVL_OK;
    };
    auto window = [&](int kind, int wkind) -> int {
        void *d_off, *d_out;
        int64_t tot = 0;
        int r;
        if ((r = result(ctx, 2 * nf + 1, sizeof(int64_t) * (n + 1), &d_off))) return r;
        if ((r = gvl_dev_variant_windows_offsets(ctx, &t, dv, n, flank_len, wkind, (int64_t *)d_off, st))) return r;
        if ((r = read_i64(ctx, (const int64_t *)d_off + n, &tot))) return r;
        if ((r = result(ctx, 2 * nf, tot * tok_bytes, &d_out))) return r;
        if ((r = gvl_dev_variant_windows(ctx, &t, dv, dc, n, flank_len, wkind, pad_char, dl, tok_bytes, (const int64_t *)d_off, tot,
                                         d_out, st)))
            return r;
        field_kind[nf] = kind, field_items[nf] = tot, field_tok[nf] = tok_bytes;
        nf++;
        return GVL_OK;
    };
    if (mode == 0) {  // windows.rs:162-219
        if ((rc = alleles(0, t.alt_alleles, t.alt_offsets, nullptr))) return rc;
        if (want_ref_bytes && d_rg && (rc = alleles(1, d_rg, d_ro, nullptr))) return rc;
        if (want_flank) {
            void *d_out;
            const int64_t tot = n * 2 * flank_len;
            if ((rc = result(ctx, 2 * nf + 1, 0, &d_out))) return rc;  // (no offsets: the caller's row_offsets)
            if ((rc = result(ctx, 2 * nf, tot * tok_bytes, &d_out))) return rc;
            if ((rc = gvl_dev_variant_windows(ctx, &t, dv, dc, n, flank_len, GVL_WINDOW_FLANKS, pad_char, dl, tok_bytes, nullptr,
                                              tot, d_out, st)))
                return rc;
            field_kind[nf] = 2, field_items[nf] = tot, field_tok[nf] = tok_bytes;
            nf++;
        }
    } else {  // windows.rs:227-296
        if (ref_mode == 1 && (rc = window(3, GVL_WINDOW_REF))) return rc;
        if (ref_mode == 2 && (rc = alleles(1, d_rg, d_ro, dl))) return rc;
        if (alt_mode == 1 && (rc = window(4, GVL_WINDOW_ALT))) return rc;
        if (alt_mode == 2 && (rc = alleles(0, t.alt_alleles, t.alt_offsets, dl))) return rc;
    }
    *n_fields = nf;
    GVL_CUDA(cudaStreamSynchronize(st));
    return GVL_OK;
}

}  // extern "C"
