// io.cu -- the two on-disk text formats of the hot path, byte-compatible with the reference's
// consumers:
//   .seq  CrossTimeGraph.java:136-137 / SpatialGraph.java:105-110  (String.join(" ", seq) + "\n")
//   .vec  WordVectorSerializer.writeWordVectors (DeepWalk.java:82); parsed by
//         python/embeddingEvaluation_tract.py:139-166 with skipheader=0
#include "dge_internal.cuh"
#include <cstdio>
#include <cstring>
#include <thread>

static inline char *put_int(char *p, int32_t v) {
    if (v < 0) { *p++ = '-'; v = -v; }
    char tmp[12];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

// formats walks [lo,hi) into buf, returns bytes written
static size_t format_walks(const int32_t *tok, int64_t lo, int64_t hi, int32_t L, const int32_t *layer,
                           const int32_t *region, int position_prefix, char *buf) {
    char *p = buf;
    for (int64_t i = lo; i < hi; i++) {
        const int32_t *s = tok + i * L;
        for (int32_t j = 0; j < L && s[j] >= 0; j++) {
            if (j) *p++ = ' ';
            p = put_int(p, position_prefix ? j : layer[s[j]]);
            *p++ = '-';
            p = put_int(p, region[s[j]]);
        }
        *p++ = '\n';
    }
    return (size_t)(p - buf);
}

extern "C" {

int dge_corpus_write_seq(const dge_corpus *c, const int32_t *label_layer, const int32_t *label_region,
                         int position_prefix, const char *path, int append) {
    if (!c) return dge_fail(nullptr, DGE_E_INVALID, "dge_corpus_write_seq: corpus is NULL");
    dge_ctx *ctx = c->ctx;
    if (!path || !label_region || (!position_prefix && !label_layer))
        return dge_fail(ctx, DGE_E_INVALID, "dge_corpus_write_seq: NULL argument");
    size_t total = (size_t)c->n * (size_t)c->L;
    int32_t *tok = (int32_t *)dge_host_alloc((total ? total : 1) * sizeof(int32_t));
    if (!tok) return dge_fail(ctx, DGE_E_CUDA, "dge_corpus_write_seq: pinned allocation failed");
    int rc = dge_corpus_tokens(c, tok);
    if (rc != DGE_OK) { dge_host_free(tok); return rc; }
    FILE *f = fopen(path, append ? "ab" : "wb");
    if (!f) {
        dge_host_free(tok);
        return dge_fail(ctx, DGE_E_IO, std::string("dge_corpus_write_seq: cannot open ") + path);
    }
    // chunked, multi-threaded formatting; chunks are written in order
    const int64_t chunk = 1 << 16;
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > 16) nt = 16;
    const size_t per_walk = (size_t)c->L * 24 + 2; // two ints of <= 11 chars + '-' + ' ' per token
    std::vector<std::vector<char>> bufs(nt);
    std::vector<size_t> lens(nt);
    bool io_ok = true;
    for (int64_t base = 0; base < c->n && io_ok; base += chunk * nt) {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; t++) {
            int64_t lo = base + (int64_t)t * chunk, hi = std::min<int64_t>(lo + chunk, c->n);
            lens[t] = 0;
            if (lo >= hi) continue;
            bufs[t].resize((size_t)(hi - lo) * per_walk);
            th.emplace_back([&, t, lo, hi]() {
                lens[t] = format_walks(tok, lo, hi, c->L, label_layer, label_region, position_prefix, bufs[t].data());
            });
        }
        for (auto &x : th) x.join();
        for (unsigned t = 0; t < nt && io_ok; t++)
            if (lens[t] && fwrite(bufs[t].data(), 1, lens[t], f) != lens[t]) io_ok = false;
    }
    if (fclose(f) != 0) io_ok = false;
    dge_host_free(tok);
    if (!io_ok) return dge_fail(ctx, DGE_E_IO, std::string("dge_corpus_write_seq: write failed: ") + path);
    return DGE_OK;
}

// `.seq` reader: what FileSentenceIterator / LineSentenceIterator + DefaultTokenizerFactory hand to Word2Vec
// (DeepWalk.java:47-59,70) when the corpus files already exist (checkInputFile :86-87,99-100 skips generation).
int dge_corpus_read_seq(dge_ctx *ctx, const char *path, const int32_t *label_layer, const int32_t *label_region,
                        int32_t n_ids, int position_prefix, dge_corpus **out) {
    if (!ctx) return dge_fail(nullptr, DGE_E_INVALID, "dge_corpus_read_seq: ctx is NULL");
    if (!out) return dge_fail(ctx, DGE_E_INVALID, "dge_corpus_read_seq: out is NULL");
    *out = nullptr;
    if (!path || n_ids < 0 || (n_ids > 0 && (!label_region || (!position_prefix && !label_layer))))
        return dge_fail(ctx, DGE_E_INVALID, "dge_corpus_read_seq: NULL argument");
    FILE *f = fopen(path, "rb");
    if (!f) return dge_fail(ctx, DGE_E_IO, std::string("dge_corpus_read_seq: cannot open ") + path);
    std::vector<char> buf;
    {
        char chunk[1 << 16];
        size_t got;
        while ((got = fread(chunk, 1, sizeof(chunk), f)) > 0) buf.insert(buf.end(), chunk, chunk + got);
    }
    bool read_ok = !ferror(f);
    fclose(f);
    if (!read_ok) return dge_fail(ctx, DGE_E_IO, std::string("dge_corpus_read_seq: read failed: ") + path);
    // (layer, region) -> id; in position-prefix mode the first number of a token is the walk position, not a label
    std::map<std::pair<int32_t, int32_t>, int32_t> ids;
    for (int32_t i = 0; i < n_ids; i++) {
        auto key = std::make_pair(position_prefix ? 0 : label_layer[i], label_region[i]);
        if (!ids.emplace(key, i).second)
            return dge_fail(ctx, DGE_E_INVALID, "dge_corpus_read_seq: duplicate label");
    }
    // pass 1: lines and the longest line
    int64_t n_lines = 0;
    int32_t L = 0, cur = 0;
    bool in_tok = false;
    for (size_t i = 0; i <= buf.size(); i++) {
        char ch = i < buf.size() ? buf[i] : '\n';
        if (ch == ' ' || ch == '\t' || ch == '\r' || ch == '\n') {
            if (in_tok) { cur++; in_tok = false; }
            if (ch == '\n') {
                if (i < buf.size() || cur > 0) { n_lines++; if (cur > L) L = cur; }
                cur = 0;
            }
        } else in_tok = true;
    }
    std::vector<int32_t> tok((size_t)n_lines * (size_t)L + 1, -1);
    // pass 2: parse "<a>-<b>"
    int64_t line = 0;
    int32_t pos = 0;
    size_t i = 0;
    const size_t n = buf.size();
    auto parse_int = [&](int32_t &v) -> bool {
        bool neg = false;
        if (i < n && buf[i] == '-') { neg = true; i++; }
        if (i >= n || buf[i] < '0' || buf[i] > '9') return false;
        int64_t x = 0;
        while (i < n && buf[i] >= '0' && buf[i] <= '9') { x = x * 10 + (buf[i] - '0'); if (x > 2147483647LL) return false; i++; }
        v = (int32_t)(neg ? -x : x);
        return true;
    };
    while (i < n && line < n_lines) {
        char ch = buf[i];
        if (ch == '\n') { line++; pos = 0; i++; continue; }
        if (ch == ' ' || ch == '\t' || ch == '\r') { i++; continue; }
        int32_t a, b;
        if (!parse_int(a) || i >= n || buf[i] != '-') return dge_fail(ctx, DGE_E_INVALID, "dge_corpus_read_seq: malformed token at line " + std::to_string(line + 1));
        i++;
        if (!parse_int(b)) return dge_fail(ctx, DGE_E_INVALID, "dge_corpus_read_seq: malformed token at line " + std::to_string(line + 1));
        auto it = ids.find(std::make_pair(position_prefix ? 0 : a, b));
        if (it == ids.end()) return dge_fail(ctx, DGE_E_INVALID, "dge_corpus_read_seq: unknown label at line " + std::to_string(line + 1));
        tok[(size_t)line * L + pos++] = it->second;
    }
    return dge_corpus_from_tokens(ctx, tok.data(), n_lines, L, n_ids, out);
}

int dge_model_write_vec(const dge_model *m, const int32_t *label_layer, const int32_t *label_region, const char *path) {
    if (!m) return dge_fail(nullptr, DGE_E_INVALID, "dge_model_write_vec: model is NULL");
    dge_ctx *ctx = m->ctx;
    if (!path || !label_layer || !label_region) return dge_fail(ctx, DGE_E_INVALID, "dge_model_write_vec: NULL argument");
    std::vector<float> syn0((size_t)m->V * m->dim + 1);
    std::vector<int32_t> ids((size_t)m->V + 1);
    int rc = dge_model_vectors(m, syn0.data(), nullptr, ids.data());
    if (rc != DGE_OK) return rc;
    FILE *f = fopen(path, "wb");
    if (!f) return dge_fail(ctx, DGE_E_IO, std::string("dge_model_write_vec: cannot open ") + path);
    bool ok = true;
    for (int32_t wd = 0; wd < m->V && ok; wd++) {
        int32_t id = ids[wd];
        if (fprintf(f, "%d-%d", label_layer[id], label_region[id]) < 0) ok = false;
        for (int32_t d = 0; d < m->dim && ok; d++)
            if (fprintf(f, " %.9g", (double)syn0[(size_t)wd * m->dim + d]) < 0) ok = false;
        if (fputc('\n', f) == EOF) ok = false;
    }
    if (fclose(f) != 0) ok = false;
    if (!ok) return dge_fail(ctx, DGE_E_IO, std::string("dge_model_write_vec: write failed: ") + path);
    return DGE_OK;
}

// ---- static flow-graph exports for the LINE / matrix-factorisation baselines.  The slot sums are
// computed on the device (dge_flows_slot_weights); formatting is the host's part, as in the Java loops.
static int flows_export_args(const dge_flows *f, const int32_t *rows, int32_t n_rows, const int32_t *cols, int32_t n_cols,
                             const char *path, const char *who) {
    if (!f) return dge_fail(nullptr, DGE_E_INVALID, std::string(who) + ": flows is NULL");
    if (!path || n_rows < 0 || n_cols < 0 || (n_rows && !rows) || (n_cols && !cols))
        return dge_fail(f->ctx, DGE_E_INVALID, std::string(who) + ": NULL argument");
    for (int32_t i = 0; i < n_rows; i++)
        if (rows[i] < 0 || rows[i] >= f->n) return dge_fail(f->ctx, DGE_E_INVALID, std::string(who) + ": row index out of range");
    for (int32_t i = 0; i < n_cols; i++)
        if (cols[i] < 0 || cols[i] >= f->n) return dge_fail(f->ctx, DGE_E_INVALID, std::string(who) + ": column index out of range");
    return DGE_OK;
}

int dge_flows_write_matrix(const dge_flows *f, int mode, int32_t lo, int32_t hi, const int32_t *rows, int32_t n_rows,
                           const int32_t *cols, int32_t n_cols, char sep, const char *path) {
    int rc = flows_export_args(f, rows, n_rows, cols, n_cols, path, "dge_flows_write_matrix");
    if (rc != DGE_OK) return rc;
    dge_ctx *ctx = f->ctx;
    std::vector<int32_t> W((size_t)f->n * f->n + 1);
    rc = dge_flows_slot_weights(f, mode, lo, hi, W.data());
    if (rc != DGE_OK) return rc;
    FILE *out = fopen(path, "wb");
    if (!out) return dge_fail(ctx, DGE_E_IO, std::string("dge_flows_write_matrix: cannot open ") + path);
    std::vector<char> line((size_t)n_cols * 12 + 2);
    bool ok = true;
    for (int32_t a = 0; a < n_rows && ok; a++) {
        char *p = line.data();
        const int32_t *w = W.data() + (size_t)rows[a] * f->n;
        for (int32_t b = 0; b < n_cols; b++) {
            if (b) *p++ = sep;
            p = put_int(p, w[cols[b]]);
        }
        *p++ = '\n';
        ok = fwrite(line.data(), 1, (size_t)(p - line.data()), out) == (size_t)(p - line.data());
    }
    if (fclose(out) != 0) ok = false;
    if (!ok) return dge_fail(ctx, DGE_E_IO, std::string("dge_flows_write_matrix: write failed: ") + path);
    return DGE_OK;
}

int dge_flows_write_od(const dge_flows *f, int mode, int32_t lo, int32_t hi, const int32_t *rows, int32_t n_rows,
                       const int32_t *cols, int32_t n_cols, const int32_t *region_ids, int keep_zero,
                       int32_t presence_hour, const char *path) {
    int rc = flows_export_args(f, rows, n_rows, cols, n_cols, path, "dge_flows_write_od");
    if (rc != DGE_OK) return rc;
    dge_ctx *ctx = f->ctx;
    if (!region_ids && f->n) return dge_fail(ctx, DGE_E_INVALID, "dge_flows_write_od: region_ids is NULL");
    if (presence_hour > 23) return dge_fail(ctx, DGE_E_INVALID, "dge_flows_write_od: presence_hour must be < 24");
    std::vector<int32_t> W((size_t)f->n * f->n + 1), P;
    rc = dge_flows_slot_weights(f, mode, lo, hi, W.data());
    if (rc == DGE_OK && presence_hour >= 0) {
        P.resize((size_t)f->n * f->n + 1);
        rc = dge_flows_slot_weights(f, 1, presence_hour, presence_hour, P.data());
    }
    if (rc != DGE_OK) return rc;
    FILE *out = fopen(path, "wb");
    if (!out) return dge_fail(ctx, DGE_E_IO, std::string("dge_flows_write_od: cannot open ") + path);
    std::vector<char> buf((size_t)n_cols * 40 + 2);
    bool ok = true;
    for (int32_t a = 0; a < n_rows && ok; a++) {
        char *p = buf.data();
        const size_t base = (size_t)rows[a] * f->n;
        for (int32_t b = 0; b < n_cols; b++) {
            const int32_t w = W[base + cols[b]];
            if (presence_hour >= 0 && P[base + cols[b]] <= 0) continue;
            if (w <= 0 && !keep_zero) continue;
            p = put_int(p, region_ids[rows[a]]); *p++ = ' ';
            p = put_int(p, region_ids[cols[b]]); *p++ = ' ';
            p = put_int(p, w); *p++ = '\n';
        }
        const size_t len = (size_t)(p - buf.data());
        if (len) ok = fwrite(buf.data(), 1, len, out) == len;
    }
    if (fclose(out) != 0) ok = false;
    if (!ok) return dge_fail(ctx, DGE_E_IO, std::string("dge_flows_write_od: write failed: ") + path);
    return DGE_OK;
}

} // extern "C"
