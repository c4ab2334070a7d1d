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

} // extern "C"
