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

// ---- `.seq` text formatted on the device
#define SEQ_BLOCK 128
static inline int int_width(int32_t v) { // characters of the decimal form, sign included
    int n = v < 0 ? 2 : 1;
    for (int64_t x = v < 0 ? -(int64_t)v : v; x >= 10; x /= 10) n++;
    return n;
}
__device__ __forceinline__ int dev_digits(uint32_t x) {
    int n = 1;
    n += x >= 10u; n += x >= 100u; n += x >= 1000u; n += x >= 10000u; n += x >= 100000u; n += x >= 1000000u;
    n += x >= 10000000u; n += x >= 100000000u; n += x >= 1000000000u;
    return n;
}
__device__ __forceinline__ int dev_width(int32_t v) { return (v < 0) + dev_digits(v < 0 ? (uint32_t)(-(int64_t)v) : (uint32_t)v); }
__device__ __forceinline__ char *dev_put_int(char *p, int32_t v) {
    uint32_t x = v < 0 ? (uint32_t)(-(int64_t)v) : (uint32_t)v;
    if (v < 0) *p++ = '-';
    char *e = p + dev_digits(x), *q = e;
    do { *--q = (char)('0' + x % 10u); x /= 10u; } while (x);
    return e;
}
// line length of walks [lo, lo + m): tokens "<a>-<b>" joined by one space, '\n' terminated (String.join, CrossTimeGraph.java:136-137)
__global__ void __launch_bounds__(SEQ_BLOCK)
k_seq_lengths(const int32_t *__restrict__ tok, int64_t n, int64_t lo, int64_t m, int32_t L, const int32_t *__restrict__ layer,
              const int32_t *__restrict__ region, int position_prefix, int32_t *__restrict__ len) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    int total = 1; // '\n'
    for (int32_t j = 0; j < L; j++) {
        const int32_t t = tok[(int64_t)j * n + lo + i];
        if (t < 0) break;
        total += (j > 0) + dev_width(position_prefix ? j : layer[t]) + 1 + dev_width(region[t]);
    }
    len[i] = total;
}
// every block formats its SEQ_BLOCK lines into shared memory (they are contiguous in the file) and copies them out coalesced
__global__ void __launch_bounds__(SEQ_BLOCK)
k_seq_format(const int32_t *__restrict__ tok, int64_t n, int64_t lo, int64_t m, int32_t L, const int32_t *__restrict__ layer,
             const int32_t *__restrict__ region, int position_prefix, const int64_t *__restrict__ pos, char *__restrict__ out) {
    extern __shared__ char s_text[];
    const int64_t b0 = blockIdx.x * (int64_t)blockDim.x, i = b0 + threadIdx.x;
    const int64_t b1 = b0 + blockDim.x < m ? b0 + blockDim.x : m;
    const int64_t base = pos[b0];
    if (i < m) {
        char *p = s_text + (pos[i] - base);
        for (int32_t j = 0; j < L; j++) {
            const int32_t t = tok[(int64_t)j * n + lo + i];
            if (t < 0) break;
            if (j) *p++ = ' ';
            p = dev_put_int(p, position_prefix ? j : layer[t]);
            *p++ = '-';
            p = dev_put_int(p, region[t]);
        }
        *p = '\n';
    }
    __syncthreads();
    const int64_t bytes = pos[b1] - base;
    for (int64_t q = threadIdx.x; q < bytes; q += blockDim.x) out[base + q] = s_text[q];
}

extern "C" {

int dge_corpus_write_seq(const dge_corpus *c, const int32_t *label_layer, const int32_t *label_region,
                         int position_prefix, const char *path, int append) {
    if (!c) return dge_fail(nullptr, DGE_E_INVALID, "dge_corpus_write_seq: corpus is NULL");
    dge_ctx *ctx = c->ctx;
    if (!path || (c->n_ids > 0 && (!label_region || (!position_prefix && !label_layer))))
        return dge_fail(ctx, DGE_E_INVALID, "dge_corpus_write_seq: NULL argument");
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    FILE *f = fopen(path, append ? "ab" : "wb");
    if (!f) return dge_fail(ctx, DGE_E_IO, std::string("dge_corpus_write_seq: cannot open ") + path);
    if (c->n == 0) { return fclose(f) == 0 ? DGE_OK : dge_fail(ctx, DGE_E_IO, std::string("dge_corpus_write_seq: write failed: ") + path); }
    // The text is formatted ON THE DEVICE (the token matrix never crosses PCIe; a line is ~2.5x the bytes of its int32
    // tokens, but the host no longer spends ~20 ns per token in put_int): per chunk of walks, line lengths -> exclusive
    // scan -> every block formats its 128 lines into shared memory and copies them out coalesced; the chunk travels to
    // one of two pinned buffers while the previous chunk is being written to the file.
    const int32_t L = c->L, n_ids = c->n_ids;
    int d1 = 1, d2 = 1;                       // widest first number ("<layer>" or "<position>") and region id
    for (int32_t i = 0; i < n_ids; i++) {
        if (!position_prefix) d1 = std::max(d1, int_width(label_layer[i]));
        d2 = std::max(d2, int_width(label_region[i]));
    }
    if (position_prefix) d1 = int_width(L > 0 ? L - 1 : 0);
    const size_t max_line = (size_t)L * (size_t)(d1 + d2 + 2) + 1;            // "<a>-<b>" + separator per token, + '\n'
    if ((size_t)SEQ_BLOCK * max_line > 200 * 1024)
        return fclose(f), dge_fail(ctx, DGE_E_LIMIT, "dge_corpus_write_seq: lines of this length exceed the formatter's shared-memory tile");
    const int64_t chunk = std::max<int64_t>(SEQ_BLOCK, (int64_t)(((size_t)64 << 20) / max_line) / SEQ_BLOCK * SEQ_BLOCK);
    const size_t chunk_bytes = (size_t)chunk * max_line;
    int32_t *d_layer = nullptr, *d_region = nullptr, *d_len = nullptr;
    int64_t *d_pos = nullptr;
    char *d_text = nullptr, *h_text[2] = {nullptr, nullptr};
    auto cleanup = [&]() {
        dge_free(ctx, d_layer); dge_free(ctx, d_region); dge_free(ctx, d_len); dge_free(ctx, d_pos); dge_free(ctx, d_text);
        if (h_text[0]) cudaFreeHost(h_text[0]);
        if (h_text[1]) cudaFreeHost(h_text[1]);
    };
    if ((!position_prefix && dge_malloc(ctx, &d_layer, (size_t)n_ids) != cudaSuccess) || dge_malloc(ctx, &d_region, (size_t)n_ids) != cudaSuccess ||
        dge_malloc(ctx, &d_len, (size_t)chunk + 1) != cudaSuccess || dge_malloc(ctx, &d_pos, (size_t)chunk + 2) != cudaSuccess ||
        dge_malloc(ctx, &d_text, chunk_bytes) != cudaSuccess || cudaMallocHost((void **)&h_text[0], chunk_bytes) != cudaSuccess ||
        cudaMallocHost((void **)&h_text[1], chunk_bytes) != cudaSuccess) {
        cleanup(); fclose(f);
        return dge_fail(ctx, DGE_E_CUDA, "dge_corpus_write_seq: allocation failed");
    }
    cudaStream_t st = ctx->stream;
    if (!position_prefix) cudaMemcpyAsync(d_layer, label_layer, sizeof(int32_t) * (size_t)n_ids, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_region, label_region, sizeof(int32_t) * (size_t)n_ids, cudaMemcpyHostToDevice, st);
    cudaFuncSetAttribute(k_seq_format, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SEQ_BLOCK * max_line));
    bool io_ok = true;
    int rc = DGE_OK;
    size_t pending_bytes = 0;   // bytes of the previous chunk, sitting in h_text[(k - 1) & 1], not yet written
    int64_t k = 0;
    dge_phase_timer t(ctx, "seq_format");
    for (int64_t lo = 0; lo < c->n && io_ok && rc == DGE_OK; lo += chunk, k++) {
        const int64_t m = std::min<int64_t>(chunk, c->n - lo);
        const unsigned blocks = (unsigned)((m + SEQ_BLOCK - 1) / SEQ_BLOCK);
        k_seq_lengths<<<blocks, SEQ_BLOCK, 0, st>>>(c->tok, c->n, lo, m, L, d_layer, d_region, position_prefix, d_len);
        rc = dge_scan_i32(ctx, d_len, (int32_t)m, d_pos);
        if (rc != DGE_OK) break;
        k_seq_format<<<blocks, SEQ_BLOCK, SEQ_BLOCK * max_line, st>>>(c->tok, c->n, lo, m, L, d_layer, d_region, position_prefix, d_pos, d_text);
        ctx->launches += 2;
        int64_t total = 0;
        cudaMemcpyAsync(&total, d_pos + m, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_text[k & 1], d_text, (size_t)total, cudaMemcpyDeviceToHost, st);
        // while this chunk travels, the previous one goes to the file
        if (pending_bytes && fwrite(h_text[(k - 1) & 1], 1, pending_bytes, f) != pending_bytes) io_ok = false;
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) { rc = dge_fail(ctx, DGE_E_CUDA, std::string("dge_corpus_write_seq: ") + cudaGetErrorString(e)); break; }
        pending_bytes = (size_t)total;
    }
    if (rc == DGE_OK && io_ok && pending_bytes && fwrite(h_text[(k - 1) & 1], 1, pending_bytes, f) != pending_bytes) io_ok = false;
    t.stop();
    if (fclose(f) != 0) io_ok = false;
    cleanup();
    if (rc != DGE_OK) return rc;
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
