// plan_host.cu — b200ols_group_plan_build: host orchestration of the device-side `.over()` key planning
// (kernels and the algorithm: group_plan.cuh).  Three short host synchronisations per plan: the varying-bit mask
// (decides the radix passes), the number of groups (sizes the tables), the offsets themselves (host metadata of
// b200ols_frame, [G+1] int64).  The permutation stays on the device.
#include "engine_ctx.h"
#include "group_plan.cuh"

using namespace b200;

namespace b200 {

struct GroupPlan {
    // device workspace (grow-only): [images A|B][idx A|B][hist][flags][cnt][raw keys][row_index][offsets|first_row][scalars]
    char *ws = nullptr;
    size_t ws_cap = 0;
    // pinned host mirror for scalars / offsets / first rows
    char *hbuf = nullptr;
    size_t hbuf_cap = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // current plan
    int64_t n_rows = 0, n_groups = 0;
    const int64_t *offsets_host = nullptr, *first_row_host = nullptr;
    const int64_t *offsets_dev = nullptr;
    const int64_t *row_index_dev = nullptr;
    bool valid = false;
};

}  // namespace b200

void group_plan_destroy(GroupPlan *g) {
    if (!g) return;
    if (g->ws) cudaFree(g->ws);
    if (g->hbuf) cudaFreeHost(g->hbuf);
    if (g->ev0) cudaEventDestroy(g->ev0);
    if (g->ev1) cudaEventDestroy(g->ev1);
    delete g;
}

static size_t al(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

static int key_bytes(int dtype) { return (dtype == KEY_I32 || dtype == KEY_U32 || dtype == KEY_F32) ? 4 : 8; }

// exclusive scan of a[0..m) in place; *total (device) receives the sum when not null
static int exclusive_scan_u32(b200ols_ctx *c, uint32_t *a, int64_t m, uint32_t *seg, uint32_t *total) {
    const int64_t ns = (m + SCAN_SEG - 1) / SCAN_SEG;
    scan_seg_sum_kernel<<<static_cast<unsigned>(ns), PLAN_THREADS, 0, c->stream>>>(a, m, seg);
    scan_seg_scan_kernel<<<1, PLAN_THREADS, 0, c->stream>>>(seg, ns, total);
    scan_seg_apply_kernel<<<static_cast<unsigned>(ns), PLAN_THREADS, 0, c->stream>>>(a, m, seg);
    c->launches += 3;
    CU(cudaGetLastError());
    return 0;
}

extern "C" int b200ols_group_plan_build(b200ols_ctx *c, const b200ols_key_column *keys, int32_t n_keys, int64_t n,
                                        int32_t memspace, b200ols_group_plan *out) {
    if (!c || !out) return fail(B200OLS_ERR_INVALID, "NULL argument");
    if (n_keys < 1 || n_keys > PLAN_MAX_KEYS) return fail(B200OLS_ERR_INVALID, "n_keys must be in [1, %d]", PLAN_MAX_KEYS);
    if (!keys) return fail(B200OLS_ERR_INVALID, "keys is NULL");
    if (n < 0) return fail(B200OLS_ERR_INVALID, "n_rows < 0");
    if (n >= (static_cast<int64_t>(1) << 31))
        return fail(B200OLS_ERR_UNSUPPORTED, "group planning of %lld rows: frames of 2^31 rows or more do not fit one device", (long long)n);
    if (memspace != B200OLS_HOST && memspace != B200OLS_DEVICE) return fail(B200OLS_ERR_INVALID, "bad memspace");
    for (int j = 0; j < n_keys; ++j) {
        if (keys[j].dtype < KEY_I64 || keys[j].dtype > KEY_F32) return fail(B200OLS_ERR_INVALID, "keys[%d]: bad dtype %d", j, keys[j].dtype);
        if (n > 0 && !keys[j].values) return fail(B200OLS_ERR_INVALID, "keys[%d].values is NULL", j);
    }
    CU(cudaSetDevice(c->device));
    if (!c->gplan) c->gplan = new GroupPlan();
    GroupPlan *g = c->gplan;
    g->valid = false;
    if (!g->ev0) {
        CU(cudaEventCreate(&g->ev0));
        CU(cudaEventCreate(&g->ev1));
    }
    std::memset(out, 0, sizeof(*out));
    out->n_rows = n;

    const int64_t nb = std::max<int64_t>(1, (n + PLAN_TILE - 1) / PLAN_TILE);
    const int64_t hist_n = 256 * nb;
    const int64_t seg_n = (std::max(hist_n, nb) + SCAN_SEG - 1) / SCAN_SEG + 1;
    size_t raw_bytes = 0;
    if (memspace == B200OLS_HOST)
        for (int j = 0; j < n_keys; ++j) raw_bytes += al(static_cast<size_t>(n) * key_bytes(keys[j].dtype) + 16);
    const size_t N = static_cast<size_t>(n);
    const size_t need = 2 * al(N * 8) + 2 * al(N * 4) + al(hist_n * 4) + al(N) + al((nb + 1) * 4) + al(seg_n * 4) + raw_bytes +
                        al(N * 8) + 2 * al((N + 1) * 8) + al(sizeof(PlanScalars) + 16) + 4096;
    if (need > g->ws_cap) {
        CU(cudaStreamSynchronize(c->stream));
        if (g->ws) cudaFree(g->ws);
        g->ws = nullptr;
        g->ws_cap = 0;
        void *p = nullptr;
        const size_t cap = need + (need >> 3);
        CU(cudaMalloc(&p, cap));
        g->ws = static_cast<char *>(p);
        g->ws_cap = cap;
    }
    const size_t hneed = al(sizeof(PlanScalars)) + 2 * al((N + 1) * 8) + 64;  // worst case: every row its own group
    // the host mirror grows with the number of groups actually found (below); start with the scalars + 64k groups
    auto ensure_hbuf = [&](size_t bytes) -> int {
        if (bytes <= g->hbuf_cap) return 0;
        if (g->hbuf) cudaFreeHost(g->hbuf);
        g->hbuf = nullptr;
        g->hbuf_cap = 0;
        void *p = nullptr;
        CU(cudaMallocHost(&p, bytes));
        g->hbuf = static_cast<char *>(p);
        g->hbuf_cap = bytes;
        return 0;
    };
    (void)hneed;
    TRY(ensure_hbuf(al(sizeof(PlanScalars)) + 2 * al((65536 + 1) * 8) + 64));

    char *w = g->ws;
    auto take = [&](size_t bytes) { char *p = w; w += al(bytes); return p; };
    uint64_t *img[2] = {reinterpret_cast<uint64_t *>(take(N * 8)), reinterpret_cast<uint64_t *>(take(N * 8))};
    uint32_t *idx[2] = {reinterpret_cast<uint32_t *>(take(N * 4)), reinterpret_cast<uint32_t *>(take(N * 4))};
    uint32_t *hist = reinterpret_cast<uint32_t *>(take(hist_n * 4));
    uint8_t *flag = reinterpret_cast<uint8_t *>(take(N));
    uint32_t *cnt = reinterpret_cast<uint32_t *>(take((nb + 1) * 4));
    uint32_t *seg = reinterpret_cast<uint32_t *>(take(seg_n * 4));
    int64_t *row_index = reinterpret_cast<int64_t *>(take(N * 8));
    int64_t *offsets = reinterpret_cast<int64_t *>(take((N + 1) * 8));
    int64_t *first_row = reinterpret_cast<int64_t *>(take((N + 1) * 8));
    PlanScalars *sc = reinterpret_cast<PlanScalars *>(take(sizeof(PlanScalars) + 16));
    uint32_t *total = reinterpret_cast<uint32_t *>(sc) + 6;  // inside the 16 spare bytes after the scalars

    PlanKeys pk;
    std::memset(&pk, 0, sizeof(pk));
    pk.n_keys = n_keys;
    std::vector<StageSeg> segs;
    for (int j = 0; j < n_keys; ++j) {
        pk.dtype[j] = keys[j].dtype;
        if (memspace == B200OLS_DEVICE) {
            pk.col[j] = keys[j].values;
        } else {
            char *d = take(static_cast<size_t>(n) * key_bytes(keys[j].dtype) + 16);
            pk.col[j] = d;
            if (n > 0) segs.push_back({keys[j].values, d, static_cast<size_t>(n) * key_bytes(keys[j].dtype)});
        }
    }
    if (!segs.empty()) TRY(stage_h2d(c, segs.data(), static_cast<int>(segs.size())));

    if (n == 0) {
        int64_t *ho = reinterpret_cast<int64_t *>(g->hbuf + al(sizeof(PlanScalars)));
        ho[0] = 0;
        g->n_rows = 0;
        g->n_groups = 0;
        g->offsets_host = ho;
        g->first_row_host = ho + 1;
        g->offsets_dev = nullptr;
        g->row_index_dev = nullptr;
        g->valid = true;
        out->n_groups = 0;
        out->group_offsets = ho;
        out->group_first_row = ho + 1;
        return 0;
    }

    const unsigned grid_stream = static_cast<unsigned>(std::min<int64_t>(nb, static_cast<int64_t>(c->sm_count) * 8));
    PlanScalars *hsc = reinterpret_cast<PlanScalars *>(g->hbuf);
    CU(cudaEventRecord(g->ev0, c->stream));

    // sort keys from the last to the first (LSD over the tuple); each key by its varying 8-bit digits, LSD
    int cur = 0;           // img[cur] / idx[cur] hold the current order
    bool permuted = false;  // idx[cur] is meaningful (false: identity)
    bool sorted_input = false;
    for (int j = n_keys - 1; j >= 0; --j) {
        const PlanScalars init = {0ull, ~0ull, 0u, 0u};
        *hsc = init;
        CU(cudaMemcpyAsync(sc, hsc, sizeof(PlanScalars), cudaMemcpyHostToDevice, c->stream));
        const int check = (j == n_keys - 1) ? 1 : 0;
        plan_prepass_kernel<<<grid_stream, PLAN_THREADS, 0, c->stream>>>(pk, j, permuted ? idx[cur] : nullptr, img[cur], n, check, sc);
        c->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(hsc, sc, sizeof(PlanScalars), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (check && !hsc->unsorted) {
            sorted_input = true;  // GroupsSlice: rows are already grouped in ascending key order
            break;
        }
        const uint64_t varying = hsc->key_or ^ hsc->key_and;
        if (!varying) continue;  // every row has the same value of this key
        const int lo = __builtin_ctzll(varying), hi = 64 - __builtin_clzll(varying);
        for (int shift = lo; shift < hi; shift += 8) {
            if (((varying >> shift) & 0xffull) == 0) continue;
            radix_hist_kernel<<<static_cast<unsigned>(nb), PLAN_THREADS, 0, c->stream>>>(img[cur], n, shift, hist, nb);
            c->launches++;
            TRY(exclusive_scan_u32(c, hist, hist_n, seg, nullptr));
            radix_scatter_kernel<<<static_cast<unsigned>(nb), PLAN_THREADS, 0, c->stream>>>(img[cur], permuted ? idx[cur] : nullptr, img[cur ^ 1],
                                                                                           idx[cur ^ 1], n, shift, hist, nb);
            c->launches++;
            CU(cudaGetLastError());
            cur ^= 1;
            permuted = true;
        }
    }

    // boundaries -> group ids
    const bool use_img = n_keys == 1;  // img[cur] holds the single key's images in the final order
    plan_flags_kernel<<<grid_stream, PLAN_THREADS, 0, c->stream>>>(pk, use_img ? img[cur] : nullptr, permuted ? idx[cur] : nullptr, n, flag);
    plan_count_kernel<<<static_cast<unsigned>(nb), PLAN_THREADS, 0, c->stream>>>(flag, n, cnt);
    c->launches += 2;
    CU(cudaGetLastError());
    TRY(exclusive_scan_u32(c, cnt, nb, seg, total));
    uint32_t *htotal = reinterpret_cast<uint32_t *>(hsc) + 6;
    CU(cudaMemcpyAsync(htotal, total, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    const int64_t G = static_cast<int64_t>(*htotal);
    plan_write_kernel<<<static_cast<unsigned>(nb), PLAN_THREADS, 0, c->stream>>>(flag, permuted ? idx[cur] : nullptr, n, cnt, G, offsets, first_row,
                                                                                permuted ? row_index : nullptr);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaEventRecord(g->ev1, c->stream));
    TRY(ensure_hbuf(al(sizeof(PlanScalars)) + 2 * al((static_cast<size_t>(G) + 1) * 8) + 64));
    int64_t *ho = reinterpret_cast<int64_t *>(g->hbuf + al(sizeof(PlanScalars)));
    int64_t *hf = reinterpret_cast<int64_t *>(g->hbuf + al(sizeof(PlanScalars)) + al((static_cast<size_t>(G) + 1) * 8));
    CU(cudaMemcpyAsync(ho, offsets, sizeof(int64_t) * (G + 1), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(hf, first_row, sizeof(int64_t) * G, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g->ev0, g->ev1);

    g->n_rows = n;
    g->n_groups = G;
    g->offsets_host = ho;
    g->first_row_host = hf;
    g->offsets_dev = offsets;
    g->row_index_dev = permuted ? row_index : nullptr;
    g->valid = true;
    (void)sorted_input;
    out->n_groups = G;
    out->group_offsets = ho;
    out->group_first_row = hf;
    out->row_index = g->row_index_dev;
    out->device_ms = ms;
    return 0;
}

static __global__ void iota_i64_kernel(int64_t *p, int64_t n) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

extern "C" int b200ols_group_plan_row_index(b200ols_ctx *c, int64_t *dst, int32_t memspace) {
    if (!c || !c->gplan || !c->gplan->valid) return fail(B200OLS_ERR_INVALID, "no group plan on this context");
    GroupPlan *g = c->gplan;
    if (g->n_rows == 0) return 0;
    if (!dst) return fail(B200OLS_ERR_INVALID, "dst is NULL");
    CU(cudaSetDevice(c->device));
    const size_t bytes = sizeof(int64_t) * static_cast<size_t>(g->n_rows);
    const int64_t *src = g->row_index_dev;
    if (!src) {  // contiguous groups: identity (materialised in the plan workspace's row_index slot is not kept; build it)
        if (memspace == B200OLS_HOST) {
            for (int64_t i = 0; i < g->n_rows; ++i) dst[i] = i;
            return 0;
        }
        iota_i64_kernel<<<static_cast<unsigned>((g->n_rows + 255) / 256), 256, 0, c->stream>>>(dst, g->n_rows);
        c->launches++;
        CU(cudaGetLastError());
        return 0;
    }
    if (memspace == B200OLS_HOST) {
        const StageSeg s = {dst, const_cast<int64_t *>(src), bytes};
        return stage_d2h(c, &s, 1);
    }
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

extern "C" int b200ols_group_plan_group_of_row(b200ols_ctx *c, int32_t *dst, int32_t memspace) {
    if (!c || !c->gplan || !c->gplan->valid) return fail(B200OLS_ERR_INVALID, "no group plan on this context");
    GroupPlan *g = c->gplan;
    if (g->n_rows == 0) return 0;
    if (!dst) return fail(B200OLS_ERR_INVALID, "dst is NULL");
    CU(cudaSetDevice(c->device));
    const int64_t n = g->n_rows;
    int32_t *d = dst;
    if (memspace == B200OLS_HOST) d = reinterpret_cast<int32_t *>(g->ws);  // the image buffers are free after the build
    const int64_t nb = (n + PLAN_TILE - 1) / PLAN_TILE;
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>(nb, static_cast<int64_t>(c->sm_count) * 8));
    plan_group_of_row_kernel<<<grid, PLAN_THREADS, 0, c->stream>>>(g->offsets_dev, g->n_groups, g->row_index_dev, n, d);
    c->launches++;
    CU(cudaGetLastError());
    if (memspace == B200OLS_HOST) {
        const StageSeg s = {dst, d, sizeof(int32_t) * static_cast<size_t>(n)};
        return stage_d2h(c, &s, 1);
    }
    return 0;
}
