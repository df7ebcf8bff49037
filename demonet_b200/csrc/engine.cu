// Engine: the whole SSDLite forward (SSD.forward eval branch, demonet/models/generalized_ssd.py:271-349)
// as a pre-planned launch sequence over a device arena, replayed from a CUDA graph.
// The layer list (dn_op[]) is produced by the Python host side from the reference's module
// structure (demonet_b200/plan.py); this file only owns memory, tensor maps and launch order.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "dwconv.cuh"
#include "pwconv.cuh"

using namespace dn;

struct GraphKey {
    int B;
    const void* images;
    void *boxes, *scores, *labels, *counts;
    bool operator<(const GraphKey& o) const {
        return std::tie(B, images, boxes, scores, labels, counts) <
               std::tie(o.B, o.images, o.boxes, o.scores, o.labels, o.counts);
    }
};
struct GraphEntry {
    bool warm = false;
    cudaGraphExec_t exec = nullptr;
    unsigned long long last_use = 0;
};
constexpr int DN_MAX_SLOTS = 4;           // engine instances of the pipeline mode (batches in flight)
constexpr size_t DN_MAX_GRAPHS = 16;      // per engine instance; least recently used entries are dropped beyond this

struct dn_engine {
    dn_model_desc desc;
    std::vector<dn_op> ops;
    std::vector<dn_buf> bufs;
    std::vector<float> anchors_host;
    int max_batch = 0;
    int device = 0;
    unsigned char* arena = nullptr;
    size_t arena_bytes = 0;
    std::vector<size_t> buf_off;
    unsigned char* weights = nullptr;
    size_t weight_bytes = 0;
    float* anchors_dev = nullptr;
    void* post_ws = nullptr;
    size_t post_ws_bytes = 0;
    void* se_ws = nullptr;                          // scratch of the squeeze-excitation layers (largest of them)
    size_t se_ws_bytes = 0;
    int n_se = 0;
    std::vector<CUtensorMap> tmap_a, tmap_w, tmap_y;       // per op (PW only)
    std::vector<char> has_tmap_y;
    std::vector<char> pw_pair;                             // per op (PW only): the weight map was built for the GEMM's pair mode
    std::vector<CUtensorMap> tmap_dw;               // per op (DW only; PWDW: the input window map)
    std::vector<DwTiling> dw_tiling;
    std::vector<DwStream> dw_stream;
    std::vector<DwPool> dw_pool;                    // per SE op: pooling done by the depthwise launch before it (parts > 0)
    int n_se_pooled = 0;                            // ... how many of them in the forward enqueued last
    std::vector<char> dw_tma;                       // per op: 0 direct, 1 TMA tiles, 2 row stream, 3 stride-2 row stream
    std::vector<int> dw_tw;                         // per op: output columns per thread of the stride-2 stream
    bool dw_ready = false;
    bool tmaps_ready = false;
    std::map<GraphKey, GraphEntry> graphs;
    unsigned long long graph_tick = 0;
    cudaStream_t capture_stream = nullptr;
    // side lanes (dn_op.lane > 0): head branches forked off the main chain.  lane_stream[l-1] carries lane l;
    // op_done[j] is recorded after op j when an op of another lane reads its output (deps[i] lists those j)
    int n_lanes = 0;                                // highest lane id in use (0 = everything on one stream)
    std::vector<cudaStream_t> lane_stream;
    std::vector<cudaEvent_t> lane_joined;           // recorded at the end of each side lane
    std::vector<cudaEvent_t> op_done;               // per op, nullptr when nobody waits for it
    std::vector<std::vector<int>> deps;             // per op: producers on other lanes
    cudaEvent_t fork_ev = nullptr;
    // pipeline mode (dn_model_desc.pipeline_slots = n in 2..4): twins[0..n-2] are further complete engines (own arena,
    // tensor maps, graphs); consecutive forwards go round the n instances on engine-owned streams, so the latency-bound
    // tail of forward i (tiny layers, NMS rounds) overlaps the bandwidth-bound head of the forwards behind it
    // (measured at B = 256, one box: 72.6k img/s with one in flight, 79.4k with two, 81.3k with three, 82.3k with four)
    std::vector<dn_engine*> twins;
    int n_slots = 1;
    int next_slot = 0;
    int last_slot = 0;                              // slot of the forward issued last (dn_engine_copy_buffer reads it)
    long long n_graph_replays = 0;                  // forwards that were one cudaGraphLaunch
    long long n_forwards = 0;
    cudaStream_t slot_stream[DN_MAX_SLOTS] = {};
    cudaEvent_t slot_in[DN_MAX_SLOTS] = {}, slot_done[DN_MAX_SLOTS] = {};
    bool slot_pending[DN_MAX_SLOTS] = {};
    // staging for dn_engine_forward_host: two input buffers so that the H2D copy of call i+1 (on copy_stream)
    // overlaps the forward of call i (on the caller's stream)
    float* stage_images = nullptr;              // [2][max_batch,3,H,W]
    size_t stage_image_bytes = 0;
    unsigned char* stage_u8 = nullptr;          // [2][max_batch,3,H,W] uint8, allocated on the first uint8 call
    size_t stage_u8_bytes = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t h2d_done[2] = {nullptr, nullptr}, fwd_done[2] = {nullptr, nullptr};
    bool fwd_recorded[2] = {false, false};
    int host_parity = 0;
    unsigned char* stage_out = nullptr;
    size_t so_boxes = 0, so_scores = 0, so_labels = 0, so_counts = 0, so_total = 0;
    size_t device_bytes = 0;
};

static void* buf_ptr(dn_engine* e, int id) { return (id >= 0) ? (void*)(e->arena + e->buf_off[id]) : nullptr; }

static void drop_graphs(dn_engine* e) {
    for (auto& kv : e->graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    e->graphs.clear();
}

extern "C" int dn_engine_create(dn_engine** out, const dn_model_desc* d, int max_batch) {
    DN_REQUIRE(out && d, DN_ERR_INVALID, "NULL argument");
    DN_REQUIRE(max_batch > 0, DN_ERR_INVALID, "max_batch must be positive");
    DN_REQUIRE(d->pipeline_slots >= 0 && d->pipeline_slots <= DN_MAX_SLOTS, DN_ERR_INVALID, "pipeline_slots must be in 0..%d (got %d)",
               DN_MAX_SLOTS, d->pipeline_slots);
    DN_REQUIRE(d->n_ops > 0 && d->n_bufs > 0 && d->ops_host && d->bufs_host && d->anchors_host, DN_ERR_INVALID,
               "model description is incomplete");
    DN_REQUIRE(d->logits_buf >= 0 && d->logits_buf < d->n_bufs && d->bbox_buf >= 0 && d->bbox_buf < d->n_bufs,
               DN_ERR_INVALID, "bad head buffer ids");
    for (int i = 0; i < d->n_ops; ++i) {
        const dn_op& o = d->ops_host[i];
        DN_REQUIRE(o.kind >= DN_OP_STEM && o.kind <= DN_OP_DWPW, DN_ERR_INVALID, "op %d: unknown kind %d", i, o.kind);
        DN_REQUIRE(o.kind != DN_OP_DWPW || (dwpw_fused_supported(o.h_in, o.w_in, o.c_in, o.c_out, o.ksize, o.stride) &&
                                            (o.res_buf == DN_BUF_NONE || o.res_buf == o.in_buf) && o.in_buf != o.out_buf),
                   DN_ERR_UNSUPPORTED, "op %d: shape not supported by the fused depthwise + project kernel", i);
        if (o.kind == DN_OP_NOP) continue;
        DN_REQUIRE(o.kind != DN_OP_PWDW || pwdw_fused_supported(o.h_in, o.w_in, o.c_in, o.c_out, o.ksize, o.stride),
                   DN_ERR_UNSUPPORTED, "op %d: shape not supported by the fused expand + depthwise kernel", i);
        DN_REQUIRE(o.in_buf == DN_BUF_IMAGES || (o.in_buf >= 0 && o.in_buf < d->n_bufs), DN_ERR_INVALID, "op %d: bad in_buf", i);
        DN_REQUIRE(o.kind == DN_OP_SE || (o.out_buf >= 0 && o.out_buf < d->n_bufs), DN_ERR_INVALID, "op %d: bad out_buf", i);
        DN_REQUIRE(o.res_buf == DN_BUF_NONE || (o.res_buf >= 0 && o.res_buf < d->n_bufs), DN_ERR_INVALID, "op %d: bad res_buf", i);
        if (o.se_fold && o.kind == DN_OP_PW) {
            DN_REQUIRE(i > 0 && d->ops_host[i - 1].kind == DN_OP_SE && d->ops_host[i - 1].se_fold && d->ops_host[i - 1].in_buf == o.in_buf &&
                           d->ops_host[i - 1].lane == o.lane && o.act == DN_ACT_NONE && o.c_in % d->ops_host[i - 1].c_in == 0,
                       DN_ERR_INVALID, "op %d: se_fold needs the squeeze-excitation of its input directly in front", i);
        }
        if (o.se_fold && o.kind == DN_OP_SE)
            DN_REQUIRE(i + 1 < d->n_ops && d->ops_host[i + 1].kind == DN_OP_PW && d->ops_host[i + 1].se_fold, DN_ERR_INVALID,
                       "op %d: a folded squeeze-excitation needs its project GEMM directly behind", i);
    }
    dn_engine* e = new dn_engine();
    e->desc = *d;
    e->ops.assign(d->ops_host, d->ops_host + d->n_ops);
    e->bufs.assign(d->bufs_host, d->bufs_host + d->n_bufs);
    e->anchors_host.assign(d->anchors_host, d->anchors_host + (size_t)d->post.num_priors * 4);
    e->desc.ops_host = nullptr;
    e->desc.bufs_host = nullptr;
    e->desc.anchors_host = nullptr;
    e->max_batch = max_batch;
    {
        static const bool lanes_off = [] {          // measurement aid: DN_LANES=0 keeps every op on the caller's stream
            const char* v = getenv("DN_LANES");
            return v && atoi(v) == 0;
        }();
        for (auto& o : e->ops) {
            if (lanes_off || o.lane < 0 || o.lane > 64) o.lane = 0;
            if (o.lane > e->n_lanes) e->n_lanes = o.lane;
        }
        e->deps.resize(e->ops.size());
        std::vector<char> need(e->ops.size(), 0);
        for (size_t i = 0; i < e->ops.size() && e->n_lanes > 0; ++i) {
            const int ins[2] = {e->ops[i].in_buf, e->ops[i].res_buf};
            for (int b : ins) {
                if (b < 0) continue;
                for (int j = (int)i - 1; j >= 0; --j) {
                    if (e->ops[j].out_buf != b && !(e->ops[j].kind == DN_OP_SE && e->ops[j].in_buf == b)) continue;
                    if (e->ops[j].lane != e->ops[i].lane) {
                        e->deps[i].push_back(j);
                        need[j] = 1;
                    }
                    break;                          // the latest writer orders everything before it on its own lane
                }
            }
        }
        e->op_done.assign(e->ops.size(), nullptr);
        for (size_t j = 0; j < e->ops.size(); ++j)
            if (need[j] && cudaEventCreateWithFlags(&e->op_done[j], cudaEventDisableTiming) != cudaSuccess) {
                set_error("cudaEventCreate failed");
                dn_engine_destroy(e);
                return DN_ERR_CUDA;
            }
    }
    auto fail = [&](int rc) {
        dn_engine_destroy(e);
        return rc;
    };
    if (cudaGetDevice(&e->device) != cudaSuccess) {
        set_error("cudaGetDevice failed");
        return fail(DN_ERR_CUDA);
    }
    size_t off = 0;
    e->buf_off.resize(e->bufs.size());
    for (size_t i = 0; i < e->bufs.size(); ++i) {
        e->buf_off[i] = off;
        size_t bytes = (size_t)e->bufs[i].elems_per_image * e->bufs[i].elem_bytes * max_batch;
        off += (bytes + 1023) & ~(size_t)1023;
    }
    e->arena_bytes = off;
    e->post_ws_bytes = dn_postprocess_workspace_bytes(max_batch, &e->desc.post);
    for (const auto& o : e->ops)
        if (o.kind == DN_OP_SE) {
            e->se_ws_bytes = std::max(e->se_ws_bytes, dn_se_workspace_bytes(max_batch, o.h_in * o.w_in, o.c_in));
            ++e->n_se;
        }
    const size_t D = e->desc.post.detections_per_img;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    e->so_boxes = 0;
    e->so_scores = al(e->so_boxes + (size_t)max_batch * D * 16);
    e->so_labels = al(e->so_scores + (size_t)max_batch * D * 4);
    e->so_counts = al(e->so_labels + (size_t)max_batch * D * 8);
    e->so_total = al(e->so_counts + (size_t)max_batch * 4);
    const size_t img_bytes = (size_t)max_batch * 3 * e->desc.image_h * e->desc.image_w * sizeof(float);
#define TRY(x)                                                                  \
    if ((x) != cudaSuccess) {                                                   \
        set_error("%s failed: %s", #x, cudaGetErrorString(cudaGetLastError())); \
        return fail(DN_ERR_CUDA);                                               \
    }
    TRY(cudaMalloc(&e->arena, e->arena_bytes));
    TRY(cudaMemset(e->arena, 0, e->arena_bytes));
    TRY(cudaMalloc(&e->anchors_dev, e->anchors_host.size() * sizeof(float)));
    TRY(cudaMemcpy(e->anchors_dev, e->anchors_host.data(), e->anchors_host.size() * sizeof(float), cudaMemcpyHostToDevice));
    TRY(cudaMalloc(&e->post_ws, e->post_ws_bytes));
    if (e->se_ws_bytes) TRY(cudaMalloc(&e->se_ws, e->se_ws_bytes));
    e->stage_image_bytes = (img_bytes + 255) & ~(size_t)255;
    TRY(cudaMalloc(&e->stage_images, 2 * e->stage_image_bytes));
    TRY(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        TRY(cudaEventCreateWithFlags(&e->h2d_done[i], cudaEventDisableTiming));
        TRY(cudaEventCreateWithFlags(&e->fwd_done[i], cudaEventDisableTiming));
    }
    TRY(cudaMalloc(&e->stage_out, e->so_total));
    TRY(cudaStreamCreateWithFlags(&e->capture_stream, cudaStreamNonBlocking));
    e->lane_stream.assign(e->n_lanes, nullptr);
    e->lane_joined.assign(e->n_lanes, nullptr);
    for (int l = 0; l < e->n_lanes; ++l) {
        TRY(cudaStreamCreateWithFlags(&e->lane_stream[l], cudaStreamNonBlocking));
        TRY(cudaEventCreateWithFlags(&e->lane_joined[l], cudaEventDisableTiming));
    }
    TRY(cudaEventCreateWithFlags(&e->fork_ev, cudaEventDisableTiming));
#undef TRY
#define TRY_AFTER(x)                                                            \
    if ((x) != cudaSuccess) {                                                   \
        set_error("%s failed: %s", #x, cudaGetErrorString(cudaGetLastError())); \
        return fail(DN_ERR_CUDA);                                               \
    }
    if (d->pipeline_slots >= 2) {
        dn_model_desc d1 = *d;
        d1.pipeline_slots = 0;
        e->n_slots = d->pipeline_slots;
        for (int i = 1; i < e->n_slots; ++i) {
            dn_engine* t = nullptr;
            int rc = dn_engine_create(&t, &d1, max_batch);
            if (rc) return fail(rc);
            e->twins.push_back(t);
        }
        for (int i = 0; i < e->n_slots; ++i) {
            TRY_AFTER(cudaStreamCreateWithFlags(&e->slot_stream[i], cudaStreamNonBlocking));
            TRY_AFTER(cudaEventCreateWithFlags(&e->slot_in[i], cudaEventDisableTiming));
            TRY_AFTER(cudaEventCreateWithFlags(&e->slot_done[i], cudaEventDisableTiming));
        }
    }
    e->device_bytes = e->arena_bytes + e->post_ws_bytes + e->se_ws_bytes + 2 * e->stage_image_bytes + e->so_total + e->anchors_host.size() * 4;
    *out = e;
    return DN_OK;
}

extern "C" int dn_engine_destroy(dn_engine* e) {
    if (!e) return DN_OK;
    if (!e->twins.empty()) cudaDeviceSynchronize();
    for (dn_engine* t : e->twins) dn_engine_destroy(t);
    for (int i = 0; i < DN_MAX_SLOTS; ++i) {
        if (e->slot_stream[i]) cudaStreamDestroy(e->slot_stream[i]);
        if (e->slot_in[i]) cudaEventDestroy(e->slot_in[i]);
        if (e->slot_done[i]) cudaEventDestroy(e->slot_done[i]);
    }
    drop_graphs(e);
    if (e->capture_stream) cudaStreamDestroy(e->capture_stream);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    for (auto st : e->lane_stream)
        if (st) cudaStreamDestroy(st);
    for (auto ev : e->lane_joined)
        if (ev) cudaEventDestroy(ev);
    for (auto ev : e->op_done)
        if (ev) cudaEventDestroy(ev);
    if (e->fork_ev) cudaEventDestroy(e->fork_ev);
    for (int i = 0; i < 2; ++i) {
        if (e->h2d_done[i]) cudaEventDestroy(e->h2d_done[i]);
        if (e->fwd_done[i]) cudaEventDestroy(e->fwd_done[i]);
    }
    cudaFree(e->arena);
    cudaFree(e->weights);
    cudaFree(e->anchors_dev);
    cudaFree(e->post_ws);
    cudaFree(e->se_ws);
    cudaFree(e->stage_images);
    cudaFree(e->stage_u8);
    cudaFree(e->stage_out);
    delete e;
    return DN_OK;
}

extern "C" int dn_engine_load_weights(dn_engine* e, const void* blob_host, size_t bytes) {
    DN_REQUIRE(e && blob_host && bytes > 0, DN_ERR_INVALID, "NULL argument");
    for (dn_engine* t : e->twins) {
        int rc = dn_engine_load_weights(t, blob_host, bytes);
        if (rc) return rc;
    }
    for (size_t i = 0; i < e->ops.size(); ++i) {
        const dn_op& o = e->ops[i];
        DN_REQUIRE(o.w_off >= 0 && (size_t)o.w_off < bytes && o.b_off >= 0 && (size_t)o.b_off < bytes, DN_ERR_INVALID,
                   "op %zu: weight offsets outside the blob", i);
        DN_REQUIRE(o.w_off % 16 == 0 && o.b_off % 16 == 0, DN_ERR_INVALID, "op %zu: weight offsets must be 16-byte aligned", i);
    }
    DN_CHECK_CUDA(cudaDeviceSynchronize());
    drop_graphs(e);
    if (bytes != e->weight_bytes) {
        cudaFree(e->weights);
        e->weights = nullptr;
        DN_CHECK_CUDA(cudaMalloc(&e->weights, bytes));
        e->device_bytes += bytes - e->weight_bytes;
        e->weight_bytes = bytes;
        e->tmaps_ready = false;
    }
    DN_CHECK_CUDA(cudaMemcpy(e->weights, blob_host, bytes, cudaMemcpyHostToDevice));
    if (!e->dw_ready) {
        e->tmap_dw.resize(e->ops.size());
        e->dw_tiling.resize(e->ops.size());
        e->dw_stream.resize(e->ops.size());
        e->dw_tma.assign(e->ops.size(), 0);
        e->dw_tw.assign(e->ops.size(), 0);
        for (size_t i = 0; i < e->ops.size(); ++i) {
            const dn_op& o = e->ops[i];
            if (o.kind != DN_OP_DW) continue;
            const DwImpl impl = dw_choose(o.h_in, o.w_in, o.c_in, o.ksize, o.stride);
            if (impl == DW_STREAM) {
                e->dw_tma[i] = 2;
                if (!dw_stream_plan(o.h_in, o.w_in, o.c_in, o.ksize, o.stride, &e->dw_stream[i])) return DN_ERR_UNSUPPORTED;
                int rc = dw_stream_make_tmap(&e->tmap_dw[i], buf_ptr(e, o.in_buf), e->max_batch, o.h_in, o.w_in, o.c_in, o.ksize,
                                             e->dw_stream[i]);
                if (rc) return rc;
                continue;
            }
            if (impl == DW_STREAM2) {
                e->dw_tma[i] = 3;
                if (!dw_stream2_plan(o.h_in, o.w_in, o.c_in, o.ksize, &e->dw_stream[i], &e->dw_tw[i])) return DN_ERR_UNSUPPORTED;
                int rc = dw_stream2_make_tmap(&e->tmap_dw[i], buf_ptr(e, o.in_buf), e->max_batch, o.h_in, o.w_in, o.c_in, o.ksize,
                                              e->dw_stream[i]);
                if (rc) return rc;
                continue;
            }
            if (impl != DW_TMA) continue;
            e->dw_tma[i] = 1;
            int rc = dw_plan(o.h_in, o.w_in, o.c_in, o.ksize, o.stride, &e->dw_tiling[i]);
            if (rc) return rc;
            rc = dw_make_tmap(&e->tmap_dw[i], buf_ptr(e, o.in_buf), e->max_batch, o.h_in, o.w_in, o.c_in, e->dw_tiling[i]);
            if (rc) return rc;
        }
        e->dw_ready = true;
    }
    if (!e->tmaps_ready && e->desc.gemm_impl == 0) {
        e->tmap_a.resize(e->ops.size());
        e->tmap_w.resize(e->ops.size());
        e->tmap_y.resize(e->ops.size());
        e->has_tmap_y.assign(e->ops.size(), 0);
        e->pw_pair.assign(e->ops.size(), 0);
        for (size_t i = 0; i < e->ops.size(); ++i) {
            const dn_op& o = e->ops[i];
            if (o.kind == DN_OP_DWPW) {
                int rc = dwpw_fused_make_tmaps(&e->tmap_a[i], &e->tmap_w[i], buf_ptr(e, o.in_buf), e->weights + o.w2_off,
                                               e->max_batch, o.h_in, o.w_in);
                if (rc) return rc;
                continue;
            }
            if (o.kind == DN_OP_PWDW) {
                int rc = pwdw_fused_make_tmaps(&e->tmap_a[i], &e->tmap_w[i], buf_ptr(e, o.in_buf), e->weights + o.w_off,
                                               e->max_batch, o.h_in, o.w_in);
                if (rc) return rc;
                continue;
            }
            if (o.kind != DN_OP_PW) continue;
            int bn, nt, st, cols, pair;
            size_t smem;
            const long long m_max = (long long)e->max_batch * o.h_in * o.w_in;
            pwconv_tc_plan(m_max, o.c_in, o.c_out, &bn, &nt, &st, &cols, &smem, nullptr, &pair);
            int rc = make_tmap_h16_2d(&e->tmap_a[i], buf_ptr(e, o.in_buf), m_max, o.c_in, 128, 64);
            if (rc) return rc;
            rc = make_tmap_h16_2d(&e->tmap_w[i], e->weights + o.w_off, o.c_out, o.c_in, pair ? bn / 2 : bn, 64);
            e->pw_pair[i] = (char)pair;
            if (rc) return rc;
            // dense bf16 outputs without a residual are written by TMA (box 32 rows x 32 columns)
            if (!o.out_fp32 && o.res_buf == DN_BUF_NONE && o.out_batch_stride == 0 && o.out_row_stride == 0 &&
                o.out_offset == 0 && o.c_out % 8 == 0) {
                rc = make_tmap_h16_2d(&e->tmap_y[i], buf_ptr(e, o.out_buf), m_max, o.c_out, 32, 32);
                if (rc) return rc;
                e->has_tmap_y[i] = 1;
            }
        }
        e->tmaps_ready = true;
    }
    return DN_OK;
}

// DN_SE_POOL=0: squeeze-excitation layers pool with their own pass instead of taking the sums from the depthwise
// row stream in front of them; 1: only the stride-1 streams pool (measurement aids)
static int se_pool_fusion() {
    static const int mode = [] {
        const char* v = getenv("DN_SE_POOL");
        return v ? atoi(v) : 2;
    }();
    return mode;
}

static int enqueue_op(dn_engine* e, size_t i, const float* images, int B, cudaStream_t s) {
    {
        const dn_op& o = e->ops[i];
        const unsigned char* W = e->weights;
        int rc = DN_OK;
        switch (o.kind) {
            case DN_OP_STEM:
                rc = dn_stem_conv(images, (const float*)(W + o.w_off), (const float*)(W + o.b_off), e->desc.image_mean,
                                  e->desc.image_std, buf_ptr(e, o.out_buf), B, o.h_in, o.w_in, o.c_out, o.act, s);
                break;
            case DN_OP_DW:
                if (e->dw_ready && (e->dw_tma[i] == 2 || e->dw_tma[i] == 3)) {
                    // a squeeze-excitation right behind: let the row stream leave its channel sums in the SE workspace
                    DwPool pool{nullptr, se_max_pool_slots(), 0, 0};
                    const bool se_next = i + 1 < e->ops.size() && e->ops[i + 1].kind == DN_OP_SE &&
                                         e->ops[i + 1].in_buf == o.out_buf && e->ops[i + 1].lane == o.lane &&
                                         se_pool_fusion() >= (e->dw_tma[i] == 2 ? 1 : 2);
                    if (se_next) pool.partial = (float*)e->se_ws;
                    if (e->dw_tma[i] == 2)
                        rc = dwconv_stream_launch(e->tmap_dw[i], e->dw_stream[i], (const float*)(W + o.w_off),
                                                  (const float*)(W + o.b_off), buf_ptr(e, o.out_buf), B, o.h_in, o.w_in, o.c_in,
                                                  o.ksize, o.act, s, se_next ? &pool : nullptr);
                    else
                        rc = dwconv_stream2_launch(e->tmap_dw[i], e->dw_stream[i], e->dw_tw[i], (const float*)(W + o.w_off),
                                                   (const float*)(W + o.b_off), buf_ptr(e, o.out_buf), B, o.h_in, o.w_in, o.c_in,
                                                   o.ksize, o.act, s, se_next ? &pool : nullptr);
                    if (se_next) e->dw_pool[i + 1] = pool;
                    break;
                }
                if (e->dw_ready && e->dw_tma[i] == 1) {
                    rc = dwconv_tma_launch(e->tmap_dw[i], e->dw_tiling[i], (const float*)(W + o.w_off), (const float*)(W + o.b_off),
                                           buf_ptr(e, o.out_buf), B, o.h_in, o.w_in, o.c_in, o.ksize, o.stride, o.act, s);
                    break;
                }
                rc = dn_dwconv(buf_ptr(e, o.in_buf), (const float*)(W + o.w_off), (const float*)(W + o.b_off),
                               buf_ptr(e, o.out_buf), B, o.h_in, o.w_in, o.c_in, o.ksize, o.stride, o.act, s);
                break;
            case DN_OP_PW: {
                PwEpilogue ep;
                const int hw = o.h_in * o.w_in;
                ep.bias = (const float*)(W + o.b_off);
                ep.residual = (const dn_half_t*)buf_ptr(e, o.res_buf);
                ep.y = (unsigned char*)buf_ptr(e, o.out_buf) + (size_t)o.out_offset * (o.out_fp32 ? 4 : 2);
                ep.N = o.c_out;
                ep.act = o.act;
                ep.out_fp32 = o.out_fp32;
                ep.hw = hw;
                ep.out_batch_stride = o.out_batch_stride ? o.out_batch_stride : (long long)hw * o.c_out;
                ep.out_row_stride = o.out_row_stride ? o.out_row_stride : o.c_out;
                if (o.se_fold && e->desc.gemm_impl == 0) {
                    // the squeeze-excitation in front left its [B][C] scales in the SE workspace: this GEMM applies them to
                    // its A operand in shared memory (C = the SE layer's channel count; K = p * C for pixel-packed layers)
                    const dn_op& se = e->ops[i - 1];
                    ep.a_scale = se_scales_ptr(e->se_ws, B, se.c_in);
                    ep.a_scale_c = se.c_in;
                }
                const int M = B * hw;
                if (e->desc.gemm_impl == 0)
                    rc = pwconv_tc_launch(e->tmap_a[i], e->tmap_w[i], e->has_tmap_y[i] ? &e->tmap_y[i] : nullptr, ep, M,
                                          (long long)e->max_batch * hw, o.c_in, o.c_out, s, e->pw_pair[i]);
                else
                    rc = pwconv_simt(buf_ptr(e, o.in_buf), W + o.w_off, ep, M, o.c_in, o.c_out, s);
                break;
            }
            case DN_OP_PWDW:
                rc = pwdw_fused_launch(e->tmap_a[i], e->tmap_w[i], (const float*)(W + o.b_off), (const float*)(W + o.w2_off),
                                       (const float*)(W + o.b2_off), buf_ptr(e, o.out_buf), B, o.h_in, o.w_in, o.act, o.act2, s);
                break;
            case DN_OP_DWPW:
                rc = dwpw_fused_launch(e->tmap_a[i], e->tmap_w[i], (const float*)(W + o.w_off), (const float*)(W + o.b_off),
                                       (const float*)(W + o.b2_off), buf_ptr(e, o.out_buf), B, o.h_in, o.w_in, o.act,
                                       o.res_buf != DN_BUF_NONE, s);
                break;
            case DN_OP_NOP:
                break;
            case DN_OP_SE: {
                const DwPool& pool = e->dw_pool[i];      // parts > 0: the depthwise launch just before pooled for us
                rc = se_inplace_pooled(buf_ptr(e, o.in_buf), (const float*)(W + o.w_off), (const float*)(W + o.b_off),
                                       (const float*)(W + o.w2_off), (const float*)(W + o.b2_off), B, o.h_in * o.w_in, o.c_in,
                                       o.c_mid, e->se_ws, e->se_ws_bytes, pool.parts, pool.slots, o.h_in, s,
                                       !(o.se_fold && e->desc.gemm_impl == 0));
                e->n_se_pooled += pool.parts > 0;
                break;
            }
        }
        return rc;
    }
}

static int enqueue_forward(dn_engine* e, const float* images, int B, float* out_boxes, float* out_scores,
                           int64_t* out_labels, int32_t* out_counts, cudaStream_t s) {
    // Lane 0 runs on the caller's stream; a side lane starts behind everything already enqueued there (fork_ev),
    // waits for the producers it reads from (op_done) and is joined again before the post-processing.  Under
    // stream capture the same event calls turn into parallel branches of the graph.
    std::vector<char> lane_used(e->n_lanes, 0);
    e->n_se_pooled = 0;
    e->dw_pool.assign(e->ops.size(), DwPool{nullptr, 0, 0, 0});
    if (e->n_lanes > 0) DN_CHECK_CUDA(cudaEventRecord(e->fork_ev, s));
    for (size_t i = 0; i < e->ops.size(); ++i) {
        const int lane = e->ops[i].lane;
        cudaStream_t ls = s;
        if (lane > 0) {
            ls = e->lane_stream[lane - 1];
            if (!lane_used[lane - 1]) {
                lane_used[lane - 1] = 1;
                DN_CHECK_CUDA(cudaStreamWaitEvent(ls, e->fork_ev, 0));
            }
        }
        for (int j : e->deps[i]) DN_CHECK_CUDA(cudaStreamWaitEvent(ls, e->op_done[j], 0));
        int rc = enqueue_op(e, i, images, B, ls);
        if (rc) return rc;
        if (e->op_done[i]) DN_CHECK_CUDA(cudaEventRecord(e->op_done[i], ls));
    }
    for (int l = 0; l < e->n_lanes; ++l) {
        if (!lane_used[l]) continue;
        DN_CHECK_CUDA(cudaEventRecord(e->lane_joined[l], e->lane_stream[l]));
        DN_CHECK_CUDA(cudaStreamWaitEvent(s, e->lane_joined[l], 0));
    }
    return dn_postprocess((const float*)buf_ptr(e, e->desc.logits_buf), (const float*)buf_ptr(e, e->desc.bbox_buf),
                          e->anchors_dev, B, &e->desc.post, e->post_ws, e->post_ws_bytes, out_boxes, out_scores, out_labels,
                          out_counts, s);
}

static int forward_one(dn_engine* e, const float* images_dev, int B, float* out_boxes, float* out_scores,
                       int64_t* out_labels, int32_t* out_counts, void* stream_);

extern "C" int dn_engine_forward(dn_engine* e, const float* images_dev, int B, float* out_boxes, float* out_scores,
                                 int64_t* out_labels, int32_t* out_counts, void* stream_) {
    DN_REQUIRE(e != nullptr, DN_ERR_INVALID, "engine is NULL");
    if (e->twins.empty()) return forward_one(e, images_dev, B, out_boxes, out_scores, out_labels, out_counts, stream_);
    // pipeline mode: the forward runs on the slot's own stream behind everything already enqueued on the caller's
    // stream; the caller's stream is NOT made to wait for it here (that would serialise the next call behind this
    // one) -- dn_engine_join / dn_engine_join_previous order the results on a stream
    cudaStream_t s = (cudaStream_t)stream_;
    const int k = e->next_slot;
    e->next_slot = (k + 1) % e->n_slots;
    e->last_slot = k;
    DN_CHECK_CUDA(cudaEventRecord(e->slot_in[k], s));
    DN_CHECK_CUDA(cudaStreamWaitEvent(e->slot_stream[k], e->slot_in[k], 0));
    int rc = forward_one(k ? e->twins[k - 1] : e, images_dev, B, out_boxes, out_scores, out_labels, out_counts, e->slot_stream[k]);
    if (rc) return rc;
    DN_CHECK_CUDA(cudaEventRecord(e->slot_done[k], e->slot_stream[k]));
    e->slot_pending[k] = true;
    return DN_OK;
}

extern "C" int dn_engine_join(dn_engine* e, void* stream_) {
    DN_REQUIRE(e != nullptr, DN_ERR_INVALID, "engine is NULL");
    for (int k = 0; k < e->n_slots; ++k)
        if (e->slot_pending[k]) {
            DN_CHECK_CUDA(cudaStreamWaitEvent((cudaStream_t)stream_, e->slot_done[k], 0));
            e->slot_pending[k] = false;
        }
    return DN_OK;
}

extern "C" int dn_engine_join_previous(dn_engine* e, void* stream_) {
    DN_REQUIRE(e != nullptr, DN_ERR_INVALID, "engine is NULL");
    if (e->twins.empty()) return DN_OK;
    const int k = e->next_slot;          // the slot the NEXT call will use = the oldest forward in flight (n - 1 calls back)
    if (e->slot_pending[k]) {
        DN_CHECK_CUDA(cudaStreamWaitEvent((cudaStream_t)stream_, e->slot_done[k], 0));
        e->slot_pending[k] = false;
    }
    return DN_OK;
}

static int forward_one(dn_engine* e, const float* images_dev, int B, float* out_boxes, float* out_scores,
                       int64_t* out_labels, int32_t* out_counts, void* stream_) {
    DN_REQUIRE(e != nullptr, DN_ERR_INVALID, "engine is NULL");
    DN_REQUIRE(e->weights != nullptr, DN_ERR_INVALID, "weights have not been loaded");
    DN_REQUIRE(B > 0 && B <= e->max_batch, DN_ERR_INVALID, "batch %d outside [1, %d]", B, e->max_batch);
    DN_REQUIRE(images_dev && out_boxes && out_scores && out_labels && out_counts, DN_ERR_INVALID, "NULL tensor pointer");
    cudaStream_t s = (cudaStream_t)stream_;
    ++e->n_forwards;
    if (!e->desc.use_cuda_graph) return enqueue_forward(e, images_dev, B, out_boxes, out_scores, out_labels, out_counts, s);
    GraphKey key{B, images_dev, out_boxes, out_scores, out_labels, out_counts};
    if (e->graphs.find(key) == e->graphs.end() && e->graphs.size() >= DN_MAX_GRAPHS) {
        // a caller that keeps handing in new addresses (a data loader's fresh batches) must not grow the cache without
        // bound: drop the least recently used entry (an executing graph is released by the driver once it has finished)
        auto lru = e->graphs.begin();
        for (auto it = e->graphs.begin(); it != e->graphs.end(); ++it)
            if (it->second.last_use < lru->second.last_use) lru = it;
        if (lru->second.exec) cudaGraphExecDestroy(lru->second.exec);
        e->graphs.erase(lru);
    }
    GraphEntry& g = e->graphs[key];
    g.last_use = ++e->graph_tick;
    if (!g.warm) {      // first call with these buffers runs eagerly (also configures kernel attributes)
        g.warm = true;
        return enqueue_forward(e, images_dev, B, out_boxes, out_scores, out_labels, out_counts, s);
    }
    if (!g.exec) {
        cudaGraph_t graph = nullptr;
        DN_CHECK_CUDA(cudaStreamBeginCapture(e->capture_stream, cudaStreamCaptureModeThreadLocal));
        int rc = enqueue_forward(e, images_dev, B, out_boxes, out_scores, out_labels, out_counts, e->capture_stream);
        cudaError_t ce = cudaStreamEndCapture(e->capture_stream, &graph);
        if (rc) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        DN_REQUIRE(ce == cudaSuccess && graph, DN_ERR_CUDA, "stream capture failed: %s", cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        DN_REQUIRE(ce == cudaSuccess, DN_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
    }
    DN_CHECK_CUDA(cudaGraphLaunch(g.exec, s));
    ++e->n_graph_replays;
    return DN_OK;
}

// Host-buffer forward, shared by the fp32 and the uint8 entry points.
static int forward_host_impl(dn_engine* e, const void* images_host, bool u8, int B, float* out_boxes_host,
                             float* out_scores_host, int64_t* out_labels_host, int32_t* out_counts_host, void* stream_) {
    DN_REQUIRE(e != nullptr, DN_ERR_INVALID, "engine is NULL");
    DN_REQUIRE(B > 0 && B <= e->max_batch, DN_ERR_INVALID, "batch %d outside [1, %d]", B, e->max_batch);
    DN_REQUIRE(images_host && out_boxes_host && out_scores_host && out_labels_host && out_counts_host, DN_ERR_INVALID,
               "NULL host pointer");
    cudaStream_t s = (cudaStream_t)stream_;
    const size_t D = e->desc.post.detections_per_img;
    const size_t n_px = (size_t)B * 3 * e->desc.image_h * e->desc.image_w;
    const size_t img_bytes = n_px * sizeof(float);
    if (u8 && !e->stage_u8) {
        e->stage_u8_bytes = ((size_t)e->max_batch * 3 * e->desc.image_h * e->desc.image_w + 255) & ~(size_t)255;
        DN_CHECK_CUDA(cudaMalloc(&e->stage_u8, 2 * e->stage_u8_bytes));
        e->device_bytes += 2 * e->stage_u8_bytes;
    }
    // H2D on the copy stream into the staging buffer the previous-but-one call used; the forward on the
    // caller's stream waits for it, so back-to-back calls overlap copy(i+1) with forward(i).
    const int par = e->host_parity;
    e->host_parity ^= 1;
    float* stage = (float*)((unsigned char*)e->stage_images + (size_t)par * e->stage_image_bytes);
    if (e->fwd_recorded[par]) DN_CHECK_CUDA(cudaStreamWaitEvent(e->copy_stream, e->fwd_done[par], 0));   // WAR on stage
    if (u8) {       // a quarter of the PCIe bytes; ToTensor's x / 255 runs on the device in front of the stem
        unsigned char* su8 = e->stage_u8 + (size_t)par * e->stage_u8_bytes;
        DN_CHECK_CUDA(cudaMemcpyAsync(su8, images_host, n_px, cudaMemcpyHostToDevice, e->copy_stream));
        DN_CHECK_CUDA(cudaEventRecord(e->h2d_done[par], e->copy_stream));
        DN_CHECK_CUDA(cudaStreamWaitEvent(s, e->h2d_done[par], 0));
        int rc = u8_to_f32_launch(su8, stage, n_px, s);
        if (rc) return rc;
    } else {
        DN_CHECK_CUDA(cudaMemcpyAsync(stage, images_host, img_bytes, cudaMemcpyHostToDevice, e->copy_stream));
        DN_CHECK_CUDA(cudaEventRecord(e->h2d_done[par], e->copy_stream));
        DN_CHECK_CUDA(cudaStreamWaitEvent(s, e->h2d_done[par], 0));
    }
    unsigned char* so = e->stage_out;
    int rc = forward_one(e, stage, B, (float*)(so + e->so_boxes), (float*)(so + e->so_scores),
                         (int64_t*)(so + e->so_labels), (int32_t*)(so + e->so_counts), s);
    if (rc) return rc;
    DN_CHECK_CUDA(cudaEventRecord(e->fwd_done[par], s));
    e->fwd_recorded[par] = true;
    DN_CHECK_CUDA(cudaMemcpyAsync(out_boxes_host, so + e->so_boxes, (size_t)B * D * 16, cudaMemcpyDeviceToHost, s));
    DN_CHECK_CUDA(cudaMemcpyAsync(out_scores_host, so + e->so_scores, (size_t)B * D * 4, cudaMemcpyDeviceToHost, s));
    DN_CHECK_CUDA(cudaMemcpyAsync(out_labels_host, so + e->so_labels, (size_t)B * D * 8, cudaMemcpyDeviceToHost, s));
    DN_CHECK_CUDA(cudaMemcpyAsync(out_counts_host, so + e->so_counts, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    return DN_OK;
}

// Pipeline mode keeps the plain contract of the host entry points (results valid once `stream` is synchronised): the
// upload, forward and download of a call run on the slot's streams, which do not depend on the caller's stream, and
// the caller's stream only collects one wait per call.
static int forward_host_slots(dn_engine* e, const void* images_host, bool u8, int B, float* ob, float* os, int64_t* ol,
                              int32_t* oc, void* stream_) {
    DN_REQUIRE(e != nullptr, DN_ERR_INVALID, "engine is NULL");
    if (e->twins.empty()) return forward_host_impl(e, images_host, u8, B, ob, os, ol, oc, stream_);
    const int k = e->next_slot;
    e->next_slot = (k + 1) % e->n_slots;
    e->last_slot = k;
    int rc = forward_host_impl(k ? e->twins[k - 1] : e, images_host, u8, B, ob, os, ol, oc, e->slot_stream[k]);
    if (rc) return rc;
    DN_CHECK_CUDA(cudaEventRecord(e->slot_done[k], e->slot_stream[k]));
    DN_CHECK_CUDA(cudaStreamWaitEvent((cudaStream_t)stream_, e->slot_done[k], 0));
    return DN_OK;
}

extern "C" int dn_engine_forward_host(dn_engine* e, const float* images_host, int B, float* out_boxes_host,
                                      float* out_scores_host, int64_t* out_labels_host, int32_t* out_counts_host,
                                      void* stream_) {
    return forward_host_slots(e, images_host, false, B, out_boxes_host, out_scores_host, out_labels_host, out_counts_host, stream_);
}

extern "C" int dn_engine_forward_host_u8(dn_engine* e, const uint8_t* images_host, int B, float* out_boxes_host,
                                         float* out_scores_host, int64_t* out_labels_host, int32_t* out_counts_host,
                                         void* stream_) {
    return forward_host_slots(e, images_host, true, B, out_boxes_host, out_scores_host, out_labels_host, out_counts_host, stream_);
}

extern "C" int dn_engine_buffer(dn_engine* e, int buf_id, void** ptr_out, int64_t* elems_per_image_out) {
    DN_REQUIRE(e && ptr_out, DN_ERR_INVALID, "NULL argument");
    DN_REQUIRE(buf_id >= 0 && buf_id < (int)e->bufs.size(), DN_ERR_INVALID, "bad buffer id %d", buf_id);
    *ptr_out = buf_ptr(e, buf_id);
    if (elems_per_image_out) *elems_per_image_out = e->bufs[buf_id].elems_per_image;
    return DN_OK;
}

extern "C" int dn_engine_copy_buffer(dn_engine* e, int buf_id, void* dst_dev, size_t bytes, void* stream_) {
    DN_REQUIRE(e && dst_dev, DN_ERR_INVALID, "NULL argument");
    DN_REQUIRE(buf_id >= 0 && buf_id < (int)e->bufs.size(), DN_ERR_INVALID, "bad buffer id %d", buf_id);
    const size_t cap = (size_t)e->bufs[buf_id].elems_per_image * e->bufs[buf_id].elem_bytes * e->max_batch;
    DN_REQUIRE(bytes <= cap, DN_ERR_INVALID, "copy of %zu bytes exceeds the buffer (%zu)", bytes, cap);
    // pipeline mode: the arena of the slot that ran the forward issued last, ordered behind that forward
    dn_engine* src = e;
    if (!e->twins.empty()) {
        if (e->last_slot > 0) src = e->twins[e->last_slot - 1];
        if (e->slot_pending[e->last_slot]) DN_CHECK_CUDA(cudaStreamWaitEvent((cudaStream_t)stream_, e->slot_done[e->last_slot], 0));
    }
    DN_CHECK_CUDA(cudaMemcpyAsync(dst_dev, buf_ptr(src, buf_id), bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream_));
    return DN_OK;
}

// Per-launch device time measured INSIDE a CUDA graph: the plan is captured on one stream (no side lanes, so that the
// brackets are meaningful) with an event-record node between consecutive ops, the graph is replayed `iters` times and
// the elapsed times of consecutive events are averaged.  Every op therefore runs once per replay, in plan order, with the
// cache state the real step gives it (its input was written by the op in front of it, not by its own previous launch).
extern "C" int dn_engine_profile(dn_engine* e, const float* images_dev, int B, int iters, float* ms_out_host, void* stream_) {
    DN_REQUIRE(e && images_dev && ms_out_host, DN_ERR_INVALID, "NULL argument");
    DN_REQUIRE(e->weights != nullptr, DN_ERR_INVALID, "weights have not been loaded");
    DN_REQUIRE(B > 0 && B <= e->max_batch && iters > 0, DN_ERR_INVALID, "bad batch / iters");
    cudaStream_t s = (cudaStream_t)stream_;
    unsigned char* so = e->stage_out;
    float* ob = (float*)(so + e->so_boxes);
    float* os = (float*)(so + e->so_scores);
    int64_t* ol = (int64_t*)(so + e->so_labels);
    int32_t* oc = (int32_t*)(so + e->so_counts);
    int rc = enqueue_forward(e, images_dev, B, ob, os, ol, oc, s);          // valid inputs for every layer, attributes set
    if (rc) return rc;
    DN_CHECK_CUDA(cudaStreamSynchronize(s));
    const size_t n = e->ops.size();
    std::vector<cudaEvent_t> ev(n + 5, nullptr);       // n + 3 op boundaries, then an EMPTY bracket (ev[n+3], ev[n+4])
    auto cleanup = [&]() {
        for (auto x : ev)
            if (x) cudaEventDestroy(x);
    };
    for (auto& x : ev)
        if (cudaEventCreate(&x) != cudaSuccess) {
            cleanup();
            set_error("cudaEventCreate failed");
            return DN_ERR_CUDA;
        }
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaStream_t cs = e->capture_stream;
    cudaError_t ce = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    if (ce == cudaSuccess) {
        e->dw_pool.assign(n, DwPool{nullptr, 0, 0, 0});
        for (size_t i = 0; i < n && !rc; ++i) {
            cudaEventRecordWithFlags(ev[i], cs, cudaEventRecordExternal);
            rc = enqueue_op(e, i, images_dev, B, cs);
        }
        cudaEventRecordWithFlags(ev[n], cs, cudaEventRecordExternal);
        if (!rc)
            rc = dn_postprocess_marked((const float*)buf_ptr(e, e->desc.logits_buf), (const float*)buf_ptr(e, e->desc.bbox_buf),
                                       e->anchors_dev, B, &e->desc.post, e->post_ws, e->post_ws_bytes, ob, os, ol, oc, cs, ev[n + 1]);
        cudaEventRecordWithFlags(ev[n + 2], cs, cudaEventRecordExternal);
        // two event nodes with nothing in between: what one bracket costs by itself (subtracted from every op below)
        cudaEventRecordWithFlags(ev[n + 3], cs, cudaEventRecordExternal);
        cudaEventRecordWithFlags(ev[n + 4], cs, cudaEventRecordExternal);
        ce = cudaStreamEndCapture(cs, &graph);
    }
    if (rc || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cleanup();
        if (!rc) set_error("profile capture failed: %s", cudaGetErrorString(ce));
        return rc ? rc : DN_ERR_CUDA;
    }
    ce = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) {
        cleanup();
        set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
        return DN_ERR_CUDA;
    }
    std::vector<double> acc(n + 2, 0.0);
    double empty = 0.0;
    for (int it = -1; it < iters && ce == cudaSuccess; ++it) {              // replay -1 is a warm-up
        ce = cudaGraphLaunch(exec, s);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
        for (size_t i = 0; i < n + 2 && ce == cudaSuccess && it >= 0; ++i) {
            float ms = 0.f;
            ce = cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            acc[i] += ms;
        }
        if (ce == cudaSuccess && it >= 0) {
            float ms = 0.f;
            ce = cudaEventElapsedTime(&ms, ev[n + 3], ev[n + 4]);
            empty += ms;
        }
    }
    cudaGraphExecDestroy(exec);
    cleanup();
    DN_REQUIRE(ce == cudaSuccess, DN_ERR_CUDA, "profile replay failed: %s", cudaGetErrorString(ce));
    auto net = [&](double a) { return (float)std::max(0.0, (a - empty) / iters); };
    for (size_t i = 0; i < n; ++i) ms_out_host[i] = e->ops[i].kind == DN_OP_NOP ? 0.f : net(acc[i]);
    ms_out_host[n] = net(acc[n]);                        // softmax + decode (+ histogram, round thresholds)
    ms_out_host[n + 1] = net(acc[n + 1]);                // class sort + NMS + top-D merge, all rounds
    ms_out_host[n + 2] = (float)(empty / iters);         // the cost of an empty event bracket (already subtracted above)
    return DN_OK;
}

// layers (a squeeze-excitation is 4 launches, 3 when the depthwise launch before it pooled) + softmax/decode + round
// thresholds + 3 rounds x (class sort, warp NMS, CTA NMS, merge)
extern "C" int dn_engine_launches_per_forward(dn_engine* e) {
    if (!e) return 0;
    int nop = 0;
    for (const auto& o : e->ops) nop += (o.kind == DN_OP_NOP);
    int folded = 0;
    for (const auto& o : e->ops) folded += (o.kind == DN_OP_SE && o.se_fold && e->desc.gemm_impl == 0);
    return (int)e->ops.size() - nop + 3 * e->n_se - e->n_se_pooled - folded + 14;
}
// Stage-level entry of the folded squeeze-excitation (parity tests): SE scales of x, then the project GEMM with the
// scales applied to its A operand in shared memory.  Equals dn_se_inplace followed by dn_pwconv bit for bit.
extern "C" int dn_se_project(const void* x, const float* se_w1, const float* se_b1, const float* se_w2t, const float* se_b2,
                             const void* w_pw, const float* b_pw, const void* residual, void* y, int B, int HW, int C, int Cs,
                             int N, void* workspace, size_t workspace_bytes, void* stream_) {
    DN_REQUIRE(x && w_pw && b_pw && y, DN_ERR_INVALID, "NULL tensor pointer");
    cudaStream_t s = (cudaStream_t)stream_;
    int rc = se_inplace_pooled(const_cast<void*>(x), se_w1, se_b1, se_w2t, se_b2, B, HW, C, Cs, workspace, workspace_bytes, 0, 0, 0, s,
                               false);
    if (rc) return rc;
    PwEpilogue ep;
    ep.bias = b_pw;
    ep.residual = (const dn_half_t*)residual;
    ep.y = y;
    ep.N = N;
    ep.act = DN_ACT_NONE;
    ep.out_fp32 = 0;
    ep.hw = HW;
    ep.out_batch_stride = (long long)HW * N;
    ep.out_row_stride = N;
    ep.a_scale = se_scales_ptr(workspace, B, C);
    ep.a_scale_c = C;
    return pwconv_tc(x, w_pw, ep, B * HW, C, N, s);
}

extern "C" int dn_engine_get_stats(dn_engine* e, dn_engine_stats* out) {
    DN_REQUIRE(e && out, DN_ERR_INVALID, "NULL argument");
    *out = dn_engine_stats{};
    const dn_engine* last = e->last_slot > 0 ? e->twins[e->last_slot - 1] : e;
    for (const auto& o : e->ops) {
        out->fused_pwdw += (o.kind == DN_OP_PWDW);
        out->fused_dwpw += (o.kind == DN_OP_DWPW);
        out->se_layers += (o.kind == DN_OP_SE);
    }
    out->se_pooled = last->n_se_pooled;
    for (const auto& o : e->ops) out->se_folded += (o.kind == DN_OP_SE && o.se_fold && e->desc.gemm_impl == 0);
    out->launches_per_forward = dn_engine_launches_per_forward(const_cast<dn_engine*>(last));
    out->pipeline_slots = e->n_slots;
    out->last_slot = e->last_slot;
    out->forwards = e->n_forwards;
    out->graph_replays = e->n_graph_replays;
    for (const dn_engine* t : e->twins) out->forwards += t->n_forwards, out->graph_replays += t->n_graph_replays;
    out->act_dtype = DN_ACT_DTYPE_ID;
    return DN_OK;
}

extern "C" size_t dn_engine_device_bytes(dn_engine* e) {
    if (!e) return 0;
    size_t n = e->device_bytes;
    for (const dn_engine* t : e->twins) n += t->device_bytes;
    return n;
}
