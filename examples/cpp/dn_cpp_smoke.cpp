// C++ caller of the C ABI without Python in the process: dlopen()s libdemonet_b200_<dtype>.so, reads a serialised
// engine file (demonet_b200/export.py), creates the engine and runs one batch.
//
// This is the counterpart of the reference's only native file, test/tracing/test_demonet_tracing.cpp:31-58, which
// loads the TorchScript `ssd_lite_mobilenet_v2.pt` with libtorch and runs it on two images, rand{3,320,320} and
// rand{3,256,275} (the second one goes through the transform's resize).  Same two images here: the second is resized
// on the device by dn_resize_bilinear and its boxes are mapped back by dn_rescale_boxes.
//
//   dn_cpp_smoke --symbols LIB                     dlopen + resolve every entry point used below (no GPU needed)
//   dn_cpp_smoke LIB ENGINE_FILE IMAGES OUT        IMAGES: float32 file [3*320*320 | 3*256*275]; OUT: detections (binary)
//
// Build: make -C examples/cpp     (g++, -ldl -lcudart; no torch, no Python)
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/demonet_b200.h"

#define DN_FN(name) decltype(&name) p_##name = nullptr
struct Api {
    void* handle = nullptr;
    DN_FN(dn_abi_version);
    DN_FN(dn_last_error);
    DN_FN(dn_engine_create);
    DN_FN(dn_engine_destroy);
    DN_FN(dn_engine_load_weights);
    DN_FN(dn_engine_forward);
    DN_FN(dn_engine_get_stats);
    DN_FN(dn_resize_bilinear);
    DN_FN(dn_rescale_boxes);
    DN_FN(dn_detections_to_coco);
};

static bool load_api(const char* path, Api* a) {
    a->handle = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!a->handle) {
        fprintf(stderr, "dlopen(%s) failed: %s\n", path, dlerror());
        return false;
    }
    bool ok = true;
#define DN_SYM(name)                                                        \
    a->p_##name = reinterpret_cast<decltype(&name)>(dlsym(a->handle, #name)); \
    if (!a->p_##name) {                                                     \
        fprintf(stderr, "missing symbol %s\n", #name);                      \
        ok = false;                                                         \
    }
    DN_SYM(dn_abi_version)
    DN_SYM(dn_last_error)
    DN_SYM(dn_engine_create)
    DN_SYM(dn_engine_destroy)
    DN_SYM(dn_engine_load_weights)
    DN_SYM(dn_engine_forward)
    DN_SYM(dn_engine_get_stats)
    DN_SYM(dn_resize_bilinear)
    DN_SYM(dn_rescale_boxes)
    DN_SYM(dn_detections_to_coco)
#undef DN_SYM
    return ok;
}

struct EngineFile {
    dn_model_desc desc;
    std::vector<dn_op> ops;
    std::vector<dn_buf> bufs;
    std::vector<float> anchors;
    std::vector<unsigned char> blob;
    int act_dtype = 0;
};

static bool read_engine_file(const char* path, EngineFile* out, int abi) {
    FILE* f = fopen(path, "rb");
    if (!f) {
        fprintf(stderr, "cannot open %s\n", path);
        return false;
    }
    char magic[8];
    int32_t hdr[6];
    int64_t blob_bytes = 0;
    bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "DNENGv1\0", 8) == 0 && fread(hdr, 4, 6, f) == 6 &&
              fread(&blob_bytes, 8, 1, f) == 1;
    if (ok && (hdr[0] != abi || hdr[2] != (int)sizeof(dn_model_desc) || hdr[3] != (int)sizeof(dn_op) || hdr[4] != (int)sizeof(dn_buf))) {
        fprintf(stderr, "engine file / header mismatch: abi %d vs %d, sizes %d/%d/%d vs %zu/%zu/%zu\n", hdr[0], abi, hdr[2], hdr[3],
                hdr[4], sizeof(dn_model_desc), sizeof(dn_op), sizeof(dn_buf));
        ok = false;
    }
    if (ok) {
        out->act_dtype = hdr[1];
        ok = fread(&out->desc, sizeof(dn_model_desc), 1, f) == 1;
    }
    if (ok) {
        out->ops.resize(out->desc.n_ops);
        out->bufs.resize(out->desc.n_bufs);
        out->anchors.resize((size_t)hdr[5] * 4);
        out->blob.resize((size_t)blob_bytes);
        ok = fread(out->ops.data(), sizeof(dn_op), out->ops.size(), f) == out->ops.size() &&
             fread(out->bufs.data(), sizeof(dn_buf), out->bufs.size(), f) == out->bufs.size() &&
             fread(out->anchors.data(), 4, out->anchors.size(), f) == out->anchors.size() &&
             fread(out->blob.data(), 1, out->blob.size(), f) == out->blob.size();
    }
    fclose(f);
    if (!ok) fprintf(stderr, "%s is not a valid engine file\n", path);
    out->desc.ops_host = out->ops.data();
    out->desc.bufs_host = out->bufs.data();
    out->desc.anchors_host = out->anchors.data();
    return ok;
}

#define CUDA_OK(x)                                                                   \
    do {                                                                             \
        cudaError_t e_ = (x);                                                        \
        if (e_ != cudaSuccess) {                                                     \
            fprintf(stderr, "%s failed: %s\n", #x, cudaGetErrorString(e_));          \
            return 2;                                                                \
        }                                                                            \
    } while (0)
#define DN_OK_OR_DIE(x)                                                              \
    do {                                                                             \
        int rc_ = (x);                                                               \
        if (rc_ != 0) {                                                              \
            fprintf(stderr, "%s -> %d: %s\n", #x, rc_, api.p_dn_last_error());       \
            return 3;                                                                \
        }                                                                            \
    } while (0)

int main(int argc, char** argv) {
    Api api;
    if (argc == 3 && std::string(argv[1]) == "--symbols") {
        if (!load_api(argv[2], &api)) return 1;
        printf("ok: %s exports the %d entry points this program binds (ABI version %d)\n", argv[2], 10, api.p_dn_abi_version());
        return api.p_dn_abi_version() == DN_ABI_VERSION ? 0 : 1;
    }
    if (argc != 5) {
        fprintf(stderr, "usage: %s --symbols LIB | %s LIB ENGINE_FILE IMAGES OUT\n", argv[0], argv[0]);
        return 1;
    }
    if (!load_api(argv[1], &api)) return 1;
    if (api.p_dn_abi_version() != DN_ABI_VERSION) {
        fprintf(stderr, "ABI version mismatch: library %d, header %d\n", api.p_dn_abi_version(), DN_ABI_VERSION);
        return 1;
    }
    EngineFile ef;
    if (!read_engine_file(argv[2], &ef, DN_ABI_VERSION)) return 1;
    const int S = ef.desc.image_h, D = ef.desc.post.detections_per_img, B = 2;
    // the two images of test_demonet_tracing.cpp:31-33
    const int h1 = 256, w1 = 275;
    std::vector<float> img0((size_t)3 * S * S), img1((size_t)3 * h1 * w1);
    {
        FILE* f = fopen(argv[3], "rb");
        if (!f || fread(img0.data(), 4, img0.size(), f) != img0.size() || fread(img1.data(), 4, img1.size(), f) != img1.size()) {
            fprintf(stderr, "cannot read the two images from %s\n", argv[3]);
            return 1;
        }
        fclose(f);
    }
    CUDA_OK(cudaSetDevice(0));
    cudaStream_t stream;
    CUDA_OK(cudaStreamCreate(&stream));
    dn_engine* eng = nullptr;
    DN_OK_OR_DIE(api.p_dn_engine_create(&eng, &ef.desc, B));
    DN_OK_OR_DIE(api.p_dn_engine_load_weights(eng, ef.blob.data(), ef.blob.size()));
    float *d_batch, *d_img1, *d_boxes, *d_scores, *d_ratio;
    int64_t* d_labels;
    int32_t* d_counts;
    CUDA_OK(cudaMalloc(&d_batch, (size_t)B * 3 * S * S * 4));
    CUDA_OK(cudaMalloc(&d_img1, img1.size() * 4));
    CUDA_OK(cudaMalloc(&d_boxes, (size_t)B * D * 16));
    CUDA_OK(cudaMalloc(&d_scores, (size_t)B * D * 4));
    CUDA_OK(cudaMalloc(&d_labels, (size_t)B * D * 8));
    CUDA_OK(cudaMalloc(&d_counts, (size_t)B * 4));
    CUDA_OK(cudaMalloc(&d_ratio, (size_t)B * 2 * 4));
    // GeneralizedRCNNTransform: image 0 is already S x S (resize is the identity), image 1 is resized to S x S
    CUDA_OK(cudaMemcpyAsync(d_batch, img0.data(), img0.size() * 4, cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaMemcpyAsync(d_img1, img1.data(), img1.size() * 4, cudaMemcpyHostToDevice, stream));
    DN_OK_OR_DIE(api.p_dn_resize_bilinear(d_img1, 0, 3, h1, w1, d_batch + (size_t)3 * S * S, S, S, stream));
    // two forwards with the same buffers: the first runs eagerly, the second is captured into a CUDA graph, the third replays it
    for (int it = 0; it < 3; ++it)
        DN_OK_OR_DIE(api.p_dn_engine_forward(eng, d_batch, B, d_boxes, d_scores, d_labels, d_counts, stream));
    // transform.postprocess: boxes back to the original image sizes (ratio = original / network size, fp32)
    const float ratio[4] = {1.0f, 1.0f, (float)h1 / (float)S, (float)w1 / (float)S};
    CUDA_OK(cudaMemcpyAsync(d_ratio, ratio, sizeof(ratio), cudaMemcpyHostToDevice, stream));
    DN_OK_OR_DIE(api.p_dn_rescale_boxes(d_boxes, d_ratio, B, D, stream));
    std::vector<float> boxes((size_t)B * D * 4), scores((size_t)B * D);
    std::vector<int64_t> labels((size_t)B * D);
    std::vector<int32_t> counts(B);
    CUDA_OK(cudaMemcpyAsync(boxes.data(), d_boxes, boxes.size() * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaMemcpyAsync(scores.data(), d_scores, scores.size() * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaMemcpyAsync(labels.data(), d_labels, labels.size() * 8, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaMemcpyAsync(counts.data(), d_counts, counts.size() * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    dn_engine_stats st;
    DN_OK_OR_DIE(api.p_dn_engine_get_stats(eng, &st));
    for (int b = 0; b < B; ++b) {
        printf("image %d: %d detections", b, counts[b]);
        for (int i = 0; i < counts[b] && i < 3; ++i) {
            const float* q = &boxes[((size_t)b * D + i) * 4];
            printf("  [label %lld score %.4f box %.1f %.1f %.1f %.1f]", (long long)labels[(size_t)b * D + i], scores[(size_t)b * D + i],
                   q[0], q[1], q[2], q[3]);
        }
        printf("\n");
    }
    printf("engine: %d launches per forward, %lld forwards, %lld graph replays, %s activations\n", st.launches_per_forward,
           (long long)st.forwards, (long long)st.graph_replays, st.act_dtype ? "fp16" : "bf16");
    FILE* f = fopen(argv[4], "wb");
    if (!f) return 1;
    const int32_t hdr[2] = {B, D};
    fwrite(hdr, 4, 2, f);
    fwrite(counts.data(), 4, counts.size(), f);
    fwrite(boxes.data(), 4, boxes.size(), f);
    fwrite(scores.data(), 4, scores.size(), f);
    fwrite(labels.data(), 8, labels.size(), f);
    fclose(f);
    api.p_dn_engine_destroy(eng);
    cudaFree(d_batch); cudaFree(d_img1); cudaFree(d_boxes); cudaFree(d_scores); cudaFree(d_labels); cudaFree(d_counts); cudaFree(d_ratio);
    cudaStreamDestroy(stream);
    return st.graph_replays >= 1 ? 0 : 4;
}
