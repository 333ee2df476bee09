// Native timing loop over the C ABI, free of Python / torch dispatch overhead.
//
// B200-native counterpart of the reference's only native source, cpp/main.cpp:69-92 (a pybind11-embed
// executable that times 10 000 is_collision_free calls from C++).  This one links libmrb200.so directly:
// it reads a compiled scene blob (scripts/export_blob.py), uploads it, and times mrb200_check_configs /
// mrb200_check_edges on uniform random configurations with CUDA events.
//
//   nvcc -O2 -o native/bench_main native/bench_main.cpp -Iinclude -Lmultirobot_pathplanning_benchmark_b200 -lmrb200
//   LD_LIBRARY_PATH=multirobot_pathplanning_benchmark_b200 native/bench_main scene.blob [B] [reps]
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "mrb200.h"

#define CK(x)                                                                                   \
    do {                                                                                        \
        int rc_ = (x);                                                                          \
        if (rc_) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, mrb200_last_error()); return 1; } \
    } while (0)

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s scene.blob [configs] [reps]\n", argv[0]); return 2; }
    const long B = argc > 2 ? atol(argv[2]) : (1 << 22);
    const int reps = argc > 3 ? atoi(argv[3]) : 10;
    // file = blob words, then D floats lower limits, D floats upper limits
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<unsigned char> file(sz);
    if (fread(file.data(), 1, sz, f) != (size_t)sz) return 2;
    fclose(f);
    const unsigned* words = (const unsigned*)file.data();
    const size_t blob_bytes = (size_t)words[15] * 4;  // MRB_H_TOTAL_WORDS
    const int D = (int)words[2];                      // MRB_H_DOF
    const float* lim = (const float*)(file.data() + blob_bytes);

    mrb200_scene_t* scene;
    CK(mrb200_scene_create(2, &scene));
    CK(mrb200_scene_set_mode(scene, 0, file.data(), blob_bytes, nullptr));

    std::mt19937 rng(0);
    std::vector<float> q((size_t)B * D);
    for (long i = 0; i < B; i++)
        for (int k = 0; k < D; k++) q[i * D + k] = std::uniform_real_distribution<float>(lim[k], lim[D + k])(rng);
    float* dq;
    unsigned char* dflags;
    cudaMalloc(&dq, q.size() * 4);
    cudaMalloc(&dflags, B);
    cudaMemcpy(dq, q.data(), q.size() * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) CK(mrb200_check_configs(scene, 0, dq, B, -1.f, dflags, nullptr, 0, nullptr));
    cudaDeviceSynchronize();  // the library settles on a kernel variant once it has seen results of the first launches
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) CK(mrb200_check_configs(scene, 0, dq, B, -1.f, dflags, nullptr, 0, nullptr));
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<unsigned char> flags(B);
    cudaMemcpy(flags.data(), dflags, B, cudaMemcpyDeviceToHost);
    long nfree = 0;
    for (long i = 0; i < B; i++) nfree += flags[i];
    printf("configs: %ld x %d dof, %.3f ms per launch, %.4g checks/s, %.3f free\n", B, D, ms / reps, B * (double)reps / (ms * 1e-3),
           nfree / (double)B);

    // edges between consecutive configurations, resolution 0.01
    const long E = B / 64;
    int* dfirst;
    cudaMalloc(&dfirst, E * 4);
    for (int i = 0; i < 2; i++)
        CK(mrb200_check_edges(scene, 0, dq, dq + (size_t)E * D, E, 0.01, nullptr, 0, -1, 0, -1.f, dflags, dfirst, nullptr));
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++)
        CK(mrb200_check_edges(scene, 0, dq, dq + (size_t)E * D, E, 0.01, nullptr, 0, -1, 0, -1.f, dflags, dfirst, nullptr));
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("edges:   %ld, %.3f ms per launch, %.4g edge checks/s\n", E, ms / reps, E * (double)reps / (ms * 1e-3));

    // planner-like local edges: free start configurations, every joint moves by at most +-0.2 (clamped to the limits)
    {
        std::vector<float> a, b;
        std::mt19937 r2(1);
        const long want = 131072;
        for (long i = 0; i < B && (long)a.size() < want * D; i++) {
            if (!flags[i]) continue;
            for (int k = 0; k < D; k++) {
                const float x = q[i * D + k];
                float y = x + std::uniform_real_distribution<float>(-0.2f, 0.2f)(r2);
                y = y < lim[k] ? lim[k] : (y > lim[D + k] ? lim[D + k] : y);
                a.push_back(x);
                b.push_back(y);
            }
        }
        const long El = (long)a.size() / D;
        if (El > 0) {
            float *da, *db;
            int* dfirst2;
            cudaMalloc(&dfirst2, El * 4);
            cudaMalloc(&da, a.size() * 4);
            cudaMalloc(&db, b.size() * 4);
            cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
            for (int i = 0; i < 2; i++)
                CK(mrb200_check_edges(scene, 0, da, db, El, 0.01, nullptr, 0, -1, 0, -1.f, dflags, dfirst2, nullptr));
            cudaEventRecord(e0);
            for (int i = 0; i < reps; i++)
                CK(mrb200_check_edges(scene, 0, da, db, El, 0.01, nullptr, 0, -1, 0, -1.f, dflags, dfirst2, nullptr));
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            printf("local:   %ld, %.3f ms per launch, %.4g edge checks/s\n", El, ms / reps, El * (double)reps / (ms * 1e-3));
        }
    }
    mrb200_scene_destroy(scene);
    return 0;
}
