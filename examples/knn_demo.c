/* Plain-C caller of the C ABI (no Python, no torch): the replacement for
 *   cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)            (MonocularSlam/slam.py:126-127)
 * Build:  gcc -std=c99 -Iinclude -I/usr/local/cuda/include examples/knn_demo.c -o knn_demo \
 *             -Lopenwork_b200/_lib -ltvg_b200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/openwork_b200/_lib
 * Run:    ./knn_demo 300 500 7      -> prints  nq nt  checksum(idx) checksum(dist) kept
 * (tests/test_gpu_c_example.py runs it and recomputes the checksums with the numpy oracle.) */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "tvg_b200.h"

static uint32_t lcg(uint32_t* s) { *s = *s * 1664525u + 1013904223u; return *s; }

int main(int argc, char** argv) {
    const int nq = argc > 1 ? atoi(argv[1]) : 300, nt = argc > 2 ? atoi(argv[2]) : 500;
    uint32_t seed = argc > 3 ? (uint32_t)atoi(argv[3]) : 7u;
    const int n_max = nq > nt ? nq : nt;
    uint8_t* hq = (uint8_t*)malloc((size_t)nq * 32);
    uint8_t* ht = (uint8_t*)malloc((size_t)nt * 32);
    for (int i = 0; i < nq * 32; ++i) hq[i] = (uint8_t)(lcg(&seed) >> 24);
    for (int i = 0; i < nt * 32; ++i) ht[i] = (uint8_t)(lcg(&seed) >> 24);
    for (int i = 0; i < nt / 4; ++i)  /* duplicate rows: exact ties must break to the lowest train index */
        for (int b = 0; b < 32; ++b) ht[(size_t)(nt - 1 - i) * 32 + b] = ht[(size_t)i * 32 + b];

    tvg_ctx* ctx = NULL;
    if (tvg_ctx_create(&ctx, 0) != 0) { fprintf(stderr, "ctx: %s\n", tvg_last_error()); return 2; }
    uint8_t *dq, *dt, *dkeep;
    int32_t *didx, *ddist;
    cudaMalloc((void**)&dq, (size_t)n_max * 32); cudaMalloc((void**)&dt, (size_t)n_max * 32);
    cudaMalloc((void**)&didx, (size_t)n_max * 8); cudaMalloc((void**)&ddist, (size_t)n_max * 8);
    cudaMalloc((void**)&dkeep, (size_t)n_max);
    cudaMemcpy(dq, hq, (size_t)nq * 32, cudaMemcpyHostToDevice);
    cudaMemcpy(dt, ht, (size_t)nt * 32, cudaMemcpyHostToDevice);
    int rc = tvg_knn2_hamming256(ctx, dq, dt, 1, n_max, nq, nt, NULL, NULL, NULL, NULL, 0.75, didx, ddist, dkeep, NULL);
    if (rc != 0) { fprintf(stderr, "knn: %s\n", tvg_last_error()); return 3; }
    int32_t* idx = (int32_t*)malloc((size_t)nq * 8);
    int32_t* dist = (int32_t*)malloc((size_t)nq * 8);
    uint8_t* keep = (uint8_t*)malloc((size_t)nq);
    cudaMemcpy(idx, didx, (size_t)nq * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(dist, ddist, (size_t)nq * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(keep, dkeep, (size_t)nq, cudaMemcpyDeviceToHost);
    uint64_t ci = 0, cd = 0;
    int kept = 0;
    for (int i = 0; i < nq; ++i) {
        ci = ci * 1000003u + (uint64_t)(uint32_t)idx[2 * i] * 65537u + (uint64_t)(uint32_t)idx[2 * i + 1];
        cd = cd * 1000003u + (uint64_t)(uint32_t)dist[2 * i] * 65537u + (uint64_t)(uint32_t)dist[2 * i + 1];
        kept += keep[i] ? 1 : 0;
    }
    printf("%d %d %llu %llu %d %lld\n", nq, nt, (unsigned long long)ci, (unsigned long long)cd, kept,
           tvg_launch_count(ctx));
    tvg_ctx_destroy(ctx);
    return 0;
}
