// Sphere-agent ("abstract") environment on the device, fp64, bit-exact with the reference:
//   AbstractEnvironment.is_collision_free       P/problems/abstract_env.py:255-276
//   Sphere.collides_with_sphere      `<`        P/problems/abstract_env.py:46-49
//   Rectangle.collides_with_sphere   `<=` on squared distance   :69-84
//   AbstractEnvironment.is_edge_collision_free  P/problems/abstract_env.py:301-354
// Rounding follows numpy on the reference's host exactly: np.linalg.norm = sqrt(ddot(x, x)) and
// OpenBLAS' ddot is a sequential FMA chain for short vectors (s = x0*x0; s = fma(xk, xk, s));
// np.sum(d**2) rounds every product and sum separately in numpy's pairwise order.  Flags are
// therefore identical to the reference's on the same fp64 inputs, boundary cases included.
#include <cuda_runtime.h>
#include <stdint.h>

#include "binary_order.cuh"
#include "kernels.h"

namespace mrb {

__device__ __forceinline__ bool abstract_free(const AbstractSceneData& sc, const double* q) {
    const int na = sc.n_agents, dim = sc.dim;
    for (int i = 0; i < na; i++)
        for (int j = i + 1; j < na; j++) {
            double s = 0.0;
            for (int k = 0; k < dim; k++) {
                const double d = __dsub_rn(q[i * dim + k], q[j * dim + k]);
                s = k == 0 ? __dmul_rn(d, d) : __fma_rn(d, d, s);
            }
            if (__dsqrt_rn(s) < __dadd_rn(sc.radii[i], sc.radii[j])) return false;
        }
    for (int i = 0; i < na; i++) {
        for (int o = 0; o < sc.n_sph; o++) {
            double s = 0.0;
            for (int k = 0; k < dim; k++) {
                const double d = __dsub_rn(sc.sph_c[o][k], q[i * dim + k]);
                s = k == 0 ? __dmul_rn(d, d) : __fma_rn(d, d, s);
            }
            if (__dsqrt_rn(s) < __dadd_rn(sc.sph_r[o], sc.radii[i])) return false;
        }
        for (int o = 0; o < sc.n_rect; o++) {
            // np.sum((clip(c) - c) ** 2): numpy's pairwise summation = plain left-to-right loop
            // below 8 terms, 8 interleaved accumulators combined as a tree from 8 terms on
            double sq[ABS_MAX_DIM];
            for (int k = 0; k < dim; k++) {
                const double c = q[i * dim + k];
                const double cp = fmin(fmax(c, sc.rect_min[o][k]), sc.rect_max[o][k]);
                const double d = __dsub_rn(cp, c);
                sq[k] = __dmul_rn(d, d);
            }
            double s;
            if (dim < 8) {
                s = sq[0];
                for (int k = 1; k < dim; k++) s = __dadd_rn(s, sq[k]);
            } else {
                double r[8];
                for (int k = 0; k < 8; k++) r[k] = sq[k];
                int k = 8;
                for (; k + 8 <= dim; k += 8)
                    for (int u = 0; u < 8; u++) r[u] = __dadd_rn(r[u], sq[k + u]);
                s = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                              __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
                for (; k < dim; k++) s = __dadd_rn(s, sq[k]);
            }
            if (s <= __dmul_rn(sc.radii[i], sc.radii[i])) return false;
        }
    }
    return true;
}

__global__ void __launch_bounds__(256) abstract_configs_kernel(const __grid_constant__ AbstractSceneData sc, const double* __restrict__ q,
                                                               int64_t B, uint8_t* __restrict__ flags) {
    const int D = sc.n_agents * sc.dim;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
        double cfg[ABS_MAX_AGENTS * ABS_MAX_DIM];
        for (int k = 0; k < D; k++) cfg[k] = q[i * D + k];
        flags[i] = abstract_free(sc, cfg) ? 1 : 0;
    }
}

// one warp per edge, 32 interpolation points per step in binary order, early exit per step
__global__ void __launch_bounds__(256) abstract_edges_kernel(const __grid_constant__ AbstractSceneData sc, const double* __restrict__ q1,
                                                             const double* __restrict__ q2, int64_t E, double resolution,
                                                             const int32_t* __restrict__ Ns, int n_start, int n_max,
                                                             int include_endpoints, uint8_t* __restrict__ flags,
                                                             int32_t* __restrict__ first_pos, int* counter) {
    const int D = sc.n_agents * sc.dim;
    const int lane = threadIdx.x & 31;
    for (;;) {
        int64_t e = 0;
        if (lane == 0) e = atomicAdd(counter, 1);
        e = __shfl_sync(0xffffffffu, e, 0);
        if (e >= E) return;
        double a[ABS_MAX_AGENTS * ABS_MAX_DIM], step[ABS_MAX_AGENTS * ABS_MAX_DIM];
        double m = 0.0;
        for (int k = 0; k < D; k++) {
            a[k] = q1[e * D + k];
            const double b = q2[e * D + k];
            step[k] = __dsub_rn(b, a[k]);
            m = fmax(m, fabs(__dsub_rn(a[k], b)));
        }
        const int N = Ns ? Ns[e] : max(2, (int)__ddiv_rn(m, resolution) + 1);
        const double inv = (double)(N - 1);
        for (int k = 0; k < D; k++) step[k] = __ddiv_rn(step[k], inv);
        const int nmax = (n_max < 0 || n_max > N) ? N : n_max;
        int first = -1;
        for (int base = n_start; base < nmax && first < 0; base += 32) {
            const int pos = base + lane;
            bool hit = false;
            if (pos < nmax) {
                const int i = binary_order_index(N, pos);
                if (include_endpoints || (i != 0 && i != N - 1)) {
                    double cfg[ABS_MAX_AGENTS * ABS_MAX_DIM];
                    for (int k = 0; k < D; k++) cfg[k] = __dadd_rn(a[k], __dmul_rn(step[k], (double)i));
                    hit = !abstract_free(sc, cfg);
                }
            }
            const unsigned mask = __ballot_sync(0xffffffffu, hit);
            if (mask) first = base + __ffs(mask) - 1;
        }
        if (lane == 0) {
            flags[e] = first < 0 ? 1 : 0;
            if (first_pos) first_pos[e] = first;
        }
    }
}

// FP32 FMA peak probe: 16 independent FMA chains per thread, unrolled so that loop control is < 1 % of the issued
// instructions (the round-1 probe, 8 chains with the compiler's own unrolling, reached 86 % of the theoretical rate).
// bench.py divides threads * iters * 16 * 2 flop by the measured time to get this GPU's SIMT roofline.
__global__ void __launch_bounds__(256) fp32_probe_kernel(int iters, float* out) {
    float a[16];
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = threadIdx.x * 1e-3f + (float)k;
    const float m = 0.999f, c = 1e-3f;
#pragma unroll 1
    for (int i = 0; i < iters; i += 16) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int k = 0; k < 16; k++) a[k] = fmaf(a[k], m, c);
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 16; k++) sum += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

static int sm_count() {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

cudaError_t launch_fp32_probe(int iters, float* out, int* n_threads, cudaStream_t st) {
    const int blocks = sm_count() * 8;
    *n_threads = blocks * 256;
    if (out) fp32_probe_kernel<<<blocks, 256, 0, st>>>(iters, out);
    return cudaGetLastError();
}

cudaError_t launch_abstract_configs(const AbstractSceneData& sc, const double* q, int64_t B, uint8_t* flags, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    int64_t blocks = (B + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    abstract_configs_kernel<<<(int)blocks, 256, 0, st>>>(sc, q, B, flags);
    return cudaGetLastError();
}

cudaError_t launch_abstract_edges(const AbstractSceneData& sc, const double* q1, const double* q2, int64_t E, double resolution,
                                  const int32_t* N, int n_start, int n_max, int include_endpoints, uint8_t* flags,
                                  int32_t* first_pos, int* counter, cudaStream_t st) {
    if (E <= 0) return cudaSuccess;
    cudaError_t err = cudaMemsetAsync(counter, 0, sizeof(int), st);
    if (err != cudaSuccess) return err;
    int64_t blocks = (E + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    abstract_edges_kernel<<<(int)blocks, 256, 0, st>>>(sc, q1, q2, E, resolution, N, n_start, n_max, include_endpoints, flags,
                                                        first_pos, counter);
    return cudaGetLastError();
}

}  // namespace mrb
