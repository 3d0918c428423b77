// streambench.cu -- practical HBM ceiling for the access mixes our kernels generate (run on the GPU box).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/streambench tools/streambench.cu
// Modes: R reads + W writes of distinct 1 GiB arrays (double2 per thread), flat grid-stride or k-streaming tiles.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Ptrs { const double2* in[8]; double2* out[8]; };

template <int R, int W>
__global__ void __launch_bounds__(256) flat_kernel(Ptrs p, size_t n2) {
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n2; t += (size_t)gridDim.x * blockDim.x) {
        double2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = p.in[r][t];
        double2 s = v[0];
#pragma unroll
        for (int r = 1; r < R; ++r) { s.x += v[r].x; s.y += v[r].y; }
#pragma unroll
        for (int w = 0; w < W; ++w) p.out[w][t] = make_double2(s.x + w, s.y - w);
    }
}

// tile (32 lanes x 8 rows) x k-chunk streaming, like the solver kernels; n = grid edge
template <int R, int W>
__global__ void __launch_bounds__(256) tile_kernel(Ptrs p, int n, int kc) {
    const int i2 = blockIdx.x * 32 + threadIdx.x;           // double2 index in the row
    const int j = blockIdx.y * 8 + threadIdx.y;
    if (i2 * 2 >= n || j >= n) return;
    const size_t plane2 = (size_t)n * n / 2, row2 = (size_t)n / 2;
    const int kb = blockIdx.z * kc;
    for (int k = kb; k < kb + kc && k < n; ++k) {
        const size_t t = (size_t)k * plane2 + (size_t)j * row2 + i2;
        double2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = p.in[r][t];
        double2 s = v[0];
#pragma unroll
        for (int r = 1; r < R; ++r) { s.x += v[r].x; s.y += v[r].y; }
#pragma unroll
        for (int w = 0; w < W; ++w) p.out[w][t] = make_double2(s.x + w, s.y - w);
    }
}

template <int R, int W>
static void run(Ptrs p, int n, const char* name) {
    const size_t n2 = (size_t)n * n * n / 2;
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const double bytes = (double)(R + W) * n2 * 16;
    for (int mode = 0; mode < 2; ++mode) {
        float best = 1e9f;
        for (int it = 0; it < 6; ++it) {
            CK(cudaEventRecord(a));
            if (mode == 0) flat_kernel<R, W><<<148 * 8, 256>>>(p, n2);
            else tile_kernel<R, W><<<dim3((n / 2 + 31) / 32, n / 8, (n + 63) / 64), dim3(32, 8)>>>(p, n, 64);
            CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
            float ms; CK(cudaEventElapsedTime(&ms, a, b));
            if (it > 0 && ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("%-10s %s  R=%d W=%d  %.3f ms  %.0f GB/s\n", name, mode == 0 ? "flat" : "tile", R, W, best, bytes / best / 1e6);
    }
}

int main() {
    const int n = 512;
    const size_t bytes = (size_t)n * n * n * 8;
    Ptrs p;
    for (int r = 0; r < 6; ++r) { CK(cudaMalloc((void**)&p.in[r], bytes)); CK(cudaMemset((void*)p.in[r], 1, bytes)); }
    for (int w = 0; w < 6; ++w) { CK(cudaMalloc((void**)&p.out[w], bytes)); CK(cudaMemset((void*)p.out[w], 0, bytes)); }
    run<1, 1>(p, n, "copy");
    run<2, 1>(p, n, "2r1w");
    run<6, 3>(p, n, "6r3w");
    run<3, 3>(p, n, "3r3w");
    run<6, 6>(p, n, "6r6w");
    run<6, 1>(p, n, "6r1w");
    run<1, 6>(p, n, "1r6w");
    return 0;
}
