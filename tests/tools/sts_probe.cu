// Micro-benchmark: shared-memory STORE throughput per SM on sm_100a for the access shapes the weight-gradient ring needs.
// One CTA per SM, NW warps, each thread issues N conflict-free stores of the given width; prints bytes per cycle per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/tools/build_probe/sts_probe tests/tools/sts_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int W>   // W = 1, 2, 4 floats per store
__global__ void __launch_bounds__(512, 1) sts_kernel(float* out, long long* cyc, int iters) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    float v = (float)tid;
    float* base = sm + tid * W;                      // consecutive threads -> consecutive W-float slots: conflict-free
    const int stride = blockDim.x * W;               // next store of the same thread
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float* p = base + k * stride;
            if (W == 1) asm volatile("st.shared.f32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "f"(v));
            else if (W == 2) asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "f"(v), "f"(v));
            else asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "f"(v), "f"(v), "f"(v), "f"(v));
        }
        v += 1.0f;
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    if (out) out[blockIdx.x * blockDim.x + tid] = sm[tid];
}

// the ring-store pattern of fbp_tc_bwd2.cuh: one image row (32 points) per warp store, 4-point core rows 144 bytes apart
__global__ void __launch_bounds__(512, 1) sts_ring_kernel(float* out, long long* cyc, int iters, int with_split) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float v = (float)tid * 1.0001f;
    float* base = sm + (warp & 1) * 16 * 1152 + (warp >> 1) * 288 + (lane >> 2) * 36 + (lane & 3);   // slot, row group (8 units)
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int img = 0; img < 16; ++img)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float x = v + (float)(img * 8 + e);
                if (with_split && (img & 1)) x = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
                base[img * 1152 + e * 4] = x;
            }
        v += 1.0f;
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    if (out) out[blockIdx.x * blockDim.x + tid] = sm[tid];
}

void run_ring(int nthreads, int with_split) {
    long long* d_c;
    float* d_o;
    cudaMalloc(&d_c, 148 * sizeof(long long));
    cudaMalloc(&d_o, 148 * 512 * sizeof(float));
    const int iters = 64;
    const size_t smem = 2 * 16 * 1152 * sizeof(float);
    cudaFuncSetAttribute(sts_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sts_ring_kernel<<<148, nthreads, smem>>>(d_o, d_c, iters, with_split);
    sts_ring_kernel<<<148, nthreads, smem>>>(d_o, d_c, iters, with_split);
    cudaDeviceSynchronize();
    long long c[148];
    cudaMemcpy(c, d_c, sizeof(c), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)c[i];
    avg /= 148;
    const double bytes = (double)nthreads * 4 * 128 * iters;
    printf("ring pattern (split %d)       threads %3d: %8.0f cycles, %6.1f B/cycle/SM, %5.2f warp-stores/cycle, %6.0f cycles per 128 stores (%s)\n",
           with_split, nthreads, avg, bytes / avg, (double)(nthreads / 32) * 128 * iters / avg, avg / iters, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_c);
    cudaFree(d_o);
}

template <int W>
void run(int nthreads, const char* name) {
    long long* d_c;
    float* d_o;
    cudaMalloc(&d_c, 148 * sizeof(long long));
    cudaMalloc(&d_o, 148 * 512 * sizeof(float));
    const int iters = 64;
    const size_t smem = (size_t)nthreads * W * 16 * sizeof(float);
    cudaFuncSetAttribute(sts_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sts_kernel<W><<<148, nthreads, smem>>>(d_o, d_c, iters);
    sts_kernel<W><<<148, nthreads, smem>>>(d_o, d_c, iters);
    cudaDeviceSynchronize();
    long long c[148];
    cudaMemcpy(c, d_c, sizeof(c), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)c[i];
    avg /= 148;
    const double bytes = (double)nthreads * W * 4 * 16 * iters;
    printf("%-28s threads %3d: %8.0f cycles, %6.1f B/cycle/SM, %5.2f warp-stores/cycle (%s)\n", name, nthreads, avg, bytes / avg,
           (double)(nthreads / 32) * 16 * iters / avg, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_c);
    cudaFree(d_o);
}

int main() {
    for (int nt : {256, 512}) {
        run<1>(nt, "STS.32  (128 B per warp)");
        run<2>(nt, "STS.64  (256 B per warp)");
        run<4>(nt, "STS.128 (512 B per warp)");
    }
    run_ring(256, 0);
    run_ring(256, 1);      // (8 warps = the two quarters that store at once; the layout has no room for 16)
    return 0;
}
