// How fast can ANY kernel move the decode kernel's bytes on this GPU?  (tools/ only; not part of the library)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/copy_floor tools/probes/copy_floor.cu && gpurun_out/copy_floor
// Streams R bytes in (three equal arrays) and W bytes out with plain coalesced 128-bit accesses, in the proportions of
// BASELINE config 3 (fp16: 261.3 MB read, 67.7 + 8.5 MB written), for several grid shapes; also read-only and write-only.
// The decode kernel's roofline fraction is quoted against MEASURED_PEAKS.json (a 2 GB torch copy); this probe shows what
// a 338 MB transfer -- launch ramp and tail included -- can reach at best.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

__device__ __forceinline__ uint4 ldg16(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg16(uint4* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// mode 0: 12 loads then 3 or 4 stores (alternating) per thread iteration = 3.43 bytes read per byte written, all
// accesses coalesced 16-byte vectors, every address touched once;   1: read only   2: write only
template <int U> __global__ void stream_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n_in, size_t n_out, int mode) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    uint4 acc = make_uint4(0, 0, 0, 0);
    if (mode != 2) {
        size_t q = 0;
        for (size_t i = tid; i < n_in; i += nthr * U, ++q) {
            uint4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { const size_t j = i + (size_t)u * nthr; v[u] = j < n_in ? ldg16(in + j) : make_uint4(0, 0, 0, 0); }
#pragma unroll
            for (int u = 0; u < U; ++u) { acc.x ^= v[u].x; acc.y += v[u].y; acc.z ^= v[u].z; acc.w += v[u].w; }
            if (mode == 0) {
                const int ns = (q & 1) ? 4 : 3;
                const size_t k = (q >> 1) * 7 * nthr + ((q & 1) ? 3 * nthr : 0) + tid;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < ns && k + j * nthr < n_out) stg16(out + k + j * nthr, acc);
            }
        }
        if (mode == 1 && acc.x == 0x12345u) out[tid % n_out] = acc;
    } else {
        for (size_t k = tid; k < n_out; k += nthr) stg16(out + k, make_uint4((unsigned)k, 1, 2, 3));
    }
}

int main() {
    const size_t R = 261273600, W = 67737600 + 8467200;
    uint4 *in[2], *out;
    for (int i = 0; i < 2; ++i) { cudaMalloc(&in[i], R); cudaMemset(in[i], 1, R); }
    cudaMalloc(&out, W);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const char* names[3] = {"read+write", "read-only", "write-only"};
    for (int mode = 0; mode < 3; ++mode)
        for (int threads : {256, 512, 1024})
            for (int per_sm : {1, 2, 4, 8}) {
                if (threads * per_sm > 2048) continue;
                const int grid = sms * per_sm;
                std::vector<float> ts;
                for (int it = 0; it < 24; ++it) {
                    cudaEventRecord(e0);
                    stream_kernel<12><<<grid, threads>>>(in[it & 1], out, R / 16, W / 16, mode);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    if (it >= 4) ts.push_back(ms * 1e3f);
                }
                std::sort(ts.begin(), ts.end());
                const double bytes = mode == 0 ? (double)R + W : (mode == 1 ? (double)R : (double)W);
                printf("%-10s threads %4d x %d/SM  median %6.1f us  min %6.1f us  -> %7.1f GB/s (median)\n", names[mode], threads, per_sm,
                       ts[ts.size() / 2], ts[0], bytes / ts[ts.size() / 2] / 1e3);
            }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
