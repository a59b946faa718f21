// Probe: how long does one round of 4 coalesced 16-byte loads per thread take when G CTAs x 512 threads all do it
// at once (the shape of nms_kernel's summary pass)?  Prints clock64 cycles for block 0 / thread 0.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(const uint4* __restrict__ buf, size_t per_cta_vec, unsigned long long* out, int rounds, uint4* sink) {
    const uint4* p = buf + (size_t)blockIdx.x * per_cta_vec;
    uint4 acc = make_uint4(0, 0, 0, 0);
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(p + (size_t)(r * 4 + u) * 512 + threadIdx.x);
#pragma unroll
        for (int u = 0; u < 4; ++u) { acc.x ^= v[u].x; acc.y += v[u].y; acc.z ^= v[u].z; acc.w += v[u].w; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc.x == 0x12345678u) sink[0] = acc;
}
int main() {
    const int G = 192, rounds = 2;
    const size_t per_cta = (size_t)rounds * 4 * 512;  // vectors
    uint4* buf; unsigned long long* out; uint4* sink;
    cudaMalloc(&buf, (size_t)G * per_cta * 16); cudaMemset(buf, 1, (size_t)G * per_cta * 16);
    cudaMalloc(&out, G * 8); cudaMalloc(&sink, 16);
    char* flush; cudaMalloc(&flush, 256 << 20);
    unsigned long long h[G];
    for (int warm = 0; warm < 2; ++warm) {
        for (int cold = 0; cold < 2; ++cold) {
            if (cold) cudaMemset(flush, 0, 256 << 20);
            probe<<<G, 512>>>(buf, per_cta, out, rounds, sink);
            cudaMemcpy(h, out, G * 8, cudaMemcpyDeviceToHost);
            unsigned long long mn = ~0ull, mx = 0, sum = 0;
            for (int i = 0; i < G; ++i) { mn = h[i] < mn ? h[i] : mn; mx = h[i] > mx ? h[i] : mx; sum += h[i]; }
            printf("%s: %d rounds of 4x16B/thread, %d CTAs: cycles block0=%llu min=%llu avg=%llu max=%llu\n", cold ? "L2 flushed" : "L2 warm  ",
                   rounds, G, h[0], mn, sum / G, mx);
        }
    }
    return 0;
}
