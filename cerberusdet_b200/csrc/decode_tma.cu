// Fused Detect-head decode, persistent TMA-pipelined variant (the one the library uses whenever the
// tensors allow 16-byte rows; decode.cu is the generic fallback and the comparison point).
//
// Why: the register-resident kernel in decode.cu is latency bound -- a warp first waits for its 16
// loads, then spends ~7000 cycles on exp/rcp work with nothing in flight, and at 146 registers only
// 12 warps fit on an SM, so on average ~20 KB per SM are in flight where HBM3e needs ~45 KB
// (profiles/r01_decode_v1.md).  Here one producer warp per SM streams [64+nc] x TA tiles into a
// shared-memory ring with cp.async.bulk (TMA, mbarrier complete_tx) several tiles ahead of the 16
// consumer warps, which only ever touch shared memory.
//
// Same arithmetic and rounding points as decode.cu (reference models/yolo.py:93-99, utils/tal.py).
//
// Tile = TA consecutive anchors of one (task, level, image), all 64+nc channels: TA = 256 (fp16) or
// 128 (fp32), i.e. 512-byte rows, one bulk copy per channel row.  Consumer thread (p, q): p = anchor
// pair (fp16) / anchor (fp32) inside the tile, q = quarter: DFL side q (l, t, r, b) plus a quarter of the
// class channels.  Sides meet through a small shared scratch (one named barrier per tile): q = 0
// finishes (cx, w), q = 1 finishes (cy, h).
#include "cerb_kernels.h"

#define TMA_CONSUMER_WARPS 16
#define TMA_CONSUMERS (TMA_CONSUMER_WARPS * 32)
#define TMA_THREADS (TMA_CONSUMERS + 32)
#define TMA_PAIRS 128  // consumer threads per quarter
#define TMA_MAX_STAGES 8
#define LOG2E_F 1.4426950408889634f

struct TmaDecodeParams {
    DecodeParams d;
    int tiles_per_image[CERB_MAX_LEVELS];
    int row_tile_start[CERB_MAX_TASKS * CERB_MAX_LEVELS + 1];  // first tile of each (task, level) row
    int ntiles;
    int stages;
    int stage_bytes;  // no_max * TA * sizeof(T)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(TMA_CONSUMERS) : "memory"); }

// LW elements (4 bytes) of one channel row as floats
template <typename T> struct Lane;
template <> struct Lane<__half> {
    static constexpr int LW = 2;
    __device__ static __forceinline__ void load(const void* p, float (&x)[2]) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(p));
        x[0] = f.x; x[1] = f.y;
    }
    __device__ static __forceinline__ void store(void* p, const float (&x)[2]) {
        *reinterpret_cast<__half2*>(p) = __floats2half2_rn(x[0], x[1]);
    }
};
template <> struct Lane<float> {
    static constexpr int LW = 1;
    __device__ static __forceinline__ void load(const void* p, float (&x)[1]) { x[0] = *reinterpret_cast<const float*>(p); }
    __device__ static __forceinline__ void store(void* p, const float (&x)[1]) { *reinterpret_cast<float*>(p) = x[0]; }
};

struct TileInfo { int task, level, b, a0, cnt; };

__device__ __forceinline__ TileInfo locate_tile(const TmaDecodeParams& P, int t, int TA) {
    int lo = 0, hi = P.d.nrows;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (t >= P.row_tile_start[mid]) lo = mid; else hi = mid;
    }
    TileInfo ti;
    ti.task = lo / P.d.L;
    ti.level = lo - ti.task * P.d.L;
    const int r = t - P.row_tile_start[lo];
    const int tpi = P.tiles_per_image[ti.level];
    ti.b = r / tpi;
    ti.a0 = (r - ti.b * tpi) * TA;
    ti.cnt = min(TA, P.d.hw[ti.level] - ti.a0);
    return ti;
}

template <typename T>
__global__ void __launch_bounds__(TMA_THREADS, 1) decode_tma_kernel(const __grid_constant__ TmaDecodeParams P) {
    constexpr int LW = Lane<T>::LW;
    constexpr int TA = TMA_PAIRS * LW;  // anchors per tile
    constexpr int V = 16 / sizeof(T);   // anchors per summary entry
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[2 * TMA_MAX_STAGES];
    __shared__ float dscr[2][4][TA];  // DFL distances of the four sides, double buffered by tile parity

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int stages = P.stages;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);                                    // full: producer's expect_tx arrive
            mbar_init(smem_u32(&bars[TMA_MAX_STAGES + s]), TMA_CONSUMER_WARPS);  // empty: one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (wid == TMA_CONSUMER_WARPS) {
        // ------------------------------------------------ producer warp
        int k = 0;
        for (int t = blockIdx.x; t < P.ntiles; t += gridDim.x, ++k) {
            const int s = k % stages;
            const uint32_t use = (uint32_t)(k / stages);
            mbar_wait(smem_u32(&bars[TMA_MAX_STAGES + s]), (use & 1u) ^ 1u);
            const TileInfo ti = locate_tile(P, t, TA);
            const int no = 4 * CERB_REG_MAX + P.d.nc[ti.task];
            const int hw = P.d.hw[ti.level];
            const uint32_t row_bytes = (uint32_t)(ti.cnt * sizeof(T));
            const uint32_t full = smem_u32(&bars[s]);
            if (lane == 0) mbar_arrive_expect_tx(full, row_bytes * (uint32_t)no);
            __syncwarp();
            const T* src = reinterpret_cast<const T*>(P.d.lvl[ti.task][ti.level]) + (size_t)ti.b * no * hw + ti.a0;
            const uint32_t dst = smem_u32(smem + (size_t)s * P.stage_bytes);
            for (int r = lane; r < no; r += 32)
                bulk_g2s(dst + (uint32_t)(r * TA * sizeof(T)), src + (size_t)r * hw, row_bytes, full);
        }
        return;
    }

    // ---------------------------------------------------- consumers
    const int q = wid >> 2;                  // quarter: DFL side q + a quarter of the classes
    const int p = (wid & 3) * 32 + lane;     // lane-column inside the tile
    int k = 0;
    for (int t = blockIdx.x; t < P.ntiles; t += gridDim.x, ++k) {
        const int s = k % stages;
        const uint32_t use = (uint32_t)(k / stages);
        const TileInfo ti = locate_tile(P, t, TA);
        const int nc = P.d.nc[ti.task];
        const int A = P.d.A;
        const int al = p * LW;                       // first anchor of this thread inside the tile
        const bool live = al < ti.cnt;               // cnt is a multiple of V >= LW: whole lanes are live or not
        const unsigned char* st = smem + (size_t)s * P.stage_bytes + (size_t)al * sizeof(T);
        T* __restrict__ out = reinterpret_cast<T*>(P.d.y[ti.task]) + (size_t)ti.b * (4 + nc) * A + P.d.aoff[ti.level] + ti.a0 + al;

        mbar_wait(smem_u32(&bars[s]), use & 1u);

        // ---- DFL side q: expectation of softmax over the 16 bins (reference models/yolo.py:57-59)
        float d[LW];
        {
            float x[CERB_REG_MAX][LW];
            float m[LW];
#pragma unroll
            for (int i = 0; i < LW; ++i) m[i] = -INFINITY;
#pragma unroll
            for (int kk = 0; kk < CERB_REG_MAX; ++kk) {
                Lane<T>::load(st + (size_t)(q * CERB_REG_MAX + kk) * TA * sizeof(T), x[kk]);
#pragma unroll
                for (int i = 0; i < LW; ++i) m[i] = fmaxf(m[i], x[kk][i]);
            }
#pragma unroll
            for (int i = 0; i < LW; ++i) {
                const float mb = m[i] * LOG2E_F;
                float ssum = 0.f;
#pragma unroll
                for (int kk = 0; kk < CERB_REG_MAX; ++kk) {
                    x[kk][i] = fast_ex2(fmaf(x[kk][i], LOG2E_F, -mb));
                    ssum += x[kk][i];
                }
                const float inv = fast_rcp(ssum);
                float acc = 0.f;
#pragma unroll
                for (int kk = 1; kk < CERB_REG_MAX; ++kk) acc = fmaf((float)kk, rnd<T>(x[kk][i] * inv), acc);
                d[i] = rnd<T>(acc);
            }
        }
#pragma unroll
        for (int i = 0; i < LW; ++i) dscr[k & 1][q][al + i] = d[i];

        // ---- this thread's quarter of the class channels: sigmoid (yolo.py:99) + score summary
        {
            const int cq = (nc + 3) >> 2;
            const int c0 = q * cq, c1 = min(nc, c0 + cq);
            const size_t srow = ((size_t)(A / V) + V - 1) / V * V;
            T* smx = nullptr;
            if (P.d.smax[ti.task] != nullptr)
                smx = reinterpret_cast<T*>(P.d.smax[ti.task]) + (size_t)ti.b * nc * srow + (P.d.aoff[ti.level] + ti.a0 + al) / V;
            const bool writer = (lane % (V / LW)) == 0;
            for (int c = c0; c < c1; ++c) {
                float sc[LW];
                float mx = -INFINITY;
                if (live) {
                    Lane<T>::load(st + (size_t)(4 * CERB_REG_MAX + c) * TA * sizeof(T), sc);
#pragma unroll
                    for (int i = 0; i < LW; ++i) {
                        sc[i] = rnd<T>(fast_rcp(1.f + fast_ex2(-sc[i] * LOG2E_F)));
                        mx = fmaxf(mx, sc[i]);
                    }
                    Lane<T>::store(out + (size_t)(4 + c) * A, sc);
                }
                if (P.d.smax[ti.task] != nullptr) {
#pragma unroll
                    for (int o = 1; o < V / LW; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    if (live && writer) smx[(size_t)c * srow] = from_f32<T>(mx);
                }
            }
        }
        // all reads of the stage are done: hand it back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[TMA_MAX_STAGES + s]));

        // ---- sides meet: q = 0 finishes (cx, w) from (l, r), q = 1 finishes (cy, h) from (t, b)
        consumer_bar();
        if (q < 2 && live) {
            const int W = P.d.w[ti.level];
            const float stride = P.d.stride[ti.level];
            float oc[LW], os[LW];
#pragma unroll
            for (int i = 0; i < LW; ++i) {
                const int a = ti.a0 + al + i;
                const int g = (q == 0) ? (a % W) : (a / W);
                const float ac = rnd<T>(rnd<T>((float)g) + 0.5f);  // arange(dtype) + 0.5, tal.py:188-189
                const float p1 = rnd<T>(ac - dscr[k & 1][q][al + i]);
                const float p2 = rnd<T>(ac + dscr[k & 1][q + 2][al + i]);
                oc[i] = rnd<T>(rnd<T>(rnd<T>(p1 + p2) * 0.5f) * stride);  // utils/tal.py:198-204, yolo.py:98
                os[i] = rnd<T>(rnd<T>(p2 - p1) * stride);
            }
            Lane<T>::store(out + (size_t)q * A, oc);
            Lane<T>::store(out + (size_t)(q + 2) * A, os);
        }
    }
}

template <typename T> static cudaError_t launch_tma_t(const DecodeParams& D, cudaStream_t stream) {
    constexpr int TA = TMA_PAIRS * Lane<T>::LW;
    TmaDecodeParams P;
    P.d = D;
    int tiles = 0, no_max = 0;
    for (int l = 0; l < D.L; ++l) P.tiles_per_image[l] = (D.hw[l] + TA - 1) / TA;
    for (int t = 0; t < D.T; ++t) {
        no_max = max(no_max, 4 * CERB_REG_MAX + D.nc[t]);
        for (int l = 0; l < D.L; ++l) {
            P.row_tile_start[t * D.L + l] = tiles;
            tiles += D.B * P.tiles_per_image[l];
        }
    }
    P.row_tile_start[D.nrows] = tiles;
    P.ntiles = tiles;
    if (tiles == 0) return cudaSuccess;
    P.stage_bytes = no_max * TA * (int)sizeof(T);
    int dev = 0, sms = 0, smem_max = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int static_smem = 2 * 4 * TA * 4 + 2 * TMA_MAX_STAGES * 8 + 256;
    int stages = (smem_max - static_smem) / P.stage_bytes;
    if (stages > TMA_MAX_STAGES) stages = TMA_MAX_STAGES;
    if (stages < 2) return cudaErrorInvalidConfiguration;  // caller falls back to the generic kernel
    P.stages = stages;
    const size_t dyn = (size_t)stages * P.stage_bytes;
    cudaError_t e = cudaFuncSetAttribute(decode_tma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    const int grid = tiles < sms ? tiles : sms;
    decode_tma_kernel<T><<<grid, TMA_THREADS, dyn, stream>>>(P);
    return cudaGetLastError();
}

// Legal when every level size is a multiple of 16 bytes worth of elements and all pointers are 16-byte
// aligned (the caller checked: vec == 16 / sizeof(T)).
cudaError_t cerb_launch_decode_tma(const DecodeParams& D, int dtype, cudaStream_t stream) {
    return dtype == CERB_DTYPE_F16 ? launch_tma_t<__half>(D, stream) : launch_tma_t<float>(D, stream);
}
