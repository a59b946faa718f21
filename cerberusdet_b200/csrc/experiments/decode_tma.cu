// Fused Detect-head decode, persistent TMA-pipelined variant.  OPT-IN (CERB_DEBUG_DECODE_TMA=1, tools/ only): correct,
// but measured slower on B200 than the cp.async-pipelined kernel in decode_pipe.cu, which is the default
// (profiles/r01_decode.md has the numbers); kept as the measured alternative.
//
// Why: the register-resident kernel in decode.cu is latency bound -- a warp first waits for its 16
// loads, then spends thousands of cycles on exp/rcp work with nothing in flight, and at ~100-146
// registers only 12-20 warps fit on an SM, so on average ~20 KB per SM are in flight where HBM3e
// needs ~45 KB (profiles/r01_decode.md).  Here one producer warp per SM streams [64+nc] x TA
// tiles into a shared-memory ring with cp.async.bulk (TMA, mbarrier complete_tx) several tiles ahead
// of the 15 consumer warps, which only ever read shared memory.
//
// Same arithmetic and rounding points as decode.cu (reference models/yolo.py:93-99, utils/tal.py).
//
// Tile = TA consecutive anchors of one (task, level, image), all 64+nc channels: TA = 256 (fp16) or
// 128 (fp32), i.e. 512-byte rows, one bulk copy per channel row.  A consumer warp owns one role of one
// tile -- lane = 16 bytes of anchors: role 0 = DFL sides l,r -> (cx, w); role 1 = sides t,b -> (cy, h);
// role 2 = class sigmoids + score summary.  The 15 consumer warps form 5 groups of 3 roles; group g
// takes the CTA's tiles k = g, g+5, ...; warps never synchronise with each other, only with the
// stage's full/empty mbarriers, so light roles run ahead inside the ring.
#include <cuda.h>  // CUtensorMap types only; the encoder is fetched through cudaGetDriverEntryPoint

#include "cerb_kernels.h"

#define TMA_ROLES 3
#define TMA_GROUPS 5
#define TMA_CONSUMER_WARPS (TMA_ROLES * TMA_GROUPS)
#define TMA_THREADS ((TMA_CONSUMER_WARPS + 1) * 32)
#define TMA_MAX_STAGES 8
#define LOG2E_F 1.4426950408889634f

struct TmaDecodeParams {
    CUtensorMap maps[CERB_MAX_TASKS * CERB_MAX_LEVELS];  // one 2-D map per (task, level): [B*no rows, hw cols]
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
// one TMA op moves the whole [no x TA] box of a tile (out-of-range columns are zero-filled and counted)
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
template <typename T> union Vec16 {
    uint4 raw;
    T e[16 / sizeof(T)];
};

// Expected DFL distance of one box side for the V anchors of this lane, read from the staged tile.
template <typename T>
__device__ __forceinline__ void dfl_side_smem(const unsigned char* side_base, int row_bytes, float (&d)[16 / sizeof(T)]) {
    constexpr int V = 16 / sizeof(T);
    Vec16<T> v[CERB_REG_MAX];
#pragma unroll
    for (int k = 0; k < CERB_REG_MAX; ++k) v[k].raw = *reinterpret_cast<const uint4*>(side_base + (size_t)k * row_bytes);
#pragma unroll
    for (int i = 0; i < V; ++i) {
        float x[CERB_REG_MAX];
#pragma unroll
        for (int k = 0; k < CERB_REG_MAX; ++k) x[k] = to_f32<T>(v[k].e[i]);
        d[i] = dfl_expectation<T>(x);
    }
}

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
    constexpr int V = 16 / sizeof(T);  // anchors per lane
    constexpr int TA = 32 * V;         // anchors per tile
    constexpr int ROWB = TA * (int)sizeof(T);
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[2 * TMA_MAX_STAGES];

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int stages = P.stages;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);                            // full: producer's expect_tx arrive
            mbar_init(smem_u32(&bars[TMA_MAX_STAGES + s]), TMA_ROLES);   // empty: one arrive per role warp
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
            const uint32_t full = smem_u32(&bars[s]);
            if (lane == 0) {
                mbar_arrive_expect_tx(full, (uint32_t)(no * ROWB));
                tma_load_2d(smem_u32(smem + (size_t)s * P.stage_bytes), &P.maps[ti.task * P.d.L + ti.level], ti.a0,
                            ti.b * no, full);
            }
            __syncwarp();
        }
        return;
    }

    // ---------------------------------------------------- consumers
    const int group = wid / TMA_ROLES, role = wid - group * TMA_ROLES;
    for (int k = group, t = blockIdx.x + group * gridDim.x; t < P.ntiles; k += TMA_GROUPS, t += TMA_GROUPS * gridDim.x) {
        const int s = k % stages;
        const uint32_t use = (uint32_t)(k / stages);
        const TileInfo ti = locate_tile(P, t, TA);
        const int nc = P.d.nc[ti.task];
        const int A = P.d.A;
        const int al = lane * V;        // first anchor of this lane inside the tile
        const bool live = al < ti.cnt;  // cnt is a multiple of V: a lane is entirely live or not
        const unsigned char* st = smem + (size_t)s * P.stage_bytes + (size_t)lane * 16;
        T* __restrict__ out = reinterpret_cast<T*>(P.d.y[ti.task]) + (size_t)ti.b * (4 + nc) * A + P.d.aoff[ti.level] + ti.a0 + al;

        mbar_wait(smem_u32(&bars[s]), use & 1u);
        if (role < 2) {
            // sides (l, r) -> cx, w   or   (t, b) -> cy, h      (utils/tal.py:198-204)
            float dlo[V], dhi[V];
            if (live) {
                dfl_side_smem<T>(st + (size_t)(role * CERB_REG_MAX) * ROWB, ROWB, dlo);
                dfl_side_smem<T>(st + (size_t)((role + 2) * CERB_REG_MAX) * ROWB, ROWB, dhi);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[TMA_MAX_STAGES + s]));  // stage no longer needed by this warp
            if (live) {
                const int W = P.d.w[ti.level];
                const float stride = P.d.stride[ti.level];
                Vec16<T> oc, os;
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    const int a = ti.a0 + al + i;
                    const int g = (role == 0) ? (a % W) : (a / W);
                    const float ac = rnd<T>(rnd<T>((float)g) + 0.5f);  // arange(dtype) + 0.5, tal.py:188-189
                    const float p1 = rnd<T>(ac - dlo[i]);
                    const float p2 = rnd<T>(ac + dhi[i]);
                    const float c = rnd<T>(rnd<T>(p1 + p2) * 0.5f);
                    const float sz = rnd<T>(p2 - p1);
                    oc.e[i] = from_f32<T>(c * stride);
                    os.e[i] = from_f32<T>(sz * stride);
                }
                stg_stream16(out + (size_t)role * A, oc.raw);
                stg_stream16(out + (size_t)(role + 2) * A, os.raw);
            }
        } else {
            // class sigmoids (yolo.py:99) + score summary (max of each 16-byte score vector)
            const size_t srow = ((size_t)(A / V) + V - 1) / V * V;
            T* smx = nullptr;
            if (P.d.smax[ti.task] != nullptr)
                smx = reinterpret_cast<T*>(P.d.smax[ti.task]) + (size_t)ti.b * nc * srow + (P.d.aoff[ti.level] + ti.a0 + al) / V;
            if (live) {
                const unsigned char* cls = st + (size_t)(4 * CERB_REG_MAX) * ROWB;
                T* cout = out + (size_t)4 * A;
                int c = 0;
                for (; c + 4 <= nc; c += 4) {
                    Vec16<T> vv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) vv[u].raw = *reinterpret_cast<const uint4*>(cls + (size_t)(c + u) * ROWB);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float mx = -INFINITY;
#pragma unroll
                        for (int i = 0; i < V; ++i) {
                            const float x = to_f32<T>(vv[u].e[i]);
                            vv[u].e[i] = from_f32<T>(fast_rcp(1.f + fast_ex2(-x * LOG2E_F)));
                            mx = fmaxf(mx, to_f32<T>(vv[u].e[i]));
                        }
                        stg_stream16(cout + (size_t)(c + u) * A, vv[u].raw);
                        if (smx != nullptr) smx[(size_t)(c + u) * srow] = from_f32<T>(mx);
                    }
                }
                for (; c < nc; ++c) {
                    Vec16<T> v1;
                    v1.raw = *reinterpret_cast<const uint4*>(cls + (size_t)c * ROWB);
                    float mx = -INFINITY;
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        const float x = to_f32<T>(v1.e[i]);
                        v1.e[i] = from_f32<T>(fast_rcp(1.f + fast_ex2(-x * LOG2E_F)));
                        mx = fmaxf(mx, to_f32<T>(v1.e[i]));
                    }
                    stg_stream16(cout + (size_t)c * A, v1.raw);
                    if (smx != nullptr) smx[(size_t)c * srow] = from_f32<T>(mx);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[TMA_MAX_STAGES + s]));
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        (void)cudaGetLastError();
    }
    return fn;
}

template <typename T> static cudaError_t launch_tma_t(const DecodeParams& D, cudaStream_t stream) {
    constexpr int TA = 32 * (16 / (int)sizeof(T));
    EncodeTiledFn encode = tensor_map_encoder();
    if (!encode) return cudaErrorInvalidConfiguration;
    TmaDecodeParams P;
    P.d = D;
    int tiles = 0, no_max = 0;
    for (int l = 0; l < D.L; ++l) P.tiles_per_image[l] = (D.hw[l] + TA - 1) / TA;
    for (int t = 0; t < D.T; ++t) {
        const int no = 4 * CERB_REG_MAX + D.nc[t];
        if (no > 256) return cudaErrorInvalidConfiguration;  // TMA box rows <= 256: wide heads use the generic kernel
        no_max = max(no_max, no);
        for (int l = 0; l < D.L; ++l) {
            P.row_tile_start[t * D.L + l] = tiles;
            tiles += D.B * P.tiles_per_image[l];
            const cuuint64_t gdim[2] = {(cuuint64_t)D.hw[l], (cuuint64_t)D.B * (cuuint64_t)no};
            const cuuint64_t gstride[1] = {(cuuint64_t)D.hw[l] * sizeof(T)};
            const cuuint32_t box[2] = {(cuuint32_t)TA, (cuuint32_t)no};
            const cuuint32_t estr[2] = {1, 1};
            const CUresult r = encode(&P.maps[t * D.L + l],
                                      sizeof(T) == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                                      const_cast<void*>(D.lvl[t][l]), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return cudaErrorInvalidConfiguration;
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
    const int static_smem = 2 * TMA_MAX_STAGES * 8 + 1024;
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
