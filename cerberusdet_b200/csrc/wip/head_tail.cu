// DRAFT for round 2 -- NOT built by build.sh, NOT part of libcerb_post.so, NEVER run on a GPU yet.
// It compiles for sm_100a (nvcc -c, see the bottom of this comment); nothing else about it is verified.
// Its checker already exists and is pinned to the reference: oracle/ref_port.head_tail_port +
// tests/golden/headtail_*.npz (oracle/gen_golden_headtail.py).
//
// SURVEY 8f row 3, head-tail fusion: the LAST 1x1 convolutions of the two towers of one level
//     box = cv2[l][-1](u2)   [B, c2, H, W] -> [B, 64, H, W]      reference models/yolo.py:81-84, 89-90
//     cls = cv3[l][-1](u3)   [B, c3, H, W] -> [B, nc, H, W]
// fused with the eval decode (models/yolo.py:93-99), so that the raw head tensor [B, 64+nc, H, W] is never written
// or re-read (522 MB of the ~1.9 GB the two steps move per config-3 batch).
//
// Shape of the problem.  Per image the activations are [K = channels][anchors] with the anchors contiguous, i.e.
// MN-major for either operand position.  One CTA = one tile of 128 consecutive anchors of one image:
//     D1[128 anchors x 64]      = U2^T[128 x c2] * W2^T[c2 x 64]        tcgen05.mma kind::f16, M = 128, N = 64
//     D2[128 anchors x NCP]     = U3^T[128 x c3] * W3^T[c3 x NCP]       NCP = nc rounded up to 16 (<= 256)
// A = activations, MN-major, SWIZZLE_128B: TMA boxes {64 anchors (inner, 128 B), 64 channels} land as rows of 128 B,
//     8 rows = one 1024-byte swizzle atom; canonical layout ((8,8,m),(8,k)):((1,8,LBO),(64,SBO)) in elements
//     (cute/atom/mma_traits_sm100.hpp): SBO = 1024 B (next 8 channels), LBO = bytes between the two 64-anchor halves.
//     One MMA consumes K = 16 channels = 2 atoms: the start address advances by 2048 B per k-step.
// B = weights [N][K] row-major = K-major; packed by the CTA into the K-major SWIZZLE_128B image (K blocks of 64
//     elements, row n at n*128 B, 16-byte chunk c stored at chunk c ^ (n & 7)); SBO = 1024 B; a k-step advances the
//     start address by 32 B inside the swizzled row.
// Accumulators: TMEM, M = 128 -> lane i = anchor i, column j = output channel j.  tcgen05.ld.32x32b hands every
//     epilogue thread (4 warps, lane quadrant = warp % 4) the 64 DFL logits of its own anchor, then the class logits:
//     the epilogue is the per-anchor decode math of decode.cu with the bias add and the half rounding of the
//     reference's conv output in front.
//
// Warp roles (192 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocation + MMA issue (one
// elected lane), warps 2-5 = epilogue.  First version: the whole K extent of a tile sits in shared memory
// ((c2 + c3) * 256 B + weights; yolov8x: 100 KB + 36 KB -> 1 CTA / SM); round 2 adds a K ring so 2-3 CTAs fit.
//
// Compile check:  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -I.. -c wip/head_tail.cu -o /dev/null
#include <cuda.h>

#include "../decode_common.cuh"

#define HT_TILE 128          // anchors per CTA = UMMA M
#define HT_THREADS 192
#define HT_KBLOCK 64         // channels per TMA box / per K block of the weight image
#define HT_TMEM_COLS 512     // 64 (D1) + up to 256 (D2), rounded to a power of two >= 32

struct HeadTailParams {
    CUtensorMap map_u2;  // [B * c2 rows, hw cols] fp16, box {64, HT_KBLOCK}, SWIZZLE_128B
    CUtensorMap map_u3;  // [B * c3 rows, hw cols]
    const __half* w2;    // [64, c2] (cv2[l][-1].weight squeezed)
    const __half* b2;    // [64]
    const __half* w3;    // [nc, c3]
    const __half* b3;    // [nc]
    __half* y;           // [B, 4 + nc, A]
    __half* smax;        // optional score summary [B, nc, srow]
    int B, hw, W, A, aoff, nc, c2, c3;
    float stride;
};

__device__ __forceinline__ uint32_t ht_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ht_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void ht_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ht_mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void ht_tma_load_2d(uint32_t dst, const CUtensorMap* map, int col, int row, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(col), "r"(row), "r"(bar)
        : "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp): start >> 4 in bits [0,14),
// LBO >> 4 in [16,30), SBO >> 4 in [32,46), version 1 in [46,48), layout type in [61,64) (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t ht_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 at [4,6), a/b format F16 = 0 at [7,10) / [10,13),
// a_major (1 = MN-major) at 15, b_major (0 = K-major) at 16, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t ht_instr_desc(int M, int N) {
    return (1u << 4) | (1u << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void ht_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void ht_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 16 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void ht_tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// weights [N_real][K] row-major (global) -> K-major SWIZZLE_128B image: K block kb at kb * n_rows * 128 B, row n at
// n * 128 B, 16-byte chunk c (8 elements) at ((c ^ (n & 7)) * 16); rows >= N_real and columns >= K are zero
__device__ __forceinline__ void ht_pack_weights(unsigned char* dst, const __half* __restrict__ w, int n_real, int n_rows, int K,
                                                int tid, int nthreads) {
    const int kblocks = (K + HT_KBLOCK - 1) / HT_KBLOCK;
    const int chunks = kblocks * n_rows * 8;
    for (int i = tid; i < chunks; i += nthreads) {
        const int kb = i / (n_rows * 8), rem = i - kb * (n_rows * 8);
        const int n = rem >> 3, c = rem & 7;
        const int k0 = kb * HT_KBLOCK + c * 8;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (n < n_real && k0 + 8 <= K) {
            v = *reinterpret_cast<const uint4*>(w + (size_t)n * K + k0);  // K % 8 == 0 and 16-byte aligned rows (host checks)
        }
        *reinterpret_cast<uint4*>(dst + (size_t)kb * n_rows * 128 + n * 128 + ((c ^ (n & 7)) << 4)) = v;
    }
}

__global__ void __launch_bounds__(HT_THREADS, 1) head_tail_kernel(const __grid_constant__ HeadTailParams P) {
    extern __shared__ __align__(1024) unsigned char ht_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_per_image = (P.hw + HT_TILE - 1) / HT_TILE;
    const int b = blockIdx.x / tiles_per_image;
    const int a_tile = (blockIdx.x - b * tiles_per_image) * HT_TILE;  // first anchor of the tile inside the level

    const int ncp = (P.nc + 15) & ~15;                                // N of the class GEMM
    const int kb2 = (P.c2 + HT_KBLOCK - 1) / HT_KBLOCK, kb3 = (P.c3 + HT_KBLOCK - 1) / HT_KBLOCK;
    // shared-memory map (every region a multiple of 1024 B)
    const uint32_t a2_half = (uint32_t)kb2 * HT_KBLOCK * 128;         // bytes of one 64-anchor half of U2's tile
    const uint32_t a3_half = (uint32_t)kb3 * HT_KBLOCK * 128;
    unsigned char* sA2 = ht_smem;
    unsigned char* sA3 = sA2 + 2 * a2_half;
    unsigned char* sW2 = sA3 + 2 * a3_half;
    unsigned char* sW3 = sW2 + (size_t)kb2 * 64 * 128;
    const uint32_t w3_rows = (uint32_t)((ncp + 7) & ~7);
    unsigned char* sEnd = sW3 + (size_t)kb3 * w3_rows * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sEnd);               // [0] U2 landed, [1] U3 landed, [2] D1 ready, [3] D2 ready
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) ht_mbar_init(ht_smem_u32(bars + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: one warp allocates, the address lands in shared memory
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ht_smem_u32(tmem_slot)),
                     "n"(HT_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // weights -> swizzled K-major images (generic-proxy stores, made visible to the tensor core's async proxy below)
    ht_pack_weights(sW2, P.w2, 64, 64, P.c2, tid, HT_THREADS);
    ht_pack_weights(sW3, P.w3, P.nc, (int)w3_rows, P.c3, tid, HT_THREADS);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---- TMA producer: boxes {64 anchors, 64 channels}; anchors past hw are zero-filled by TMA, channel rows
            // past c2 / c3 belong to the next image (or are zero-filled at the very end) and are never multiplied
            ht_mbar_expect_tx(ht_smem_u32(bars + 0), 2u * a2_half);
            for (int h = 0; h < 2; ++h)
                for (int kb = 0; kb < kb2; ++kb)
                    ht_tma_load_2d(ht_smem_u32(sA2 + h * a2_half + kb * HT_KBLOCK * 128), &P.map_u2, a_tile + h * 64,
                                   b * P.c2 + kb * HT_KBLOCK, ht_smem_u32(bars + 0));
            ht_mbar_expect_tx(ht_smem_u32(bars + 1), 2u * a3_half);
            for (int h = 0; h < 2; ++h)
                for (int kb = 0; kb < kb3; ++kb)
                    ht_tma_load_2d(ht_smem_u32(sA3 + h * a3_half + kb * HT_KBLOCK * 128), &P.map_u3, a_tile + h * 64,
                                   b * P.c3 + kb * HT_KBLOCK, ht_smem_u32(bars + 1));
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---- MMA issue: D1 in TMEM columns [0, 64), D2 in [64, 64 + ncp)
            ht_mbar_wait(ht_smem_u32(bars + 0), 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t id1 = ht_instr_desc(HT_TILE, 64);
            for (int k = 0; k < P.c2 / 16; ++k) {
                const uint64_t ad = ht_smem_desc(ht_smem_u32(sA2) + k * 2048, a2_half, 1024);
                const uint64_t bd = ht_smem_desc(ht_smem_u32(sW2) + (k >> 2) * 64 * 128 + (k & 3) * 32, 0, 1024);
                ht_mma(tmem, ad, bd, id1, k > 0);
            }
            ht_commit(ht_smem_u32(bars + 2));
            ht_mbar_wait(ht_smem_u32(bars + 1), 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t id2 = ht_instr_desc(HT_TILE, ncp);
            for (int k = 0; k < P.c3 / 16; ++k) {
                const uint64_t ad = ht_smem_desc(ht_smem_u32(sA3) + k * 2048, a3_half, 1024);
                const uint64_t bd = ht_smem_desc(ht_smem_u32(sW3) + (k >> 2) * w3_rows * 128 + (k & 3) * 32, 0, 1024);
                ht_mma(tmem + 64, ad, bd, id2, k > 0);
            }
            ht_commit(ht_smem_u32(bars + 3));
        }
    } else {
        // ---- epilogue: thread = TMEM lane = anchor.  A warp may only read the lane quadrant warp % 4.
        const int q = warp & 3;
        const int a = a_tile + q * 32 + lane;            // anchor inside the level
        const bool live = a < P.hw;
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        __half* __restrict__ yb = P.y + (size_t)b * (4 + P.nc) * P.A + P.aoff + a;
        ht_mbar_wait(ht_smem_u32(bars + 2), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float d[4];
#pragma unroll
        for (int side = 0; side < 4; ++side) {
            float x[CERB_REG_MAX];
            ht_tmem_ld16(trow + side * 16, x);
#pragma unroll
            for (int k = 0; k < CERB_REG_MAX; ++k)       // conv output = half(acc + bias), like the reference's half conv
                x[k] = rnd<__half>(x[k] + __half2float(P.b2[side * 16 + k]));
            d[side] = dfl_expectation<__half>(x);
        }
        if (live) {
            const int gx = a % P.W, gy = a / P.W;
#pragma unroll
            for (int axis = 0; axis < 2; ++axis) {       // dist2bbox(xywh) per axis, * stride (tal.py:198-204, yolo.py:98)
                const float ac = rnd<__half>(rnd<__half>((float)(axis == 0 ? gx : gy)) + 0.5f);
                const float p1 = rnd<__half>(ac - d[axis]);
                const float p2 = rnd<__half>(ac + d[axis + 2]);
                const float c = rnd<__half>(rnd<__half>(p1 + p2) * 0.5f);
                const float sz = rnd<__half>(p2 - p1);
                yb[(size_t)axis * P.A] = from_f32<__half>(c * P.stride);
                yb[(size_t)(axis + 2) * P.A] = from_f32<__half>(sz * P.stride);
            }
        }
        ht_mbar_wait(ht_smem_u32(bars + 3), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const size_t srow = ((size_t)(P.A / 8) + 7) / 8 * 8;
        for (int c0 = 0; c0 < ncp; c0 += 16) {
            float x[16];
            ht_tmem_ld16(trow + 64 + c0, x);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int c = c0 + k;
                if (c < P.nc) {                          // (warp-uniform)
                    const float logit = rnd<__half>(x[k] + __half2float(P.b3[c]));
                    const __half s = from_f32<__half>(fast_rcp(1.f + fast_ex2(-logit * LOG2E_F)));
                    if (live) yb[(size_t)(4 + c) * P.A] = s;
                    if (P.smax != nullptr) {             // maximum over the 8 anchors of a 16-byte score vector
                        float m = live ? __half2float(s) : 0.f;
                        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
                        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
                        if ((lane & 7) == 0 && live)
                            P.smax[(size_t)b * P.nc * srow + (size_t)c * srow + (P.aoff + a) / 8] = __float2half_rn(m);
                    }
                }
            }
        }
    }
    // ---- teardown: every TMEM read is done before the allocation is released
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(HT_TMEM_COLS));
    }
}

// Host side (round 2): encode map_u2 / map_u3 with cuTensorMapEncodeTiled (2-D, FLOAT16, gdim {hw, B*c},
// gstride {hw*2}, box {64, HT_KBLOCK}, CU_TENSOR_MAP_SWIZZLE_128B, OOB fill NONE = zeros), require c2 % 16 == 0,
// c3 % 16 == 0, hw % 8 == 0 (16-byte global strides), dynamic shared memory =
// 2*(kb2+kb3)*64*128 + kb2*64*128 + kb3*w3_rows*128 + 64 bytes, cudaFuncAttributeMaxDynamicSharedMemorySize,
// grid = B * ceil(hw / 128), one launch per (task, level).  TMEM: 512 columns per CTA -> exactly one CTA per SM may hold
// an allocation; shrink HT_TMEM_COLS to 128 when ncp <= 64 so that several CTAs can share an SM.

// ------------------------------------------------------------------ host side of the draft (one launch per (task, level))
#include <stdio.h>

typedef CUresult (*HtEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int ht_encode(HtEncodeTiledFn enc, CUtensorMap* map, const void* base, int rows, int hw) {
    const cuuint64_t gdim[2] = {(cuuint64_t)hw, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)hw * 2};
    const cuuint32_t box[2] = {64, HT_KBLOCK};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
               ? 0
               : -1;
}

// fp16 only.  u2 [B, c2, H, W], u3 [B, c3, H, W], w2 [64, c2], b2 [64], w3 [nc, c3], b3 [nc]; y [B, 4+nc, A] (this level's
// anchors start at aoff), smax optional.  Returns 0, or a negative code with a message on stderr (draft: no error plumbing).
extern "C" int cerb_wip_head_tail(const void* u2, const void* u3, const void* w2, const void* b2, const void* w3, const void* b3,
                                  int B, int c2, int c3, int nc, int H, int W, float stride, int A, int aoff, void* y,
                                  void* smax, void* stream) {
    const int hw = H * W;
    if (c2 % 16 || c3 % 16 || hw % 8 || nc < 1 || nc > 256) {
        fprintf(stderr, "cerb_wip_head_tail: need c2 %% 16 == 0, c3 %% 16 == 0, H*W %% 8 == 0, 1 <= nc <= 256\n");
        return -1;
    }
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        return -2;
    HtEncodeTiledFn enc = reinterpret_cast<HtEncodeTiledFn>(p);
    HeadTailParams P;
    if (ht_encode(enc, &P.map_u2, u2, B * c2, hw) || ht_encode(enc, &P.map_u3, u3, B * c3, hw)) return -3;
    P.w2 = (const __half*)w2; P.b2 = (const __half*)b2; P.w3 = (const __half*)w3; P.b3 = (const __half*)b3;
    P.y = (__half*)y; P.smax = (__half*)smax;
    P.B = B; P.hw = hw; P.W = W; P.A = A; P.aoff = aoff; P.nc = nc; P.c2 = c2; P.c3 = c3; P.stride = stride;
    const int ncp = (nc + 15) & ~15, kb2 = (c2 + HT_KBLOCK - 1) / HT_KBLOCK, kb3 = (c3 + HT_KBLOCK - 1) / HT_KBLOCK;
    const int w3_rows = (ncp + 7) & ~7;
    const size_t smem = (size_t)2 * (kb2 + kb3) * HT_KBLOCK * 128 + (size_t)kb2 * 64 * 128 + (size_t)kb3 * w3_rows * 128 + 64;
    int dev = 0, smem_max = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (smem > (size_t)smem_max) {
        fprintf(stderr, "cerb_wip_head_tail: tile needs %zu B of shared memory (K ring not written yet)\n", smem);
        return -4;
    }
    if (cudaFuncSetAttribute(head_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -5;
    const int tiles = B * ((hw + HT_TILE - 1) / HT_TILE);
    head_tail_kernel<<<tiles, HT_THREADS, smem, (cudaStream_t)stream>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : -6;
}
