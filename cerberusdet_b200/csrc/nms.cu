// Per-task confidence filter, candidate expansion, lazy top-k ordering and class-offset
// greedy NMS -- one CTA per (task, image) segment, all segments in ONE launch.
//
// Replaces reference non_max_suppression (utils/general.py:360-481) incl. xywh2xyxy
// (:272-288) and the third-party torchvision.ops.nms it calls (:464).  The reference
// runs ~25 ATen launches and ~4 host syncs per (image, task) and materialises every
// candidate; this kernel never materialises the candidate table:
//
//  * Candidates are implicit: (anchor a, class c) pairs of the prediction whose score
//    passes the threshold (multi-label, :444-446) or the best class of an anchor
//    (:447-449).  Each gets a unique 64-bit key  score_bits << 32 | ~(a*nc + c), so that
//    descending key order == "score descending, candidate index ascending", the
//    canonical form of the reference's sort at :459 (ties there are unspecified).
//  * Greedy NMS is order-causal and stops at max_det (:465), so only a prefix of the
//    sorted order is ever needed.  A 4096-bin histogram of the top 12 key bits (one pass
//    over the scores, in shared memory) locates successive chunks of <= CAP keys; each
//    chunk is gathered, bitonic-sorted in shared memory and consumed; the kernel stops
//    as soon as max_det boxes are kept or max_nms candidates (:459) were consumed.
//    A bin that alone exceeds CAP (massive ties) is refined by deeper radix levels.
//  * Suppression (:462-465): candidates are taken in tiles of 256; a tile is first
//    tested against the kept list (shared memory), then resolved internally with a
//    256x256 IoU bitmask and a ballot fixpoint iteration (equal to the sequential sweep).  IoU arithmetic reproduces
//    torchvision's CPU kernel: separately rounded fp32 ops on class-offset boxes, strict
//    '>' against the threshold (the double-precision compare is folded into iou_thr).
#include "cerb_kernels.h"

#define NMS_THREADS 512
#define NMS_TILE 256
#define NMS_TILE_WORDS (NMS_TILE / 64)
#define NMS_CAP 4096         // max keys sorted at once
#define NMS_BINS 4096        // 12-bit radix digits
#define NMS_KEPT_SMEM 1024   // kept list lives in smem up to this max_det

typedef unsigned long long u64;

struct __align__(16) NmsSmem {
    unsigned g0[NMS_BINS + 8];   // level-0: G0[d] = #keys with digit >= d ; G0[4096] = 0
    unsigned g1[NMS_BINS + 8];   // scratch for deeper levels
    u64 keys[NMS_CAP];
    float4 tbox[NMS_TILE];       // class-offset corners of the tile's candidates
    float4 traw[NMS_TILE];       // un-offset corners (output)
    float tarea[NMS_TILE];
    float tscore[NMS_TILE];
    int tcls[NMS_TILE];
    int tcode[NMS_TILE];         // class id if the box is "tame" (see class shortcut), else -1
    u64 mask[NMS_TILE][NMS_TILE_WORDS];
    u64 keepmask[NMS_TILE_WORDS];
    unsigned keep32[2][NMS_TILE / 32];
    unsigned char tdead[NMS_TILE];
    unsigned warp_tot[NMS_THREADS / 32];
    unsigned counter;
    float4 kbox[NMS_KEPT_SMEM];
    float karea[NMS_KEPT_SMEM];
    int kcode[NMS_KEPT_SMEM];
};

// ---- IoU test exactly as torchvision's CPU kernel does it (std::max/std::min semantics,
// one rounding per operation, no FMA).  box i is the earlier (kept) box.
__device__ __forceinline__ bool suppresses(const float4 bi, const float ai, const float4 bj, const float aj,
                                           const float thr) {
    const float xx1 = (bi.x < bj.x) ? bj.x : bi.x;
    const float yy1 = (bi.y < bj.y) ? bj.y : bi.y;
    const float xx2 = (bj.z < bi.z) ? bj.z : bi.z;
    const float yy2 = (bj.w < bi.w) ? bj.w : bi.w;
    const float dw = __fsub_rn(xx2, xx1), dh = __fsub_rn(yy2, yy1);
    const float w = (0.f < dw) ? dw : 0.f;
    const float h = (0.f < dh) ? dh : 0.f;
    const float inter = __fmul_rn(w, h);
    if (!(inter > 0.f)) return false;  // quotient would be 0 (or NaN): never > thr >= 0
    const float uni = __fsub_rn(__fadd_rn(ai, aj), inter);
    return __fdiv_rn(inter, uni) > thr;
}

// ---- enumerate the candidates of one segment; f(score_bits, anchor, class) is called once per
// candidate.  Scores are read with 128-bit loads (the class planes of one image are one contiguous
// [nc*A] array) whenever A is a multiple of the vector width and the base is 16-byte aligned.
__device__ __forceinline__ unsigned make_sbits(float s) { return __float_as_uint(s); }
__device__ __forceinline__ u64 make_key(unsigned sbits, unsigned idx) {
    return ((u64)sbits << 32) | (u64)(0xFFFFFFFFu - idx);
}

template <typename T> struct ScoreVec {
    static constexpr int V = 16 / sizeof(T);
    union { uint4 raw; T e[16 / sizeof(T)]; };
};

#define SCAN_UNROLL 4

template <typename T, bool MULTI, typename F>
__device__ __forceinline__ void for_each_candidate(const T* __restrict__ img, int nc, int A, float thr,
                                                   const NmsParams& P, F f, const unsigned vstride = 1) {
    // vstride > 1 visits only every vstride-th 16-byte vector (multi-label vector path only): used for the
    // *estimating* histogram; every exact pass uses vstride == 1.
    constexpr int V = ScoreVec<T>::V;
    const T* __restrict__ sc = img + (size_t)4 * A;
    const bool vec_ok = (A % V == 0) && ((reinterpret_cast<uintptr_t>(sc) & 15) == 0);
    const bool filt = P.use_class_filter != 0;
    if (MULTI) {
        if (vec_ok) {
            const unsigned nvec = (unsigned)(((u64)nc * (u64)A) / V);
            const unsigned nvis = (nvec + vstride - 1) / vstride;  // vectors visited
            for (unsigned q0 = threadIdx.x; q0 < nvis; q0 += NMS_THREADS * SCAN_UNROLL) {
                ScoreVec<T> v[SCAN_UNROLL];
#pragma unroll
                for (int u = 0; u < SCAN_UNROLL; ++u) {
                    const unsigned q = q0 + u * NMS_THREADS;
                    if (q < nvis) v[u].raw = __ldg(reinterpret_cast<const uint4*>(sc) + (size_t)q * vstride);
                }
#pragma unroll
                for (int u = 0; u < SCAN_UNROLL; ++u) {
                    const unsigned q = q0 + u * NMS_THREADS;
                    if (q >= nvis) break;
                    const unsigned i = q * vstride;
                    const unsigned e0 = i * V;
                    const unsigned c = e0 / (unsigned)A;
                    const unsigned a0 = e0 - c * (unsigned)A;
                    if (filt && !((P.class_mask[c >> 5] >> (c & 31)) & 1u)) continue;
#pragma unroll
                    for (int k = 0; k < V; ++k) {
                        const float s = to_f32<T>(v[u].e[k]);
                        if (s > thr) f(make_sbits(s), (int)(a0 + k), (int)c);
                    }
                }
            }
        } else {
            for (int c = 0; c < nc; ++c) {
                if (filt && !((P.class_mask[c >> 5] >> (c & 31)) & 1u)) continue;
                const T* __restrict__ row = sc + (size_t)c * A;
                for (int a = threadIdx.x; a < A; a += NMS_THREADS) {
                    const float s = to_f32<T>(__ldg(row + a));
                    if (s > thr) f(make_sbits(s), a, c);
                }
            }
        }
    } else {
        if (vec_ok) {
            const int nv = A / V;
            for (int i = threadIdx.x; i < nv; i += NMS_THREADS) {
                const uint4* __restrict__ col = reinterpret_cast<const uint4*>(sc) + i;
                float best[V];
                int bc[V];
                {
                    ScoreVec<T> v0;
                    v0.raw = __ldg(col);
#pragma unroll
                    for (int k = 0; k < V; ++k) { best[k] = to_f32<T>(v0.e[k]); bc[k] = 0; }
                }
                int c = 1;
                for (; c + SCAN_UNROLL <= nc; c += SCAN_UNROLL) {
                    ScoreVec<T> v[SCAN_UNROLL];
#pragma unroll
                    for (int u = 0; u < SCAN_UNROLL; ++u) v[u].raw = __ldg(col + (size_t)(c + u) * nv);
#pragma unroll
                    for (int u = 0; u < SCAN_UNROLL; ++u)
#pragma unroll
                        for (int k = 0; k < V; ++k) {
                            const float s = to_f32<T>(v[u].e[k]);
                            if (s > best[k]) { best[k] = s; bc[k] = c + u; }  // lowest index wins ties (torch.max)
                        }
                }
                for (; c < nc; ++c) {
                    ScoreVec<T> v1;
                    v1.raw = __ldg(col + (size_t)c * nv);
#pragma unroll
                    for (int k = 0; k < V; ++k) {
                        const float s = to_f32<T>(v1.e[k]);
                        if (s > best[k]) { best[k] = s; bc[k] = c; }
                    }
                }
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    if (best[k] > thr) {
                        if (filt && !((P.class_mask[bc[k] >> 5] >> (bc[k] & 31)) & 1u)) continue;
                        f(make_sbits(best[k]), i * V + k, bc[k]);
                    }
                }
            }
        } else {
            for (int a = threadIdx.x; a < A; a += NMS_THREADS) {
                float best = to_f32<T>(__ldg(sc + a));
                int bc = 0;
                for (int c = 1; c < nc; ++c) {
                    const float s = to_f32<T>(__ldg(sc + (size_t)c * A + a));
                    if (s > best) { best = s; bc = c; }  // lowest index wins ties (torch.max)
                }
                if (best > thr) {
                    if (filt && !((P.class_mask[bc >> 5] >> (bc & 31)) & 1u)) continue;
                    f(make_sbits(best), a, bc);
                }
            }
        }
    }
}

__device__ __forceinline__ int level_shift(int lvl) { return lvl < 5 ? 52 - 12 * lvl : 0; }
__device__ __forceinline__ unsigned level_digit(u64 key, int lvl) {
    return lvl < 5 ? (unsigned)(key >> (52 - 12 * lvl)) & 0xFFFu : (unsigned)key & 0xFu;
}

// in-place: h[d] (counts) -> G[d] = sum_{d' >= d} h[d'], G[NMS_BINS] = 0.  All threads call.
__device__ void suffix_scan(unsigned* h, unsigned* warp_tot) {
    constexpr int PER = NMS_BINS / NMS_THREADS;  // 8
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    // thread t owns reversed positions r in [PER*t, PER*t+PER), bin = NMS_BINS-1-r
    unsigned v[PER], sum = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) { v[i] = h[NMS_BINS - 1 - (PER * t + i)]; sum += v[i]; }
    unsigned inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    unsigned base = 0;
    for (int w = 0; w < wid; ++w) base += warp_tot[w];
    unsigned run = base + inc - sum;
#pragma unroll
    for (int i = 0; i < PER; ++i) { run += v[i]; h[NMS_BINS - 1 - (PER * t + i)] = run; }
    if (t == 0) h[NMS_BINS] = 0;
    __syncthreads();
}

// largest digit d in [0, hi) with G[d] - base >= need   (G non-increasing in d; caller guarantees G[0]-base >= need)
__device__ __forceinline__ int find_digit(const unsigned* G, int hi, unsigned base, unsigned need) {
    int lo = 0, up = hi;  // invariant: G[lo]-base >= need ; (up == hi or G[up]-base < need)
    while (up - lo > 1) {
        const int mid = (lo + up) >> 1;
        if (G[mid] - base >= need) lo = mid; else up = mid;
    }
    return lo;
}

template <typename T, bool MULTI>
__global__ void __launch_bounds__(NMS_THREADS, 2) nms_kernel(const __grid_constant__ NmsParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NmsSmem& S = *reinterpret_cast<NmsSmem*>(smem_raw);

    const int seg = blockIdx.x;
    const int task = seg / P.B, b = seg - task * P.B;
    const int nc = P.nc[task], A = P.A;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const T* __restrict__ img = reinterpret_cast<const T*>(P.pred[task]) + (size_t)b * (4 + nc) * A;
    float* __restrict__ dets = P.dets + (size_t)seg * P.max_det * 6;
    const float thr = P.conf_thr, iou_thr = P.iou_thr;
    const int max_det = P.max_det;
    const unsigned cap = (unsigned)P.chunk_cap;

    float4* kbox = S.kbox;
    float* karea = S.karea;
    int* kcode = S.kcode;
    if (max_det > NMS_KEPT_SMEM) {
        float* ws = P.kept_ws + (size_t)seg * max_det * 6;
        kbox = reinterpret_cast<float4*>(ws);
        karea = ws + (size_t)max_det * 4;
        kcode = reinterpret_cast<int*>(ws + (size_t)max_det * 5);
    }
    // Class shortcut: with class-offset boxes (general.py:462-463) two boxes of different classes
    // whose un-offset corners all lie in [0, max_wh) can never intersect, so their IoU test is
    // skipped.  The host verified that the fp32 offsets make this exact (P.class_shortcut).
    const bool shortcut = P.class_shortcut != 0;

    // ---------------- level-0 histogram (top 12 key bits).  For large multi-label segments it is an
    // ESTIMATE built from every hstride-th score vector: chunk boundaries only steer how much is
    // gathered at once; what a chunk contains, its order and every count that matters (collect)
    // are exact.  A chunk that turns out too large switches the segment to the exact histogram.
    const unsigned max_nms = (unsigned)max(P.max_nms, 0);
    unsigned hstride = 1;
    if (MULTI && P.hist_sample > 1) {
        const u64 nvec = ((u64)nc * (u64)A) / ScoreVec<T>::V;
        if (nvec >= 8192 && (A % ScoreVec<T>::V) == 0) hstride = (unsigned)P.hist_sample;
    }
    auto build_hist = [&](unsigned stride, unsigned below_digit) {
        for (int i = tid; i < NMS_BINS; i += NMS_THREADS) S.g0[i] = 0;
        __syncthreads();
        for_each_candidate<T, MULTI>(img, nc, A, thr, P, [&](unsigned sb, int, int) {
            const unsigned d = sb >> 20;
            if (d < below_digit) atomicAdd(&S.g0[d], 1u);
        }, stride);
        __syncthreads();
        suffix_scan(S.g0, S.warp_tot);
    };
    build_hist(hstride, NMS_BINS);

    unsigned consumed = 0;
    int kept = 0;
    unsigned target = min((unsigned)max(P.chunk_first, 1), cap);

    // -------- consume the sorted chunk in S.keys[0..n) (first `take` entries) ; returns updated kept
    auto consume_chunk = [&](unsigned n, unsigned take) {
        // bitonic sort, descending, padded with zeros (keys are never 0)
        unsigned n2 = 1;
        while (n2 < n) n2 <<= 1;
        for (unsigned i = n + tid; i < n2; i += NMS_THREADS) S.keys[i] = 0ull;
        __syncthreads();
        for (unsigned k = 2; k <= n2; k <<= 1) {
            for (unsigned j = k >> 1; j > 0; j >>= 1) {
                for (unsigned i = tid; i < n2; i += NMS_THREADS) {
                    const unsigned ixj = i ^ j;
                    if (ixj > i) {
                        const u64 x = S.keys[i], y = S.keys[ixj];
                        const bool desc = ((i & k) == 0);
                        if (desc ? (x < y) : (x > y)) { S.keys[i] = y; S.keys[ixj] = x; }
                    }
                }
                __syncthreads();
            }
        }
        // tiles
        for (unsigned t0 = 0; t0 < take && kept < max_det; t0 += NMS_TILE) {
            const int nt = (int)min((unsigned)NMS_TILE, take - t0);
            // A: materialise the tile's boxes  (general.py:443 xywh2xyxy in dtype, :462-463 class offset in fp32)
            if (tid < NMS_TILE) {
                unsigned char dead = 1;
                if (tid < nt) {
                    const u64 key = S.keys[t0 + tid];
                    const unsigned idx = 0xFFFFFFFFu - (unsigned)key;
                    const int a = idx / nc, c = idx - a * nc;
                    const float cx = to_f32<T>(__ldg(img + a));
                    const float cy = to_f32<T>(__ldg(img + (size_t)A + a));
                    const float bw = to_f32<T>(__ldg(img + (size_t)2 * A + a));
                    const float bh = to_f32<T>(__ldg(img + (size_t)3 * A + a));
                    const float hw_ = rnd<T>(bw * 0.5f), hh_ = rnd<T>(bh * 0.5f);
                    float4 r;
                    r.x = rnd<T>(__fsub_rn(cx, hw_));
                    r.y = rnd<T>(__fsub_rn(cy, hh_));
                    r.z = rnd<T>(__fadd_rn(cx, hw_));
                    r.w = rnd<T>(__fadd_rn(cy, hh_));
                    const float off = __fmul_rn((float)c, P.class_gap);
                    float4 o;
                    o.x = __fadd_rn(r.x, off);
                    o.y = __fadd_rn(r.y, off);
                    o.z = __fadd_rn(r.z, off);
                    o.w = __fadd_rn(r.w, off);
                    S.traw[tid] = r;
                    S.tbox[tid] = o;
                    S.tarea[tid] = __fmul_rn(__fsub_rn(o.z, o.x), __fsub_rn(o.w, o.y));
                    S.tscore[tid] = __uint_as_float((unsigned)(key >> 32));
                    S.tcls[tid] = c;
                    const float g = P.class_gap;
                    const bool tame = shortcut && r.x >= 0.f && r.y >= 0.f && r.z >= 0.f && r.w >= 0.f &&
                                      r.x < g && r.y < g && r.z < g && r.w < g;
                    S.tcode[tid] = tame ? c : -1;
                    dead = 0;
                }
                S.tdead[tid] = dead;
            }
            __syncthreads();
            // B: against the kept list (two threads per candidate, each takes half of the list)
            {
                const int j = tid & (NMS_TILE - 1), half = tid >> 8;
                if (j < nt && kept > 0) {
                    const int mid = (kept + 1) >> 1;
                    const int k0 = half ? mid : 0, k1 = half ? kept : mid;
                    const float4 bj = S.tbox[j];
                    const float aj = S.tarea[j];
                    const int cj = S.tcode[j];
                    for (int k = k0; k < k1; ++k) {
                        const int ck = kcode[k];
                        if (ck != cj && (ck | cj) >= 0) continue;  // different classes, both tame
                        if (suppresses(kbox[k], karea[k], bj, aj, iou_thr)) { S.tdead[j] = 1; break; }
                    }
                }
            }
            __syncthreads();
            // C: who suppresses whom inside the tile.  mask[j][w] = bits of the earlier candidates
            //    i in word w (i < j, both alive) that would suppress j.  item = (j, word).
            if (tid < NMS_TILE) {
                const unsigned al = __ballot_sync(0xffffffffu, S.tdead[tid] == 0);
                if (lane == 0) S.keep32[0][wid] = al;
            }
            for (int item = tid; item < NMS_TILE * NMS_TILE_WORDS; item += NMS_THREADS) {
                const int j = item & (NMS_TILE - 1), w = item >> 8;
                u64 bits = 0;
                if (j < nt && !S.tdead[j] && w * 64 < j) {
                    const float4 bj = S.tbox[j];
                    const float aj = S.tarea[j];
                    const int cj = S.tcode[j];
                    const int i0 = w * 64;
                    const int iend = min(64, j - i0);
                    for (int ii = 0; ii < iend; ++ii) {
                        const int i = i0 + ii;
                        const int ci = S.tcode[i];
                        if (ci != cj && (ci | cj) >= 0) continue;  // different classes, both tame
                        if (!S.tdead[i] && suppresses(S.tbox[i], S.tarea[i], bj, aj, iou_thr)) bits |= 1ull << ii;
                    }
                }
                S.mask[j][w] = bits;
            }
            __syncthreads();
            // D: greedy result as the fixpoint of  keep[j] = alive[j] && no kept earlier i suppresses j.
            //    It is unique (keep[j] depends only on lower indices) and after r rounds the first r
            //    candidates are final, so the loop ends in <= nt rounds -- in practice a handful.
            {
                u64 m0 = 0, m1 = 0, m2 = 0, m3 = 0;
                bool alive = false;
                if (tid < NMS_TILE) {
                    m0 = S.mask[tid][0]; m1 = S.mask[tid][1]; m2 = S.mask[tid][2]; m3 = S.mask[tid][3];
                    alive = S.tdead[tid] == 0;
                }
                int cur = 0;
                for (;;) {
                    int changed = 0;
                    if (tid < NMS_TILE) {
                        const unsigned* K = S.keep32[cur];
                        const u64 k0 = ((u64)K[1] << 32) | K[0], k1 = ((u64)K[3] << 32) | K[2];
                        const u64 k2 = ((u64)K[5] << 32) | K[4], k3 = ((u64)K[7] << 32) | K[6];
                        const bool kj = alive && (((m0 & k0) | (m1 & k1) | (m2 & k2) | (m3 & k3)) == 0ull);
                        const unsigned nw = __ballot_sync(0xffffffffu, kj);
                        if (lane == 0) { S.keep32[cur ^ 1][wid] = nw; changed = (nw != K[wid]); }
                    }
                    cur ^= 1;
                    if (!__syncthreads_or(changed)) break;
                }
                // first `budget` kept candidates only (greedy stops at max_det, general.py:465)
                if (tid == 0) {
                    int budget = max_det - kept;
                    for (int w = 0; w < NMS_TILE_WORDS; ++w) {
                        u64 kw = ((u64)S.keep32[cur][2 * w + 1] << 32) | S.keep32[cur][2 * w];
                        int c = __popcll(kw);
                        while (c > budget) { kw &= ~(1ull << (63 - __clzll((long long)kw))); --c; }
                        budget -= c;
                        S.keepmask[w] = kw;
                    }
                }
            }
            __syncthreads();
            // E: append the kept ones (in order) to the kept list and to the output
            int added = 0;
#pragma unroll
            for (int w = 0; w < NMS_TILE_WORDS; ++w) added += __popcll(S.keepmask[w]);
            if (tid < nt) {
                const int w = tid >> 6, bit = tid & 63;
                if ((S.keepmask[w] >> bit) & 1ull) {
                    int pos = kept + __popcll(S.keepmask[w] & ((1ull << bit) - 1ull));
                    for (int ww = 0; ww < w; ++ww) pos += __popcll(S.keepmask[ww]);
                    kbox[pos] = S.tbox[tid];
                    karea[pos] = S.tarea[tid];
                    kcode[pos] = S.tcode[tid];
                    const float4 r = S.traw[tid];
                    float* o = dets + (size_t)pos * 6;  // general.py:474 rows (x1,y1,x2,y2,conf,cls)
                    o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
                    o[4] = S.tscore[tid];
                    o[5] = (float)S.tcls[tid];
                }
            }
            kept += added;
            __syncthreads();
        }
    };

    // -------- gather the keys in [lo, hi) into S.keys ; returns count (uniform)
    auto collect = [&](u64 lo, u64 hi) -> unsigned {
        if (tid == 0) S.counter = 0;
        __syncthreads();
        const unsigned sb_lo = (unsigned)(lo >> 32), sb_hi = (unsigned)((hi - 1) >> 32);
        for_each_candidate<T, MULTI>(img, nc, A, thr, P, [&](unsigned sb, int a, int c) {
            if (sb >= sb_lo && sb <= sb_hi) {
                const u64 key = make_key(sb, (unsigned)(a * nc + c));
                if (key >= lo && key < hi) {
                    const unsigned p = atomicAdd(&S.counter, 1u);
                    if (p < NMS_CAP) S.keys[p] = key;
                }
            }
        });
        __syncthreads();
        const unsigned n = S.counter;
        __syncthreads();
        return n;
    };

    // -------- a level-0 bin that alone exceeds the chunk capacity: refine by deeper digits
    auto overflow_bin = [&](int bin) {
        const u64 lowlim = (u64)bin << 52;
        u64 bnd = (u64)(bin + 1) << 52;
        while (consumed < max_nms && kept < max_det) {
            u64 lo = lowlim, hi = bnd, L = lowlim;
            int lvl = 1;
            bool empty = false;
            for (;;) {
                for (int i = tid; i < NMS_BINS; i += NMS_THREADS) S.g1[i] = 0;
                __syncthreads();
                const unsigned sb_lo = (unsigned)(lo >> 32), sb_hi = (unsigned)((hi - 1) >> 32);
                for_each_candidate<T, MULTI>(img, nc, A, thr, P, [&](unsigned sb, int a, int c) {
                    if (sb >= sb_lo && sb <= sb_hi) {
                        const u64 key = make_key(sb, (unsigned)(a * nc + c));
                        if (key >= lo && key < hi) atomicAdd(&S.g1[level_digit(key, lvl)], 1u);
                    }
                });
                __syncthreads();
                suffix_scan(S.g1, S.warp_tot);
                const unsigned tot = S.g1[0];
                if (tot == 0) { empty = true; break; }
                if (tot <= cap) { L = lo; break; }
                const unsigned need = min(cap, max_nms - consumed);
                const int nb = lvl < 5 ? NMS_BINS : 16;
                const int sh = level_shift(lvl);
                const u64 parent = lvl < 5 ? (lo >> (sh + 12)) << (sh + 12) : (lo >> 4) << 4;
                const int d = find_digit(S.g1, nb, 0u, min(need, tot));
                const unsigned cnt = S.g1[d];
                if (cnt <= cap) { L = parent | ((u64)d << sh); break; }
                const unsigned c1 = S.g1[d + 1];
                if (c1 > 0) { L = parent | ((u64)(d + 1) << sh); break; }
                const u64 blo = parent | ((u64)d << sh);
                const u64 bhi = blo + (1ull << sh);
                lo = lo > blo ? lo : blo;
                hi = hi < bhi ? hi : bhi;
                ++lvl;
                __syncthreads();
            }
            __syncthreads();
            if (empty) return;
            if (L < lowlim) L = lowlim;
            const unsigned n = collect(L, bnd);
            const unsigned take = min(n, max_nms - consumed);
            consume_chunk(n, take);
            consumed += take;
            bnd = L;
            if (L == lowlim) return;
        }
    };

    // ---------------- main loop over level-0 digit ranges, from the top
    int hi0 = NMS_BINS / 2;  // float sign bit is 0: digits < 2048, so (hi0 << 52) never overflows
    while (hi0 > 0 && consumed < max_nms && kept < max_det) {
        const unsigned base = S.g0[hi0];
        const unsigned rem_s = S.g0[0] - base;  // candidates left below hi0, in histogram units
        if (hstride == 1 && rem_s == 0) break;
        const unsigned need = min(target, max_nms - consumed);
        const unsigned need_s = hstride == 1 ? need : (need + need / 4 + hstride - 1) / hstride;  // +25% margin
        int lo0 = 0;
        bool single_heavy = false;
        if (rem_s > need_s) {
            const int d = find_digit(S.g0, hi0, base, need_s);
            lo0 = d;
            const unsigned est = (S.g0[d] - base) * hstride;
            if (est > (hstride == 1 ? cap : cap - cap / 4)) {
                if (S.g0[d + 1] - base > 0) lo0 = d + 1;
                else single_heavy = (hstride == 1);  // exact: digit d alone exceeds the capacity
            }
        }
        if (single_heavy) {
            overflow_bin(lo0);
            hi0 = lo0;
            target = cap;
            continue;
        }
        const unsigned n = collect((u64)lo0 << 52, (u64)hi0 << 52);
        if (n > cap) {
            // estimate was off (or ties piled up): count exactly what is left and retry
            build_hist(1u, (unsigned)hi0);
            hstride = 1;
            continue;
        }
        const unsigned take = min(n, max_nms - consumed);
        consume_chunk(n, take);
        consumed += take;
        hi0 = lo0;
        target = min(target * 2u, cap);
    }

    if (tid == 0) P.counts[seg] = kept;
}

size_t cerb_nms_kept_ws_bytes(int T, int B, int max_det) {
    if (max_det <= NMS_KEPT_SMEM) return 0;
    return (size_t)T * B * max_det * 6 * sizeof(float);
}

template <typename T, bool MULTI> static cudaError_t launch_nms_t(const NmsParams& P, cudaStream_t stream) {
    const size_t smem = sizeof(NmsSmem);
    cudaError_t e = cudaFuncSetAttribute(nms_kernel<T, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    nms_kernel<T, MULTI><<<P.T * P.B, NMS_THREADS, smem, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t cerb_launch_nms(const NmsParams& P, int dtype, cudaStream_t stream) {
    if (P.T * P.B == 0) return cudaSuccess;
    // multi_label &= nc > 1 (general.py:419): with nc == 1 both modes select the same candidates
    const bool multi = P.multi_label != 0;
    if (dtype == CERB_DTYPE_F16)
        return multi ? launch_nms_t<__half, true>(P, stream) : launch_nms_t<__half, false>(P, stream);
    return multi ? launch_nms_t<float, true>(P, stream) : launch_nms_t<float, false>(P, stream);
}
