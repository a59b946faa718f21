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
//  * Suppression (:462-465): candidates are taken in tiles of 512, one per thread; a
//    candidate is first tested against the kept list (shared memory), then the tile is
//    resolved internally with a per-thread suppressor bitmask and a ballot fixpoint
//    iteration (equal to the sequential sweep).  Per-class chains skip the pairs that the
//    class offset makes disjoint.  IoU arithmetic reproduces
//    torchvision's CPU kernel: separately rounded fp32 ops on class-offset boxes, strict
//    '>' against the threshold (the double-precision compare is folded into iou_thr).
#include "cerb_kernels.h"

#define NMS_THREADS 512
#define NMS_TILE 512         // == NMS_THREADS: one candidate per thread
#define NMS_TILE_WORDS (NMS_TILE / 64)
#define NMS_CAP 4096         // max keys sorted at once
#define NMS_BINS 4096        // 12-bit radix digits
#define NMS_KEPT_SMEM 1024   // kept list lives in smem up to this max_det
#define NMS_BUCKETS 1024     // class buckets of the kept-list chains (class & 1023)

typedef unsigned long long u64;

struct __align__(16) NmsSmem {
    unsigned g0[NMS_BINS + 8];   // level-0: G0[d] = #keys with digit >= d (histogram units); G0[4096] = 0
    unsigned g1[NMS_BINS + 8];   // scratch for deeper levels
    u64 keys[NMS_CAP];
    float4 tbox[NMS_TILE];       // class-offset corners of the tile's candidates
    float tarea[NMS_TILE];
    int tcode[NMS_TILE];         // class id if the box is "tame" (see class shortcut), -1 wild, -2 padding
    int tprev[NMS_TILE];         // previous candidate of the same class inside the tile, or -1
    unsigned keep32[2][NMS_TILE / 32];
    unsigned char tdead[NMS_TILE];
    unsigned warp_tot[NMS_THREADS / 32];
    unsigned counter;
    int tile_wild, kept_wild;
    int khead[NMS_BUCKETS];      // newest kept box of each class bucket, or -1
    int knext_s[NMS_KEPT_SMEM];  // next kept box of the same bucket
    float4 kbox[NMS_KEPT_SMEM];
    float karea[NMS_KEPT_SMEM];
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
    if (max_det > NMS_KEPT_SMEM) {
        float* ws = P.kept_ws + (size_t)seg * max_det * 5;
        kbox = reinterpret_cast<float4*>(ws);
        karea = ws + (size_t)max_det * 4;
    }
    // Class shortcut.  Boxes are offset by class * max_wh before NMS (general.py:462-463), so two boxes
    // of different classes whose un-offset corners all lie in the window [tame_lo, tame_hi) of width
    // max_wh can never intersect (the host checked that the fp32 offsets are exact, P.class_shortcut).
    // While every box involved is such a "tame" box, a candidate is only tested against the kept /
    // earlier boxes of its own class, found through per-class chains; one wild box anywhere (or
    // agnostic mode, or a kept list too large for shared memory) switches to testing all pairs.
    const bool shortcut = P.class_shortcut != 0 && max_det <= NMS_KEPT_SMEM;
    bool kept_wild = false;
    for (int i = tid; i < NMS_BUCKETS; i += NMS_THREADS) S.khead[i] = -1;
    if (tid == 0) S.kept_wild = 0;

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
        // tiles: one candidate per thread
        for (unsigned t0 = 0; t0 < take && kept < max_det; t0 += NMS_TILE) {
            const int nt = (int)min((unsigned)NMS_TILE, take - t0);
            // A: materialise the tile's boxes  (general.py:443 xywh2xyxy in dtype, :462-463 class offset in fp32)
            float4 raw = make_float4(0.f, 0.f, 0.f, 0.f), box = raw;
            float area = 0.f, score = 0.f;
            int cls = 0, code = -1;
            bool dead = true;
            if (tid == 0) S.tile_wild = 0;
            __syncthreads();
            if (tid < nt) {
                const u64 key = S.keys[t0 + tid];
                const unsigned idx = 0xFFFFFFFFu - (unsigned)key;
                const int a = idx / nc;
                cls = idx - a * nc;
                const float cx = to_f32<T>(__ldg(img + a));
                const float cy = to_f32<T>(__ldg(img + (size_t)A + a));
                const float bw = to_f32<T>(__ldg(img + (size_t)2 * A + a));
                const float bh = to_f32<T>(__ldg(img + (size_t)3 * A + a));
                const float hw_ = rnd<T>(bw * 0.5f), hh_ = rnd<T>(bh * 0.5f);
                raw.x = rnd<T>(__fsub_rn(cx, hw_));
                raw.y = rnd<T>(__fsub_rn(cy, hh_));
                raw.z = rnd<T>(__fadd_rn(cx, hw_));
                raw.w = rnd<T>(__fadd_rn(cy, hh_));
                const float off = __fmul_rn((float)cls, P.class_gap);
                box.x = __fadd_rn(raw.x, off);
                box.y = __fadd_rn(raw.y, off);
                box.z = __fadd_rn(raw.z, off);
                box.w = __fadd_rn(raw.w, off);
                area = __fmul_rn(__fsub_rn(box.z, box.x), __fsub_rn(box.w, box.y));
                score = __uint_as_float((unsigned)(key >> 32));
                const float lo = P.tame_lo, hi = P.tame_hi;
                const bool tame = raw.x >= lo && raw.y >= lo && raw.z >= lo && raw.w >= lo &&
                                  raw.x < hi && raw.y < hi && raw.z < hi && raw.w < hi;
                code = tame ? cls : -1;
                if (!tame) S.tile_wild = 1;
                dead = false;
            }
            S.tbox[tid] = box;
            S.tarea[tid] = area;
            S.tcode[tid] = dead ? -2 : code;  // -2: padding, never equal to a class
            __syncthreads();
            // Per-class chains are valid while every box involved is tame (see the class shortcut above).
            const bool fast = shortcut && !S.tile_wild && !kept_wild;
            int prev = -1;
            if (fast && !dead) {  // previous candidate of the same class inside the tile
                int p = tid - 1;
                while (p >= 0 && S.tcode[p] != code) --p;
                prev = p;
            }
            S.tprev[tid] = prev;
            // B: against the kept list
            if (!dead && kept > 0) {
                if (fast) {
                    for (int k = S.khead[cls & (NMS_BUCKETS - 1)]; k >= 0; k = S.knext_s[k]) {
                        if (suppresses(kbox[k], karea[k], box, area, iou_thr)) { dead = true; break; }
                    }
                } else {
                    for (int k = 0; k < kept; ++k) {
                        if (suppresses(kbox[k], karea[k], box, area, iou_thr)) { dead = true; break; }
                    }
                }
            }
            S.tdead[tid] = dead ? 1 : 0;
            __syncthreads();
            // C: m[w] = bits of the earlier candidates i in word w (alive) that would suppress this one
            u64 m[NMS_TILE_WORDS];
#pragma unroll
            for (int w = 0; w < NMS_TILE_WORDS; ++w) m[w] = 0ull;
            if (!dead) {
                if (fast) {
                    for (int i = prev; i >= 0; i = S.tprev[i]) {
                        if (!S.tdead[i] && suppresses(S.tbox[i], S.tarea[i], box, area, iou_thr)) {
                            const u64 bit = 1ull << (i & 63);
#pragma unroll
                            for (int w = 0; w < NMS_TILE_WORDS; ++w)
                                if (w == (i >> 6)) m[w] |= bit;
                        }
                    }
                } else {
                    for (int i = 0; i < tid; ++i) {
                        if (!S.tdead[i] && suppresses(S.tbox[i], S.tarea[i], box, area, iou_thr)) {
                            const u64 bit = 1ull << (i & 63);
#pragma unroll
                            for (int w = 0; w < NMS_TILE_WORDS; ++w)
                                if (w == (i >> 6)) m[w] |= bit;
                        }
                    }
                }
            }
            {
                const unsigned al = __ballot_sync(0xffffffffu, !dead);
                if (lane == 0) S.keep32[0][wid] = al;
            }
            __syncthreads();
            // D: greedy result as the fixpoint of  keep[j] = alive[j] && no kept earlier i suppresses j.
            //    It is unique (keep[j] depends only on lower indices) and after r rounds the first r
            //    candidates are final, so the loop ends in <= nt rounds -- in practice a handful.
            int cur = 0;
            for (;;) {
                const unsigned* K = S.keep32[cur];
                u64 hit = 0ull;
#pragma unroll
                for (int w = 0; w < NMS_TILE_WORDS; ++w) hit |= m[w] & (((u64)K[2 * w + 1] << 32) | (u64)K[2 * w]);
                const bool kj = !dead && hit == 0ull;
                const unsigned nw = __ballot_sync(0xffffffffu, kj);
                int changed = 0;
                if (lane == 0) { S.keep32[cur ^ 1][wid] = nw; changed = (nw != K[wid]); }
                cur ^= 1;
                if (!__syncthreads_or(changed)) break;
            }
            // E: the first `budget` kept candidates (greedy stops at max_det, general.py:465) are appended,
            //    in order, to the kept list and to the output
            {
                const unsigned* K = S.keep32[cur];
                int before = 0, total = 0;
#pragma unroll
                for (int w = 0; w < NMS_THREADS / 32; ++w) {
                    const int c = __popc(K[w]);
                    if (w < wid) before += c;
                    total += c;
                }
                const bool mine = (K[wid] >> lane) & 1u;
                const int pos = kept + before + __popc(K[wid] & ((1u << lane) - 1u));
                if (mine && pos < max_det) {
                    kbox[pos] = box;
                    karea[pos] = area;
                    float* o = dets + (size_t)pos * 6;  // general.py:474 rows (x1,y1,x2,y2,conf,cls)
                    o[0] = raw.x; o[1] = raw.y; o[2] = raw.z; o[3] = raw.w;
                    o[4] = score;
                    o[5] = (float)cls;
                    if (code < 0) S.kept_wild = 1;
                    if (max_det <= NMS_KEPT_SMEM) {  // chain the box into its class bucket (any order is fine)
                        const int old = atomicExch(&S.khead[cls & (NMS_BUCKETS - 1)], pos);
                        S.knext_s[pos] = old;
                    }
                }
                kept = min(kept + total, max_det);
            }
            __syncthreads();
            kept_wild = kept_wild || (S.kept_wild != 0);
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
    return (size_t)T * B * max_det * 5 * sizeof(float);
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
