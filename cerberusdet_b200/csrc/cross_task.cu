// Cross-task merge of the per-task NMS results, one CTA per image.
//
// Replaces, for a whole batch in one launch, the per-image host loop of the reference
// CerberusDetInference.predict (cerberusdet_inference.py:140-155):
//   _combine_output      (:72-83)   concatenate the tasks' rows, local -> global class ids
//   nms_between_tasks    (utils/general.py:484-554) with box_iou (utils/metrics.py:415-433)
//   scale_boxes(...).round() (utils/general.py:313-357), optional
// The reference does this on the CPU with a Python O(n^2) loop per image and one .cpu() per task per image.
//
// Semantics kept exactly (pinned by tests/test_host_logic.py against the reference and by the GPU tests against
// cross_task.py): rows stay in task order; IoU is only taken between boxes of different tasks, as fp32
// inter / (area_a + area_b - inter + 1e-7) with one rounding per operation; rows are scanned top to bottom;
// an undeleted row with overlaps > thr keeps the FIRST arg-max score among {overlapping columns in ascending
// order, then the row itself} and deletes the others (already deleted columns still take part); if every row
// would be deleted nothing is.
#include "cerb_kernels.h"

#define XT_THREADS 256
#define XT_MAX_ROWS 1024  // rows of one image (sum of the tasks' counts) the shared-memory tables hold; images with more
                          // rows (max_det = 1000 in the reference's detect.py:124) use the caller's global workspace

// per-image tables, in shared memory (n <= XT_MAX_ROWS) or in the global workspace; [rows] each, mask [rows][words]
struct XtTables {
    float4* box;
    float* score;
    float* area;
    unsigned short* task;
    unsigned short* src;   // position inside the task's padded rows
    unsigned* deleted;     // [words]
    unsigned* mask;        // [n][words]
};
__host__ __device__ inline size_t xt_table_bytes(size_t rows) {
    const size_t words = (rows + 31) / 32;
    // box 16, score 4, area 4, task 2, src 2 per row; deleted words; mask rows * words; 16-byte aligned pieces
    return rows * 28 + ((words * 4 + 15) / 16) * 16 + rows * words * 4 + 64;
}
__device__ __forceinline__ XtTables xt_carve(unsigned char* base, size_t rows) {
    const size_t words = (rows + 31) / 32;
    XtTables t;
    t.box = reinterpret_cast<float4*>(base); base += rows * 16;
    t.score = reinterpret_cast<float*>(base); base += rows * 4;
    t.area = reinterpret_cast<float*>(base); base += rows * 4;
    t.task = reinterpret_cast<unsigned short*>(base); base += rows * 2;
    t.src = reinterpret_cast<unsigned short*>(base); base += rows * 2;
    base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(base) + 15) & ~(uintptr_t)15);
    t.deleted = reinterpret_cast<unsigned*>(base); base += ((words * 4 + 15) / 16) * 16;
    t.mask = reinterpret_cast<unsigned*>(base);
    return t;
}
struct XtSmem {
    int start[CERB_MAX_TASKS + 1];
    int any_overlap;
    int pad[2];
};

__global__ void __launch_bounds__(XT_THREADS) cross_task_kernel(const __grid_constant__ CrossTaskParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    XtSmem& H = *reinterpret_cast<XtSmem*>(smem_raw);
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int T = P.T, md = P.max_det;

    if (tid == 0) {
        int n = 0;
        for (int t = 0; t < T; ++t) { H.start[t] = n; n += min(max(P.counts[t * P.B + b], 0), md); }
        H.start[T] = n;
        H.any_overlap = 0;
    }
    __syncthreads();
    const int n = H.start[T];
    const int words = (n + 31) >> 5;
    // the tables: shared memory when this image's rows fit (sized by the launch for min(T * max_det, XT_MAX_ROWS) rows),
    // else this image's slice of the global workspace (the host guarantees it is there whenever T * max_det > XT_MAX_ROWS)
    const int smem_rows = min(T * md, XT_MAX_ROWS);
    const XtTables S_ = (n <= smem_rows)
                            ? xt_carve(smem_raw + sizeof(XtSmem), (size_t)max(n, 1))
                            : xt_carve(P.workspace + (size_t)b * xt_table_bytes((size_t)T * md), (size_t)n);
    struct { float4* box; float* score; float* area; unsigned short* task; unsigned short* src; unsigned* deleted; unsigned* mask; int* start; int& any_overlap; }
        S = {S_.box, S_.score, S_.area, S_.task, S_.src, S_.deleted, S_.mask, H.start, H.any_overlap};
    // ---- combine: rows in task order (cerberusdet_inference.py:72-83)
    for (int t = 0; t < T; ++t) {
        const int s0 = S.start[t], cnt = S.start[t + 1] - s0;
        const float* src = P.dets + ((size_t)t * P.B + b) * md * 6;
        for (int i = tid; i < cnt; i += XT_THREADS) {
            const float* r = src + (size_t)i * 6;
            const float4 bx = make_float4(r[0], r[1], r[2], r[3]);
            S.box[s0 + i] = bx;
            S.score[s0 + i] = r[4];
            S.area[s0 + i] = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));  // (a2 - a1).prod(2)
            S.task[s0 + i] = (unsigned short)t;
            S.src[s0 + i] = (unsigned short)i;
        }
    }
    for (int i = tid; i < words; i += XT_THREADS) S.deleted[i] = 0;
    __syncthreads();

    // ---- IoU bitmask between boxes of different tasks: bit (r, c) for c in a LATER task  (general.py:509-531)
    const float thr = P.iou_thr;
    int local_any = 0;
    for (int item = tid; item < n * words; item += XT_THREADS) {
        const int r = item / words, w = item - r * words;
        unsigned bits = 0;
        const int c_begin = max(w * 32, S.start[S.task[r] + 1]);  // first column of the next task
        const int c_end = min(n, w * 32 + 32);
        if (c_begin < c_end) {
            const float4 a = S.box[r];
            const float aa = S.area[r];
            for (int c = c_begin; c < c_end; ++c) {
                const float4 q = S.box[c];
                const float iw = fmaxf(__fsub_rn(fminf(a.z, q.z), fmaxf(a.x, q.x)), 0.f);  // (min(a2,b2) - max(a1,b1)).clamp(0)
                const float ih = fmaxf(__fsub_rn(fminf(a.w, q.w), fmaxf(a.y, q.y)), 0.f);
                const float inter = __fmul_rn(iw, ih);
                const float den = __fadd_rn(__fsub_rn(__fadd_rn(aa, S.area[c]), inter), 1e-7f);
                if (__fdiv_rn(inter, den) > thr) bits |= 1u << (c & 31);
            }
        }
        S.mask[item] = bits;
        local_any |= (bits != 0);
    }
    if (local_any) S.any_overlap = 1;
    __syncthreads();

    // ---- sequential row scan by one warp (general.py:536-549); lane l owns the deleted-word l, l+32 (n <= 1024)
    if (wid == 0 && S.any_overlap) {
        for (int r = 0; r < n; ++r) {
            if ((S.deleted[r >> 5] >> (r & 31)) & 1u) continue;  // uniform: shared memory, warp-synchronous
            // best column of this row: highest score, lowest index on ties
            float best = -INFINITY;
            int best_c = -1;
            bool any = false;
            for (int w = lane; w < words; w += 32) {
                unsigned bits = S.mask[r * words + w];
                any |= bits != 0;
                while (bits) {
                    const int c = w * 32 + __ffs(bits) - 1;
                    bits &= bits - 1;
                    const float sc = S.score[c];
                    if (sc > best || (sc == best && c < best_c) || best_c < 0) { best = sc; best_c = c; }
                }
            }
            if (!__any_sync(0xffffffffu, any)) continue;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
                if (oc >= 0 && (best_c < 0 || ob > best || (ob == best && oc < best_c))) { best = ob; best_c = oc; }
            }
            // idxs = cat(cols, [row]): the row wins only if strictly greater than every column (torch.argmax: first max)
            const int winner = (S.score[r] > best) ? r : best_c;
            for (int w = lane; w < words; w += 32) {
                unsigned del = S.mask[r * words + w];
                if ((r >> 5) == w) del |= 1u << (r & 31);
                if ((winner >> 5) == w) del &= ~(1u << (winner & 31));
                S.deleted[w] |= del;
            }
            __syncwarp();
        }
    }
    __syncthreads();

    // ---- survivors in order; "if len(bboxes) == len(to_delete): return bboxes" (general.py:551-552)
    int ndel = 0;
    for (int w = 0; w < words; ++w) ndel += __popc(S.deleted[w]);
    const bool keep_all = (ndel == n);
    float* out = P.out + (size_t)b * T * md * 6;
    const float* sc5 = P.scale ? P.scale + (size_t)b * 5 : nullptr;
    for (int r = tid; r < n; r += XT_THREADS) {
        const bool dead = !keep_all && ((S.deleted[r >> 5] >> (r & 31)) & 1u);
        if (dead) continue;
        int pos = r;
        if (!keep_all) {
            pos = 0;
            for (int w = 0; w < (r >> 5); ++w) pos += 32 - __popc(S.deleted[w]);
            pos += (r & 31) - __popc(S.deleted[r >> 5] & ((1u << (r & 31)) - 1u));
        }
        float4 bx = S.box[r];
        if (sc5) {  // scale_boxes(ratio_pad=None) + clip_boxes + .round()  (general.py:313-357, inference :153)
            const float gain = sc5[0], px = sc5[1], py = sc5[2], ow = sc5[3], oh = sc5[4];
            bx.x = rintf(fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.x, px), gain), 0.f), ow));
            bx.y = rintf(fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.y, py), gain), 0.f), oh));
            bx.z = rintf(fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.z, px), gain), 0.f), ow));
            bx.w = rintf(fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.w, py), gain), 0.f), oh));
        }
        const int t = S.task[r];
        const float* srow = P.dets + (((size_t)t * P.B + b) * md + S.src[r]) * 6;
        float* o = out + (size_t)pos * 6;
        o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
        o[4] = S.score[r];
        o[5] = srow[5] + (float)P.class_offset[t];  // local -> global class id (inference :56-70,:80)
    }
    if (tid == 0) P.out_counts[b] = keep_all ? n : n - ndel;
}

size_t cerb_cross_task_smem(int rows) { return sizeof(XtSmem) + xt_table_bytes((size_t)(rows < XT_MAX_ROWS ? rows : XT_MAX_ROWS)); }

// bytes of global workspace a launch needs: nothing while every image's rows fit the shared-memory tables for sure
size_t cerb_cross_task_ws_bytes(int T, int B, int max_det) {
    const size_t rows = (size_t)T * (size_t)max_det;
    return rows > XT_MAX_ROWS ? (size_t)B * xt_table_bytes(rows) : 0;
}

cudaError_t cerb_launch_cross_task(const CrossTaskParams& P, cudaStream_t stream) {
    if (P.B == 0) return cudaSuccess;
    const int rows = P.T * P.max_det;
    if (rows > XT_MAX_ROWS && P.workspace == nullptr) return cudaErrorInvalidConfiguration;
    const size_t smem = cerb_cross_task_smem(rows);
    cudaError_t e = cudaFuncSetAttribute(cross_task_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cross_task_kernel<<<P.B, XT_THREADS, smem, stream>>>(P);
    return cudaGetLastError();
}
