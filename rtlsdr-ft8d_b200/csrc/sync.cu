// sync.cu -- Costas 7x7 sync scoring over the whole waterfall + exact top-K candidate selection.
// Replaces ft8_find_sync() / ft8_sync_score() / heapify_*(), /root/reference/ft8_lib/ft8/decode.c:35-108,
// 173-234, 388-435.
//
// Two kernels.  sync_score_kernel: a few CTAs per slot, each stages the slot's waterfall (94 KB for the daemon
// geometry) in shared memory and scores its share of the tosr*fosr*36*(bins-7) positions; scores are stored in the
// reference's loop order (time_sub, freq_sub, time_offset, freq_offset).  sync_select_kernel: one CTA per slot
// compacts the positions with score >= min_score IN THAT ORDER (ballot + prefix) and one thread replays the
// reference's min-heap insertions and the final heap sort over the survivors only.  The replay is what makes the
// retained set at the cut score and the order among equal scores identical to the reference (they depend on heap
// history).  The surviving candidates are also appended to a flat work list for the decode kernel.
#include "common.cuh"

namespace ft8b200 {
namespace {

constexpr int kSyncThreads = 1024;
__constant__ uint8_t c_costas[7] = {3, 1, 4, 0, 6, 5, 2};

struct Geo { int nb, nbins, tosr, fosr, stride, nfo, npos; };

// ref: ft8_sync_score(), decode.c:44-108
__device__ __forceinline__ int sync_score(const uint8_t *__restrict__ mag, const Geo &g, int ts, int fs, int to, int fo) {
    const long origin = (((long)to * g.tosr + ts) * g.fosr + fs) * g.nbins + fo;
    int score = 0, terms = 0;
#pragma unroll
    for (int grp = 0; grp < 3; ++grp) {
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const int rel = 36 * grp + k;
            const int row = to + rel;
            if (row < 0) continue;
            if (row >= g.nb) break;  // leaves this group only, like the reference's inner `break`
            const uint8_t *p = mag + origin + (long)rel * g.stride;
            const int tone = c_costas[k];
            const int centre = p[tone];
            if (tone > 0) { score += centre - p[tone - 1]; ++terms; }
            if (tone < 7) { score += centre - p[tone + 1]; ++terms; }
            if (k > 0 && row > 0) { score += centre - p[tone - g.stride]; ++terms; }
            if (k + 1 < 7 && row + 1 < g.nb) { score += centre - p[tone + g.stride]; ++terms; }
        }
    }
    if (terms > 0) score /= terms;  // truncating division
    return score;
}

// candidate_t as one 64-bit word: score | time_offset<<16 | freq_offset<<32 | time_sub<<48 | freq_sub<<56
__device__ __forceinline__ int cand_score(unsigned long long c) { return (int)(short)(c & 0xffffull); }

__device__ void sift_down(unsigned long long *h, int n) {  // ref: heapify_down, decode.c:388-415
    int cur = 0;
    for (;;) {
        int pick = cur;
        const int l = 2 * cur + 1, r = l + 1;
        if (l < n && cand_score(h[l]) < cand_score(h[pick])) pick = l;
        if (r < n && cand_score(h[r]) < cand_score(h[pick])) pick = r;
        if (pick == cur) return;
        const unsigned long long t = h[pick]; h[pick] = h[cur]; h[cur] = t;
        cur = pick;
    }
}
__device__ void sift_up(unsigned long long *h, int n) {  // ref: heapify_up, decode.c:417-435
    int cur = n - 1;
    while (cur > 0) {
        const int par = (cur - 1) / 2;
        if (cand_score(h[cur]) >= cand_score(h[par])) return;
        const unsigned long long t = h[par]; h[par] = h[cur]; h[cur] = t;
        cur = par;
    }
}

// Phase 1: score every position.  grid = (chunks, slots): each CTA stages the slot's waterfall in shared memory
// (when it fits) and scores its share of the positions; scores go to global memory as int16 in position order.
template <bool kStage>
__global__ void __launch_bounds__(kSyncThreads)
sync_score_kernel(const uint8_t *__restrict__ mag_all, size_t slot_stride, Geo g, int pos_per_chunk, int16_t *__restrict__ scores_all) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x;
    const int slot = blockIdx.y;
    const int wf_bytes = g.nb * g.stride;
    const uint8_t *gmag = mag_all + (size_t)slot * slot_stride;
    const uint8_t *mag = gmag;
    if (kStage) {
        if ((((size_t)gmag) & 15) == 0 && (wf_bytes & 15) == 0) {
            const uint4 *src = reinterpret_cast<const uint4 *>(gmag);
            uint4 *dst = reinterpret_cast<uint4 *>(smem);
            for (int k = tid; k < wf_bytes / 16; k += kSyncThreads) dst[k] = __ldg(src + k);
        } else {
            for (int k = tid; k < wf_bytes; k += kSyncThreads) smem[k] = gmag[k];
        }
        mag = smem;
        __syncthreads();
    }
    int16_t *scores = scores_all + (size_t)slot * g.npos;
    const int p0 = blockIdx.x * pos_per_chunk;
    int p1 = p0 + pos_per_chunk;
    if (p1 > g.npos) p1 = g.npos;
    for (int p = p0 + tid; p < p1; p += kSyncThreads) {
        const int fo = p % g.nfo;
        int q = p / g.nfo;
        const int to = q % 36 - 12;
        q /= 36;
        const int fs = q % g.fosr, ts = q / g.fosr;
        scores[p] = (int16_t)sync_score(mag, g, ts, fs, to, fo);  // stored as int16_t in candidate_t
    }
}

// Phase 2: one CTA per slot (looping over slots): ordered compaction of the positions with score >= min_score,
// then the exact heap replay by one thread, then the candidates are appended to the decode work list.
__global__ void __launch_bounds__(kSyncThreads)
sync_select_kernel(const int16_t *__restrict__ scores_all, int n_slots, Geo g, int max_cand, int min_score, candidate_t *__restrict__ cand_out,
                   int *__restrict__ ncand_out, uint32_t *__restrict__ scratch_all, uint32_t *__restrict__ work, unsigned int *__restrict__ work_total) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_warp_cnt[2][32];
    __shared__ int s_total, s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long *heap = reinterpret_cast<unsigned long long *>(smem);
    uint32_t *scratch = scratch_all + (size_t)blockIdx.x * g.npos;

    for (int slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
        const int16_t *scores = scores_all + (size_t)slot * g.npos;
        // Ordered compaction with ONE block barrier: warp w owns the contiguous positions [w*span, (w+1)*span); it counts
        // its survivors, the warp totals are prefix-summed, then it writes its survivors at its offset (position order ==
        // the reference's loop order).  Scores are re-read in the second sweep (L1/L2 hits).
        const int span = ((g.npos + 31) / 32 + 31) / 32 * 32;  // positions per warp, multiple of 32
        const int w0 = warp * span;
        int mine = 0;
#pragma unroll 4
        for (int o = 0; o < span; o += 32) {
            const int p = w0 + o + lane;
            const bool pass = p < g.npos && scores[p] >= min_score;
            mine += __popc(__ballot_sync(0xffffffffu, pass));
        }
        if (lane == 0) s_warp_cnt[0][warp] = mine;
        __syncthreads();
        const int cnt = s_warp_cnt[0][lane];
        int running = __reduce_add_sync(0xffffffffu, lane < warp ? cnt : 0);
        const int n_pass_total = __reduce_add_sync(0xffffffffu, cnt);
#pragma unroll 4
        for (int o = 0; o < span; o += 32) {
            const int p = w0 + o + lane;
            int score = 0;
            bool pass = false;
            if (p < g.npos) { score = scores[p]; pass = score >= min_score; }
            const unsigned ballot = __ballot_sync(0xffffffffu, pass);
            if (pass) scratch[running + __popc(ballot & ((1u << lane) - 1u))] = ((uint32_t)p << 12) | ((uint32_t)score & 0xfffu);
            running += __popc(ballot);
        }
        __syncthreads();

        if (tid == 0) {  // exact replay of the reference's heap (decode.c:198-231) over the survivors
            const int n_pass = n_pass_total;
            int n = 0;
            for (int e = 0; e < n_pass; ++e) {
                const uint32_t v = scratch[e];
                const int score = ((int)(v << 20)) >> 20;
                const int p = (int)(v >> 12);
                if (n == max_cand && score > cand_score(heap[0])) {
                    heap[0] = heap[n - 1];
                    --n;
                    sift_down(heap, n);
                }
                if (n < max_cand) {
                    const int fo = p % g.nfo;
                    int q = p / g.nfo;
                    const int to = q % 36 - 12;
                    q /= 36;
                    const int fs = q % g.fosr, ts = q / g.fosr;
                    heap[n] = ((unsigned long long)(uint16_t)(short)score) | ((unsigned long long)(uint16_t)(short)to << 16) |
                              ((unsigned long long)(uint16_t)(short)fo << 32) | ((unsigned long long)(uint8_t)ts << 48) |
                              ((unsigned long long)(uint8_t)fs << 56);
                    ++n;
                    sift_up(heap, n);
                }
            }
            for (int rest = n; rest > 1;) {  // heap sort -> descending score
                const unsigned long long t = heap[rest - 1]; heap[rest - 1] = heap[0]; heap[0] = t;
                --rest;
                sift_down(heap, rest);
            }
            s_total = n;
            ncand_out[slot] = n;
            s_base = (work && n > 0) ? (int)atomicAdd(work_total, (unsigned int)n) : 0;
        }
        __syncthreads();
        {
            const int n = s_total;
            unsigned long long *dst = reinterpret_cast<unsigned long long *>(cand_out + (size_t)slot * max_cand);
            for (int k = tid; k < max_cand; k += kSyncThreads) dst[k] = (k < n) ? heap[k] : 0ull;
            if (work)
                for (int k = tid; k < n; k += kSyncThreads) work[s_base + k] = (uint32_t)slot * (uint32_t)max_cand + (uint32_t)k;
        }
        __syncthreads();
    }
}

}  // namespace

cudaError_t launch_find_sync(const uint8_t *d_mag, size_t slot_stride, int n_slots, int num_blocks, int num_bins, int time_osr, int freq_osr,
                             int max_cand, int min_score, candidate_t *d_cand, int *d_ncand, int16_t *d_scores, uint32_t *d_scratch,
                             int scratch_slots, uint32_t *d_work, unsigned int *d_work_total, int sm_count, cudaStream_t st, int *launches) {
    Geo g;
    g.nb = num_blocks; g.nbins = num_bins; g.tosr = time_osr; g.fosr = freq_osr;
    g.stride = time_osr * freq_osr * num_bins;
    g.nfo = num_bins - 7;
    g.npos = time_osr * freq_osr * 36 * g.nfo;
    const int wf_bytes = g.nb * g.stride;
    const size_t staged = (size_t)((wf_bytes + 15) & ~15);
    // enough CTAs to fill the machine even for small batches: ~4 per SM, at least 1024 positions each
    int chunks = (4 * sm_count + n_slots - 1) / n_slots;
    const int max_chunks = (g.npos + kSyncThreads - 1) / kSyncThreads;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks > 12) chunks = 12;
    if (chunks < 1) chunks = 1;
    const int per = (g.npos + chunks - 1) / chunks;
    dim3 grid(chunks, n_slots);
    if (staged <= 200 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(sync_score_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged);
        if (e != cudaSuccess) return e;
        sync_score_kernel<true><<<grid, kSyncThreads, staged, st>>>(d_mag, slot_stride, g, per, d_scores);
    } else {
        sync_score_kernel<false><<<grid, kSyncThreads, 0, st>>>(d_mag, slot_stride, g, per, d_scores);
    }
    ++*launches;
    if (d_work_total) {
        cudaError_t e = cudaMemsetAsync(d_work_total, 0, sizeof(unsigned int), st);
        if (e != cudaSuccess) return e;
    }
    const int sgrid = n_slots < scratch_slots ? n_slots : scratch_slots;
    sync_select_kernel<<<sgrid, kSyncThreads, (size_t)max_cand * 8, st>>>(d_scores, n_slots, g, max_cand, min_score, d_cand, d_ncand, d_scratch,
                                                                          d_work, d_work_total);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace ft8b200
