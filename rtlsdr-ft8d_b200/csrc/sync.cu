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

constexpr int kSyncThreads = 256;  // small CTAs: the heap replay is one thread's latency, so many slots should be resident per SM
constexpr int kSurvSmem = 1024;    // survivors of a slot mirrored in shared memory for the serial replay (4 KB)

struct Geo { int nb, nbins, tosr, fosr, stride, nfo, npos; };

// A heap entry is the survivor word itself: (position << 12) | (score & 0xfff) -- position = index in the reference's loop
// order (< 2^20), score sign-extended from 12 bits.  Comparisons look at the score only, exactly like the reference's
// heap[a].score < heap[b].score, so equal scores compare equal whatever their position.
__device__ __forceinline__ int ent_score(uint32_t v) { return ((int)(v << 20)) >> 20; }

// ref: heapify_down, decode.c:388-415 -- the element at the root sinks while a child is STRICTLY smaller (left child first,
// the right one only if smaller than the left).  "Hole" form of the reference's swap chain: the sinking element is always
// the one compared against, so moving children up and storing it once at the end performs the same comparisons and leaves
// the same array.  Both children are loaded before either is examined (one shared-memory latency per level, not two).
__device__ __forceinline__ void sift_down(uint32_t *h, int n) {
    const uint32_t x = h[0];
    const int xs = ent_score(x);
    int cur = 0;
    for (;;) {
        const int l = 2 * cur + 1, r = l + 1;
        if (l >= n) break;
        const uint32_t vl = h[l], vr = h[r < n ? r : l];
        int ps = xs, pick = cur;
        uint32_t pv = x;
        if (ent_score(vl) < ps) { ps = ent_score(vl); pick = l; pv = vl; }
        if (r < n && ent_score(vr) < ps) { pick = r; pv = vr; }
        if (pick == cur) break;
        h[cur] = pv;
        cur = pick;
    }
    h[cur] = x;
}
// ref: heapify_up, decode.c:417-435 -- the last element rises while STRICTLY smaller than its parent
__device__ __forceinline__ void sift_up(uint32_t *h, int n, uint32_t x) {
    const int xs = ent_score(x);
    int cur = n - 1;
    while (cur > 0) {
        const int par = (cur - 1) >> 1;
        const uint32_t pv = h[par];
        if (xs >= ent_score(pv)) break;
        h[cur] = pv;
        cur = par;
    }
    h[cur] = x;
}

// Phase 1: score every position.  For a fixed (time_sub, freq_sub) every byte a score touches lies in ONE sub-plane of
// the waterfall: mag[((row*tosr + ts)*fosr + fs)*nbins + bin] for row < nb, bin < nbins (nb x nbins bytes: 23.5 KB for
// the daemon geometry, 89 KB for the 12 kHz monitor).  grid = (planes * splits, slots): a CTA stages its plane compactly
// in shared memory ([row][bin]) and scores the positions of that plane for its share of the 36 time offsets; thread t
// owns frequency offsets t, t + blockDim, ... so a warp reads 32 consecutive bytes per access (one wavefront) and the
// row-range tests are uniform across the CTA.  Scores go to global memory as int16 in the reference's loop order.
constexpr int kScoreThreads = 256;

// ref: ft8_sync_score(), decode.c:44-108 / ft4_sync_score(), decode.c:110-171, on the compact plane (`nbins` = row pitch)
// FT8: 3 groups of 7 symbols at 0, 36, 72 sharing one Costas array, 8 tones.  FT4: 4 groups of 4 symbols at 1, 34, 67,
// 100, one Costas array per group, 4 tones.
template <int kBins, bool kFt4>
__device__ __forceinline__ int sync_score_plane(const uint8_t *__restrict__ plane, int nb, int nbins_rt, int to, int fo) {
    const int nbins = kBins > 0 ? kBins : nbins_rt;
    constexpr int kGroups = kFt4 ? 4 : 3, kLen = kFt4 ? 4 : 7, kFirst = kFt4 ? 1 : 0, kStep = kFt4 ? 33 : 36, kTop = kFt4 ? 3 : 7;
    constexpr int kCostas8[7] = {3, 1, 4, 0, 6, 5, 2};
    constexpr int kCostas4[4][4] = {{0, 1, 3, 2}, {1, 0, 2, 3}, {2, 3, 1, 0}, {3, 2, 0, 1}};
    int score = 0, terms = 0;
#pragma unroll
    for (int grp = 0; grp < kGroups; ++grp) {
#pragma unroll
        for (int k = 0; k < kLen; ++k) {
            const int row = to + kFirst + kStep * grp + k;
            if (row < 0) continue;
            if (row >= nb) break;  // leaves this group only, like the reference's inner `break`
            const int tone = kFt4 ? kCostas4[grp][k < 4 ? k : 0] : kCostas8[k];
            const uint8_t *p = plane + row * nbins + fo + tone;
            const int centre = p[0];
            if (tone > 0) { score += centre - p[-1]; ++terms; }
            if (tone < kTop) { score += centre - p[1]; ++terms; }
            if (k > 0 && row > 0) { score += centre - p[-nbins]; ++terms; }
            if (k + 1 < kLen && row + 1 < nb) { score += centre - p[nbins]; ++terms; }
        }
    }
    if (terms > 0) score /= terms;  // truncating division
    return score;
}

template <int kBins, bool kStage, bool kFt4>
__global__ void __launch_bounds__(kScoreThreads)
sync_score_kernel(const uint8_t *__restrict__ mag_all, size_t slot_stride, Geo g, int splits, int to_per_cta, int16_t *__restrict__ scores_all) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x;
    const int slot = blockIdx.y;
    const int plane_id = blockIdx.x / splits, split = blockIdx.x - plane_id * splits;  // plane_id = ts*fosr + fs
    const int nbins = kBins > 0 ? kBins : g.nbins;
    const uint8_t *gplane = mag_all + (size_t)slot * slot_stride + (size_t)plane_id * nbins;  // row r at gplane + r*stride
    const int to0 = -12 + split * to_per_cta;
    int to1 = to0 + to_per_cta;
    if (to1 > 24) to1 = 24;
    const uint8_t *plane;
    int pitch;
    if (kStage) {
        // rows this CTA can touch: [to0 - 1, (to1 - 1) + last sync symbol + 1], clipped (last sync symbol: 78 FT8, 103 FT4)
        int r0 = to0 - 1, r1 = to1 + (kFt4 ? 104 : 79);
        if (r0 < 0) r0 = 0;
        if (r1 > g.nb) r1 = g.nb;
        if (((((size_t)gplane) | (size_t)g.stride | (size_t)nbins) & 15) == 0) {
            const int vec_per_row = nbins >> 4;
            for (int v = tid; v < (r1 - r0) * vec_per_row; v += kScoreThreads) {
                const int r = r0 + v / vec_per_row, c = v - (v / vec_per_row) * vec_per_row;
                reinterpret_cast<uint4 *>(smem + (size_t)r * nbins)[c] = __ldg(reinterpret_cast<const uint4 *>(gplane + (size_t)r * g.stride) + c);
            }
        } else {
            for (int v = tid; v < (r1 - r0) * nbins; v += kScoreThreads) {
                const int r = r0 + v / nbins, c = v - (v / nbins) * nbins;
                smem[(size_t)r * nbins + c] = gplane[(size_t)r * g.stride + c];
            }
        }
        plane = smem;
        pitch = nbins;
        __syncthreads();
    } else {
        plane = gplane;
        pitch = g.stride;
    }
    int16_t *scores = scores_all + (size_t)slot * g.npos + (size_t)plane_id * 36 * g.nfo;
    for (int to = to0; to < to1; ++to) {
        for (int fo = tid; fo < g.nfo; fo += kScoreThreads) {
            const int sc = kStage ? sync_score_plane<kBins, kFt4>(plane, g.nb, pitch, to, fo) : sync_score_plane<0, kFt4>(plane, g.nb, pitch, to, fo);
            scores[(to + 12) * g.nfo + fo] = (int16_t)sc;  // stored as int16_t in candidate_t
        }
    }
}

// FT8 fast path.  Every term of a score is a difference between a cell P[r][c] and one of its four neighbours, and WHICH
// neighbours take part depends only on the Costas index k (k = 0 has no earlier symbol, k = 6 no later one, k = 3 is tone 0
// and has no lower neighbour; tone 7 never occurs) and on the row being the first/last of the waterfall.  So the CTA first
// turns its tile of the plane into four 16-bit planes, with a = P - P[c+1], b = P - P[c-1], u = r > 0 ? P - P[r-1] : 0,
// d = r + 1 < nb ? P - P[r+1] : 0:
//     W0 = a + b + d   (k = 0)     Wm = a + b + u + d   (k = 1,2,4,5)     W3 = a + u + d   (k = 3)     W6 = a + b + u   (k = 6)
// each stored with a bias of +1024 (so two cells are computed per 32-bit integer operation without borrows between the halves)
// and with 12 rows before and >= 10 rows after the waterfall holding the biased zero: the reference's `row < 0 -> continue` and
// `row >= nb -> break` become zero contributions, and a score is 21 loads at compile-time offsets plus 21 adds, minus 21 * 1024.
// The number of terms it is divided by depends on the time offset only (exact truncating division by multiply-high).
// The sums are the same integers as the reference's, so the scores are too.
// grid = (planes * tiles, slots); a tile is kTileF frequency offsets (all 36 time offsets, all rows).
constexpr int kTileF = 64;
constexpr int kTilePitch = kTileF + 8;   // 16-bit elements per derived row (71 needed: fo .. fo + 6), 18 groups of 4
constexpr int kRawPitch = kTileF + 16;   // raw bytes per row: columns fo0 - 4 .. fo0 + 75, whole aligned words
constexpr int kPadBefore = 12;           // time offsets start at -12
constexpr uint32_t kBias2 = 0x04000400u; // biased zero, two cells
__host__ __device__ constexpr int fast_rows(int nb) { return kPadBefore + (nb > 102 ? nb : 102); }  // last row touched: 23 + 72 + 6

// two 16-bit lanes per word, every lane stays non-negative: a' = P + 256 - N in [1, 511]
__device__ __forceinline__ uint32_t diff2(uint32_t p2, uint32_t n2) { return p2 + 0x01000100u - n2; }

__global__ void __launch_bounds__(kScoreThreads)
sync_score_ft8_kernel(const uint8_t *__restrict__ mag_all, size_t slot_stride, Geo g, int tiles, int16_t *__restrict__ scores_all) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint32_t s_magic[36];
    const int tid = threadIdx.x, slot = blockIdx.y;
    const int plane_id = blockIdx.x / tiles, tile = blockIdx.x - plane_id * tiles;
    const int fo0 = tile * kTileF;
    const int nb = g.nb, nbins = g.nbins;
    const int rows = fast_rows(nb);
    uint8_t *raw = smem;                                                   // [nb][kRawPitch]
    uint16_t *w0 = reinterpret_cast<uint16_t *>(smem + (((size_t)nb * kRawPitch + 15) & ~(size_t)15));
    const int plane_elems = rows * kTilePitch;
    const uint8_t *gplane = mag_all + (size_t)slot * slot_stride + (size_t)plane_id * nbins;  // row r at gplane + r*stride

    // raw tile: words covering columns [fo0 - 4, fo0 + 76), zero outside [0, nbins)
    constexpr int kWordsPerRow = kRawPitch / 4;
    for (int v = tid; v < nb * kWordsPerRow; v += kScoreThreads) {
        const int r = v / kWordsPerRow, wi = v - r * kWordsPerRow;
        const int c = fo0 - 4 + 4 * wi;
        uint32_t word = 0;
        if (c >= 0 && c + 4 <= nbins) word = __ldg(reinterpret_cast<const uint32_t *>(gplane + (size_t)r * g.stride + c));
        reinterpret_cast<uint32_t *>(raw + (size_t)r * kRawPitch)[wi] = word;
    }
    if (tid < 36) {  // number of terms of a score at time offset to = tid - 12 (ft8_sync_score's num_average) -> 2^32 / terms + 1
        const int to = tid - 12;
        int terms = 0;
        for (int grp = 0; grp < 3; ++grp)
            for (int k = 0; k < 7; ++k) {
                const int row = to + 36 * grp + k;
                if (row < 0) continue;
                if (row >= nb) break;
                terms += (k == 3 ? 1 : 2) + ((k > 0 && row > 0) ? 1 : 0) + ((k < 6 && row + 1 < nb) ? 1 : 0);
            }
        s_magic[tid] = terms > 1 ? (0xffffffffu / (uint32_t)terms + 1u) : 0u;  // 0: divide by 1 (or nothing to divide)
    }
    __syncthreads();
    constexpr int kGroups = kTilePitch / 4;  // 18 groups of 4 cells per row
    for (int v = tid; v < rows * kGroups; v += kScoreThreads) {
        const int rp = v / kGroups, q = v - rp * kGroups;
        const int r = rp - kPadBefore;
        uint2 o0, om, o3, o6;
        if (r < 0 || r >= nb) {
            o0 = om = o3 = o6 = make_uint2(kBias2, kBias2);
        } else {
            const uint32_t *row = reinterpret_cast<const uint32_t *>(raw + (size_t)r * kRawPitch) + q;  // words: [0] left, [1] the 4 cells, [2] right
            const uint32_t wl = row[0], wc = row[1], wr = row[2];
            const uint32_t p01 = __byte_perm(wc, 0u, 0x4140), p23 = __byte_perm(wc, 0u, 0x4342);    // (P0,P1) (P2,P3) as 16-bit lanes
            const uint32_t l01 = __byte_perm(p01, wl, 0x1017), l23 = __byte_perm(wc, 0u, 0x4241);   // (P-1,P0) (P1,P2); p01/p23 supply the zero bytes
            const uint32_t r23 = __byte_perm(p23, wr, 0x1412);                                      // (P3,P4)
            const uint32_t a01 = diff2(p01, l23), a23 = diff2(p23, r23);   // P - right neighbour (right of (P0,P1) is (P1,P2))
            const uint32_t b01 = diff2(p01, l01), b23 = diff2(p23, l23);   // P - left neighbour
            uint32_t u01 = 0x01000100u, u23 = 0x01000100u, d01 = 0x01000100u, d23 = 0x01000100u;
            if (r > 0) {
                const uint32_t wu = row[-kWordsPerRow + 1];
                u01 = diff2(p01, __byte_perm(wu, 0u, 0x4140)); u23 = diff2(p23, __byte_perm(wu, 0u, 0x4342));
            }
            if (r + 1 < nb) {
                const uint32_t wd = row[kWordsPerRow + 1];
                d01 = diff2(p01, __byte_perm(wd, 0u, 0x4140)); d23 = diff2(p23, __byte_perm(wd, 0u, 0x4342));
            }
            const uint32_t k1 = 0x01000100u;
            const uint32_t ab01 = a01 + b01, ab23 = a23 + b23, ud01 = u01 + d01, ud23 = u23 + d23;
            o0 = make_uint2(ab01 + d01 + k1, ab23 + d23 + k1);
            om = make_uint2(ab01 + ud01, ab23 + ud23);
            o3 = make_uint2(a01 + ud01 + k1, a23 + ud23 + k1);
            o6 = make_uint2(ab01 + u01 + k1, ab23 + u23 + k1);
        }
        uint2 *dst = reinterpret_cast<uint2 *>(w0 + (size_t)rp * kTilePitch) + q;
        dst[0] = o0;
        dst[plane_elems / 4] = om;
        dst[2 * (plane_elems / 4)] = o3;
        dst[3 * (plane_elems / 4)] = o6;
    }
    __syncthreads();
    const int fl = tid & (kTileF - 1), fo = fo0 + fl;
    if (fo >= g.nfo) return;
    int16_t *scores = scores_all + (size_t)slot * g.npos + (size_t)plane_id * 36 * g.nfo;
    constexpr int kCostas8[7] = {3, 1, 4, 0, 6, 5, 2};
    for (int ti = tid / kTileF; ti < 36; ti += kScoreThreads / kTileF) {  // warp-uniform time offset to = ti - 12: padded row = ti + 36 grp + k
        const uint16_t *b0 = w0 + ti * kTilePitch + fl, *bm = b0 + plane_elems, *b3 = bm + plane_elems, *b6 = b3 + plane_elems;
        int sum = 0;
#pragma unroll
        for (int grp = 0; grp < 3; ++grp) {
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const uint16_t *w = k == 0 ? b0 : (k == 3 ? b3 : (k == 6 ? b6 : bm));
                sum += w[(36 * grp + k) * kTilePitch + kCostas8[k]];
            }
        }
        int score = sum - 21 * 1024;
        const uint32_t magic = s_magic[ti];
        if (magic) {  // truncating division by the number of terms
            const uint32_t mag_q = __umulhi((uint32_t)(score < 0 ? -score : score), magic);
            score = score < 0 ? -(int)mag_q : (int)mag_q;
        }
        scores[ti * g.nfo + fo] = (int16_t)score;
    }
}

// Phase 2: one CTA per slot (looping over slots): ordered compaction of the positions with score >= min_score, then the
// exact heap replay, then the candidates are appended to the decode work list.
//
// The replay is the reference's loop (decode.c:198-231): push while the heap has room; once it is full a survivor enters
// only if its score is STRICTLY above the root's, evicting the root.  A survivor that fails that test is a no-op in the
// reference, so warp 0 tests 32 survivors at a time against the current root score and skips the failures wholesale: the
// serial work is the pushes/evictions that really happen (about K (1 + ln(survivors / K)) on noise), not one step per
// survivor -- 35 856 of them with min_score = 0, 137 232 on the 12 kHz waterfall.  Position -> (time_sub, freq_sub,
// time_offset, freq_offset) is decoded after the sort by all threads, off the serial path.
__global__ void __launch_bounds__(kSyncThreads)
sync_select_kernel(const int16_t *__restrict__ scores_all, int n_slots, Geo g, int max_cand, int min_score, candidate_t *__restrict__ cand_out,
                   int *__restrict__ ncand_out, uint32_t *__restrict__ scratch_all, uint32_t *__restrict__ work, unsigned int *__restrict__ work_total,
                   int mask_bytes) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *gmask = smem + (((size_t)max_cand * 4 + 15) & ~(size_t)15);   // one pass mask per group of 8 positions (mask_bytes > 0)
    // the first kSurvSmem survivors are ALSO kept in shared memory: the replay below is one thread reading them one after the other,
    // and from the global scratch list each read was an L2 round trip on the kernel's critical path
    uint32_t *surv = reinterpret_cast<uint32_t *>(gmask + (((size_t)(mask_bytes > 0 ? mask_bytes : 0) + 15) & ~(size_t)15));
    __shared__ int s_warp_cnt[2][32];
    __shared__ int s_total, s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *heap = reinterpret_cast<uint32_t *>(smem);
    uint32_t *scratch = scratch_all + (size_t)blockIdx.x * g.npos;

    for (int slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
        const int16_t *scores = scores_all + (size_t)slot * g.npos;
        // Ordered compaction with ONE block barrier: warp w owns the contiguous positions [w*span, (w+1)*span); it counts
        // its survivors, the warp totals are prefix-summed, then it writes its survivors at its offset (position order ==
        // the reference's loop order).  Scores are re-read in the second sweep (L1/L2 hits).  A lane reads EIGHT consecutive
        // scores per step (one 128-bit load) when the slot's score array allows it: the sweeps are chains of dependent L2
        // loads, and with one score per lane and step they were the kernel's whole duration (37 of its 42 us at 15 survivors).
        constexpr int kSelWarps = kSyncThreads / 32;
        int running, n_pass;
        if (mask_bytes > 0 && (g.npos & 7) == 0 && (((size_t)scores) & 15) == 0) {
            // One sweep over the scores (loads batched six deep) leaves a pass mask per group in shared memory; the second pass reads
            // the masks and fetches scores only for the few groups that hold a survivor.
            const int n8 = g.npos >> 3;                                            // groups of 8 positions
            const int span8 = (n8 + kSelWarps - 1) / kSelWarps;                    // groups per warp
            const int g0 = warp * span8, g1 = (g0 + span8 < n8) ? g0 + span8 : n8;
            const uint4 *sv = reinterpret_cast<const uint4 *>(scores);
            auto pass_mask = [&](const uint4 v) -> unsigned {                      // bit j: score j of the group passes
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                unsigned m = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if ((int)(short)(w[j] & 0xffffu) >= min_score) m |= 1u << (2 * j);
                    if ((int)(short)(w[j] >> 16) >= min_score) m |= 2u << (2 * j);
                }
                return m;
            };
            int mine = 0;
            for (int qb = g0 + lane; qb < g1; qb += 32 * 6) {
                uint4 v[6];
#pragma unroll
                for (int u = 0; u < 6; ++u) v[u] = (qb + 32 * u < g1) ? __ldg(sv + qb + 32 * u) : make_uint4(0x80008000u, 0x80008000u, 0x80008000u, 0x80008000u);
#pragma unroll
                for (int u = 0; u < 6; ++u) {
                    if (qb + 32 * u < g1) {
                        const unsigned m = pass_mask(v[u]);
                        gmask[qb + 32 * u] = (uint8_t)m;
                        mine += __popc(m);
                    }
                }
            }
            mine = __reduce_add_sync(0xffffffffu, mine);
            if (lane == 0) s_warp_cnt[0][warp] = mine;
            __syncthreads();
            const int cnt = lane < kSelWarps ? s_warp_cnt[0][lane] : 0;
            running = __reduce_add_sync(0xffffffffu, lane < warp ? cnt : 0);
            n_pass = __reduce_add_sync(0xffffffffu, cnt);
            if (mine > 0) {                                                        // warp-uniform: most warps hold no survivor at all
                for (int qb = g0; qb < g1; qb += 32) {                             // warp-uniform trip count
                    const int q = qb + lane;
                    unsigned m = q < g1 ? gmask[q] : 0u;
                    if (__ballot_sync(0xffffffffu, m != 0) == 0) continue;
                    const int c = __popc(m);
                    int incl = c;                                                  // inclusive prefix over the lanes: lane order == position order
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int t2 = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t2; }
                    int at = running + incl - c;
                    while (m) {
                        const int j = __ffs((int)m) - 1;
                        m &= m - 1;
                        const int p = 8 * q + j;
                        const uint32_t ent = ((uint32_t)p << 12) | ((uint32_t)(int)scores[p] & 0xfffu);
                        if (at < kSurvSmem) surv[at] = ent;
                        scratch[at++] = ent;
                    }
                    running += __shfl_sync(0xffffffffu, incl, 31);
                }
            }
        } else {
            const int span = ((g.npos + kSelWarps - 1) / kSelWarps + 31) / 32 * 32;  // positions per warp, multiple of 32
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
            const int cnt = lane < kSelWarps ? s_warp_cnt[0][lane] : 0;
            running = __reduce_add_sync(0xffffffffu, lane < warp ? cnt : 0);
            n_pass = __reduce_add_sync(0xffffffffu, cnt);
#pragma unroll 4
            for (int o = 0; o < span; o += 32) {
                const int p = w0 + o + lane;
                int score = 0;
                bool pass = false;
                if (p < g.npos) { score = scores[p]; pass = score >= min_score; }
                const unsigned ballot = __ballot_sync(0xffffffffu, pass);
                if (pass) {
                    const int at = running + __popc(ballot & ((1u << lane) - 1u));
                    const uint32_t ent = ((uint32_t)p << 12) | ((uint32_t)score & 0xfffu);
                    if (at < kSurvSmem) surv[at] = ent;
                    scratch[at] = ent;
                }
                running += __popc(ballot);
            }
        }
        __syncthreads();

        if (warp == 0) {
            int n = 0;
            // (a) room in the heap: every survivor is pushed (one thread; nothing to skip)
            const int n_fill = n_pass < max_cand ? n_pass : max_cand;
            if (lane == 0)
                for (; n < n_fill; ++n) sift_up(heap, n + 1, n < kSurvSmem ? surv[n] : scratch[n]);
            n = n_fill;
            // (b) heap full: 32 survivors per step against the root score, which only ever rises
            if (n_pass > max_cand) {
                __syncwarp();
                int root = ent_score(heap[0]);
                for (int base = max_cand; base < n_pass; base += 32) {
                    const int e = base + lane;
                    const uint32_t v = e < n_pass ? (e < kSurvSmem ? surv[e] : scratch[e]) : 0u;
                    const int sc = ent_score(v);
                    unsigned todo = __ballot_sync(0xffffffffu, e < n_pass && sc > root);
                    while (todo) {
                        const int src = __ffs((int)todo) - 1;
                        const uint32_t vv = __shfl_sync(0xffffffffu, v, src);
                        if (lane == 0) {  // pop the root (last element to the root, sift down), push the newcomer (decode.c:203-216)
                            heap[0] = heap[max_cand - 1];
                            sift_down(heap, max_cand - 1);
                            sift_up(heap, max_cand, vv);
                            root = ent_score(heap[0]);
                        }
                        root = __shfl_sync(0xffffffffu, root, 0);
                        todo &= ~((2u << src) - 1u);                                   // lanes up to src are settled
                        todo &= __ballot_sync(0xffffffffu, e < n_pass && sc > root);   // the others face the new root
                    }
                }
            }
            if (lane == 0) {
                for (int rest = n; rest > 1;) {  // heap sort -> descending score (decode.c:219-231)
                    const uint32_t t = heap[rest - 1]; heap[rest - 1] = heap[0]; heap[0] = t;
                    --rest;
                    sift_down(heap, rest);
                }
                s_total = n;
                ncand_out[slot] = n;
                s_base = (work && n > 0) ? (int)atomicAdd(work_total, (unsigned int)n) : 0;
            }
        }
        __syncthreads();
        {
            const int n = s_total;
            unsigned long long *dst = reinterpret_cast<unsigned long long *>(cand_out + (size_t)slot * max_cand);
            for (int k = tid; k < max_cand; k += kSyncThreads) {
                unsigned long long c = 0ull;
                if (k < n) {  // candidate_t as one 64-bit word: score | time_offset<<16 | freq_offset<<32 | time_sub<<48 | freq_sub<<56
                    const uint32_t v = heap[k];
                    const int score = ent_score(v), p = (int)(v >> 12);
                    const int fo = p % g.nfo;
                    int q = p / g.nfo;
                    const int to = q % 36 - 12;
                    q /= 36;
                    const int fs = q % g.fosr, ts = q / g.fosr;
                    c = ((unsigned long long)(uint16_t)(short)score) | ((unsigned long long)(uint16_t)(short)to << 16) |
                        ((unsigned long long)(uint16_t)(short)fo << 32) | ((unsigned long long)(uint8_t)ts << 48) | ((unsigned long long)(uint8_t)fs << 56);
                }
                dst[k] = c;
            }
            if (work)
                for (int k = tid; k < n; k += kSyncThreads) work[s_base + k] = (uint32_t)slot * (uint32_t)max_cand + (uint32_t)k;
        }
        __syncthreads();
    }
}

}  // namespace

cudaError_t launch_find_sync(const uint8_t *d_mag, size_t slot_stride, int n_slots, int num_blocks, int num_bins, int time_osr, int freq_osr,
                             int protocol, int max_cand, int min_score, candidate_t *d_cand, int *d_ncand, int16_t *d_scores, uint32_t *d_scratch,
                             int scratch_slots, uint32_t *d_work, unsigned int *d_work_total, int sm_count, cudaStream_t st, int *launches) {
    Geo g;
    g.nb = num_blocks; g.nbins = num_bins; g.tosr = time_osr; g.fosr = freq_osr;
    g.stride = time_osr * freq_osr * num_bins;
    g.nfo = num_bins - 7;
    g.npos = time_osr * freq_osr * 36 * g.nfo;
    // one CTA per (slot, sub-plane, share of the 36 time offsets); enough CTAs to fill the machine for small batches
    const int planes = time_osr * freq_osr;
    int splits = (4 * sm_count + n_slots * planes - 1) / (n_slots * planes);
    if (splits > 6) splits = 6;
    if (splits < 1) splits = 1;
    const int to_per_cta = (36 + splits - 1) / splits;
    splits = (36 + to_per_cta - 1) / to_per_cta;
    dim3 grid(planes * splits, n_slots);
    const size_t staged = ((size_t)g.nb * g.nbins + 15) & ~(size_t)15;
    const bool ft4 = protocol == PROTO_FT4;
    const size_t fast_smem = (((size_t)g.nb * kRawPitch + 15) & ~(size_t)15) + (size_t)4 * fast_rows(g.nb) * kTilePitch * sizeof(int16_t);
    if (!ft4 && fast_smem <= 96 * 1024 && (((size_t)d_mag | slot_stride | (size_t)g.stride | (size_t)g.nbins) & 3) == 0) {
        // FT8 fast path: derived int16 planes per (plane, tile of 64 frequency offsets)
        const int tiles = (g.nfo + kTileF - 1) / kTileF;
        cudaError_t e = cudaFuncSetAttribute(sync_score_ft8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        if (e != cudaSuccess) return e;
        sync_score_ft8_kernel<<<dim3(planes * tiles, n_slots), kScoreThreads, fast_smem, st>>>(d_mag, slot_stride, g, tiles, d_scores);
    } else if (staged <= 200 * 1024) {
        if (g.nbins == 256 && !ft4) {
            sync_score_kernel<256, true, false><<<grid, kScoreThreads, staged, st>>>(d_mag, slot_stride, g, splits, to_per_cta, d_scores);
        } else {
            auto kern = ft4 ? sync_score_kernel<0, true, true> : sync_score_kernel<0, true, false>;
            if (staged > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged);
                if (e != cudaSuccess) return e;
            }
            kern<<<grid, kScoreThreads, staged, st>>>(d_mag, slot_stride, g, splits, to_per_cta, d_scores);
        }
    } else {
        auto kern = ft4 ? sync_score_kernel<0, false, true> : sync_score_kernel<0, false, false>;
        kern<<<grid, kScoreThreads, 0, st>>>(d_mag, slot_stride, g, splits, to_per_cta, d_scores);
    }
    ++*launches;
    if (d_work_total) {
        cudaError_t e = cudaMemsetAsync(d_work_total, 0, 4 * sizeof(unsigned int), st);  // [0] items, [1] next item (decode_kernel)
        if (e != cudaSuccess) return e;
    }
    const int sgrid = n_slots < scratch_slots ? n_slots : scratch_slots;
    // heap words + (when they fit under the default 48 KB) one pass-mask byte per group of 8 positions
    const size_t heap_bytes = ((size_t)max_cand * 4 + 15) & ~(size_t)15;
    int mask_bytes = (g.npos & 7) == 0 ? g.npos >> 3 : 0;
    const size_t surv_bytes = (size_t)kSurvSmem * sizeof(uint32_t);
    if (heap_bytes + (size_t)mask_bytes + 16 + surv_bytes > 48 * 1024) mask_bytes = 0;
    sync_select_kernel<<<sgrid, kSyncThreads, heap_bytes + (((size_t)mask_bytes + 15) & ~(size_t)15) + surv_bytes, st>>>(d_scores, n_slots, g, max_cand, min_score, d_cand, d_ncand, d_scratch,
                                                                                    d_work, d_work_total, mask_bytes);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace ft8b200
