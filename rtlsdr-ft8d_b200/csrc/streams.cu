// streams.cu -- receiver streams: rtlsdr_callback() with persistent decimator state and the daemon's
// double-buffered 15 s slots.  Replaces the function-static state of rtlsdr_callback()
// (/root/reference/rtlsdr_ft8d.c:80-86,113-114), rx_state's double buffer (rtlsdr_ft8d.h:82-98), the buffer
// flip of main() (:1339-1354) and decoder()'s skip/condition/decode sequence (:221-285).
//
// Host-side logic only; the arithmetic is the decimator kernels.  A stream owns a device ring of raw bytes.
// rtlsdr_callback() copies the caller's buffer into pinned staging (the caller may reuse it on return) and
// queues an async H2D copy.  Decimation is deferred to "pump" time (buffer flip, fetch, or ring nearly full):
// all complete 751-sample blocks received so far are reduced -- aligned super-blocks by the fast kernel, the
// ragged edges by the generic one -- and comb+FIR appends their outputs to the slot being filled.  The only
// state carried between pumps is the last 59 block sums (56 FIR taps + 3 comb delays) and the < 1 block of
// unconsumed raw bytes; no integrator state is needed because the closed form has none.
#include "common.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

using namespace ft8b200;

namespace ft8b200 {
ft8b200_ctx_t *default_ctx();
void default_stream_release();
}

namespace {
constexpr size_t kSuper = 12016;                       // bytes per super-block (8 blocks)
constexpr size_t kBlockBytes = 1502;
constexpr size_t kRingBytes = 80ull * 1000 * 1000;     // one 72 MB slot + slack, multiple of 16
constexpr int kStage = 16;                             // pinned staging buffers
constexpr size_t kStageBytes = 1 << 20;                // each 1 MiB (a callback delivers 64 KiB)
constexpr int kMaxPumpBlocks = (int)(kRingBytes / kBlockBytes) + 8;

void die(const char *what, cudaError_t e) {
    fprintf(stderr, "libft8b200: stream: %s failed: %s\n", what, cudaGetErrorString(e));
    abort();
}
#define CK(call)                                   \
    do {                                           \
        cudaError_t e__ = (call);                  \
        if (e__ != cudaSuccess) die(#call, e__);   \
    } while (0)
}  // namespace

struct ft8b200_stream {
    ft8b200_ctx_t *ctx = nullptr;
    int device = 0;
    cudaStream_t st = nullptr;
    uint8_t *d_ring = nullptr;          // bytes [ring_base, bytes_in) of the stream live at d_ring[0 ...]
    uint64_t ring_base = 0;             // global byte offset of d_ring[0]; always a multiple of kSuper
    uint64_t bytes_in = 0;              // bytes received so far
    uint64_t blocks_done = 0;           // blocks already reduced + filtered
    BlockSums *d_sums = nullptr;        // [kHistBlocks history][new blocks of this pump]
    float *d_i[2] = {nullptr, nullptr}, *d_q[2] = {nullptr, nullptr};
    float *d_peak = nullptr;            // [2]
    uint32_t iq_index[2] = {0, 0};      // rx_state.iqIndex
    int buffer_index = 0;               // rx_state.bufferIndex
    uint8_t *h_stage[kStage] = {};
    cudaEvent_t ev[kStage] = {};
    int stage_next = 0;
    int launches = 0;
    std::mutex mu;
};

namespace {

std::mutex g_default_mu;
ft8b200_stream_t *g_default_stream = nullptr;

void pump(ft8b200_stream_t *s) {
    const uint64_t blocks_total = s->bytes_in / kBlockBytes;
    if (blocks_total <= s->blocks_done) return;
    const uint64_t b0 = s->blocks_done, b1 = blocks_total;
    const int n_new = (int)(b1 - b0);
    BlockSums *out = s->d_sums + kHistBlocks;  // entry of block b0
    // aligned super-blocks [a0, a1) go through the 128-bit kernel, the edges through the generic one
    uint64_t a0 = (b0 + 7) / 8 * 8, a1 = b1 / 8 * 8;
    if (a1 <= a0) { a0 = b1; a1 = b1; }
    auto ptr = [&](uint64_t block) { return s->d_ring + (block * kBlockBytes - s->ring_base); };
    if (a0 > b0)
        CK(launch_cic_block_sums_generic(ptr(b0), 0, 1, (uint32_t)((b0 * kDecim) & 3u), (int)(a0 - b0), out, 0, s->st, &s->launches));
    if (a1 > a0)
        CK(launch_cic_block_sums(ptr(a0), 0, 1, (int)(a1 - a0), out + (a0 - b0), 0, 0, 0, s->st, &s->launches));
    if (b1 > a1)
        CK(launch_cic_block_sums_generic(ptr(a1), 0, 1, (uint32_t)((a1 * kDecim) & 3u), (int)(b1 - a1), out + (a1 - b0), 0, s->st, &s->launches));
    const int buf = s->buffer_index;
    CK(launch_cic_comb_fir(out, 0, n_new, (int)s->iq_index[buf], false, 1, s->d_i[buf], s->d_q[buf], nullptr, s->d_peak + buf, nullptr,
                           s->st, &s->launches));
    CK(launch_shift_history(s->d_sums, n_new, s->st, &s->launches));
    const uint64_t idx = (uint64_t)s->iq_index[buf] + (uint64_t)n_new;
    s->iq_index[buf] = (uint32_t)(idx < (uint64_t)kSlot ? idx : (uint64_t)kSlot);  // the reference stops counting at 48000 (:196-200)
    s->blocks_done = b1;
    // drop consumed super-blocks from the ring: keep bytes from the super-block that holds the next block
    const uint64_t keep_from = (b1 * kBlockBytes) / kSuper * kSuper;
    if (keep_from > s->ring_base) {
        const size_t live = (size_t)(s->bytes_in - keep_from);
        const size_t shift = (size_t)(keep_from - s->ring_base);
        if (live > 0) {
            if (shift >= live) {
                CK(cudaMemcpyAsync(s->d_ring, s->d_ring + shift, live, cudaMemcpyDeviceToDevice, s->st));
            } else {  // overlapping move (only when a pump consumed less than it kept): go through the tail of the ring
                CK(cudaMemcpyAsync(s->d_ring + kRingBytes - live, s->d_ring + shift, live, cudaMemcpyDeviceToDevice, s->st));
                CK(cudaMemcpyAsync(s->d_ring, s->d_ring + kRingBytes - live, live, cudaMemcpyDeviceToDevice, s->st));
            }
        }
        s->ring_base = keep_from;
    }
}

void append(ft8b200_stream_t *s, const unsigned char *samples, uint32_t count) {
    uint32_t done = 0;
    while (done < count) {
        if ((size_t)(s->bytes_in - s->ring_base) + kStageBytes + 2 * kSuper > kRingBytes) pump(s);
        const uint32_t n = (count - done) < kStageBytes ? (count - done) : (uint32_t)kStageBytes;
        const int k = s->stage_next;
        CK(cudaEventSynchronize(s->ev[k]));  // the previous copy out of this staging buffer has finished
        memcpy(s->h_stage[k], samples + done, n);
        CK(cudaMemcpyAsync(s->d_ring + (s->bytes_in - s->ring_base), s->h_stage[k], n, cudaMemcpyHostToDevice, s->st));
        CK(cudaEventRecord(s->ev[k], s->st));
        s->stage_next = (k + 1) % kStage;
        s->bytes_in += n;
        done += n;
    }
}

ft8b200_stream_t *default_stream() {
    std::lock_guard<std::mutex> lk(g_default_mu);
    if (!g_default_stream) {
        g_default_stream = ft8b200_stream_create(default_ctx());
        if (!g_default_stream) {
            fprintf(stderr, "libft8b200: cannot create the default receiver stream: %s\n", ft8b200_last_error());
            abort();
        }
    }
    return g_default_stream;
}

}  // namespace

namespace ft8b200 {
// freeFFTW() tears the default context down: the default stream lives on it and must go first
void default_stream_release() {
    std::lock_guard<std::mutex> lk(g_default_mu);
    if (g_default_stream) {
        ft8b200_stream_destroy(g_default_stream);
        g_default_stream = nullptr;
    }
}
}  // namespace ft8b200

extern "C" {

ft8b200_stream_t *ft8b200_stream_create(ft8b200_ctx_t *ctx) {
    if (!ctx) return nullptr;
    ft8b200_stream_t *s = new ft8b200_stream();
    s->ctx = ctx;
    s->st = (cudaStream_t)ft8b200_cuda_stream(ctx);
    s->device = ctx_device(ctx);   // the stream lives on its context's device, whatever the caller's current device is
    CK(cudaSetDevice(s->device));
    CK(upload_fir_constants());
    CK(cudaMalloc(&s->d_ring, kRingBytes + 16));
    CK(cudaMalloc(&s->d_sums, sizeof(BlockSums) * (size_t)(kHistBlocks + kMaxPumpBlocks)));
    CK(cudaMemsetAsync(s->d_sums, 0, sizeof(BlockSums) * kHistBlocks, s->st));  // zero filter state
    for (int b = 0; b < 2; ++b) {
        CK(cudaMalloc(&s->d_i[b], sizeof(float) * kSlot));
        CK(cudaMalloc(&s->d_q[b], sizeof(float) * kSlot));
        CK(cudaMemsetAsync(s->d_i[b], 0, sizeof(float) * kSlot, s->st));
        CK(cudaMemsetAsync(s->d_q[b], 0, sizeof(float) * kSlot, s->st));
    }
    CK(cudaMalloc(&s->d_peak, 2 * sizeof(float)));
    CK(cudaMemsetAsync(s->d_peak, 0, 2 * sizeof(float), s->st));
    for (int k = 0; k < kStage; ++k) {
        CK(cudaHostAlloc(&s->h_stage[k], kStageBytes, cudaHostAllocDefault));
        CK(cudaEventCreateWithFlags(&s->ev[k], cudaEventDisableTiming));
    }
    return s;
}

void ft8b200_stream_destroy(ft8b200_stream_t *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->st);
    cudaFree(s->d_ring); cudaFree(s->d_sums); cudaFree(s->d_peak);
    for (int b = 0; b < 2; ++b) { cudaFree(s->d_i[b]); cudaFree(s->d_q[b]); }
    for (int k = 0; k < kStage; ++k) { cudaFreeHost(s->h_stage[k]); cudaEventDestroy(s->ev[k]); }
    delete s;
}

void rtlsdr_callback(unsigned char *samples, uint32_t samples_count, void *ctx) {
    ft8b200_stream_t *s = ctx ? (ft8b200_stream_t *)ctx : default_stream();
    if (samples_count & 7u) {
        // the reference's mixer walks the buffer 8 bytes at a time (rtlsdr_ft8d.c:129) and would overrun it
        fprintf(stderr, "libft8b200: rtlsdr_callback: samples_count %u is not a multiple of 8; buffer dropped\n", samples_count);
        return;
    }
    std::lock_guard<std::mutex> lk(s->mu);
    CK(cudaSetDevice(s->device));  // librtlsdr calls from its own thread, whose current device is 0 until told otherwise
    append(s, samples, samples_count);
}

// For the four calls below s == NULL means the process-wide default stream, the one rtlsdr_callback(..., ctx = NULL)
// feeds -- which is how the reference's daemon registers its callback (rtlsdr_ft8d.c:214).
uint32_t ft8b200_stream_count(ft8b200_stream_t *s) {
    if (!s) s = default_stream();
    std::lock_guard<std::mutex> lk(s->mu);
    // outputs the reference would have stored so far = complete blocks received since the slot started
    const uint64_t pending = s->bytes_in / kBlockBytes - s->blocks_done;
    const uint64_t idx = (uint64_t)s->iq_index[s->buffer_index] + pending;
    return (uint32_t)(idx < (uint64_t)kSlot ? idx : (uint64_t)kSlot);
}

int ft8b200_stream_flip(ft8b200_stream_t *s) {
    if (!s) s = default_stream();
    std::lock_guard<std::mutex> lk(s->mu);
    CK(cudaSetDevice(s->device));
    pump(s);
    s->buffer_index ^= 1;
    const int b = s->buffer_index;
    s->iq_index[b] = 0;
    CK(cudaMemsetAsync(s->d_peak + b, 0, sizeof(float), s->st));
    // decoder() clears the tail of whatever it decodes (rtlsdr_ft8d.c:243-246); clearing the new buffer now is equivalent
    CK(cudaMemsetAsync(s->d_i[b], 0, sizeof(float) * kSlot, s->st));
    CK(cudaMemsetAsync(s->d_q[b], 0, sizeof(float) * kSlot, s->st));
    return 0;
}

int ft8b200_stream_fetch(ft8b200_stream_t *s, float *h_i, float *h_q, uint32_t *n_valid) {
    if (!h_i || !h_q) return FT8B200_BAD_ARG();
    if (!s) s = default_stream();
    std::lock_guard<std::mutex> lk(s->mu);
    CK(cudaSetDevice(s->device));
    const int prev = s->buffer_index ^ 1;
    CK(cudaMemcpyAsync(h_i, s->d_i[prev], sizeof(float) * kSlot, cudaMemcpyDeviceToHost, s->st));
    CK(cudaMemcpyAsync(h_q, s->d_q[prev], sizeof(float) * kSlot, cudaMemcpyDeviceToHost, s->st));
    CK(cudaStreamSynchronize(s->st));
    if (n_valid) *n_valid = s->iq_index[prev];
    return 0;
}

int ft8b200_stream_decode(ft8b200_stream_t *s, struct decoder_results *h_results, int32_t *h_nresults) {
    if (!h_results || !h_nresults) return FT8B200_BAD_ARG();
    if (!s) s = default_stream();
    int prev;
    {
        std::lock_guard<std::mutex> lk(s->mu);
        prev = s->buffer_index ^ 1;
        if (s->iq_index[prev] < (uint32_t)((15 - 3) * 3200)) {  // "Signal too short, skipping!", rtlsdr_ft8d.c:235-238
            *h_nresults = -1;
            return 0;
        }
    }
    // conditioning (0.5 / peak) is applied on load by the waterfall kernel; then sync, LDPC, spots
    int rc = ft8b200_process_conditioned(s->ctx, s->d_i[prev], s->d_q[prev], s->d_peak + prev, 1, s->st);
    if (rc) return rc;
    return ft8b200_fetch_results(s->ctx, 1, h_results, h_nresults, s->st);
}

}  // extern "C"
