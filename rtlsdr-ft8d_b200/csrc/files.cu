// files.cu -- the on-disk formats either side of the path, and whole recordings through it in batches.
//   * .iq  (float32 interleaved I, -Q; up to 48000 pairs)        ref: readRawIQfile,  rtlsdr_ft8d.c:744-783
//   * .c2  (14-byte name, int type, double dial frequency, .iq)   ref: readC2file,     rtlsdr_ft8d.c:810-856
//   * WAV  (RIFF, PCM s16 mono)                                    ref: load_wav,       ft8_lib/common/wave.c:66-128
//   * ft8b200_decode_iq_files:  decodeRecordedFile() (rtlsdr_ft8d.c:859-887) for a batch of files
//   * ft8b200_decode_audio / ft8b200_decode_wav_files:  decode_ft8's main() (ft8_lib/decode_ft8.c:272-409) for a batch
// Host code here only parses files and formats records.  Every numeric step of the path runs on the device: the
// readers return the samples UNSCALED together with their peak, and the reference's "normalise @ -3 dB" is applied by the
// waterfall kernel on load (same float expression: sample * (float)(0.5 / max(1e-24f, peak))); s16 -> float is x / 32768.0f,
// exact in binary, done on the device as well.
#include "common.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <mutex>
#include <vector>

using namespace ft8b200;

namespace {

__global__ void s16_to_float_kernel(const int16_t *__restrict__ in, float *__restrict__ out, size_t n) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = __fdiv_rn((float)in[k], 32768.0f);  // wave.c:118-121
}

// float pairs (I, -Q) -> planar I, Q and max(|I|, |Q|); returns the number of pairs
int deinterleave(const float *buf, size_t nread, float *h_i, float *h_q, float *peak) {
    const int rec = (int)(nread / 2);
    float m = 0.0f;
    for (int k = 0; k < rec; ++k) {
        const float a = buf[2 * k], b = -buf[2 * k + 1];  // "neg, convention used by wsprsim"
        h_i[k] = a;
        h_q[k] = b;
        const float fa = fabsf(a), fb = fabsf(b);
        if (fa > m) m = fa;
        if (fb > m) m = fb;
    }
    for (int k = rec; k < kSlot; ++k) { h_i[k] = 0.0f; h_q[k] = 0.0f; }
    if (peak) *peak = m;
    return rec;
}

bool ends_with(const char *s, const char *suffix) {
    const size_t a = strlen(s), b = strlen(suffix);
    return a >= b && strcmp(s + a - b, suffix) == 0;
}

// Scratch device buffers of one call, from the device's stream-ordered pool: after the first call the pool holds the
// memory (release threshold raised below), so a batch does not pay cudaMalloc/cudaFree -- which also synchronise the whole
// device -- thirteen times over.
struct DevFree {
    cudaStream_t st;
    std::vector<void *> ptrs;
    explicit DevFree(cudaStream_t s) : st(s) {
        static std::once_flag once[64];
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64)
            std::call_once(once[dev], [dev] {
                cudaMemPool_t pool;
                unsigned long long keep = ~0ull;
                if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            });
    }
    ~DevFree() { for (void *p : ptrs) cudaFreeAsync(p, st); }
    template <class T> cudaError_t alloc(T **p, size_t bytes) {
        cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 1, st);
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

}  // namespace

extern "C" {

int ft8b200_read_iq_file(const char *path, float *h_i, float *h_q, float *peak) {
    if (!path || !h_i || !h_q) return 0;
    FILE *fd = fopen(path, "rb");
    if (!fd) return 0;  // the reference prints "Cannot open data file..." and returns 0
    std::vector<float> buf(2 * (size_t)kSlot);
    const size_t nread = fread(buf.data(), sizeof(float), buf.size(), fd);
    fclose(fd);
    return deinterleave(buf.data(), nread, h_i, h_q, peak);
}

int ft8b200_read_c2_file(const char *path, float *h_i, float *h_q, float *peak, double *dial_freq, int *type, char *name15) {
    if (!path || !h_i || !h_q) return 0;
    FILE *fd = fopen(path, "rb");
    if (!fd) return 0;
    char name[15] = {0};
    int ty = 0;
    double fr = 0.0;
    size_t got = fread(name, 1, 14, fd);
    got += fread(&ty, sizeof(int), 1, fd);
    got += fread(&fr, sizeof(double), 1, fd);
    (void)got;  // like the reference, a short header just yields a short (possibly empty) recording
    std::vector<float> buf(2 * (size_t)kSlot);
    const size_t nread = fread(buf.data(), sizeof(float), buf.size(), fd);
    fclose(fd);
    if (dial_freq) *dial_freq = fr;
    if (type) *type = ty;
    if (name15) memcpy(name15, name, 15);
    return deinterleave(buf.data(), nread, h_i, h_q, peak);
}

// Same contract and return codes as load_wav() (wave.c:66-128): -1 = not 16-bit mono PCM with a 16-byte fmt chunk,
// -2 = more samples than *num_samples; additionally -3 = cannot open / truncated (the reference would crash).
// signal may be NULL when raw_s16 is given: the samples are then returned unconverted (for the device-side conversion).
int ft8b200_load_wav_s16(int16_t *raw_s16, float *signal, int *num_samples, int *sample_rate, const char *path) {
    if (!path || !num_samples || !sample_rate || (!raw_s16 && !signal)) return -3;
    FILE *f = fopen(path, "rb");
    if (!f) return -3;
    char id[4];
    uint32_t chunk_size = 0, sub1 = 0, rate = 0, byte_rate = 0, sub2 = 0;
    uint16_t fmt = 0, channels = 0, align = 0, bits = 0;
    bool ok = fread(id, 4, 1, f) == 1 && fread(&chunk_size, 4, 1, f) == 1 && fread(id, 4, 1, f) == 1 && fread(id, 4, 1, f) == 1 &&
              fread(&sub1, 4, 1, f) == 1;
    if (!ok) { fclose(f); return -3; }
    if (sub1 != 16) { fclose(f); return -1; }
    ok = fread(&fmt, 2, 1, f) == 1 && fread(&channels, 2, 1, f) == 1 && fread(&rate, 4, 1, f) == 1 && fread(&byte_rate, 4, 1, f) == 1 &&
         fread(&align, 2, 1, f) == 1 && fread(&bits, 2, 1, f) == 1;
    if (!ok) { fclose(f); return -3; }
    if (fmt != 1 || channels != 1 || bits != 16) { fclose(f); return -1; }
    ok = fread(id, 4, 1, f) == 1 && fread(&sub2, 4, 1, f) == 1;
    if (!ok || align == 0) { fclose(f); return -3; }
    if ((long)(sub2 / align) > (long)*num_samples) { fclose(f); return -2; }
    const int n = (int)(sub2 / align);
    std::vector<int16_t> tmp;
    int16_t *dst = raw_s16;
    if (!dst) { tmp.resize((size_t)n); dst = tmp.data(); }
    if (align == 2) {
        const size_t got = fread(dst, 2, (size_t)n, f);
        for (size_t k = got; k < (size_t)n; ++k) dst[k] = 0;
    } else {
        // a header whose blockAlign is not 2 although it says mono 16-bit: the reference reads n * blockAlign bytes into a buffer of that
        // size and then takes its first n int16 (wave.c:104-121); the same here, through a buffer of our own (the caller's holds n samples)
        std::vector<uint8_t> bytes((size_t)n * (align > 2 ? align : 2), 0);
        const size_t got = fread(bytes.data(), align, (size_t)n, f);
        (void)got;   // what a short file did not deliver stays zero
        memcpy(dst, bytes.data(), (size_t)n * sizeof(int16_t));
    }
    fclose(f);
    if (signal)
        for (int k = 0; k < n; ++k) signal[k] = dst[k] / 32768.0f;
    *num_samples = n;
    *sample_rate = (int)rate;
    return 0;
}

int ft8b200_load_wav(float *signal, int *num_samples, int *sample_rate, const char *path) {
    return ft8b200_load_wav_s16(nullptr, signal, num_samples, sample_rate, path);
}

// decodeRecordedFile() for a batch: h_results = n x max_messages records, h_nresults / h_samples = n ints.
// Unreadable files and unknown extensions give 0 samples and 0 results (the reference prints a message and returns).
int ft8b200_decode_iq_files(ft8b200_ctx_t *ctx, const char *const *paths, int n, struct decoder_results *h_results, int32_t *h_nresults,
                            int32_t *h_samples) {
    if (!ctx || !paths || n < 1 || !h_results || !h_nresults) return FT8B200_BAD_ARG();
    if (cudaSetDevice(ctx_device(ctx)) != cudaSuccess) return FT8B200_CUDA_FAIL();  // the caller's current device may be another one
    std::vector<float> hi((size_t)n * kSlot), hq((size_t)n * kSlot), peak((size_t)n, 0.0f);
    for (int k = 0; k < n; ++k) {
        int rec = 0;
        if (paths[k] && ends_with(paths[k], ".iq")) rec = ft8b200_read_iq_file(paths[k], &hi[(size_t)k * kSlot], &hq[(size_t)k * kSlot], &peak[(size_t)k]);
        else if (paths[k] && ends_with(paths[k], ".c2")) rec = ft8b200_read_c2_file(paths[k], &hi[(size_t)k * kSlot], &hq[(size_t)k * kSlot], &peak[(size_t)k], nullptr, nullptr, nullptr);
        else { memset(&hi[(size_t)k * kSlot], 0, sizeof(float) * kSlot); memset(&hq[(size_t)k * kSlot], 0, sizeof(float) * kSlot); }
        if (h_samples) h_samples[k] = rec;
    }
    cudaStream_t st = (cudaStream_t)ft8b200_cuda_stream(ctx);
    DevFree pool(st);
    float *d_i = nullptr, *d_q = nullptr, *d_peak = nullptr;
    const size_t bytes = (size_t)n * kSlot * sizeof(float);
    if (pool.alloc(&d_i, bytes) != cudaSuccess || pool.alloc(&d_q, bytes) != cudaSuccess || pool.alloc(&d_peak, sizeof(float) * n) != cudaSuccess)
        return FT8B200_ENOMEM;
    if (cudaMemcpyAsync(d_i, hi.data(), bytes, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(d_q, hq.data(), bytes, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(d_peak, peak.data(), sizeof(float) * n, cudaMemcpyHostToDevice, st) != cudaSuccess)
        return FT8B200_CUDA_FAIL();
    int rc = ft8b200_process_conditioned(ctx, d_i, d_q, d_peak, n, nullptr);
    if (rc) return rc;
    rc = ft8b200_fetch_results(ctx, n, h_results, h_nresults, nullptr);
    if (rc) return rc;
    if (h_samples)
        for (int k = 0; k < n; ++k) if (h_samples[k] == 0) h_nresults[k] = 0;  // `if (samples_len)` guard, rtlsdr_ft8d.c:875
    return 0;
}

// decode_ft8 main() for n device-resident recordings (float audio, `stride` samples apart, n_samples each):
// monitor waterfall -> ft8_find_sync(120, min_score 10) -> ft8_decode(20 iterations) -> first-seen unique messages in
// candidate order (hash table of 50, decode_ft8.c:336-406).  h_out: n x max_messages entries, h_count: n.
// The context must have been created with max_candidates 120 / max_messages 50 to reproduce decode_ft8 exactly.
int ft8b200_decode_audio(ft8b200_ctx_t *ctx, const float *d_audio, size_t stride, int n_samples, int n, int sample_rate, int protocol,
                         ft8b200_decoded_t *h_out, int32_t *h_count, int max_out_per_recording) {
    if (!ctx || !d_audio || n < 1 || !h_out || !h_count || (protocol != PROTO_FT4 && protocol != PROTO_FT8)) return FT8B200_BAD_ARG();
    if (cudaSetDevice(ctx_device(ctx)) != cudaSuccess) return FT8B200_CUDA_FAIL();
    ft8b200_config_t cfg;
    if (ft8b200_get_config(ctx, &cfg)) return FT8B200_BAD_ARG();
    const int K = cfg.max_candidates, M = cfg.max_messages;
    if (max_out_per_recording < M) return FT8B200_BAD_ARG();
    const int tosr = 2, fosr = 2;  // kTime_osr, kFreq_osr (decode_ft8.c:27-28)
    const float symbol_period = (protocol == PROTO_FT4) ? 0.048f : 0.160f;
    const float slot_time = (protocol == PROTO_FT4) ? 7.5f : 15.0f;
    const int max_blocks = (int)(slot_time / symbol_period);
    const int num_bins = (int)(sample_rate * symbol_period / 2);
    const size_t bstride = (size_t)tosr * fosr * num_bins;
    const size_t mag_stride = ((size_t)max_blocks * bstride + 15) & ~(size_t)15;
    cudaStream_t st = (cudaStream_t)ft8b200_cuda_stream(ctx);
    DevFree pool(st);
    uint8_t *d_mag = nullptr, *d_ok = nullptr, *d_stage = nullptr;
    candidate_t *d_cand = nullptr;
    int *d_ncand = nullptr;
    decode_status_t *d_status = nullptr;
    message_t *d_msg = nullptr, *d_umsg = nullptr;
    struct decoder_results *d_res = nullptr;
    int32_t *d_nres = nullptr, *d_uscore = nullptr, *d_ucand = nullptr;
    float *d_ufreq = nullptr;
    const size_t S = (size_t)n;
    bool okc = pool.alloc(&d_mag, S * mag_stride) == cudaSuccess && pool.alloc(&d_cand, S * K * sizeof(candidate_t)) == cudaSuccess &&
               pool.alloc(&d_ncand, S * sizeof(int)) == cudaSuccess && pool.alloc(&d_ok, S * K) == cudaSuccess && pool.alloc(&d_stage, S * K) == cudaSuccess &&
               pool.alloc(&d_status, S * K * sizeof(decode_status_t)) == cudaSuccess && pool.alloc(&d_msg, S * K * sizeof(message_t)) == cudaSuccess &&
               pool.alloc(&d_umsg, S * M * sizeof(message_t)) == cudaSuccess && pool.alloc(&d_res, S * M * sizeof(struct decoder_results)) == cudaSuccess &&
               pool.alloc(&d_nres, S * sizeof(int32_t)) == cudaSuccess && pool.alloc(&d_uscore, S * M * sizeof(int32_t)) == cudaSuccess &&
               pool.alloc(&d_ucand, S * M * sizeof(int32_t)) == cudaSuccess && pool.alloc(&d_ufreq, S * M * sizeof(float)) == cudaSuccess;
    if (!okc) return FT8B200_ENOMEM;
    // read back whole below, written only up to each recording's count: defined bytes behind the counts
    if (cudaMemsetAsync(d_umsg, 0, S * M * sizeof(message_t), st) != cudaSuccess || cudaMemsetAsync(d_ucand, 0, S * M * sizeof(int32_t), st) != cudaSuccess ||
        cudaMemsetAsync(d_cand, 0, S * K * sizeof(candidate_t), st) != cudaSuccess)
        return FT8B200_CUDA_FAIL();
    int nb = 0, rc;
    if ((rc = ft8b200_monitor_waterfall(ctx, d_audio, stride, n_samples, n, sample_rate, tosr, fosr, protocol, d_mag, mag_stride, &nb, nullptr))) return rc;
    for (int k = 0; k < n; ++k) h_count[k] = 0;
    if (nb == 0) return 0;
    // explicit protocol: the one selected with ft8b200_set_protocol for the stage-wise API is neither used nor changed
    rc = find_sync_proto(ctx, protocol, d_mag, mag_stride, n, nb, num_bins, tosr, fosr, d_cand, d_ncand, nullptr);
    if (!rc) rc = decode_proto(ctx, protocol, d_mag, mag_stride, n, nb, num_bins, tosr, fosr, d_cand, d_ncand, d_ok, d_stage, d_status, d_msg, nullptr, nullptr, nullptr);
    if (!rc) rc = ft8b200_spots(ctx, n, fosr, d_cand, d_ncand, d_ok, d_msg, d_res, d_nres, d_umsg, d_ufreq, d_uscore, d_ucand, nullptr);
    if (rc) return rc;
    std::vector<message_t> umsg(S * M);
    std::vector<int32_t> ucand(S * M), nres(S);
    std::vector<candidate_t> cand(S * K);
    okc = cudaMemcpyAsync(umsg.data(), d_umsg, S * M * sizeof(message_t), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
          cudaMemcpyAsync(ucand.data(), d_ucand, S * M * sizeof(int32_t), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
          cudaMemcpyAsync(nres.data(), d_nres, S * sizeof(int32_t), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
          cudaMemcpyAsync(cand.data(), d_cand, S * K * sizeof(candidate_t), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
          cudaStreamSynchronize(st) == cudaSuccess;
    if (!okc) return FT8B200_CUDA_FAIL();
    for (int s = 0; s < n; ++s) {
        h_count[s] = nres[(size_t)s];
        for (int k = 0; k < nres[(size_t)s] && k < M; ++k) {
            const candidate_t &c = cand[(size_t)s * K + ucand[(size_t)s * M + k]];
            ft8b200_decoded_t &o = h_out[(size_t)s * max_out_per_recording + k];
            memset(&o, 0, sizeof(o));
            memcpy(o.text, umsg[(size_t)s * M + k].text, sizeof(o.text));
            o.hash = umsg[(size_t)s * M + k].hash;
            o.score = c.score;
            o.freq_hz = (c.freq_offset + (float)c.freq_sub / fosr) / symbol_period;  // decode_ft8.c:349-350
            o.time_sec = (c.time_offset + (float)c.time_sub / tosr) * symbol_period;
        }
    }
    return 0;
}

int ft8b200_decode_wav_files(ft8b200_ctx_t *ctx, const char *const *paths, int n, int protocol, ft8b200_decoded_t *h_out, int32_t *h_count,
                             int max_out_per_recording, int32_t *h_status) {
    if (!ctx || !paths || n < 1 || !h_out || !h_count) return FT8B200_BAD_ARG();
    if (cudaSetDevice(ctx_device(ctx)) != cudaSuccess) return FT8B200_CUDA_FAIL();
    const int cap = 15 * 12000;  // decode_ft8.c:271-273: float signal[15 * sample_rate]
    std::vector<int16_t> raw((size_t)n * cap, 0);
    std::vector<int> ns((size_t)n, 0), status((size_t)n, 0), rates((size_t)n, 12000);
    int max_n = 0;
    for (int k = 0; k < n; ++k) {
        int num = cap, sr = 12000;
        status[(size_t)k] = ft8b200_load_wav_s16(&raw[(size_t)k * cap], nullptr, &num, &sr, paths[k]);
        if (status[(size_t)k] < 0) num = 0;
        else rates[(size_t)k] = sr;  // decode_ft8 takes every file at its own rate (decode_ft8.c:277-285)
        ns[(size_t)k] = num;
        if (num > max_n) max_n = num;
        if (h_status) h_status[k] = status[(size_t)k];
    }
    for (int k = 0; k < n; ++k) h_count[k] = 0;
    if (max_n == 0) return 0;
    cudaStream_t st = (cudaStream_t)ft8b200_cuda_stream(ctx);
    DevFree pool(st);
    int16_t *d_raw = nullptr;
    float *d_audio = nullptr;
    if (pool.alloc(&d_raw, (size_t)n * cap * sizeof(int16_t)) != cudaSuccess || pool.alloc(&d_audio, (size_t)n * cap * sizeof(float)) != cudaSuccess)
        return FT8B200_ENOMEM;
    if (cudaMemcpyAsync(d_raw, raw.data(), (size_t)n * cap * sizeof(int16_t), cudaMemcpyHostToDevice, st) != cudaSuccess) return FT8B200_CUDA_FAIL();
    const size_t total = (size_t)n * cap;
    s16_to_float_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_raw, d_audio, total);
    if (cudaGetLastError() != cudaSuccess) return FT8B200_CUDA_FAIL();
    // recordings of different lengths or rates: neighbours of equal length and rate are decoded together (usually all are 15 s at 12 kHz)
    std::vector<char> done((size_t)n, 0);
    for (int k = 0; k < n; ++k) {
        if (done[(size_t)k] || ns[(size_t)k] == 0) continue;
        int run = 1;
        while (k + run < n && ns[(size_t)(k + run)] == ns[(size_t)k] && rates[(size_t)(k + run)] == rates[(size_t)k]) ++run;
        int rc = ft8b200_decode_audio(ctx, d_audio + (size_t)k * cap, (size_t)cap, ns[(size_t)k], run, rates[(size_t)k], protocol,
                                      h_out + (size_t)k * max_out_per_recording, h_count + k, max_out_per_recording);
        if (rc) return rc;
        for (int j = 0; j < run; ++j) done[(size_t)(k + j)] = 1;
    }
    return 0;
}

}  // extern "C"
