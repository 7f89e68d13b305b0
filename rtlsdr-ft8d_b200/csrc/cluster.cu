// cluster.cu -- ft8b200_cluster_t: every visible GPU of one box driven from ONE process through the C ABI.
//
// The work shards by independent 15 s slot or by receiver stream: no data-path collective exists.  A cluster
// owns one ft8b200_pipe_t per device; a step is one batch per device, submitted back to back from the caller's thread (submission is
// asynchronous: a few hundred microseconds of launches per device against milliseconds of device work).  The only exchange is the one
// the north star names: the decoded-spot records (decoder_results[max_messages] + a count per slot, 1.4 KB per slot) of a step are
// gathered with ONE grouped ncclAllGather over NVLink on a high-priority side stream per device, after which device 0's copy is read
// into pinned host memory and handed to the caller in (device, slot) order.  Device-side ordering protects the lanes' buffers
// (ft8b200_pipe_depend_on), so the next step's kernels never wait for the host.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2, the version already in the process if there is one): libft8b200.so keeps
// linking against the CUDA runtime only, and a single-GPU host needs no NCCL at all.  With more than one device and no NCCL the
// cluster cannot be created -- there is no staged-through-the-host fallback.
//
// Reference counterpart: none (the daemon is one receiver, one decoder thread, rtlsdr_ft8d.c:1293-1378); this is the box-level form
// of its "receive slot n+1 while slot n decodes" loop for BASELINE configs #4 and #5.
#include "common.cuh"

#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <deque>
#include <string>
#include <vector>

namespace {

// the few NCCL entry points used, declared here so that no NCCL header is needed to build the library
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;   // ncclSuccess == 0
constexpr int kNcclUint8 = 1;  // ncclUint8 in every NCCL 2.x
struct NcclApi {
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    void *handle = nullptr;
    bool ok = false;
};

const NcclApi &nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
        }
        if (!a.handle) return a;
        bool ok = true;
        auto get = [&](const char *sym, void **fn) { *fn = dlsym(a.handle, sym); if (!*fn) ok = false; };
        get("ncclCommInitAll", reinterpret_cast<void **>(&a.CommInitAll));
        get("ncclCommDestroy", reinterpret_cast<void **>(&a.CommDestroy));
        get("ncclAllGather", reinterpret_cast<void **>(&a.AllGather));
        get("ncclGroupStart", reinterpret_cast<void **>(&a.GroupStart));
        get("ncclGroupEnd", reinterpret_cast<void **>(&a.GroupEnd));
        get("ncclGetErrorString", reinterpret_cast<void **>(&a.GetErrorString));
        get("ncclGetVersion", reinterpret_cast<void **>(&a.GetVersion));
        a.ok = ok;
        return a;
    }();
    return api;
}

struct Step { std::vector<int> rows; };  // result rows (slots) per device of one submitted step

}  // namespace

struct ft8b200_cluster {
    ft8b200_config_t cfg;
    int n = 0;
    std::vector<int> dev;                   // CUDA ordinals
    std::vector<ft8b200_pipe_t *> pipes;
    std::vector<ft8b200_ctx_t *> util;      // one utility context per device (synthesis, allocations of the caller)
    std::vector<cudaStream_t> gst;          // gather stream per device (high priority, non-blocking)
    std::vector<cudaEvent_t> gev;
    std::vector<uint8_t *> d_stage, d_all;  // per device: its rows [records | counts], and every device's
    size_t cap_rows = 0;                    // rows per device the gather buffers hold
    uint8_t *h_all = nullptr;               // pinned
    std::vector<ncclComm_t> comms;
    std::deque<Step> steps;
    uint64_t gathers = 0;
    bool broken = false;                    // a step failed part-way (some devices submitted / handed out, others not): no further steps
    std::string err;
};

namespace {

int cfail(ft8b200_cluster_t *c, int code, const std::string &msg) {
    c->err = msg;
    return code;
}
#define CCU(call)                                                                                                     \
    do {                                                                                                              \
        cudaError_t e__ = (call);                                                                                     \
        if (e__ != cudaSuccess) { (void)cudaGetLastError(); return cfail(c, FT8B200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } \
    } while (0)

// a failure in the middle of a step leaves the devices' pipes out of step with each other: the cluster refuses further work
int cbreak(ft8b200_cluster_t *c, int code, const std::string &msg) {
    c->broken = true;
    return cfail(c, code, msg);
}
int refuse_if_broken(ft8b200_cluster_t *c) {
    if (!c->broken) return 0;
    if (c->err.rfind("cluster stopped", 0) != 0) c->err = "cluster stopped after a failed step: " + c->err;
    return FT8B200_ECUDA;
}

size_t row_bytes(const ft8b200_cluster_t *c) { return (size_t)c->cfg.max_messages * sizeof(struct decoder_results) + sizeof(int32_t); }

void release_gather_buffers(ft8b200_cluster_t *c) {
    for (int d = 0; d < (int)c->dev.size(); ++d) {
        cudaSetDevice(c->dev[(size_t)d]);
        if ((size_t)d < c->d_stage.size() && c->d_stage[(size_t)d]) cudaFree(c->d_stage[(size_t)d]);
        if ((size_t)d < c->d_all.size() && c->d_all[(size_t)d]) cudaFree(c->d_all[(size_t)d]);
    }
    c->d_stage.assign((size_t)c->n, nullptr);
    c->d_all.assign((size_t)c->n, nullptr);
    if (c->h_all) cudaFreeHost(c->h_all);
    c->h_all = nullptr;
    c->cap_rows = 0;
}

// stage layout per device: cap_rows x max_messages records, then cap_rows int32 counts (-1 = no such slot)
int ensure_gather_buffers(ft8b200_cluster_t *c, size_t rows) {
    if (rows <= c->cap_rows) return 0;
    for (int d = 0; d < c->n; ++d) {  // nothing may still be reading the old buffers
        CCU(cudaSetDevice(c->dev[(size_t)d]));
        CCU(cudaStreamSynchronize(c->gst[(size_t)d]));
    }
    release_gather_buffers(c);
    const size_t per = rows * row_bytes(c);
    for (int d = 0; d < c->n; ++d) {
        CCU(cudaSetDevice(c->dev[(size_t)d]));
        CCU(cudaMalloc(&c->d_stage[(size_t)d], per));
        CCU(cudaMalloc(&c->d_all[(size_t)d], per * (size_t)c->n));
    }
    CCU(cudaMallocHost(&c->h_all, per * (size_t)c->n));
    c->cap_rows = rows;
    return 0;
}

// contiguous blocks of ceil(n / devices) items per device (tools/shard.py::shard_range)
void shard(int n_items, int n_dev, int d, int *lo, int *hi) {
    const int per = (n_items + n_dev - 1) / n_dev;
    *lo = d * per < n_items ? d * per : n_items;
    *hi = *lo + per < n_items ? *lo + per : n_items;
}

}  // namespace

extern "C" {

ft8b200_cluster_t *ft8b200_cluster_create(const ft8b200_config_t *cfg_in, int n_devices, int depth) {
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1) {
        (void)ft8b200_create(cfg_in);  // fails the same way and leaves the reason in ft8b200_last_error()
        return nullptr;
    }
    if (n_devices <= 0 || n_devices > visible) n_devices = visible;
    ft8b200_cluster_t *c = new ft8b200_cluster();
    ft8b200_default_config(&c->cfg);
    if (cfg_in) c->cfg = *cfg_in;
    c->n = n_devices;
    c->d_stage.assign((size_t)n_devices, nullptr);
    c->d_all.assign((size_t)n_devices, nullptr);
    bool ok = true;
    for (int d = 0; d < n_devices && ok; ++d) {
        ft8b200_config_t cfg = c->cfg;
        cfg.device = d;
        c->dev.push_back(d);
        ft8b200_pipe_t *p = ft8b200_pipe_create(&cfg, depth);
        ft8b200_ctx_t *u = p ? ft8b200_create(&cfg) : nullptr;
        c->pipes.push_back(p);
        c->util.push_back(u);
        cudaStream_t st = nullptr;
        cudaEvent_t ev = nullptr;
        int lo = 0, hi = 0;
        ok = p && u && cudaSetDevice(d) == cudaSuccess && cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess &&
             cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, hi) == cudaSuccess &&
             cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess;
        c->gst.push_back(st);
        c->gev.push_back(ev);
    }
    if (ok && n_devices > 1) {
        const NcclApi &api = nccl_api();
        if (!api.ok) {
            c->err = "libnccl.so.2 could not be loaded: a cluster of more than one GPU gathers its spot records with NCCL (no host-staged fallback)";
            ok = false;
        } else {
            c->comms.assign((size_t)n_devices, nullptr);
            const ncclResult_t r = api.CommInitAll(c->comms.data(), n_devices, c->dev.data());
            if (r != 0) { c->err = std::string("ncclCommInitAll: ") + api.GetErrorString(r); c->comms.clear(); ok = false; }
        }
    }
    if (!ok) {
        if (c->err.empty()) c->err = ft8b200_last_error();
        fprintf(stderr, "libft8b200: ft8b200_cluster_create: %s\n", c->err.c_str());
        ft8b200_cluster_destroy(c);
        return nullptr;
    }
    return c;
}

void ft8b200_cluster_destroy(ft8b200_cluster_t *c) {
    if (!c) return;
    for (int d = 0; d < (int)c->gst.size(); ++d) {
        cudaSetDevice(c->dev[(size_t)d]);
        if (c->gst[(size_t)d]) cudaStreamSynchronize(c->gst[(size_t)d]);
    }
    if (!c->comms.empty()) for (ncclComm_t comm : c->comms) if (comm) nccl_api().CommDestroy(comm);
    release_gather_buffers(c);
    for (int d = 0; d < (int)c->pipes.size(); ++d) {
        cudaSetDevice(c->dev[(size_t)d]);
        if (c->pipes[(size_t)d]) ft8b200_pipe_destroy(c->pipes[(size_t)d]);
        if ((size_t)d < c->util.size() && c->util[(size_t)d]) ft8b200_destroy(c->util[(size_t)d]);
        if ((size_t)d < c->gst.size() && c->gst[(size_t)d]) cudaStreamDestroy(c->gst[(size_t)d]);
        if ((size_t)d < c->gev.size() && c->gev[(size_t)d]) cudaEventDestroy(c->gev[(size_t)d]);
    }
    delete c;
}

const char *ft8b200_cluster_error(ft8b200_cluster_t *c) { return c ? c->err.c_str() : "null cluster"; }
int ft8b200_cluster_devices(ft8b200_cluster_t *c) { return c ? c->n : 0; }
int ft8b200_cluster_in_flight(ft8b200_cluster_t *c) { return c ? (int)c->steps.size() : 0; }
ft8b200_pipe_t *ft8b200_cluster_pipe(ft8b200_cluster_t *c, int device_index) { return (c && device_index >= 0 && device_index < c->n) ? c->pipes[(size_t)device_index] : nullptr; }
ft8b200_ctx_t *ft8b200_cluster_ctx(ft8b200_cluster_t *c, int device_index) { return (c && device_index >= 0 && device_index < c->n) ? c->util[(size_t)device_index] : nullptr; }
uint64_t ft8b200_cluster_gathers(ft8b200_cluster_t *c) { return c ? c->gathers : 0; }
int ft8b200_cluster_nccl_version(ft8b200_cluster_t *c) {
    int v = 0;
    if (c && !c->comms.empty() && nccl_api().GetVersion) nccl_api().GetVersion(&v);
    return v;
}
uint64_t ft8b200_cluster_kernel_launches(ft8b200_cluster_t *c) {
    uint64_t n = 0;
    if (c) for (ft8b200_pipe_t *p : c->pipes) n += ft8b200_pipe_kernel_launches(p);
    return n;
}

int ft8b200_cluster_shard(ft8b200_cluster_t *c, int n_items, int device_index, int *first, int *count) {
    if (!c || device_index < 0 || device_index >= c->n || n_items < 0) return FT8B200_BAD_ARG();
    int lo, hi;
    shard(n_items, c->n, device_index, &lo, &hi);
    if (first) *first = lo;
    if (count) *count = hi - lo;
    return 0;
}

// one step: device d gets n_per_device[d] streams of its own device-resident input d_iq[d] (0 streams: the device sits the step out)
int ft8b200_cluster_submit_streams(ft8b200_cluster_t *c, const uint8_t *const *d_iq, size_t bytes_per_stream, size_t stream_stride_bytes,
                                   const int *n_streams_per_device, int slots_per_stream, size_t bytes_per_slot) {
    if (!c || !d_iq || !n_streams_per_device || slots_per_stream < 1) return c ? cfail(c, FT8B200_EINVAL, "ft8b200_cluster_submit: bad argument") : FT8B200_EINVAL;
    if (int rb = refuse_if_broken(c)) return rb;
    for (int d = 0; d < c->n; ++d)
        if (n_streams_per_device[d] > 0 && !d_iq[d]) return cfail(c, FT8B200_EINVAL, "ft8b200_cluster_submit: null input for a device with work");
    Step st;
    for (int d = 0; d < c->n; ++d) {
        const int n = n_streams_per_device[d];
        int rc = 0;
        if (n > 0) {
            rc = slots_per_stream > 1 ? ft8b200_pipe_submit_streams(c->pipes[(size_t)d], d_iq[d], bytes_per_stream, stream_stride_bytes, n, slots_per_stream, bytes_per_slot)
                                      : ft8b200_pipe_submit(c->pipes[(size_t)d], d_iq[d], bytes_per_stream, stream_stride_bytes, n);
        }
        if (rc) return (d > 0 ? cbreak : cfail)(c, rc, std::string("device ") + std::to_string(d) + ": " + ft8b200_pipe_error(c->pipes[(size_t)d]));
        st.rows.push_back(n > 0 ? n * slots_per_stream : 0);
    }
    c->steps.push_back(st);
    return 0;
}

int ft8b200_cluster_submit(ft8b200_cluster_t *c, const uint8_t *const *d_iq, size_t bytes_per_stream, size_t stream_stride_bytes, const int *n_slots_per_device) {
    return ft8b200_cluster_submit_streams(c, d_iq, bytes_per_stream, stream_stride_bytes, n_slots_per_device, 1, 0);
}

// n_slots independent slots in (pinned) host memory, sharded in contiguous blocks over the devices
int ft8b200_cluster_submit_host(ft8b200_cluster_t *c, const uint8_t *h_iq, size_t bytes_per_stream, int n_slots) {
    if (!c || !h_iq || n_slots < 1) return c ? cfail(c, FT8B200_EINVAL, "ft8b200_cluster_submit_host: bad argument") : FT8B200_EINVAL;
    if (int rb = refuse_if_broken(c)) return rb;
    Step st;
    for (int d = 0; d < c->n; ++d) {
        int lo, hi;
        shard(n_slots, c->n, d, &lo, &hi);
        if (hi > lo) {
            const int rc = ft8b200_pipe_submit_host(c->pipes[(size_t)d], h_iq + (size_t)lo * bytes_per_stream, bytes_per_stream, hi - lo);
            if (rc) return (d > 0 ? cbreak : cfail)(c, rc, std::string("device ") + std::to_string(d) + ": " + ft8b200_pipe_error(c->pipes[(size_t)d]));
        }
        st.rows.push_back(hi - lo);
    }
    c->steps.push_back(st);
    return 0;
}

// Oldest step: wait for every device's batch, gather the records of all devices (NCCL, one grouped all-gather), read device 0's copy
// and unpack it in (device, slot) order.  Returns the number of slots written, or a negative error.
int ft8b200_cluster_collect(ft8b200_cluster_t *c, struct decoder_results *h_results, int32_t *h_nresults, int capacity_slots) {
    if (!c || !h_results || !h_nresults) return c ? cfail(c, FT8B200_EINVAL, "ft8b200_cluster_collect: null result buffer") : FT8B200_EINVAL;
    if (int rb = refuse_if_broken(c)) return rb;
    if (c->steps.empty()) return cfail(c, FT8B200_EINVAL, "ft8b200_cluster_collect: nothing in flight");
    const Step st = c->steps.front();
    int total = 0, max_rows = 0;
    for (int r : st.rows) { total += r; if (r > max_rows) max_rows = r; }
    if (capacity_slots < total) return cfail(c, FT8B200_EINVAL, "ft8b200_cluster_collect: result buffers too small");
    const size_t M = (size_t)c->cfg.max_messages, rec_bytes = M * sizeof(struct decoder_results);
    int rc = ensure_gather_buffers(c, (size_t)(max_rows > 0 ? max_rows : 1));
    if (rc) return rc;
    const size_t cap = c->cap_rows, per = cap * row_bytes(c);
    c->broken = true;   // until the step is through: any failure below leaves some pipes a batch ahead of the others
    // stage every device's rows on its gather stream, ordered behind the batch by the host-side wait of collect_device
    for (int d = 0; d < c->n; ++d) {
        CCU(cudaSetDevice(c->dev[(size_t)d]));
        uint8_t *stage = c->d_stage[(size_t)d];
        const int rows = st.rows[(size_t)d];
        CCU(cudaMemsetAsync(stage + cap * rec_bytes, 0xff, cap * sizeof(int32_t), c->gst[(size_t)d]));  // counts = -1: no such slot
        if (rows > 0) {
            struct decoder_results *d_res = nullptr;
            int32_t *d_n = nullptr;
            const int got = ft8b200_pipe_collect_device(c->pipes[(size_t)d], &d_res, &d_n);
            if (got != rows) return cfail(c, got < 0 ? got : FT8B200_ECUDA, std::string("device ") + std::to_string(d) + ": " + ft8b200_pipe_error(c->pipes[(size_t)d]));
            CCU(cudaMemcpyAsync(stage, d_res, (size_t)rows * rec_bytes, cudaMemcpyDeviceToDevice, c->gst[(size_t)d]));
            CCU(cudaMemcpyAsync(stage + cap * rec_bytes, d_n, (size_t)rows * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->gst[(size_t)d]));
        }
    }
    if (c->n > 1) {
        const NcclApi &api = nccl_api();
        ncclResult_t r = api.GroupStart();
        for (int d = 0; d < c->n && r == 0; ++d)
            r = api.AllGather(c->d_stage[(size_t)d], c->d_all[(size_t)d], per, kNcclUint8, c->comms[(size_t)d], c->gst[(size_t)d]);
        const ncclResult_t r2 = api.GroupEnd();
        if (r != 0 || r2 != 0) return cfail(c, FT8B200_ECUDA, std::string("ncclAllGather: ") + api.GetErrorString(r != 0 ? r : r2));
        ++c->gathers;
    }
    for (int d = 0; d < c->n; ++d) {
        CCU(cudaSetDevice(c->dev[(size_t)d]));
        CCU(cudaEventRecord(c->gev[(size_t)d], c->gst[(size_t)d]));
        // the lane just handed out is rewritten only after its records have left it (device-side ordering, no host wait)
        if (st.rows[(size_t)d] > 0) ft8b200_pipe_depend_on(c->pipes[(size_t)d], c->gev[(size_t)d]);
    }
    CCU(cudaSetDevice(c->dev[0]));
    const uint8_t *src = c->n > 1 ? c->d_all[0] : c->d_stage[0];
    CCU(cudaMemcpyAsync(c->h_all, src, per * (size_t)c->n, cudaMemcpyDeviceToHost, c->gst[0]));
    CCU(cudaStreamSynchronize(c->gst[0]));
    int out = 0;
    for (int d = 0; d < c->n; ++d) {
        const uint8_t *blk = c->h_all + (size_t)d * per;
        const int32_t *cnt = reinterpret_cast<const int32_t *>(blk + cap * rec_bytes);
        for (int k = 0; k < st.rows[(size_t)d]; ++k, ++out) {
            if (cnt[k] < 0) return cfail(c, FT8B200_ECUDA, "ft8b200_cluster_collect: a gathered slot carries no count");
            memcpy(h_results + (size_t)out * M, blk + (size_t)k * rec_bytes, rec_bytes);
            h_nresults[out] = cnt[k];
        }
    }
    c->steps.pop_front();
    c->broken = false;
    return out;
}

void *ft8b200_device_malloc(ft8b200_ctx_t *ctx, size_t bytes) {
    if (!ctx || cudaSetDevice(ft8b200::ctx_device(ctx)) != cudaSuccess) return nullptr;
    void *p = nullptr;
    return cudaMalloc(&p, bytes ? bytes : 1) == cudaSuccess ? p : nullptr;
}

void ft8b200_device_free(ft8b200_ctx_t *ctx, void *p) {
    if (!ctx || !p) return;
    cudaSetDevice(ft8b200::ctx_device(ctx));
    cudaFree(p);
}

}  // extern "C"
