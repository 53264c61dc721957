/* clive_b200.cu — libclive_b200.so: the C ABI of include/clive_b200.h on sm_100a.
 *
 * Build (see __graft_entry__.build / clive2_b200/build.py):
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC ...
 * -fmad=false is part of the arithmetic contract (clv_device.cuh): results equal the oracle's bit for bit.
 *
 * There is no host execution path in this library: every entry point that computes launches a kernel and reports
 * CLV_ERR_NO_DEVICE / CLV_ERR_CUDA when that is impossible.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "clv_device.cuh"
#include "clv_accel_build.h"

using namespace clv;

/* ---------------------------------------------------------------------------------------------------------------
 * errors
 * ------------------------------------------------------------------------------------------------------------- */
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(e_ == cudaErrorMemoryAllocation ? CLV_ERR_NOMEM : CLV_ERR_CUDA, "%s: %s (%s:%d)", \
                        #call, cudaGetErrorString(e_), __FILE__, __LINE__);                              \
    } while (0)
#define REQUIRE(cond, msg)                                           \
    do {                                                             \
        if (!(cond)) return fail(CLV_ERR_INVALID, "%s: %s", __func__, msg); \
    } while (0)

extern "C" const char *clv_last_error(void) { return g_err; }
extern "C" const char *clv_version(void) { return "clive_b200 0.1 (sm_100a, fmad=false)"; }

/* ---------------------------------------------------------------------------------------------------------------
 * context + memory
 * ------------------------------------------------------------------------------------------------------------- */
#ifndef CLV_WALK_TAIL_BUDGET
/* k_walk_trace suspends the searches still running this many iterations after the round's queue ran dry (WalkRecords).
 * C4, ms per pass, trace rounds alone / whole pass: off 12.71 / 29.30, 256: 12.64 / 29.61, 128: 11.87 / 29.48,
 * 64: 11.28 / 29.16, 32: 10.87 / 28.97, 16: 11.04 / 30.44 (more suspended searches, more shade work); half-frame shard
 * off / 32 / 64: 16.49 / 17.47 / 16.12; C3 off / 32 / 64: 13.54 / 13.44 / 13.39 */
#define CLV_WALK_TAIL_BUDGET 64
#endif
#define CLV_MAX_SLICES 8 /* <= 16: one task counter per slice in the 64-byte header of the connect scratch */
struct clv_ctx {
    int device;
    cudaStream_t stream;
    Counters *d_counters;
    int counting;
    unsigned long long launches;
    int sm_count;
    int walk_mode;       /* fused pass, stages 1-4: 0 = automatic, 1 = megakernel (k_fused_walks), 2 = wavefront (k_walk_*) */
    int walk_refill_below, walk_schedule; /* k_walk_trace: refill threshold and vote weight, as for k_connect_rays */
    int walk_grid_divisor;                /* k_walk_trace: persistent grid = resident blocks / divisor (slices co-reside) */
    int trace_blocks_per_sm[2];           /* resident blocks of k_walk_trace (cache-resident / large flavour) */
    int connect_dynamic_fetch, rays_blocks_per_sm[2]; /* k_connect_rays: dynamic work fetch on a persistent grid */
    int finish_cooperative;                           /* k_connect_finish: warp-cooperative form */
    int walk_tail_budget;                 /* k_walk_trace: iterations past the end of a round's queue before searches are suspended; 0 = never */
    int pipe_shade_every;                             /* k_walk_pipeline: every k-th block only shades (0: dual-role warps) */
    int refill_below;    /* k_connect_rays: hand out new rays when fewer lanes than this hold one */
    int ray_schedule;    /* k_connect_rays: 0 = rounds; w > 0 = majority vote, inner steps win when n_inner*w >= n_leaf*4 */
    int ray_schedule_large; /* ... of the large-scene flavour, whose leaf steps are warp-cooperative */
    /* scratch of the stage-level clv_connect_paths (grown on demand) */
    void *cs_mem;
    uint32_t cs_capacity;
    /* the fused pass runs as `pass_slices` independent slices of the samples, slice 0 on `stream`, the others on
     * `slice_stream[]`, joined before the finalize kernel: the tail of one slice's kernel (SMs draining while the
     * last blocks finish) is filled by blocks of another slice's kernel instead of idling */
    int pass_slices;
    cudaStream_t slice_stream[CLV_MAX_SLICES];
    cudaEvent_t slice_fork, slice_join[CLV_MAX_SLICES];
};

extern "C" int clv_device_count(int *count) {
    REQUIRE(count, "null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(CLV_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return CLV_OK;
}

/* Device memory comes from the device's stream-ordered pool (cudaMallocAsync on the context's stream) with the
 * release threshold lifted, so that a freed block is handed to the next allocation instead of going back to the
 * driver: the movie driver creates and destroys a scene, an accel and a renderer (~1 GB at 720p) for every frame, and
 * cudaMalloc/cudaFree of those blocks cost more than half of the frame's 32 passes (measured: 0.23 s of 0.61 s).
 * Every use of the memory is ordered on the same stream, which is what makes the asynchronous free safe. */
template <class T>
static inline cudaError_t dev_alloc(clv_ctx *c, T **p, size_t nbytes) { return cudaMallocAsync((void **)p, nbytes, c->stream); }
static inline cudaError_t dev_free(clv_ctx *c, void *p) { return p ? cudaFreeAsync(p, c->stream) : cudaSuccess; }

extern "C" int clv_ctx_create(int device, clv_ctx **out) {
    REQUIRE(out, "null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(CLV_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
    REQUIRE(device >= 0 && device < n, "device index out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(CLV_ERR_NO_DEVICE, "device %d is sm_%d%d; this build is sm_100a only", device, prop.major, prop.minor);
    clv_ctx *c = new (std::nothrow) clv_ctx();
    if (!c) return fail(CLV_ERR_NOMEM, "host allocation");
    c->device = device;
    c->counting = 0;
    c->launches = 0;
    c->sm_count = prop.multiProcessorCount;
    /* measured with 512-task chunks at C2: 8 / 12 / 16 / 20 / 24 / 28 / 32 -> 1.028 / 0.968 / 0.953 / 0.956 / 0.983 / 0.992 /
     * 1.093 ms for k_connect_rays; C3: 12 / 16 / 24 -> 6.70 / 6.44 / 6.44 ms */
    c->refill_below = getenv("CLV_REFILL_BELOW") ? atoi(getenv("CLV_REFILL_BELOW")) : 16;
    c->ray_schedule = getenv("CLV_RAY_SCHEDULE") ? atoi(getenv("CLV_RAY_SCHEDULE")) : 8; /* measured: C2 1.15 -> 1.10 ms, C3 8.07 -> 7.10 ms */
    c->walk_mode = getenv("CLV_WALK_MODE") ? atoi(getenv("CLV_WALK_MODE")) : 0;
    c->walk_refill_below = getenv("CLV_WALK_REFILL_BELOW") ? atoi(getenv("CLV_WALK_REFILL_BELOW")) : 24;
    if (c->walk_refill_below < 1) c->walk_refill_below = 1;
    /* with leaves tested by the whole warp a leaf step is cheap per triangle, so it may win the vote earlier.  C4, ms per
     * pass (k_walk_trace / k_connect_rays) at weight 1 / 2 / 4 / 6 / 8 / 16 / 32: 10.69 / 10.39 / 10.32 / 10.41 / 10.52 /
     * 11.08 / 12.05 and 9.61 / 9.54 / 9.60 / 9.76 / 9.96 / 10.81 / 12.39; C3 at 2 / 4 / 8 / 16: 4.14 / 4.12 / 4.23 / 4.47
     * and 5.22 / 5.26 / 5.47 / 5.98 */
    c->walk_schedule = getenv("CLV_WALK_SCHEDULE") ? atoi(getenv("CLV_WALK_SCHEDULE")) : 4;
    c->ray_schedule_large = getenv("CLV_RAY_SCHEDULE") ? atoi(getenv("CLV_RAY_SCHEDULE")) : 4;
    /* with the pass cut into slices, a trace grid of HALF the resident blocks lets the kernels of two slices share the SMs,
     * so one slice's tail (a few long rays, latency bound) overlaps another's work: C4, 3 slices, divisor 1 / 2 / 3 ->
     * 29.7 / 28.8 / 30.2 ms per pass; 2 slices divisor 1: 28.9; 4 slices divisor 1 / 2: 30.6 / 29.5 */
    c->walk_grid_divisor = getenv("CLV_WALK_GRID_DIVISOR") ? std::max(1, atoi(getenv("CLV_WALK_GRID_DIVISOR"))) : 2;
    c->trace_blocks_per_sm[0] = c->trace_blocks_per_sm[1] = 0; /* asked from the occupancy calculator on first use */
    c->rays_blocks_per_sm[0] = c->rays_blocks_per_sm[1] = 0;
    /* measured (k_connect_rays, ms per pass, static 512-task chunks -> dynamic fetch): C4 12.99 -> 11.07, C3 6.47 -> 5.87,
     * C2 0.95 -> 0.85; C1 126 -> 136 Msamples/s */
    c->connect_dynamic_fetch = getenv("CLV_CONNECT_DYNAMIC") ? atoi(getenv("CLV_CONNECT_DYNAMIC")) : 1;
    c->finish_cooperative = getenv("CLV_FINISH_COOPERATIVE") ? atoi(getenv("CLV_FINISH_COOPERATIVE")) : 0;
    c->walk_tail_budget = getenv("CLV_WALK_TAIL_BUDGET") ? std::max(0, atoi(getenv("CLV_WALK_TAIL_BUDGET"))) : CLV_WALK_TAIL_BUDGET;
    c->pipe_shade_every = getenv("CLV_PIPE_SHADE_EVERY") ? atoi(getenv("CLV_PIPE_SHADE_EVERY")) : 5;
    c->cs_mem = nullptr;
    c->cs_capacity = 0;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    {
        cudaMemPool_t pool;
        uint64_t keep = ~0ull; /* never trim: see dev_alloc */
        CK(cudaDeviceGetDefaultMemPool(&pool, device));
        CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    /* 0 = automatic.  Measured: C2 (262 144 samples = 1.7 waves of resident threads) 1/2/3/4/6/8 slices ->
     * 102.0/100.9/106.5/106.1/104.7/102.9 Msamples/s; C1 (65 536 samples, less than one wave) 125.7/119.8/119.7/119.2;
     * C3 (1 M samples) 68.2/-/68.7/68.3.  Automatic = 3 slices from one full wave of samples up, else 1. */
    c->pass_slices = getenv("CLV_PASS_SLICES") ? atoi(getenv("CLV_PASS_SLICES")) : 0;
    if (c->pass_slices < 0) c->pass_slices = 0;
    if (c->pass_slices > CLV_MAX_SLICES) c->pass_slices = CLV_MAX_SLICES;
    CK(cudaEventCreateWithFlags(&c->slice_fork, cudaEventDisableTiming));
    for (int k = 0; k < CLV_MAX_SLICES; k++) {
        c->slice_stream[k] = nullptr;
        c->slice_join[k] = nullptr;
        if (k == 0) continue;
        CK(cudaStreamCreateWithFlags(&c->slice_stream[k], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c->slice_join[k], cudaEventDisableTiming));
    }
    CK(dev_alloc(c, &c->d_counters, COUNTERS_CLASSES * sizeof(Counters)));
    CK(cudaMemsetAsync(c->d_counters, 0, COUNTERS_CLASSES * sizeof(Counters), c->stream));
    *out = c;
    return CLV_OK;
}

extern "C" int clv_ctx_destroy(clv_ctx *c) {
    if (!c) return CLV_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    dev_free(c, c->d_counters);
    dev_free(c, c->cs_mem);
    cudaStreamSynchronize(c->stream);
    cudaEventDestroy(c->slice_fork);
    for (int k = 1; k < CLV_MAX_SLICES; k++) {
        if (c->slice_stream[k]) { cudaStreamSynchronize(c->slice_stream[k]); cudaStreamDestroy(c->slice_stream[k]); }
        if (c->slice_join[k]) cudaEventDestroy(c->slice_join[k]);
    }
    cudaStreamDestroy(c->stream);
    delete c;
    return CLV_OK;
}

extern "C" int clv_ctx_synchronize(clv_ctx *c) {
    REQUIRE(c, "null ctx");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return CLV_OK;
}

extern "C" int clv_ctx_stream(clv_ctx *c, void **stream) {
    REQUIRE(c && stream, "null");
    *stream = (void *)c->stream;
    return CLV_OK;
}

extern "C" int clv_mem_alloc(clv_ctx *c, size_t nbytes, void **d_ptr) {
    REQUIRE(c && d_ptr, "null");
    CK(cudaSetDevice(c->device));
    *d_ptr = nullptr;
    if (nbytes == 0) nbytes = 16;
    CK(dev_alloc(c, d_ptr, nbytes));
    CK(cudaMemsetAsync(*d_ptr, 0, nbytes, c->stream));
    return CLV_OK;
}
extern "C" int clv_mem_free(clv_ctx *c, void *d_ptr) {
    REQUIRE(c, "null ctx");
    if (!d_ptr) return CLV_OK;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(dev_free(c, d_ptr));
    return CLV_OK;
}
extern "C" int clv_mem_upload(clv_ctx *c, void *d_dst, const void *h_src, size_t nbytes) {
    REQUIRE(c && (nbytes == 0 || (d_dst && h_src)), "null");
    if (!nbytes) return CLV_OK;
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(d_dst, h_src, nbytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream)); /* the host array is only borrowed for the duration of the call */
    return CLV_OK;
}
extern "C" int clv_mem_download(clv_ctx *c, void *h_dst, const void *d_src, size_t nbytes) {
    REQUIRE(c && (nbytes == 0 || (h_dst && d_src)), "null");
    if (!nbytes) return CLV_OK;
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(h_dst, d_src, nbytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CLV_OK;
}

__global__ void k_fill32(uint32_t *dst, uint32_t value, size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) dst[i] = value;
}
static inline unsigned grid_for(size_t n, unsigned block, unsigned cap = 148u * 32u) {
    size_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    return (unsigned)std::min<size_t>(g, cap);
}
#define LAUNCHED(c)                  \
    do {                             \
        (c)->launches++;             \
        CK(cudaGetLastError());      \
    } while (0)

extern "C" int clv_mem_fill32(clv_ctx *c, void *d_dst, uint32_t value, size_t count) {
    REQUIRE(c && (count == 0 || d_dst), "null");
    if (!count) return CLV_OK;
    CK(cudaSetDevice(c->device));
    k_fill32<<<grid_for(count, 256), 256, 0, c->stream>>>((uint32_t *)d_dst, value, count);
    LAUNCHED(c);
    return CLV_OK;
}
extern "C" int clv_host_alloc(size_t nbytes, void **h_ptr) {
    REQUIRE(h_ptr, "null");
    CK(cudaHostAlloc(h_ptr, nbytes ? nbytes : 16, cudaHostAllocDefault));
    return CLV_OK;
}
extern "C" int clv_host_free(void *h_ptr) {
    if (h_ptr) CK(cudaFreeHost(h_ptr));
    return CLV_OK;
}

/* ---------------------------------------------------------------------------------------------------------------
 * acceleration structure: reference Box[]/Triangle[] -> compact traversal layout (clv_device.cuh AccelView)
 * ------------------------------------------------------------------------------------------------------------- */
/* Scenes whose traversal records (nodes + triangles) do not fit the cache hierarchy are bound by the latency of
 * dependent record fetches, and more resident warps hide more of it even at the price of spills; cache-resident scenes
 * are issue-bound and lose with fewer registers.  Measured (walks at 64 / 48 / 40 / 32 registers per thread): C2 (2.8 MB)
 * 0.98 / 1.01 / - / - ms, C3 (60 MB) 6.81 / 6.32 / 6.13 / 6.17 ms, C4 (700 MB) 25.5 / 22.5 / 21.2 / 19.6 ms; occlusion
 * rays at 63 / 48 / 40 / 32 registers: C2 0.99 / 1.00 ms, C3 6.73 / 6.60 / - / 8.71 ms, C4 14.3 / 13.5 / 13.9 / 16.4 ms. */
/* node layout k_walk_trace walks: the binary nodes (0) or their 4-wide collapse (1), per kernel flavour */
#ifndef CLV_TRACE_WIDE
#define CLV_TRACE_WIDE 0
#endif
#ifndef CLV_TRACE_WIDE_LARGE
#define CLV_TRACE_WIDE_LARGE 1 /* half the dependent record fetches per ray: C4 13.2 -> 12.2 ms, C3 4.51 -> 4.33 ms */
#endif
#define CLV_TRACE_WIDE_VALUE (CLV_TRACE_WIDE != 0)
#define CLV_TRACE_WIDE_LARGE_VALUE (CLV_TRACE_WIDE_LARGE != 0)
/* which scenes run stages 1-4 as wavefront rounds when the walk mode is automatic (measured: profiles/r02_notes.md) */
#ifndef CLV_WAVEFRONT_BY_DEFAULT
#define CLV_WAVEFRONT_BY_DEFAULT(large_scene) (large_scene)
#endif
/* ... and how many samples a pass (of this renderer: a tile shard has fewer) needs for the wavefront form to pay: every
 * round ends with a tail of a few long rays whose length does not shrink with the shard.  C4, samples per pass 2 073 600 /
 * 1 036 800 / 259 200: wavefront 14.7 / 9.2 / 5.2 ms of walks, megakernel 18.9 / ~9.5 / 3.4 */
#ifndef CLV_WAVEFRONT_MIN_SAMPLES
#define CLV_WAVEFRONT_MIN_SAMPLES (600u * 1024u)
#endif
/* ... and whether the wavefront walks then run as one pipelined kernel (1) or as 12 rounds of trace + shade (0) */
#ifndef CLV_PIPELINE_BY_DEFAULT
#define CLV_PIPELINE_BY_DEFAULT 0
#endif
#ifndef CLV_LARGE_SCENE_BYTES
#define CLV_LARGE_SCENE_BYTES (32u << 20)
#endif
struct clv_accel {
    clv_ctx *ctx;
    int refs; /* 1 for the creator + 1 per renderer built on it: clv_accel_destroy only frees at zero, so releasing the
               * scene (or its buffers) before a renderer that still traverses the structure is harmless */
    V4 *d_nodes, *d_wnodes, *d_tris;
    V4 *d_shade; /* compact shading records (AccelView::shade); the reference-layout Triangle[] is not kept on the device */
    int32_t *d_camera_tris;
    AccelView view;
    uint64_t n_nodes, n_wide, n_tris;
    int large; /* traversal records exceed CLV_LARGE_SCENE_BYTES: the fused pass uses its high-occupancy kernel flavours */
    int binary_depth, wide_depth; /* inner levels of the two node layouts (bounded by the traversal stacks) */
};

extern "C" int clv_accel_destroy(clv_accel *a);
extern "C" int clv_accel_create(clv_ctx *c, const clv_Box *h_boxes_, int32_t n_boxes, const clv_Triangle *h_tris_,
                                int32_t n_tris, clv_accel **out) {
    REQUIRE(c && out && h_boxes_ && h_tris_, "null");
    REQUIRE(n_boxes > 0 && n_tris > 0, "empty scene");
    REQUIRE(n_tris < (1 << 27), "more than 2^27 triangles");
    *out = nullptr;
    CK(cudaSetDevice(c->device));
    const Box *boxes = reinterpret_cast<const Box *>(h_boxes_);
    const Triangle *tris = reinterpret_cast<const Triangle *>(h_tris_);

    HostAccel H;
    const char *why = build_host_accel(boxes, n_boxes, tris, n_tris, H);
    if (why) return fail(CLV_ERR_INVALID, "clv_accel_create: %s", why);
    std::vector<NodeRec> &nodes = H.nodes;
    std::vector<V4> &ctris = H.tris;
    std::vector<int32_t> &camera_tris = H.camera_tris;
    int32_t root_ref = H.root_ref;

    clv_accel *a = new (std::nothrow) clv_accel();
    if (!a) return fail(CLV_ERR_NOMEM, "host allocation");
    memset(a, 0, sizeof *a);
    a->ctx = c;
    a->refs = 1;
    a->n_nodes = nodes.size();
    a->n_tris = (uint64_t)n_tris;
    size_t nb_nodes = std::max<size_t>(nodes.size(), 1) * sizeof(NodeRec);
    size_t nb_tris = ctris.size() * sizeof(V4);
    size_t nb_ref = (size_t)n_tris * 4 * sizeof(V4);
    size_t nb_cam = std::max<size_t>(camera_tris.size(), 1) * sizeof(int32_t);
    /* on failure free what was allocated so far: the pool never trims, a half-built 10 M-triangle structure would strand
     * its blocks for the life of the process */
#define CKA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { clv_accel_destroy(a); \
        return fail(e_ == cudaErrorMemoryAllocation ? CLV_ERR_NOMEM : CLV_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } } while (0)
    CKA(dev_alloc(c, &a->d_nodes, nb_nodes));
    size_t nb_wide = std::max<size_t>(H.wide.size(), 1) * sizeof(WideRec);
    CKA(dev_alloc(c, &a->d_wnodes, nb_wide));
    if (!H.wide.empty()) CKA(cudaMemcpyAsync(a->d_wnodes, H.wide.data(), H.wide.size() * sizeof(WideRec), cudaMemcpyHostToDevice, c->stream));
    CKA(dev_alloc(c, &a->d_tris, nb_tris));
    CKA(dev_alloc(c, &a->d_shade, nb_ref));
    CKA(dev_alloc(c, &a->d_camera_tris, nb_cam));
    if (!nodes.empty()) CKA(cudaMemcpyAsync(a->d_nodes, nodes.data(), nodes.size() * sizeof(NodeRec), cudaMemcpyHostToDevice, c->stream));
    CKA(cudaMemcpyAsync(a->d_tris, ctris.data(), nb_tris, cudaMemcpyHostToDevice, c->stream));
    CKA(cudaMemcpyAsync(a->d_shade, H.shade.data(), nb_ref, cudaMemcpyHostToDevice, c->stream));
    if (!camera_tris.empty()) CKA(cudaMemcpyAsync(a->d_camera_tris, camera_tris.data(), camera_tris.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CKA(cudaStreamSynchronize(c->stream));
#undef CKA
    a->view.nodes = a->d_nodes;
    a->view.wnodes = a->d_wnodes;
    a->view.wroot_ref = H.wide_root_ref;
    a->n_wide = H.wide.size();
    /* (env CLV_LARGE_SCENE_BYTES: test hook — lets a small scene run the large-scene kernel flavours) */
    a->large = (nb_nodes + nb_wide + nb_tris) > (getenv("CLV_LARGE_SCENE_BYTES") ? (size_t)strtoull(getenv("CLV_LARGE_SCENE_BYTES"), nullptr, 10) : (size_t)CLV_LARGE_SCENE_BYTES);
    a->binary_depth = H.binary_depth;
    a->wide_depth = H.wide_depth;
    a->view.tris = a->d_tris;
    a->view.shade = a->d_shade;
    a->view.root_min = boxes[0].min;
    a->view.root_max = boxes[0].max;
    a->view.root_ref = root_ref;
    a->view.n_camera_tris = (int32_t)camera_tris.size();
    a->view.camera_tris = a->d_camera_tris;
    *out = a;
    return CLV_OK;
}

extern "C" int clv_accel_destroy(clv_accel *a) {
    if (!a) return CLV_OK;
    if (--a->refs > 0) return CLV_OK; /* a renderer still uses it: freed when the last user lets go */
    cudaSetDevice(a->ctx->device);
    cudaStreamSynchronize(a->ctx->stream);
    dev_free(a->ctx, a->d_nodes);
    dev_free(a->ctx, a->d_wnodes);
    dev_free(a->ctx, a->d_tris);
    dev_free(a->ctx, a->d_shade);
    dev_free(a->ctx, a->d_camera_tris);
    delete a;
    return CLV_OK;
}

extern "C" int clv_accel_stats(clv_accel *a, uint64_t out[12]) {
    REQUIRE(a && out, "null");
    out[0] = a->n_nodes * sizeof(NodeRec); /* binary nodes */
    out[1] = a->n_tris * 3 * sizeof(V4);
    out[2] = a->n_nodes;
    out[3] = a->n_tris;
    out[4] = a->n_wide * sizeof(WideRec);  /* 4-wide nodes */
    out[5] = a->n_wide;
    out[6] = CLV_WIDE_CLOSEST;             /* which of the two each search kind traverses in the ORDERED flavour */
    out[7] = CLV_WIDE_OCCLUSION;
    out[8] = (uint64_t)a->large;
    out[9] = (uint64_t)(a->large ? CLV_TRACE_WIDE_LARGE_VALUE : CLV_TRACE_WIDE_VALUE); /* k_walk_trace steps over 4-wide nodes */
    out[10] = (uint64_t)a->binary_depth;
    out[11] = (uint64_t)a->wide_depth;
    return CLV_OK;
}

/* ---------------------------------------------------------------------------------------------------------------
 * kernels
 * ------------------------------------------------------------------------------------------------------------- */
/* threads per block of the per-sample kernels: 64 measured 1 % faster than 128 at C2 (a block holds its resources until
 * its slowest warp is done; two warps wait less for each other than four) */
#ifndef CLV_BLOCK
#define CLV_BLOCK 64
#endif

template <bool COUNT> __device__ __forceinline__ void flush_counters(Cnt<COUNT> &, Counters *) {}
template <> __device__ __forceinline__ void flush_counters<true>(Cnt<true> &cnt, Counters *dst) {
    /* warp-aggregate, then one atomic per counter per warp */
    unsigned rr = cnt.rr, rt = cnt.rt, nf = cnt.nf, tf = cnt.tf, vx = cnt.vx;
    for (int o = 16; o > 0; o >>= 1) {
        rr += __shfl_xor_sync(0xffffffffu, rr, o);
        rt += __shfl_xor_sync(0xffffffffu, rt, o);
        nf += __shfl_xor_sync(0xffffffffu, nf, o);
        tf += __shfl_xor_sync(0xffffffffu, tf, o);
        vx += __shfl_xor_sync(0xffffffffu, vx, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&dst->rays_ref, (unsigned long long)rr);
        atomicAdd(&dst->rays_traced, (unsigned long long)rt);
        atomicAdd(&dst->node_fetches, (unsigned long long)nf);
        atomicAdd(&dst->tri_fetches, (unsigned long long)tf);
        atomicAdd(&dst->vertices, (unsigned long long)vx);
    }
}

__global__ void __launch_bounds__(CLV_BLOCK)
k_generate_light_rays(uint32_t n, const Triangle *light_triangles, const float *surface_areas,
                      const int32_t *light_triangle_indices, const Material *materials, uint32_t *random_buffer,
                      Ray *out, const int32_t *counts) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    uint2 sd = reinterpret_cast<uint2 *>(random_buffer)[id];
    uint32_t seed0 = sd.x, seed1 = sd.y;
    out[id] = make_light_ray(light_triangles, surface_areas, light_triangle_indices, materials, counts[0], seed0, seed1);
    reinterpret_cast<uint2 *>(random_buffer)[id] = make_uint2(seed0, seed1);
}

__global__ void __launch_bounds__(CLV_BLOCK)
k_generate_camera_rays(uint32_t n, const Camera *camera, uint32_t *random_buffer, const uint32_t *indices, Ray *out) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    Camera c = camera[0];
    uint2 sd = reinterpret_cast<uint2 *>(random_buffer)[id];
    uint32_t seed0 = sd.x, seed1 = sd.y;
    out[id] = make_camera_ray(c, (int)indices[id], seed0, seed1);
    reinterpret_cast<uint2 *>(random_buffer)[id] = make_uint2(seed0, seed1);
}

template <int MODE, bool COUNT>
__global__ void __launch_bounds__(CLV_BLOCK)
k_generate_paths(uint32_t n, const Ray *rays, AccelView A, const Material *materials, uint32_t *random_buffer, V4 *out,
                 Path *output_paths, V4 *float_debug, Counters *counters) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    Cnt<COUNT> cnt;
    if (id < n) {
        uint2 sd = reinterpret_cast<uint2 *>(random_buffer)[id];
        uint32_t seed0 = sd.x, seed1 = sd.y;
        V4 uni;
        AosSink sink;
        sink.path = output_paths + id;
        random_walk<MODE, COUNT>(rays[id], A, materials, seed0, seed1, sink, &uni, cnt);
        out[id] = uni;
        if (float_debug) float_debug[id] = v4(100.0f, 100.0f, 100.0f, 100.0f);
        reinterpret_cast<uint2 *>(random_buffer)[id] = make_uint2(seed0, seed1);
    }
    flush_counters<COUNT>(cnt, counters);
}

__global__ void k_reset_light_indices(uint32_t n, int32_t *pix, int32_t *path, int32_t *ray, float *w, float *shade) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    pix[id] = -1; path[id] = 0; ray[id] = 0; w[id] = 0.0f; shade[id] = 0.0f;
}

/* connect_paths (trace.metal:620-869) as three kernels:
 *   k_connect_enumerate  one thread per sample: cheap rejections, sensor pre-test, append the surviving strategies as
 *                        32-bit tasks (sample << 6 | strategy) to one dense list (one atomicAdd per warp)
 *   k_connect_rays       the occlusion queries of the list.  Small register footprint, no shared memory => high
 *                        occupancy and the whole L1 for BVH nodes.  Each warp owns a chunk of the list; lanes take the
 *                        next task of the chunk as soon as their ray is decided (ballot-based refill, no atomics), one
 *                        traversal round per loop iteration, so lanes stay busy whatever the lengths of their rays
 *   k_connect_finish     one thread per sample: stage MIS terms in shared memory, evaluate the surviving strategies
 *                        in the reference's order, splat weights, aggregator
 */
#ifndef CLV_RAY_CHUNK
#define CLV_RAY_CHUNK 512 /* tasks per warp in k_connect_rays; measured 128 / 256 / 384 / 512 / 1024: 1.068 / 0.993 / 0.980 / 0.983 / 1.120 ms at C2, 256 -> 512: 6.60 -> 6.44 ms at C3 */
#endif
#define CLV_WARPS_PER_BLOCK (CLV_BLOCK / 32)

struct ConnectScratch { /* device buffers shared by the three kernels */
    uint32_t *tasks;      /* [n * 36] */
    uint32_t *task_count; /* [1] */
    uint32_t *cursor;     /* [1]: next task not handed out yet (dynamic work fetch of k_connect_rays) */
    unsigned int *vis;    /* [n][2]: visibility bits of strategies 0..41 */
    float *sensor;        /* [n][CLV_SENSOR_WORDS] */
};

/* this sample's stored vertices in whichever layout the launch uses */
template <bool SOA> struct PathViewOf;
template <> struct PathViewOf<false> {
    typedef AosPaths type;
    static __device__ __forceinline__ AosPaths make(const PathStore &st, uint32_t slot, uint32_t) {
        return aos_paths(st.camera_paths, st.light_paths, slot);
    }
};
template <> struct PathViewOf<true> {
    typedef SoaPaths type;
    static __device__ __forceinline__ SoaPaths make(const PathStore &st, uint32_t slot, uint32_t pixel) {
        SoaPaths P;
        P.v = st.soa; P.len = st.soa_len; P.n = st.n; P.slot = slot; P.pixel0 = (int32_t)pixel;
        return P;
    }
};

template <int MODE, bool COUNT, bool SOA>
__global__ void __launch_bounds__(CLV_BLOCK)
k_connect_enumerate(uint32_t first, uint32_t n, PathStore store, AccelView A, const Material *materials,
                    const Camera *camera, ConnectScratch cs, Counters *counters) {
    __shared__ float cam_scratch[CLV_ENUM_SCRATCH * CLV_BLOCK];
    const uint32_t id = first + blockIdx.x * blockDim.x + threadIdx.x; /* samples [first, n) of the launch's slice */
    const int lane = threadIdx.x & 31;
    Cnt<COUNT> cnt;
    unsigned long long cand = 0ull;
    if (id < n) {
        Camera c = camera[0];
        TArr Sens;
        Sens.base = cs.sensor + (size_t)id * CLV_SENSOR_WORDS;
        Sens.stride = 1;
        TArr Cam;
        Cam.base = cam_scratch + threadIdx.x;
        Cam.stride = CLV_BLOCK;
        cand = connect_enumerate<MODE, COUNT>(PathViewOf<SOA>::make(store, id, 0u), A, materials, c, Sens, Cam, cnt);
        cs.vis[2 * (size_t)id] = 0u;
        cs.vis[2 * (size_t)id + 1] = 0u;
    }
    const int mine = __popcll(cand);
    int incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(cs.task_count, (uint32_t)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    uint32_t pos = base + (uint32_t)(incl - mine);
    while (cand) {
        int q = __ffsll((long long)cand) - 1;
        cand &= cand - 1;
        cs.tasks[pos++] = (id << 6) | (uint32_t)q;
    }
    flush_counters<COUNT>(cnt, counters);
}

/* the warp primitives behind the cooperative leaf steps of clv_device.cuh (tests/hostemu supplies 32 host threads in
 * lockstep instead) */
#ifndef CLV_COOP_LEAVES
#define CLV_COOP_LEAVES 1 /* leaves tested by the whole warp (hit_leaf_step_warp / occ_leaf_step_warp, clv_device.cuh) */
#endif
#ifndef CLV_COOP_LEAVES_RESIDENT
#define CLV_COOP_LEAVES_RESIDENT 0 /* k_connect_rays of the cache-resident flavour (C1 / C2) keeps the per-lane leaf loop */
#endif
struct DeviceWarp {
    int lane_;
    __device__ __forceinline__ int lane() const { return lane_; }
    __device__ __forceinline__ unsigned ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
    __device__ __forceinline__ int shfl(int v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    __device__ __forceinline__ float shfl(float v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
};

#ifndef CLV_RAYS_MIN_BLOCKS
#define CLV_RAYS_MIN_BLOCKS 1 /* no register cap: the kernel takes 63 */
#endif
#ifndef CLV_RAYS_MIN_BLOCKS_LARGE
/* C4, ms per pass of k_connect_rays at 40 / 48 / 56 / 64 registers: 11.22 / 10.80 / 10.55 / 11.20 (before the nearest-child
 * selection of the occlusion step stopped indexing its arrays at run time, which had put them in local memory: 11.07 at 48) */
#define CLV_RAYS_MIN_BLOCKS_LARGE (1152 / CLV_BLOCK) /* = 56 registers per thread */
#endif
template <int MODE, bool COUNT, bool SOA, int LARGE = 0>
__global__ void __launch_bounds__(CLV_BLOCK, LARGE ? CLV_RAYS_MIN_BLOCKS_LARGE : CLV_RAYS_MIN_BLOCKS)
k_connect_rays(PathStore store, AccelView A, const Camera *camera, ConnectScratch cs,
               Counters *counters, int refill_below, int schedule, int dynamic_fetch) {
    const uint32_t count = *cs.task_count;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    /* two ways to deal the task list to the warps: static = warp w owns the CLV_RAY_CHUNK tasks starting at w * chunk
     * (grid sized for the worst case, surplus warps leave at once); dynamic = persistent warps take the next tasks of
     * the list whenever they refill (one atomicAdd per warp and refill), so no warp is left with a long chunk at the end */
    const uint64_t first = dynamic_fetch ? 0ull : (uint64_t)warp * CLV_RAY_CHUNK;
    if (first >= count) return; /* whole warp */
    int total = dynamic_fetch ? 0x7fffffff : (int)min((uint64_t)CLV_RAY_CHUNK, (uint64_t)count - first);
    const uint32_t *tasks = cs.tasks + first;
    Cnt<COUNT> cnt;
    Camera c = camera[0];
    OccRay ray;
    int stack[CLV_OCC_STACK];
    bool have = false;
    uint32_t task = 0;
    int next = 0; /* warp-uniform: first task of the chunk not handed out yet */
    for (;;) {
        const unsigned holding = __ballot_sync(0xffffffffu, have);
        if (next < total && __popc(holding) < refill_below) {
            const unsigned idle = ~holding;
            if (dynamic_fetch) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(cs.cursor, (uint32_t)__popc(idle));
                base = __shfl_sync(0xffffffffu, base, 0);
                next = (int)min(base, count);            /* this refill hands out [base, base + idle lanes) */
                if ((uint64_t)base + (uint64_t)__popc(idle) >= (uint64_t)count) total = (int)count; /* list exhausted */
                else total = 0x7fffffff;
            }
            const int my = next + __popc(idle & ((1u << lane) - 1u));
            if (!have && my < (dynamic_fetch ? (int)count : total)) {
                task = tasks[my];
                const uint32_t slot = task >> 6;
                const int q = (int)(task & 63u);
                TArr Sens;
                Sens.base = cs.sensor + (size_t)slot * CLV_SENSOR_WORDS;
                Sens.stride = 1;
                int st = connect_ray_begin<MODE, COUNT>(q, PathViewOf<SOA>::make(store, slot, 0u), Sens, A, c, ray, cnt);
                have = st == OCC_RUNNING;
                if (st == OCC_CLEAR) atomicOr(&cs.vis[2 * (size_t)slot + (q >> 5)], 1u << (q & 31));
            }
            next += __popc(idle);
            if (dynamic_fetch && total != 0x7fffffff) next = total; /* nothing left to fetch: stop refilling */
        } else if (holding == 0u) {
            break;
        }
        int st = OCC_RUNNING;
        if (schedule == 0) {
            /* round form: every lane walks down to its next leaf, then every lane tests its leaf */
            if (have) st = occ_round<COUNT>(A, ray, stack, cnt);
        } else {
            /* majority vote: per iteration the warp runs ONE kind of step — an inner node for the lanes that sit at
             * one, or a whole leaf for the lanes that hold one — whichever more lanes wait for */
            const bool at_inner = have && ray.cur >= 0;
            const int n_inner = __popc(__ballot_sync(0xffffffffu, at_inner));
            const int n_leaf = __popc(__ballot_sync(0xffffffffu, have && !at_inner));
            if (n_inner * schedule >= n_leaf * 4) { if (at_inner) st = occ_inner_step<COUNT>(A, ray, stack, cnt); }
            else if (CLV_COOP_LEAVES && (LARGE || CLV_COOP_LEAVES_RESIDENT)) { DeviceWarp w{lane}; st = occ_leaf_step_warp<COUNT>(w, A, ray, stack, have && !at_inner, cnt); }
            else if (have && !at_inner) st = occ_leaf_step<COUNT>(A, ray, stack, cnt);
        }
        if (st != OCC_RUNNING) {
            have = false;
            const uint32_t slot = task >> 6;
            const int q = (int)(task & 63u);
            if (st == OCC_CLEAR) atomicOr(&cs.vis[2 * (size_t)slot + (q >> 5)], 1u << (q & 31));
        }
    }
    flush_counters<COUNT>(cnt, counters);
}

/* no minimum-blocks bound here: measured 0.81 ms unbounded (90 registers) vs 0.83 / 0.94 / 1.32 ms at 5 / 6 / 7 blocks */
#ifdef CLV_FINISH_MIN_BLOCKS
#define CLV_FINISH_BOUNDS __launch_bounds__(CLV_BLOCK, CLV_FINISH_MIN_BLOCKS)
#else
#define CLV_FINISH_BOUNDS __launch_bounds__(CLV_BLOCK)
#endif
template <bool SPLAT, bool COUNT, bool SOA>
__global__ void CLV_FINISH_BOUNDS
k_connect_finish(uint32_t first, uint32_t n, const uint32_t *sample_ids, PathStore store,
                 const Material *materials, const Camera *camera, ConnectScratch cs, WeightAggregator *aggs, V4 *out,
                 LightSink sink, Counters *counters) {
    __shared__ float scratch[CLV_CONNECT_SCRATCH * CLV_BLOCK];
    const uint32_t id = first + blockIdx.x * blockDim.x + threadIdx.x;
    Cnt<COUNT> cnt;
    if (id == first) { *cs.task_count = 0u; *cs.cursor = 0u; } /* every ray of this pass is done (stream order): reset the list for the next */
    const unsigned warp_lanes = __ballot_sync(0xffffffffu, id < n);
    if (id < n) {
        Camera c = camera[0];
        TArr S, Sens;
        S.base = scratch + threadIdx.x;
        S.stride = CLV_BLOCK;
        Sens.base = cs.sensor + (size_t)id * CLV_SENSOR_WORDS;
        Sens.stride = 1;
        /* staged ABI: aggregator slot == id.  Fused renderer: aggregators are per pixel */
        uint32_t agg_slot = sample_ids ? sample_ids[id] : id;
        typedef typename PathViewOf<SOA>::type PV;
        ConnectCtx<PV> X;
        unsigned long long valid0;
        connect_stage<COUNT>(PathViewOf<SOA>::make(store, id, agg_slot), materials, S, Sens, X, valid0, cnt);
        unsigned long long visible = (unsigned long long)cs.vis[2 * (size_t)id] | ((unsigned long long)cs.vis[2 * (size_t)id + 1] << 32);
        connect_finish<SPLAT>(X, valid0 | visible, id, agg_slot, materials, c, aggs, out, sink, warp_lanes);
    }
    flush_counters<COUNT>(cnt, counters);
}

/* k_connect_finish, warp-cooperative form.  The per-sample form above leaves lanes idle twice over: samples have different
 * numbers of surviving strategies (0..36) and strategies have different loop lengths (s, t), ncu: 13.8 of 32 lanes.  Here a
 * warp owns 32 samples and, for t = 1..6 in turn, deals the surviving (sample, s) strategies of that t — listed s-major, so
 * neighbouring lanes run the same (s, t) — to its 32 lanes, one strategy per lane (connect_strategy: same operations as the
 * sequential form); the terms land in shared memory and every sample then adds ITS OWN terms in ascending s.  Sums are
 * therefore formed in the reference's order (t outer, s inner): same bits. */
template <bool SPLAT, bool COUNT, bool SOA>
__global__ void __launch_bounds__(CLV_BLOCK)
k_connect_finish_coop(uint32_t first, uint32_t n, const uint32_t *sample_ids, PathStore store, const Material *materials,
                      const Camera *camera, ConnectScratch cs, WeightAggregator *aggs, V4 *out, LightSink sink,
                      Counters *counters) {
    __shared__ float scratch[CLV_CONNECT_SCRATCH * CLV_BLOCK];
    __shared__ float pback[6 * CLV_BLOCK];
    __shared__ float4 results[CLV_WARPS_PER_BLOCK][7 * 32];
    __shared__ int ctx_words[4 * CLV_BLOCK]; /* len_c, len_l, spec_l, spec_c of every sample of the block */
    const uint32_t id = first + blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp_in_block = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    Cnt<COUNT> cnt;
    if (id == first) { *cs.task_count = 0u; *cs.cursor = 0u; }
    const bool active = id < n;
    const Camera c = camera[0];
    typedef typename PathViewOf<SOA>::type PV;
    ConnectCtx<PV> X;
    unsigned long long survivors = 0ull;
    uint32_t agg_slot = 0;
    if (active) {
        TArr S, Sens;
        S.base = scratch + threadIdx.x;
        S.stride = CLV_BLOCK;
        Sens.base = cs.sensor + (size_t)id * CLV_SENSOR_WORDS;
        Sens.stride = 1;
        agg_slot = sample_ids ? sample_ids[id] : id;
        unsigned long long valid0;
        connect_stage<COUNT>(PathViewOf<SOA>::make(store, id, agg_slot), materials, S, Sens, X, valid0, cnt);
        survivors = valid0 | (unsigned long long)cs.vis[2 * (size_t)id] | ((unsigned long long)cs.vis[2 * (size_t)id + 1] << 32);
        ctx_words[4 * threadIdx.x + 0] = X.len_c;
        ctx_words[4 * threadIdx.x + 1] = X.len_l;
        ctx_words[4 * threadIdx.x + 2] = (int)X.spec_l;
        ctx_words[4 * threadIdx.x + 3] = (int)X.spec_c;
    }
    __syncwarp();
    const bool spec_sensor = materials[7].type > 0;
    V3 total = v3(0, 0, 0);
    float weight_sum = 0.0f;
    TArr PB;
    PB.base = pback + threadIdx.x;
    PB.stride = CLV_BLOCK;
    const uint32_t warp_first_id = id - (uint32_t)lane;
    const int warp_first_tid = (int)threadIdx.x - lane;
    for (int t = 1; t <= CLV_MAX_BOUNCES; t++) {
        const unsigned bits = (unsigned)(survivors >> ((t - 1) * 7)) & 0x7fu;
        unsigned m[7];
        int off[8];
        off[0] = 0;
#pragma unroll
        for (int s = 0; s < 7; s++) {
            m[s] = __ballot_sync(0xffffffffu, (bits >> s) & 1u);
            off[s + 1] = off[s] + __popc(m[s]);
        }
        const int n_tasks = off[7]; /* warp-uniform */
        if (n_tasks == 0) continue;
        for (int j = lane; j < n_tasks; j += 32) {
            int s = 0;
#pragma unroll
            for (int k = 1; k < 7; k++) s += (j >= off[k]) ? 1 : 0;
            /* off[] and m[] are indexed with a lane-dependent s: select without local-memory arrays */
            unsigned ms = m[0];
            int os = off[0];
#pragma unroll
            for (int k = 1; k < 7; k++) { if (s == k) { ms = m[k]; os = off[k]; } }
            const int src = (int)__fns(ms, 0u, j - os + 1);
            const uint32_t src_id = warp_first_id + (uint32_t)src;
            const int src_tid = warp_first_tid + src;
            ConnectCtx<PV> Y;
            Y.P = PathViewOf<SOA>::make(store, src_id, 0u);
            Y.len_c = ctx_words[4 * src_tid + 0];
            Y.len_l = ctx_words[4 * src_tid + 1];
            Y.spec_l = (unsigned)ctx_words[4 * src_tid + 2];
            Y.spec_c = (unsigned)ctx_words[4 * src_tid + 3];
            Y.spec_sensor = spec_sensor;
            Y.S.base = scratch + src_tid;
            Y.S.stride = CLV_BLOCK;
            Y.Sens.base = cs.sensor + (size_t)src_id * CLV_SENSOR_WORDS;
            Y.Sens.stride = 1;
            V3 term;
            float w;
            bool add;
            connect_strategy<SPLAT>(Y, (t - 1) * 7 + s, PB, src_id, materials, c, sink, term, w, add);
            results[warp_in_block][j] = make_float4(term.x, term.y, term.z, add ? w : -1.0f); /* an accepted w is > 0 (or NaN) */
        }
        __syncwarp();
        if (bits) {
#pragma unroll
            for (int s = 0; s < 7; s++) {
                if ((bits >> s) & 1u) {
                    const float4 r = results[warp_in_block][off[s] + __popc(m[s] & lt_mask)];
                    if (!(r.w < 0.0f)) {
                        total = total + v3(r.x, r.y, r.z);
                        weight_sum += r.w;
                    }
                }
            }
        }
        __syncwarp();
    }
    if (active) connect_write_aggregator(X.P, total, weight_sum, agg_slot, c, aggs, out);
    flush_counters<COUNT>(cnt, counters);
}

__global__ void __launch_bounds__(CLV_BLOCK)
k_adaptive_finalize_samples(uint32_t n, const WeightAggregator *aggs, const Camera *camera, V4 *out,
                            uint32_t *sample_counts, const uint32_t *bin_offsets, float *sample_weights) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    Camera c = camera[0];
    V3 total;
    float wsum;
    uint32_t count;
    finalize_pixel(id, aggs, c, bin_offsets, total, wsum, count);
    sample_counts[id] = count;
    out[id] = v4(total.x, total.y, total.z, 1.0f);
    sample_weights[id] = wsum;
}

/* one bitonic compare-exchange step, trace.metal:872-934 */
__global__ void k_light_sort(uint32_t n_threads, int32_t *pix, int32_t *path, int32_t *ray, float *w, float *shade,
                             uint32_t stage, uint32_t pass_of_stage, uint32_t n, uint32_t pairs_per_thread) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_threads) return;
    for (uint32_t p = 0; p < pairs_per_thread; ++p) {
        uint32_t gp = id * pairs_per_thread + p;
        uint32_t dist = 1u << (pass_of_stage - 1);
        uint32_t block_width = 1u << stage;
        uint32_t l = (gp / dist) * dist * 2u + (gp % dist);
        uint32_t r = l + dist;
        if (r >= n || l >= n) continue;
        bool ascending = ((gp & (block_width >> 1)) == 0u);
        int32_t lp = pix[l], rp = pix[r];
        if ((ascending && lp > rp) || (!ascending && lp < rp)) {
            pix[l] = rp; pix[r] = lp;
            int32_t ti = path[l]; path[l] = path[r]; path[r] = ti;
            ti = ray[l]; ray[l] = ray[r]; ray[r] = ti;
            float tf = w[l]; w[l] = w[r]; w[r] = tf;
            tf = shade[l]; shade[l] = shade[r]; shade[r] = tf;
        }
    }
}

/* light_image_gather, trace.metal:937-964 */
__global__ void k_light_image_gather(uint32_t n, const Path *light_paths, const Material *materials,
                                     const int32_t *path_indices, const int32_t *ray_indices, const int32_t *bins,
                                     uint32_t offset, const float *weights, const float *shades, V4 *light_image,
                                     float *sum_weights) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    uint32_t start = (uint32_t)bins[id] + offset, end = (uint32_t)bins[id + 1] + offset;
    V3 total = v3(0, 0, 0);
    float weight_sum = 0.0f;
    for (uint32_t i = start; i < end; i++) {
        const Path *path = light_paths + path_indices[i];
        int ray_idx = ray_indices[i];
        const Ray *ray = path->rays + ray_idx;
        const Ray *prior = path->rays + (ray_idx - 1 > 0 ? ray_idx - 1 : 0);
        V3 mc = xyz(ld4(&materials[ray->material].color));
        total = total + (xyz(prior->color) * (weights[i] * shades[i])) * mc;
        weight_sum += weights[i];
    }
    light_image[id] = v4(total.x, total.y, total.z, 1.0f);
    sum_weights[id] += weight_sum;
}

/* device-side replacement of sort + host bins + gather for the staged arrays: one thread per entry, atomics */
__global__ void k_light_image_splat(uint32_t n_entries, uint32_t n_pixels, const Path *light_paths,
                                    const Material *materials, const int32_t *pixel_indices, const int32_t *path_indices,
                                    const int32_t *ray_indices, const float *weights, const float *shades,
                                    V4 *light_image, float *sum_weights) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_entries) return;
    int32_t px = pixel_indices[i];
    if (px < 0 || (uint32_t)px >= n_pixels) return;
    const Path *path = light_paths + path_indices[i];
    int ray_idx = ray_indices[i];
    const Ray *ray = path->rays + ray_idx;
    const Ray *prior = path->rays + (ray_idx - 1 > 0 ? ray_idx - 1 : 0);
    V3 mc = xyz(ld4(&materials[ray->material].color));
    V3 contrib = (xyz(prior->color) * (weights[i] * shades[i])) * mc;
    atomicAdd(&light_image[px].x, contrib.x);
    atomicAdd(&light_image[px].y, contrib.y);
    atomicAdd(&light_image[px].z, contrib.z);
    atomicAdd(&sum_weights[px], weights[i]);
}
__global__ void k_set_light_image(uint32_t n, V4 *light_image) { /* gather writes (0,0,0,1) for empty pixels */
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) light_image[i] = v4(0.0f, 0.0f, 0.0f, 1.0f);
}

/* ---- fused per-pass kernels (Renderer.run_sample without host round trips) ---------------------------------------- */
/* stages 1-4: both ray generators and both random walks of one sample, RNG streams in the reference's order
 * (light gen, camera gen, light walk, camera walk; renderer.py:281-285) */
/* measured at C2: 1.36 ms unbounded (94 regs), 1.19 ms at 6 blocks/SM (80 regs), 1.07 ms at 8 (64 regs, some spills) */
#ifndef CLV_WALK_MIN_BLOCKS
#define CLV_WALK_MIN_BLOCKS (1024 / CLV_BLOCK) /* = 64 registers per thread */
#endif
#ifndef CLV_WALK_MIN_BLOCKS_LARGE
#define CLV_WALK_MIN_BLOCKS_LARGE (2048 / CLV_BLOCK) /* = 32 registers per thread, every warp slot of the SM filled */
#endif
template <int MODE, bool COUNT, bool SOA, int LARGE = 0>
__global__ void __launch_bounds__(CLV_BLOCK, LARGE ? CLV_WALK_MIN_BLOCKS_LARGE : CLV_WALK_MIN_BLOCKS)
k_fused_walks(uint32_t first, uint32_t n_local, const uint32_t *sample_ids, const Triangle *light_triangles, const float *surface_areas,
              const int32_t *light_triangle_indices, int light_count, const Material *materials, const Camera *camera,
              AccelView A, uint32_t *seeds, PathStore store, Accum *accum, Counters *counters) {
    uint32_t id = first + blockIdx.x * blockDim.x + threadIdx.x;
    Cnt<COUNT> cnt;
    if (id < n_local) {
        Camera c = camera[0];
        uint32_t gid = sample_ids ? sample_ids[id] : id;
        uint2 sd = reinterpret_cast<uint2 *>(seeds)[id];
        uint32_t seed0 = sd.x, seed1 = sd.y;
        fused_walk_sample<MODE, COUNT, SOA>(id, gid, light_triangles, surface_areas, light_triangle_indices, light_count,
                                            materials, c, A, seed0, seed1, store, accum, cnt);
        reinterpret_cast<uint2 *>(seeds)[id] = make_uint2(seed0, seed1);
    }
    flush_counters<COUNT>(cnt, counters);
}

/* ---- wavefront form of stages 1-4: k_walk_init, then per round k_walk_trace + k_walk_shade ----------- *
 * The megakernel above runs a sample's <= 12 closest-hit searches inside one thread, bounce-synchronously: a warp waits for
 * its slowest ray at every bounce and ncu counts 5-11 of 32 lanes per instruction.  Here a round traces ray number r of
 * every sample that still has one:
 *   k_walk_trace  persistent warps.  A lane that finishes its search takes the next ray of the round's queue (one
 *                 atomicAdd per warp and refill: warp-level work fetch), and per iteration the warp runs ONE kind of
 *                 step — an inner node or a whole leaf — for the lanes that wait for it, chosen by majority vote
 *   k_walk_shade  one thread per traced ray: BSDF sampling, vertex store, next ray; survivors are appended to the next
 *                 round's queue (one atomicAdd per warp)
 * Same per-ray operations in the same order as the megakernel: bit-identical results. */
/* control words of the pipelined form (k_walk_pipeline, below) */
enum { PIPE_TRACE_TAIL = 0, PIPE_TRACE_HEAD, PIPE_TRACE_AVAIL, PIPE_DONE_TAIL, PIPE_DONE_HEAD, PIPE_DONE_AVAIL, PIPE_LIVE, PIPE_ERROR };
struct WalkQueues {
    uint32_t *q[2];   /* ping-pong lists of sample slots */
    uint32_t *count;  /* [CLV_WALK_LAUNCH_ROUNDS + 1]: entries of round r's list (round 0 = every sample of the slice, implicit) */
    uint32_t *cursor; /* [CLV_WALK_LAUNCH_ROUNDS]: next unfetched entry of round r's list */
    uint32_t *pool;   /* [CLV_WALK_LAUNCH_ROUNDS]: continuation records round r has taken */
};
#define CLV_WALK_SAMPLES_PER_RECORD 32
#define CLV_WALK_RECORDS_MIN 1024
#define CLV_WALK_COUNTER_WORDS 64 /* per slice: count[17] + cursor[16] + pool[16], padded; k_walk_init clears them */
static_assert(3 * CLV_WALK_LAUNCH_ROUNDS + 1 <= CLV_WALK_COUNTER_WORDS && CLV_WALK_COUNTER_WORDS <= CLV_BLOCK, "walk counters");
/* continuation records of one slice (clv_device.cuh, hit_save): round r writes into half r & 1 and resumes from the other */
struct WalkRecords {
    V4 *half[2];
    uint32_t capacity; /* records per half; a search that finds the pool full simply keeps going */
    int tail_budget;   /* iterations a warp keeps stepping after the round's queue ran dry before it suspends; 0 = never */
};

template <bool SOA>
__global__ void __launch_bounds__(CLV_BLOCK)
k_walk_init(uint32_t first, uint32_t n_end, const Triangle *light_triangles, const float *surface_areas,
            const int32_t *light_triangle_indices, int light_count, const Material *materials, uint32_t *seeds,
            PathStore store, WalkState ws, WalkQueues wq, uint32_t *pipe_ctrl) {
    const uint32_t id = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x < CLV_WALK_COUNTER_WORDS) wq.count[threadIdx.x] = 0u; /* count[] and cursor[] are one block */
    if (pipe_ctrl && blockIdx.x == 0 && threadIdx.x < 8) /* pipelined form: every sample's first ray is an implicit entry */
        pipe_ctrl[threadIdx.x] = (threadIdx.x == PIPE_TRACE_TAIL || threadIdx.x == PIPE_TRACE_AVAIL || threadIdx.x == PIPE_LIVE) ? n_end - first : 0u;
    if (id >= n_end) return;
    uint2 sd = reinterpret_cast<uint2 *>(seeds)[id];
    uint32_t seed0 = sd.x, seed1 = sd.y;
    wavefront_init_sample<SOA>(id, light_triangles, surface_areas, light_triangle_indices, light_count, materials, seed0, seed1,
                               store, ws);
    reinterpret_cast<uint2 *>(seeds)[id] = make_uint2(seed0, seed1);
}

#ifndef CLV_TRACE_MIN_BLOCKS
#define CLV_TRACE_MIN_BLOCKS (1024 / CLV_BLOCK) /* <= 64 registers per thread */
#endif
#ifndef CLV_TRACE_MIN_BLOCKS_LARGE
/* measured at C4 / C3 (ms per pass of all trace rounds; binary nodes): 64 / 48 / 40 / 32 registers -> 13.9 / 13.2 / 13.6 /
 * 14.9 and 4.64 / 4.51 / 4.81 / 5.62; over the 4-wide nodes: 64 / 48 / 40 -> 12.2 / 12.4 / 13.5 and 4.33 / 4.52 / 5.07:
 * the kernel does not respond to occupancy (profiles/r02_notes.md) */
#define CLV_TRACE_MIN_BLOCKS_LARGE (1024 / CLV_BLOCK) /* <= 64 registers per thread */
#endif
template <bool COUNT, int LARGE>
__global__ void __launch_bounds__(CLV_BLOCK, LARGE ? CLV_TRACE_MIN_BLOCKS_LARGE : CLV_TRACE_MIN_BLOCKS)
k_walk_trace(uint32_t first, uint32_t n_identity, const uint32_t *queue, const uint32_t *count_ptr, uint32_t *cursor,
             WalkState ws, AccelView A, Counters *counters, int refill_below, int schedule, const V4 *rec_in,
             const uint32_t *rec_in_count, V4 *rec_out, uint32_t *pool, uint32_t pool_capacity, int tail_budget) {
    /* the round's work: first the searches the previous round suspended, then the round's queue */
    const uint32_t n_resume = rec_in_count ? min(*rec_in_count, pool_capacity) : 0u;
    const uint32_t count = n_resume + (queue ? *count_ptr : n_identity);
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    constexpr bool WIDE = LARGE ? (CLV_TRACE_WIDE_LARGE != 0) : (CLV_TRACE_WIDE != 0);
    Cnt<COUNT> cnt;
    HitRay ray;
    StackEntry st[CLV_HIT_STACK(WIDE)];
    bool have = false, exhausted = false; /* exhausted is warp-uniform */
    uint32_t slot = 0;
    int tail = 0;
    for (;;) {
        unsigned holding = __ballot_sync(0xffffffffu, have);
        if (exhausted && tail_budget > 0 && ++tail == tail_budget) {
            /* the round is over for everybody else: suspend the searches still running (a sample may wait
             * CLV_WALK_MAX_LAG rounds in total; after that its searches run to their end as they always did) */
            const bool defer = have && ws.lag[slot] < CLV_WALK_MAX_LAG;
            const unsigned deferring = __ballot_sync(0xffffffffu, defer);
            if (deferring) {
                uint32_t base = 0;
                if (lane == __ffs(deferring) - 1) base = atomicAdd(pool, (uint32_t)__popc(deferring));
                base = __shfl_sync(0xffffffffu, base, __ffs(deferring) - 1);
                const uint32_t rec = base + (uint32_t)__popc(deferring & lt_mask);
                if (defer && rec < pool_capacity) {
                    hit_save(rec_out + (size_t)rec * CLV_HIT_RECORD_V4, slot, ray, st);
                    stcg4(ws.hit + slot, v4(i2f(CLV_HIT_PENDING), 0.0f, 0.0f, 0.0f));
                    ws.lag[slot]++;
                    have = false;
                }
                holding = __ballot_sync(0xffffffffu, have);
            }
        }
        if (!exhausted && __popc(holding) < refill_below) {
            /* warp-level work fetch: the idle lanes take the next entries of the round's queue */
            const unsigned idle = ~holding;
            const int n_idle = __popc(idle);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(cursor, (uint32_t)n_idle);
            base = __shfl_sync(0xffffffffu, base, 0);
            const uint32_t my = base + (uint32_t)__popc(idle & lt_mask);
            if (!have && my < count) {
                if (my < n_resume) { /* a search the previous round suspended: go on where it stopped */
                    slot = hit_restore(rec_in + (size_t)my * CLV_HIT_RECORD_V4, ray, st);
                    have = true;
                } else {
                    slot = queue ? queue[my - n_resume] : first + my;
                    const V4 a0 = ldcg4(ws.s0 + slot), a1 = ldcg4(ws.s1 + slot);
                    ray.o = xyz(a0); ray.d = xyz(a1); ray.inv = rcp(ray.d);
                    cnt.ray_ref();
                    if (hit_begin<COUNT, WIDE>(A, ray, cnt) == HIT_RUNNING) {
                        have = true;
                        if (CLV_PREFETCH_NEXT) prefetch_ref<WIDE>(WIDE ? A.wnodes : A.nodes, A.tris, ray.cur);
                    }
                    else stcg4(ws.hit + slot, v4(i2f(-1), CLV_INF, 0.0f, 0.0f));
                }
            }
            if ((uint64_t)base + (uint64_t)n_idle >= (uint64_t)count) exhausted = true;
        } else if (holding == 0u) {
            break;
        }
        /* majority vote: one inner-node step for the lanes at inner nodes, or one whole leaf for the lanes at leaves */
        const bool at_inner = have && ray.cur >= 0;
        const int n_inner = __popc(__ballot_sync(0xffffffffu, at_inner));
        const int n_leaf = __popc(__ballot_sync(0xffffffffu, have && !at_inner));
        int status = HIT_RUNNING;
        if (n_inner * schedule >= n_leaf * 4) {
            if (at_inner) status = WIDE ? hit_inner_step_wide<COUNT>(A, ray, st, cnt) : hit_inner_step<COUNT>(A, ray, st, cnt);
        } else if (CLV_COOP_LEAVES) { DeviceWarp w{lane}; status = hit_leaf_step_warp<COUNT>(w, A, ray, st, have && !at_inner, cnt); }
        else if (have && !at_inner) status = hit_leaf_step<COUNT, WIDE>(A, ray, st, cnt);
        if (status == HIT_DONE) {
            have = false;
            stcg4(ws.hit + slot, v4(i2f(ray.best_i), ray.best_t, ray.bu, ray.bv));
        }
    }
    flush_counters<COUNT>(cnt, counters);
}

#ifndef CLV_SHADE_MIN_BLOCKS
/* measured at C4 / C3 (ms per pass of all shade rounds): uncapped (117 registers) / 96 / 80 / 64 / 48 / 40 / 32 ->
 * 4.17 / 3.46 / 3.42 / 2.70 / 3.57 / 3.90 / 4.18 and 2.05 / 1.66 / 1.65 / 1.24 / 1.70 / 1.81 / 1.98 */
#define CLV_SHADE_MIN_BLOCKS (1024 / CLV_BLOCK) /* <= 64 registers per thread */
#endif
template <bool COUNT, bool SOA>
__global__ void __launch_bounds__(CLV_BLOCK, CLV_SHADE_MIN_BLOCKS)
k_walk_shade(uint32_t first, uint32_t n_identity, const uint32_t *queue, const uint32_t *count_ptr, uint32_t *queue_out,
             uint32_t *count_out, const uint32_t *sample_ids, const Material *materials, const Camera *camera, AccelView A,
             uint32_t *seeds, PathStore store, WalkState ws, Accum *accum, Counters *counters, const V4 *rec_in,
             const uint32_t *rec_in_count, uint32_t pool_capacity) {
    /* the same work list as the round's k_walk_trace: the samples of the resumed searches, then the queue */
    const uint32_t n_resume = rec_in_count ? min(*rec_in_count, pool_capacity) : 0u;
    const uint32_t count = n_resume + (queue ? *count_ptr : n_identity);
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    Cnt<COUNT> cnt;
    const Camera c = camera[0];
    for (uint64_t w0 = (uint64_t)warp * 32u; w0 < count; w0 += (uint64_t)n_warps * 32u) { /* warp-uniform trip count */
        const uint32_t idx = (uint32_t)w0 + (uint32_t)lane;
        int next = WALK_OVER;
        uint32_t slot = 0;
        if (idx < count) {
            slot = idx < n_resume ? hit_record_slot(rec_in + (size_t)idx * CLV_HIT_RECORD_V4) : (queue ? queue[idx - n_resume] : first + idx);
            const uint32_t gid = sample_ids ? sample_ids[slot] : slot;
            uint2 sd = reinterpret_cast<uint2 *>(seeds)[slot];
            uint32_t seed0 = sd.x, seed1 = sd.y;
            next = wavefront_shade_sample<COUNT, SOA>(slot, gid, materials, c, A, seed0, seed1, store, ws, accum, cnt);
            if (seed0 != sd.x || seed1 != sd.y) reinterpret_cast<uint2 *>(seeds)[slot] = make_uint2(seed0, seed1);
        }
        /* WALK_WAIT: the search was suspended again; its record is already on the next round's list */
        const unsigned staying = __ballot_sync(0xffffffffu, next == WALK_NEXT);
        if (staying) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(count_out, (uint32_t)__popc(staying));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (next == WALK_NEXT) queue_out[base + (uint32_t)__popc(staying & ((1u << lane) - 1u))] = slot;
        }
    }
    flush_counters<COUNT>(cnt, counters);
}

/* ---- pipelined form of the wavefront walks: ONE persistent kernel per slice instead of 12 x (trace, shade) ------------- *
 * The round form above ends every round with a tail: the last rays of a round are long (hundreds of steps, each a
 * dependent memory round trip) and run alone while every other warp has left — ncu: the SMs are active 75 % of a trace
 * launch — and a shard of the frame (multi-GPU) pays the same absolute tail on fewer rays.  Here nothing waits for a
 * round: a warp that still holds rays keeps stepping them; a warp that holds none takes work from one of two queues —
 *   trace queue  sample slots whose next ray is ready (entries 0..n0-1 are implicit: every sample's first ray)
 *   done queue   sample slots whose ray has been traced and whose hit record waits for its bounce
 * — shading a batch (which appends the survivors to the trace queue) or refilling its lanes with rays.  Both queues are
 * linear (a sample contributes at most 12 entries per pass to each), reserved with one atomic per warp; an entry is valid
 * once its slot word is written (the buffer starts as EMPTY).  Records that cross SMs inside the kernel (walk state, hit
 * records, seeds) are read with L1-bypassing loads after a __threadfence handshake through the queue entry.  The kernel
 * ends when the count of live samples reaches zero.  Same per-ray / per-bounce operations as the other two forms:
 * bit-identical results. */
struct WalkPipe {
    uint32_t *ctrl;    /* [8]: per queue (trace, done) tail = reserved, head = claimed, avail = published and not claimed;
                        * live samples; error */
    uint32_t *trace_q; /* explicit entries: index i >= n0 lives at trace_q[i - n0]; capacity 11 * n0 */
    uint32_t *done_q;  /* capacity 12 * n0 */
    uint32_t n0, first;
};
#define CLV_PIPE_EMPTY 0xffffffffu

/* Claim up to `want` entries of a queue (q = PIPE_TRACE_TAIL or PIPE_DONE_TAIL: tail, head, avail follow each other):
 * returns the first claimed index and the number claimed.  `avail` counts entries that are published and not claimed; a
 * claimer subtracts what it wants, keeps what the old value covered and gives the rest back — two or three atomics, no
 * retry loop.  (A compare-and-swap on the head, the first version, lets ONE of thousands of contending warps succeed per
 * memory round trip: the whole kernel then ran at one claim per microsecond.) */
__device__ __forceinline__ uint32_t pipe_claim(volatile uint32_t *ctrl, int q, int want, int lane, int *got) {
    uint32_t base = 0;
    int k = 0;
    if (lane == 0 && (int)ctrl[q + 2] > 0) { /* a look first: no atomics on an empty queue */
        const int old = atomicSub((int *)const_cast<uint32_t *>(ctrl) + q + 2, want);
        k = old < 0 ? 0 : (old < want ? old : want);
        if (k < want) atomicAdd((int *)const_cast<uint32_t *>(ctrl) + q + 2, want - k);
        if (k > 0) base = atomicAdd(const_cast<uint32_t *>(ctrl) + q + 1, (uint32_t)k);
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    *got = __shfl_sync(0xffffffffu, k, 0);
    return base;
}
/* publish `count` entries the warp has just written (every writing lane fenced before its store) */
__device__ __forceinline__ void pipe_publish(volatile uint32_t *ctrl, int q, int count, int lane) {
    __syncwarp();
    if (lane == 0) {
        __threadfence();
        atomicAdd((int *)const_cast<uint32_t *>(ctrl) + q + 2, count);
    }
}
/* the slot word of a claimed entry (the producer reserved it before us and writes it a few instructions later) */
__device__ __forceinline__ uint32_t pipe_wait(const volatile uint32_t *q, uint32_t index) {
    uint32_t v = q[index];
    while (v == CLV_PIPE_EMPTY) { __nanosleep(64); v = q[index]; }
    __threadfence();
    return v;
}

template <bool COUNT, bool SOA, int LARGE>
__global__ void __launch_bounds__(CLV_BLOCK, 1024 / CLV_BLOCK)
k_walk_pipeline(WalkPipe wp, const uint32_t *sample_ids, const Material *materials, const Camera *camera, AccelView A,
                uint32_t *seeds, PathStore store, WalkState ws, Accum *accum, Counters *counters, int refill_below, int schedule,
                int shade_every) {
    constexpr bool WIDE = LARGE ? (CLV_TRACE_WIDE_LARGE != 0) : (CLV_TRACE_WIDE != 0);
    /* Roles.  shade_every = 0: every warp does both jobs, shading only while it holds no ray — each batch of rays then ends
     * with the warp draining lane by lane before it may shade (measured: 1.5x slower than the round form).  shade_every =
     * k > 0: every k-th block only shades, the others only trace and refill their lanes whenever the trace queue has
     * entries, as in the round form; blocks of both roles alternate in launch order, so any resident prefix of the grid
     * holds both (the grid never exceeds what the device can hold at once). */
    const bool shade_role = shade_every > 0 && (blockIdx.x % (unsigned)shade_every) == (unsigned)shade_every - 1u;
    const bool trace_role = shade_every <= 0 || !shade_role;
    /* per warp: rays fetched from the trace queue 32 at a time and handed to the lanes without further atomics, and the
     * slots of finished rays collected until a batch is worth one reservation in the done queue.  (One global atomic per
     * refill and per finished ray — 10^9 a second on two addresses — made the first version of this kernel 80x slower.) */
    __shared__ uint32_t in_buf[CLV_WARPS_PER_BLOCK][32], out_buf[CLV_WARPS_PER_BLOCK][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    volatile uint32_t *ctrl = wp.ctrl;
    const volatile uint32_t *trace_q = wp.trace_q, *done_q = wp.done_q;
    const Camera c = camera[0];
    Cnt<COUNT> cnt;
    HitRay ray;
    StackEntry st[CLV_HIT_STACK(WIDE)];
    bool have = false;
    uint32_t slot = 0;
    unsigned idle_polls = 0, idle_sleep = 128;
    int in_pos = 0, in_count = 0, out_count = 0; /* warp-uniform */
    int skip_polls = 0, backoff = 1;             /* warp-uniform: iterations until the global trace queue is looked at again */
    auto flush_done = [&]() { /* publish the collected slots of finished rays */
        if (out_count == 0) return;
        /* ONE fence per batch, by every lane (each may have written hit records since the last flush), instead of one per
         * finished ray: a fence stalls the warp until its stores have landed, and in steady state some lane finishes a ray
         * in most iterations */
        __threadfence();
        __syncwarp();
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd((uint32_t *)&ctrl[PIPE_DONE_TAIL], (uint32_t)out_count);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < out_count) wp.done_q[base + (uint32_t)lane] = out_buf[wib][lane];
        pipe_publish(ctrl, PIPE_DONE_TAIL, out_count, lane);
        out_count = 0;
    };
    for (;;) {
        const unsigned holding = __ballot_sync(0xffffffffu, have);
        bool finished = false; /* this lane's ray ended in this iteration */
        /* ---- refill the lanes from the warp's fetched entries; fetch 32 more when those are used up ---- */
        if (trace_role && __popc(holding) < refill_below) {
            if (in_pos == in_count) {
                if (skip_polls > 0) skip_polls--;
                else {
                    int got;
                    const uint32_t base = pipe_claim(ctrl, PIPE_TRACE_TAIL, 32, lane, &got);
                    if (got) {
                        if (lane < got) {
                            const uint32_t index = base + (uint32_t)lane;
                            in_buf[wib][lane] = index < wp.n0 ? wp.first + index : pipe_wait(trace_q, index - wp.n0);
                        }
                        __syncwarp();
                        in_pos = 0; in_count = got;
                        idle_polls = 0; idle_sleep = 128; backoff = 1;
                    } else {
                        skip_polls = backoff;
                        backoff = min(backoff * 2, 64);
                    }
                }
            }
            if (in_pos < in_count) {
                const unsigned idle = ~holding;
                const int rank = __popc(idle & lt_mask);
                const int take = min(__popc(idle), in_count - in_pos);
                if (!have && rank < take) {
                    slot = in_buf[wib][in_pos + rank];
                    const V4 a0 = ldcg4(ws.s0 + slot), a1 = ldcg4(ws.s1 + slot);
                    ray.o = xyz(a0); ray.d = xyz(a1); ray.inv = rcp(ray.d);
                    cnt.ray_ref();
                    if (hit_begin<COUNT, WIDE>(A, ray, cnt) == HIT_RUNNING) have = true;
                    else { stcg4(ws.hit + slot, v4(i2f(-1), CLV_INF, 0.0f, 0.0f)); finished = true; }
                }
                in_pos += take;
            }
        }
        const unsigned tracing = __ballot_sync(0xffffffffu, have);
        if (tracing) {
            /* ---- one traversal step, kind chosen by majority vote ---- */
            const bool at_inner = have && ray.cur >= 0;
            const int n_inner = __popc(__ballot_sync(0xffffffffu, at_inner));
            const int n_leaf = __popc(tracing) - n_inner;
            int status = HIT_RUNNING;
            if (n_inner * schedule >= n_leaf * 4) {
                if (at_inner) status = WIDE ? hit_inner_step_wide<COUNT>(A, ray, st, cnt) : hit_inner_step<COUNT>(A, ray, st, cnt);
            } else if (have && !at_inner) status = hit_leaf_step<COUNT, WIDE>(A, ray, st, cnt);
            if (status == HIT_DONE) {
                have = false;
                stcg4(ws.hit + slot, v4(i2f(ray.best_i), ray.best_t, ray.bu, ray.bv));
                finished = true;
            }
        }
        /* ---- collect the rays that ended ---- */
        const unsigned fin = __ballot_sync(0xffffffffu, finished);
        if (fin) {
            if (out_count + __popc(fin) > 32) flush_done();
            if (finished) out_buf[wib][out_count + __popc(fin & lt_mask)] = slot;
            __syncwarp();
            out_count += __popc(fin);
            if (out_count >= 24) flush_done();
        }
        if (tracing) continue;
        /* ---- this warp holds no ray ---- */
        flush_done();
        if (in_pos < in_count) continue; /* fetched entries left: hand them out first */
        /* shade a batch of traced samples, if there is one */
        int got = 0;
        uint32_t dbase = 0;
        if (shade_every <= 0 || shade_role) dbase = pipe_claim(ctrl, PIPE_DONE_TAIL, 32, lane, &got);
        if (got) {
            idle_polls = 0; idle_sleep = 128;
            skip_polls = 0; backoff = 1; /* new rays are about to appear */
            bool more = false, mine = lane < got;
            uint32_t s = 0;
            if (mine) {
                s = pipe_wait(done_q, dbase + (uint32_t)lane);
                const uint32_t gid = sample_ids ? sample_ids[s] : s;
                uint2 sd = __ldcg(reinterpret_cast<const uint2 *>(seeds) + s);
                uint32_t seed0 = sd.x, seed1 = sd.y;
                more = wavefront_shade_sample<COUNT, SOA>(s, gid, materials, c, A, seed0, seed1, store, ws, accum, cnt) != WALK_OVER;
                if (seed0 != sd.x || seed1 != sd.y) reinterpret_cast<uint2 *>(seeds)[s] = make_uint2(seed0, seed1);
            }
            const unsigned staying = __ballot_sync(0xffffffffu, more);
            const unsigned ended = __ballot_sync(0xffffffffu, mine && !more);
            if (staying) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd((uint32_t *)&ctrl[PIPE_TRACE_TAIL], (uint32_t)__popc(staying));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (more) {
                    __threadfence(); /* walk state, seeds and vertex records before the entry that announces the next ray */
                    wp.trace_q[base + (uint32_t)__popc(staying & lt_mask) - wp.n0] = s;
                }
                pipe_publish(ctrl, PIPE_TRACE_TAIL, __popc(staying), lane);
            }
            if (ended && lane == 0) atomicSub((uint32_t *)&ctrl[PIPE_LIVE], (uint32_t)__popc(ended));
            continue;
        }
        /* ---- nothing to do right now: back off (thousands of idle warps polling one line starve the warps that work) ---- */
        {
            uint2 le = make_uint2(0u, 0u);
            if (lane == 0) le = __ldcv(reinterpret_cast<const uint2 *>(const_cast<const uint32_t *>(ctrl) + PIPE_LIVE));
            le.x = __shfl_sync(0xffffffffu, le.x, 0);
            le.y = __shfl_sync(0xffffffffu, le.y, 0);
            if (le.x == 0u || le.y != 0u) break;
        }
        idle_polls += idle_sleep >> 7;
        if (idle_polls > (1u << 24)) { /* watchdog: ~2 s asleep without work while samples are alive = a lost entry */
            if (lane == 0) ctrl[PIPE_ERROR] = 1u;
            break;
        }
        skip_polls = 0; /* idle: look at the trace queue on the next iteration */
        __nanosleep(idle_sleep);
        idle_sleep = min(idle_sleep * 2u, 16384u);
    }
    flush_counters<COUNT>(cnt, counters);
}

/* stages 6-8: finalize (identity sample bins) + light image + accumulation; clears the per-pass light image */
__global__ void __launch_bounds__(CLV_BLOCK)
k_fused_finalize(uint32_t n_pixels, const WeightAggregator *aggs, const Camera *camera, V4 *light_image, Accum *accum,
                 float *pass_weight_sum) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    float wpass = 0.0f;
    if (id < n_pixels) {
        Camera c = camera[0];
        wpass = fused_finalize_pixel(id, aggs, c, light_image, accum);
    }
    if (pass_weight_sum) {
        wpass = nan_to_num0(wpass);
        for (int o = 16; o > 0; o >>= 1) wpass += __shfl_xor_sync(0xffffffffu, wpass, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(pass_weight_sum, wpass);
    }
}

/* ---- device tone map: Renderer.image (renderer.py:293-300) + camera.tone_map (camera.py:73-82) ----------------------- *
 * estimate = nan_to_num(summed_image / summed_sample_weights) in float32; log-average luminance, scaling and the
 * 255 * s / (s + white^2) curve in float64 exactly as numpy promotes them (float32 array * python float stays float32,
 * dividing by the float64 log-average promotes).  Only the order of the float64 sum of logs differs from numpy's. */
__global__ void k_tone_reduce(uint32_t n_pixels, const Accum *accum, double *sum_log) {
    double acc = 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += gridDim.x * blockDim.x) {
        const Accum a = accum[i];
        float b = nan_to_num0(a.image[0] / a.weight), g = nan_to_num0(a.image[1] / a.weight), r = nan_to_num0(a.image[2] / a.weight);
        double lum = __dadd_rn(__dadd_rn(__dmul_rn((double)b, 0.0722), __dmul_rn((double)g, 0.7152)), __dmul_rn((double)r, 0.2126));
        acc += log(0.1 + lum);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(sum_log, acc);
}
__global__ void k_tone_map(uint32_t n_pixels, const Accum *accum, const double *sum_log, float exposure, double white2,
                           uint8_t *out_bgr) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    const double log_avg = exp(*sum_log / (double)n_pixels);
    const Accum a = accum[i];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        float e = nan_to_num0(a.image[ch] / a.weight) * exposure;
        double s = (double)e / log_avg;
        double m = 255.0 * s / (s + white2);
        out_bgr[3 * (size_t)i + ch] = (uint8_t)(int)m; /* astype(uint8): truncation */
    }
}

/* ---- parity hooks -------------------------------------------------------------------------------------------------- */
template <int MODE, bool COUNT>
__global__ void __launch_bounds__(CLV_BLOCK)
k_traverse(uint32_t n, const Ray *rays, AccelView A, int32_t *tri, float *t, float *u, float *v, Counters *counters) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    Cnt<COUNT> cnt;
    if (id < n) {
        const Ray *r = rays + id;
        int best_i = -1;
        float best_t = CLV_INF, bu = 0.0f, bv = 0.0f;
        cnt.ray_ref();
        closest_hit<MODE, COUNT>(A, xyz(ld4(&r->origin)), xyz(ld4(&r->direction)), xyz(ld4(&r->inv_direction)), best_i,
                                 best_t, bu, bv, cnt, -CLV_INF, -1);
        tri[id] = best_i; t[id] = best_t; u[id] = bu; v[id] = bv;
    }
    flush_counters<COUNT>(cnt, counters);
}

template <int MODE, bool COUNT>
__global__ void __launch_bounds__(CLV_BLOCK)
k_visibility(uint32_t n, const Ray *a, const Ray *b, AccelView A, int32_t *visible, Counters *counters) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    Cnt<COUNT> cnt;
    if (id < n)
        visible[id] = visibility_test<MODE, COUNT>(A, xyz(ld4(&a[id].origin)), a[id].triangle, xyz(ld4(&b[id].origin)),
                                                   b[id].triangle, cnt) ? 1 : 0;
    flush_counters<COUNT>(cnt, counters);
}

/* ---- host side of the connection stage --------------------------------------------------------------------------- */
static size_t connect_scratch_bytes(uint32_t n) {
    return 64 + (size_t)n * (36 * 4 + 2 * 4 + CLV_SENSOR_WORDS * 4);
}
static ConnectScratch connect_scratch_at(void *mem, uint32_t n) {
    ConnectScratch cs;
    char *p = (char *)mem;
    cs.task_count = (uint32_t *)p;   /* header: task_count[CLV_MAX_SLICES] then cursor[CLV_MAX_SLICES] */
    cs.cursor = cs.task_count + CLV_MAX_SLICES;
    p += 64;
    cs.tasks = (uint32_t *)p;
    p += (size_t)n * 36 * 4;
    cs.vis = (unsigned int *)p;
    p += (size_t)n * 2 * 4;
    cs.sensor = (float *)p;
    return cs;
}

/* ---------------------------------------------------------------------------------------------------------------
 * section 4: stage-level entry points (reference argument order)
 * ------------------------------------------------------------------------------------------------------------- */
#define GRID(n) (((n) + CLV_BLOCK - 1) / CLV_BLOCK)
/* dispatch on (traversal mode, counting) */
#define DISPATCH_MC(mode, counting, CALL)                                              \
    do {                                                                               \
        if ((mode) == CLV_TRAVERSE_REFERENCE_ORDER) {                                  \
            if (counting) { CALL(TRAVERSE_REFERENCE, true); } else { CALL(TRAVERSE_REFERENCE, false); } \
        } else {                                                                       \
            if (counting) { CALL(TRAVERSE_ORDERED, true); } else { CALL(TRAVERSE_ORDERED, false); }     \
        }                                                                              \
    } while (0)

/* ... and on the path layout */
#define DISPATCH_MCS(mode, counting, soa, CALL)                                      \
    do {                                                                             \
        if (soa) {                                                                   \
            if ((mode) == CLV_TRAVERSE_REFERENCE_ORDER) {                            \
                if (counting) { CALL(TRAVERSE_REFERENCE, true, true); } else { CALL(TRAVERSE_REFERENCE, false, true); } \
            } else {                                                                 \
                if (counting) { CALL(TRAVERSE_ORDERED, true, true); } else { CALL(TRAVERSE_ORDERED, false, true); }     \
            }                                                                        \
        } else {                                                                     \
            if ((mode) == CLV_TRAVERSE_REFERENCE_ORDER) {                            \
                if (counting) { CALL(TRAVERSE_REFERENCE, true, false); } else { CALL(TRAVERSE_REFERENCE, false, false); } \
            } else {                                                                 \
                if (counting) { CALL(TRAVERSE_ORDERED, true, false); } else { CALL(TRAVERSE_ORDERED, false, false); }     \
            }                                                                        \
        }                                                                            \
    } while (0)

static int g_stage_mode = CLV_TRAVERSE_ORDERED;
extern "C" int clv_set_stage_traverse_mode(int mode) {
    REQUIRE(mode == CLV_TRAVERSE_REFERENCE_ORDER || mode == CLV_TRAVERSE_ORDERED, "bad mode");
    g_stage_mode = mode;
    return CLV_OK;
}

struct clv_renderer;
static int mark_connect(clv_renderer *r, cudaStream_t st, int which); /* timing mark after enumerate / rays / finish */
/* enumerate -> rays -> finish for the samples [first, n) on `stream`; `cs` holds this slice's own task list */
static int launch_connect(clv_ctx *c, cudaStream_t stream, int mode, bool splat, bool large, uint32_t first, uint32_t n,
                          const uint32_t *sample_ids, const PathStore &store,
                          bool soa, const AccelView &view, const Material *mats, const Camera *cam,
                          ConnectScratch cs, WeightAggregator *aggs, V4 *out, const LightSink &sink,
                          struct clv_renderer *timed /* NULL, or the renderer whose timing marks follow each kernel */) {
    Counters *ctr = c->d_counters + COUNTERS_CONNECT;
    const uint32_t len = n - first;
#define CALL(M, CT, SO) k_connect_enumerate<M, CT, SO><<<GRID(len), CLV_BLOCK, 0, stream>>>(first, n, store, view, mats, cam, cs, ctr)
    DISPATCH_MCS(mode, c->counting, soa, CALL);
#undef CALL
    LAUNCHED(c);
    if (timed && mark_connect(timed, stream, 0) != CLV_OK) return CLV_ERR_CUDA;
    const uint64_t warps = ((uint64_t)len * 36 + CLV_RAY_CHUNK - 1) / CLV_RAY_CHUNK;
    unsigned ray_blocks = (unsigned)((warps + CLV_WARPS_PER_BLOCK - 1) / CLV_WARPS_PER_BLOCK);
    const bool large_flavour = large && mode == CLV_TRAVERSE_ORDERED && !c->counting && soa;
    const int dyn = c->connect_dynamic_fetch && (uint64_t)len * 36 < 0x7fffffffull; /* task positions are ints in the kernel */
    if (dyn) {
        /* persistent grid: the SMs filled once (resident blocks of the flavour in use), never more warps than the worst-case
         * number of 32-task groups */
        int &per_sm = c->rays_blocks_per_sm[large_flavour ? 1 : 0];
        if (!per_sm) {
            if (large_flavour) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_connect_rays<TRAVERSE_ORDERED, false, true, 1>, CLV_BLOCK, 0));
            else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_connect_rays<TRAVERSE_ORDERED, false, true, 0>, CLV_BLOCK, 0));
            if (per_sm < 1) per_sm = 1;
        }
        const uint64_t need = ((uint64_t)len * 36 + CLV_BLOCK - 1) / CLV_BLOCK;
        ray_blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)c->sm_count * (uint64_t)per_sm, need));
    }
    if (large_flavour) {
        /* the renderer's default configuration on a scene larger than the caches: high-occupancy flavour */
        k_connect_rays<TRAVERSE_ORDERED, false, true, 1><<<ray_blocks, CLV_BLOCK, 0, stream>>>(store, view, cam, cs, ctr, c->refill_below, c->ray_schedule_large, dyn);
    } else {
#define CALL(M, CT, SO) k_connect_rays<M, CT, SO><<<ray_blocks, CLV_BLOCK, 0, stream>>>(store, view, cam, cs, ctr, c->refill_below, c->ray_schedule, dyn)
        DISPATCH_MCS(mode, c->counting, soa, CALL);
#undef CALL
    }
    LAUNCHED(c);
    if (timed && mark_connect(timed, stream, 1) != CLV_OK) return CLV_ERR_CUDA;
#define FIN(SP, CT, SO) do { if (c->finish_cooperative) k_connect_finish_coop<SP, CT, SO><<<GRID(len), CLV_BLOCK, 0, stream>>>(first, n, sample_ids, store, mats, cam, cs, aggs, out, sink, ctr); \
                            else k_connect_finish<SP, CT, SO><<<GRID(len), CLV_BLOCK, 0, stream>>>(first, n, sample_ids, store, mats, cam, cs, aggs, out, sink, ctr); } while (0)
    if (soa) {
        if (splat) { if (c->counting) FIN(true, true, true); else FIN(true, false, true); }
        else { if (c->counting) FIN(false, true, true); else FIN(false, false, true); }
    } else {
        if (splat) { if (c->counting) FIN(true, true, false); else FIN(true, false, false); }
        else { if (c->counting) FIN(false, true, false); else FIN(false, false, false); }
    }
#undef FIN
    LAUNCHED(c);
    if (timed && mark_connect(timed, stream, 2) != CLV_OK) return CLV_ERR_CUDA;
    return CLV_OK;
}

extern "C" int clv_generate_light_rays(clv_ctx *c, uint32_t n, const clv_Triangle *lt, const float *areas,
                                       const int32_t *lti, const clv_Material *mats, uint32_t *rnd, clv_Ray *out,
                                       const int32_t *counts) {
    REQUIRE(c && lt && areas && lti && mats && rnd && out && counts, "null");
    if (!n) return CLV_OK;
    CK(cudaSetDevice(c->device));
    k_generate_light_rays<<<GRID(n), CLV_BLOCK, 0, c->stream>>>(n, (const Triangle *)lt, areas, lti, (const Material *)mats, rnd, (Ray *)out, counts);
    LAUNCHED(c);
    return CLV_OK;
}

extern "C" int clv_generate_camera_rays(clv_ctx *c, uint32_t n, const clv_Camera *cam, uint32_t *rnd,
                                        const uint32_t *indices, clv_Ray *out) {
    REQUIRE(c && cam && rnd && indices && out, "null");
    if (!n) return CLV_OK;
    CK(cudaSetDevice(c->device));
    k_generate_camera_rays<<<GRID(n), CLV_BLOCK, 0, c->stream>>>(n, (const Camera *)cam, rnd, indices, (Ray *)out);
    LAUNCHED(c);
    return CLV_OK;
}

extern "C" int clv_generate_paths(clv_ctx *c, uint32_t n, const clv_Ray *rays, const clv_accel *accel,
                                  const clv_Material *mats, uint32_t *rnd, float *out, clv_Path *paths, float *debug) {
    REQUIRE(c && rays && accel && mats && rnd && out && paths, "null");
    if (!n) return CLV_OK;
    CK(cudaSetDevice(c->device));
#define CALL(M, CT) k_generate_paths<M, CT><<<GRID(n), CLV_BLOCK, 0, c->stream>>>(n, (const Ray *)rays, accel->view, (const Material *)mats, rnd, (V4 *)out, (Path *)paths, (V4 *)debug, c->d_counters)
    DISPATCH_MC(g_stage_mode, c->counting, CALL);
#undef CALL
    LAUNCHED(c);
    return CLV_OK;
}

extern "C" int clv_reset_light_indices(clv_ctx *c, uint32_t n, int32_t *pix, int32_t *path, int32_t *ray, float *w,
                                       float *shade) {
    REQUIRE(c && pix && path && ray && w && shade, "null");
    if (!n) return CLV_OK;
    CK(cudaSetDevice(c->device));
    k_reset_light_indices<<<(n + 255) / 256, 256, 0, c->stream>>>(n, pix, path, ray, w, shade);
    LAUNCHED(c);
    return CLV_OK;
}

extern "C" int clv_connect_paths(clv_ctx *c, uint32_t n, const clv_Path *cp, const clv_Path *lp, const clv_accel *accel,
                                 const clv_Material *mats, const clv_Camera *cam, clv_WeightAggregator *aggs, float *out,
                                 int32_t *pix, int32_t *path, int32_t *ray, float *w, float *shade) {
    REQUIRE(c && cp && lp && accel && mats && cam && aggs && out && pix && path && ray && w && shade, "null");
    if (!n) return CLV_OK;
    REQUIRE(n < (1u << 26), "more than 2^26 samples per dispatch (connection tasks are sample << 6 | strategy)");
    CK(cudaSetDevice(c->device));
    LightSink sink;
    memset(&sink, 0, sizeof sink);
    sink.pixel_indices = pix; sink.path_indices = path; sink.ray_indices = ray; sink.weights = w; sink.shade = shade;
    if (c->cs_capacity < n) {
        CK(cudaStreamSynchronize(c->stream));
        if (c->cs_mem) CK(dev_free(c, c->cs_mem));
        c->cs_mem = nullptr;
        c->cs_capacity = 0;
        CK(dev_alloc(c, &c->cs_mem, connect_scratch_bytes(n)));
        c->cs_capacity = n;
    }
    ConnectScratch cs = connect_scratch_at(c->cs_mem, n);
    CK(cudaMemsetAsync(cs.task_count, 0, 64, c->stream));
    PathStore store;
    memset(&store, 0, sizeof store);
    store.camera_paths = (Path *)cp; /* read only here */
    store.light_paths = (Path *)lp;
    store.n = n;
    return launch_connect(c, c->stream, g_stage_mode, false, false, 0u, n, nullptr, store, false, accel->view,
                          (const Material *)mats, (const Camera *)cam, cs, (WeightAggregator *)aggs, (V4 *)out, sink, nullptr);
}

extern "C" int clv_adaptive_finalize_samples(clv_ctx *c, uint32_t n, const clv_WeightAggregator *aggs,
                                             const clv_Camera *cam, float *out, uint32_t *counts,
                                             const uint32_t *bin_offsets, float *weights) {
    REQUIRE(c && aggs && cam && out && counts && bin_offsets && weights, "null");
    if (!n) return CLV_OK;
    CK(cudaSetDevice(c->device));
    k_adaptive_finalize_samples<<<GRID(n), CLV_BLOCK, 0, c->stream>>>(n, (const WeightAggregator *)aggs, (const Camera *)cam, (V4 *)out, counts, bin_offsets, weights);
    LAUNCHED(c);
    return CLV_OK;
}

extern "C" int clv_light_sort(clv_ctx *c, uint32_t n_threads, int32_t *pix, int32_t *path, int32_t *ray, float *w,
                              float *shade, uint32_t stage, uint32_t pass_of_stage, uint32_t n, uint32_t ppt) {
    REQUIRE(c && pix && path && ray && w && shade, "null");
    REQUIRE(stage >= 1 && stage < 32 && pass_of_stage >= 1 && pass_of_stage <= stage, "bad stage");
    if (!n_threads) return CLV_OK;
    CK(cudaSetDevice(c->device));
    k_light_sort<<<(n_threads + 255) / 256, 256, 0, c->stream>>>(n_threads, pix, path, ray, w, shade, stage, pass_of_stage, n, ppt);
    LAUNCHED(c);
    return CLV_OK;
}

extern "C" int clv_light_image_gather(clv_ctx *c, uint32_t n, const clv_Path *lp, const clv_Material *mats,
                                      const int32_t *path_idx, const int32_t *ray_idx, const int32_t *bins,
                                      uint32_t offset, const float *weights, const float *shades, float *light_image,
                                      float *sum_weights) {
    REQUIRE(c && lp && mats && path_idx && ray_idx && bins && weights && shades && light_image && sum_weights, "null");
    if (!n) return CLV_OK;
    CK(cudaSetDevice(c->device));
    k_light_image_gather<<<GRID(n), CLV_BLOCK, 0, c->stream>>>(n, (const Path *)lp, (const Material *)mats, path_idx, ray_idx, bins, offset, weights, shades, (V4 *)light_image, sum_weights);
    LAUNCHED(c);
    return CLV_OK;
}

extern "C" int clv_light_image_splat(clv_ctx *c, uint32_t n_entries, uint32_t n_pixels, const clv_Path *lp,
                                     const clv_Material *mats, const int32_t *pix, const int32_t *path_idx,
                                     const int32_t *ray_idx, const float *weights, const float *shades,
                                     float *light_image, float *sum_weights) {
    REQUIRE(c && lp && mats && pix && path_idx && ray_idx && weights && shades && light_image && sum_weights, "null");
    CK(cudaSetDevice(c->device));
    if (n_pixels) {
        k_set_light_image<<<(n_pixels + 255) / 256, 256, 0, c->stream>>>(n_pixels, (V4 *)light_image);
        LAUNCHED(c);
    }
    if (n_entries) {
        k_light_image_splat<<<(n_entries + 255) / 256, 256, 0, c->stream>>>(n_entries, n_pixels, (const Path *)lp, (const Material *)mats, pix, path_idx, ray_idx, weights, shades, (V4 *)light_image, sum_weights);
        LAUNCHED(c);
    }
    return CLV_OK;
}

/* ---------------------------------------------------------------------------------------------------------------
 * section 5: fused renderer
 * ------------------------------------------------------------------------------------------------------------- */
/* kernels of a fused pass as the timing interface names them (clv_renderer_kernel_times) */
enum { TK_FUSED_WALKS = 0, TK_ENUMERATE, TK_RAYS, TK_FINISH, TK_FINALIZE, TK_WALK_TRACE, TK_WALK_SHADE, TK_WALK_PIPELINE, CLV_TIMED_KERNELS };
#define CLV_PASS_KERNELS CLV_TIMED_KERNELS
struct clv_renderer {
    clv_ctx *ctx;
    const clv_accel *accel;
    uint32_t width, height, n_pixels, n_local;
    int light_count;
    int mode;
    /* scene copies */
    Material *d_materials;
    Camera *d_camera;
    Triangle *d_light_triangles;
    float *d_light_areas;
    int32_t *d_light_tri_idx;
    /* work buffers */
    uint32_t *d_seeds, *d_sample_ids;
    Path *d_light_paths, *d_camera_paths; /* reference-layout path records: only while path_records is on */
    V4 *d_soa;                            /* compact SoA vertex records (the default) */
    int32_t *d_soa_len;
    int path_records;
    WeightAggregator *d_aggs;
    V4 *d_light_image;
    Accum *d_accum;
    float *d_pass_weight;
    void *d_connect_scratch;
    double *d_tone_sum;
    uint8_t *d_tone_image;
    std::vector<uint32_t> sample_ids;
    unsigned long long passes;
    /* optional per-kernel timing: events around each of the five kernels of every pass */
    int timing;
    /* timing marks: an event after every launch of a timed pass, tagged with the kernel it follows (-1 = start of a
     * pass); the time between two consecutive marks is the later one's kernel (one slice, one stream: launches serialise) */
    std::vector<cudaEvent_t> events;
    std::vector<int> event_kernel;
    double kernel_ms[CLV_PASS_KERNELS];
    /* wavefront walks (k_walk_*) */
    V4 *d_walk_state;      /* 5 parts (s0..s3, hit) x n_local */
    float *d_walk_jitter;  /* [n_local][2] */
    uint32_t *d_walk_queue; /* 2 x n_local */
    uint32_t *d_walk_counters; /* CLV_MAX_SLICES x CLV_WALK_COUNTER_WORDS */
    uint8_t *d_walk_lag;       /* [n_local] rounds a sample has waited for a suspended search */
    V4 *d_walk_records;        /* continuation records: 2 halves x walk_record_capacity x CLV_HIT_RECORD_V4 */
    size_t walk_record_capacity;
    uint32_t *d_pipe_queues;   /* pipelined walks: trace entries (11 per sample) then done entries (12 per sample) */
    uint32_t *d_pipe_ctrl;     /* CLV_MAX_SLICES x 8 */
    int pipe_blocks_per_sm;
    unsigned long long timed_passes;
};

extern "C" int clv_renderer_destroy(clv_renderer *r) {
    if (!r) return CLV_OK;
    cudaSetDevice(r->ctx->device);
    cudaStreamSynchronize(r->ctx->stream);
    dev_free(r->ctx, r->d_materials); dev_free(r->ctx, r->d_camera); dev_free(r->ctx, r->d_light_triangles); dev_free(r->ctx, r->d_light_areas);
    dev_free(r->ctx, r->d_light_tri_idx); dev_free(r->ctx, r->d_seeds); dev_free(r->ctx, r->d_sample_ids); dev_free(r->ctx, r->d_light_paths);
    dev_free(r->ctx, r->d_camera_paths); dev_free(r->ctx, r->d_soa); dev_free(r->ctx, r->d_soa_len); dev_free(r->ctx, r->d_aggs); dev_free(r->ctx, r->d_light_image); dev_free(r->ctx, r->d_accum);
    dev_free(r->ctx, r->d_pass_weight);
    dev_free(r->ctx, r->d_connect_scratch);
    dev_free(r->ctx, r->d_tone_sum);
    dev_free(r->ctx, r->d_tone_image);
    dev_free(r->ctx, r->d_walk_state); dev_free(r->ctx, r->d_walk_jitter); dev_free(r->ctx, r->d_walk_queue);
    dev_free(r->ctx, r->d_walk_counters); dev_free(r->ctx, r->d_walk_lag); dev_free(r->ctx, r->d_walk_records);
    dev_free(r->ctx, r->d_pipe_queues); dev_free(r->ctx, r->d_pipe_ctrl);
    for (cudaEvent_t e : r->events) cudaEventDestroy(e);
    clv_accel_destroy(const_cast<clv_accel *>(r->accel)); /* drops the renderer's reference */
    delete r;
    return CLV_OK;
}

static int upload_new(clv_ctx *c, void **dst, const void *src, size_t nbytes) {
    CK(dev_alloc(c, dst, nbytes ? nbytes : 16));
    if (nbytes) CK(cudaMemcpyAsync(*dst, src, nbytes, cudaMemcpyHostToDevice, c->stream));
    return CLV_OK;
}

static int load_seeds(clv_renderer *r, const uint32_t *h_seeds) {
    clv_ctx *c = r->ctx;
    if (r->sample_ids.empty()) {
        CK(cudaMemcpyAsync(r->d_seeds, h_seeds, (size_t)r->n_local * 8, cudaMemcpyHostToDevice, c->stream));
    } else {
        std::vector<uint32_t> rows((size_t)r->n_local * 2);
        for (uint32_t i = 0; i < r->n_local; i++) {
            rows[2 * (size_t)i] = h_seeds[2 * (size_t)r->sample_ids[i]];
            rows[2 * (size_t)i + 1] = h_seeds[2 * (size_t)r->sample_ids[i] + 1];
        }
        CK(cudaMemcpyAsync(r->d_seeds, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return CLV_OK;
}

extern "C" int clv_renderer_create(clv_ctx *c, const clv_accel *accel, const clv_scene_desc *s, const uint32_t *h_seeds,
                                   const uint32_t *h_sample_ids, uint32_t n_local, clv_renderer **out) {
    REQUIRE(c && accel && s && h_seeds && out, "null");
    REQUIRE(s->materials && s->n_materials >= 8 && s->camera && s->light_triangles && s->light_surface_areas &&
                s->light_triangle_indices && s->n_lights > 0, "incomplete scene description (needs >= 8 materials, >= 1 light)");
    *out = nullptr;
    CK(cudaSetDevice(c->device));
    const Camera *cam = reinterpret_cast<const Camera *>(s->camera);
    REQUIRE(cam->pixel_width > 0 && cam->pixel_height > 0, "bad camera resolution");
    uint32_t n_pixels = (uint32_t)cam->pixel_width * (uint32_t)cam->pixel_height;
    if (!h_sample_ids) n_local = n_pixels;
    REQUIRE(n_local > 0 && n_local <= n_pixels, "bad n_local");
    REQUIRE(n_local < (1u << 26), "more than 2^26 samples per pass (connection tasks are sample << 6 | strategy)");
    if (h_sample_ids)
        for (uint32_t i = 0; i < n_local; i++) REQUIRE(h_sample_ids[i] < n_pixels, "sample id out of range");

    clv_renderer *r = new (std::nothrow) clv_renderer();
    if (!r) return fail(CLV_ERR_NOMEM, "host allocation");
    r->ctx = c; r->accel = accel;
    const_cast<clv_accel *>(accel)->refs++;
    r->width = cam->pixel_width; r->height = cam->pixel_height; r->n_pixels = n_pixels; r->n_local = n_local;
    r->light_count = s->n_lights;
    r->mode = CLV_TRAVERSE_ORDERED;
    r->passes = 0;
    r->timing = 0;
    for (int k = 0; k < CLV_PASS_KERNELS; k++) r->kernel_ms[k] = 0.0;
    r->timed_passes = 0;
    if (h_sample_ids) r->sample_ids.assign(h_sample_ids, h_sample_ids + n_local);
    int rc;
#define TRY(x) do { rc = (x); if (rc != CLV_OK) { clv_renderer_destroy(r); return rc; } } while (0)
    TRY(upload_new(c, (void **)&r->d_materials, s->materials, (size_t)s->n_materials * sizeof(Material)));
    TRY(upload_new(c, (void **)&r->d_camera, s->camera, sizeof(Camera)));
    TRY(upload_new(c, (void **)&r->d_light_triangles, s->light_triangles, (size_t)s->n_lights * sizeof(Triangle)));
    TRY(upload_new(c, (void **)&r->d_light_areas, s->light_surface_areas, (size_t)s->n_lights * 4));
    TRY(upload_new(c, (void **)&r->d_light_tri_idx, s->light_triangle_indices, (size_t)s->n_lights * 4));
    TRY(upload_new(c, (void **)&r->d_sample_ids, h_sample_ids, h_sample_ids ? (size_t)n_local * 4 : 0));
    auto zalloc = [&](void **p, size_t nbytes) -> int {
        CK(dev_alloc(c, p, nbytes));
        CK(cudaMemsetAsync(*p, 0, nbytes, c->stream));
        return CLV_OK;
    };
    TRY(zalloc((void **)&r->d_seeds, (size_t)n_local * 8));
    r->d_light_paths = r->d_camera_paths = nullptr;
    r->path_records = 0;
    TRY(zalloc((void **)&r->d_soa, (size_t)n_local * 2 * 6 * 4 * sizeof(V4)));
    TRY(zalloc((void **)&r->d_soa_len, (size_t)n_local * 2 * sizeof(int32_t)));
    TRY(zalloc((void **)&r->d_aggs, (size_t)n_pixels * sizeof(WeightAggregator)));
    TRY(zalloc((void **)&r->d_light_image, (size_t)n_pixels * sizeof(V4)));
    TRY(zalloc((void **)&r->d_accum, (size_t)n_pixels * sizeof(Accum)));
    TRY(zalloc((void **)&r->d_pass_weight, 16));
    TRY(zalloc((void **)&r->d_connect_scratch, connect_scratch_bytes(n_local)));
    TRY(zalloc((void **)&r->d_tone_sum, 16));
    TRY(zalloc((void **)&r->d_tone_image, (size_t)n_pixels * 3));
    r->d_walk_state = nullptr; r->d_walk_jitter = nullptr; r->d_walk_queue = nullptr; r->d_walk_counters = nullptr;
    r->d_walk_lag = nullptr; r->d_walk_records = nullptr; r->walk_record_capacity = 0;
    r->d_pipe_queues = nullptr; r->d_pipe_ctrl = nullptr; r->pipe_blocks_per_sm = 0;
    TRY(load_seeds(r, h_seeds));
#undef TRY
    CK(cudaStreamSynchronize(c->stream));
    *out = r;
    return CLV_OK;
}

/* keep the reference-layout Path records of every pass (1040 B per subpath and sample) instead of the compact SoA
 * vertex records: for differential tests and tools that read CLV_BUF_LIGHT_PATHS / CLV_BUF_CAMERA_PATHS */
extern "C" int clv_renderer_set_path_records(clv_renderer *r, int enabled) {
    REQUIRE(r, "null");
    clv_ctx *c = r->ctx;
    CK(cudaSetDevice(c->device));
    if (enabled && !r->d_light_paths) {
        Path *lp = nullptr, *cp = nullptr;
        CK(dev_alloc(c, &lp, (size_t)r->n_local * sizeof(Path)));
        cudaError_t e2 = dev_alloc(c, &cp, (size_t)r->n_local * sizeof(Path));
        if (e2 != cudaSuccess) { /* all or nothing: a renderer with one of the two records would read a null pointer */
            dev_free(c, lp);
            return fail(e2 == cudaErrorMemoryAllocation ? CLV_ERR_NOMEM : CLV_ERR_CUDA, "path records: %s", cudaGetErrorString(e2));
        }
        r->d_light_paths = lp;
        r->d_camera_paths = cp;
        CK(cudaMemsetAsync(r->d_light_paths, 0, (size_t)r->n_local * sizeof(Path), c->stream));
        CK(cudaMemsetAsync(r->d_camera_paths, 0, (size_t)r->n_local * sizeof(Path), c->stream));
    }
    r->path_records = enabled ? 1 : 0;
    return CLV_OK;
}

extern "C" int clv_renderer_set_traverse_mode(clv_renderer *r, int mode) {
    REQUIRE(r, "null");
    REQUIRE(mode == CLV_TRAVERSE_REFERENCE_ORDER || mode == CLV_TRAVERSE_ORDERED, "bad mode");
    r->mode = mode;
    return CLV_OK;
}

static int mark(clv_renderer *r, cudaStream_t st, int kernel);
static int mark_connect(clv_renderer *r, cudaStream_t st, int which) { return mark(r, st, TK_ENUMERATE + which); }
static int mark(clv_renderer *r, cudaStream_t st, int kernel) {
    if (!r->timing) return CLV_OK;
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    CK(cudaEventRecord(e, st));
    r->events.push_back(e);
    r->event_kernel.push_back(kernel);
    return CLV_OK;
}

/* buffers of the wavefront walks, allocated the first time a pass runs in that mode */
static int ensure_walk_buffers(clv_renderer *r) {
    if (r->d_walk_records) return CLV_OK; /* the last one allocated: all or nothing */
    clv_ctx *c = r->ctx;
    const size_t n = r->n_local;
    /* continuation records: a slice [first, end) owns records [first / 32 + slice * 1024, ...) of each half, 1024 +
     * (end - first) / 32 of them (about 1 % of a round's rays are suspended at the default budget; a full pool is harmless) */
    r->walk_record_capacity = n / CLV_WALK_SAMPLES_PER_RECORD + (size_t)(CLV_MAX_SLICES + 1) * CLV_WALK_RECORDS_MIN;
    cudaError_t e = dev_alloc(c, &r->d_walk_state, n * 5 * sizeof(V4));
    if (e == cudaSuccess) e = dev_alloc(c, &r->d_walk_jitter, n * 2 * sizeof(float));
    if (e == cudaSuccess) e = dev_alloc(c, &r->d_walk_queue, n * 2 * sizeof(uint32_t));
    if (e == cudaSuccess) e = dev_alloc(c, &r->d_walk_counters, (size_t)CLV_MAX_SLICES * CLV_WALK_COUNTER_WORDS * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemsetAsync(r->d_walk_counters, 0, (size_t)CLV_MAX_SLICES * CLV_WALK_COUNTER_WORDS * sizeof(uint32_t), c->stream);
    if (e == cudaSuccess) e = dev_alloc(c, &r->d_walk_lag, n);
    if (e == cudaSuccess) e = dev_alloc(c, &r->d_walk_records, 2 * r->walk_record_capacity * CLV_HIT_RECORD_V4 * sizeof(V4));
    if (e != cudaSuccess) { /* a renderer with some of them would read null pointers on the next pass */
        dev_free(c, r->d_walk_state); dev_free(c, r->d_walk_jitter); dev_free(c, r->d_walk_queue); dev_free(c, r->d_walk_counters);
        dev_free(c, r->d_walk_lag); dev_free(c, r->d_walk_records);
        r->d_walk_state = nullptr; r->d_walk_jitter = nullptr; r->d_walk_queue = nullptr; r->d_walk_counters = nullptr;
        r->d_walk_lag = nullptr; r->d_walk_records = nullptr;
        cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? CLV_ERR_NOMEM : CLV_ERR_CUDA, "wavefront walk buffers: %s", cudaGetErrorString(e));
    }
    return CLV_OK;
}
static int ensure_pipe_buffers(clv_renderer *r) {
    if (r->d_pipe_ctrl) return CLV_OK; /* the last one allocated: all or nothing */
    clv_ctx *c = r->ctx;
    cudaError_t e = dev_alloc(c, &r->d_pipe_queues, (size_t)r->n_local * 23 * sizeof(uint32_t));
    if (e == cudaSuccess) e = dev_alloc(c, &r->d_pipe_ctrl, (size_t)CLV_MAX_SLICES * 8 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemsetAsync(r->d_pipe_ctrl, 0, (size_t)CLV_MAX_SLICES * 8 * sizeof(uint32_t), c->stream);
    if (e != cudaSuccess) {
        dev_free(c, r->d_pipe_queues); dev_free(c, r->d_pipe_ctrl);
        r->d_pipe_queues = nullptr; r->d_pipe_ctrl = nullptr;
        cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? CLV_ERR_NOMEM : CLV_ERR_CUDA, "pipelined walk buffers: %s", cudaGetErrorString(e));
    }
    return CLV_OK;
}

/* persistent grid of k_walk_trace: every SM filled with the resident blocks of the flavour, never more warps than rays */
template <typename K>
static int trace_grid(clv_ctx *c, K kernel, int flavour, uint32_t n_rays, bool sliced, unsigned *grid_out) {
    if (!c->trace_blocks_per_sm[flavour]) {
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, CLV_BLOCK, 0));
        c->trace_blocks_per_sm[flavour] = per_sm > 0 ? per_sm : 1;
    }
    /* a pass that runs as one slice (per-kernel timing) has nobody to share the SMs with: full grid */
    const uint64_t divisor = sliced ? (uint64_t)std::max(1, c->walk_grid_divisor) : 1u;
    const uint64_t resident = std::max<uint64_t>(1, (uint64_t)c->sm_count * (uint64_t)c->trace_blocks_per_sm[flavour] / divisor);
    const uint64_t needed = ((uint64_t)n_rays + CLV_BLOCK - 1) / CLV_BLOCK;
    *grid_out = (unsigned)std::max<uint64_t>(1, std::min(resident, needed));
    return CLV_OK;
}

/* stages 1-4 of the samples [first, end) as wavefront rounds on `st` */
static int launch_wavefront_walks(clv_renderer *r, cudaStream_t st, int slice, uint32_t first, uint32_t end, const uint32_t *ids,
                                  const PathStore &store, bool soa, bool /*large: the trace flavour follows the scene*/, bool sliced) {
    clv_ctx *c = r->ctx;
    const uint32_t len = end - first;
    const size_t n = r->n_local;
    WalkState ws;
    ws.s0 = r->d_walk_state; ws.s1 = ws.s0 + n; ws.s2 = ws.s1 + n; ws.s3 = ws.s2 + n; ws.hit = ws.s3 + n;
    ws.jitter = r->d_walk_jitter;
    ws.lag = r->d_walk_lag;
    WalkQueues wq;
    wq.q[0] = r->d_walk_queue + first;      /* a slice's lists never hold more than its own samples */
    wq.q[1] = r->d_walk_queue + n + first;
    wq.count = r->d_walk_counters + (size_t)slice * CLV_WALK_COUNTER_WORDS;
    wq.cursor = wq.count + (CLV_WALK_LAUNCH_ROUNDS + 1);
    wq.pool = wq.cursor + CLV_WALK_LAUNCH_ROUNDS;
    WalkRecords wr;
    {
        const size_t rec0 = (size_t)first / CLV_WALK_SAMPLES_PER_RECORD + (size_t)slice * CLV_WALK_RECORDS_MIN;
        wr.capacity = (uint32_t)(len / CLV_WALK_SAMPLES_PER_RECORD + CLV_WALK_RECORDS_MIN);
        wr.half[0] = r->d_walk_records + rec0 * CLV_HIT_RECORD_V4;
        wr.half[1] = wr.half[0] + r->walk_record_capacity * CLV_HIT_RECORD_V4;
        wr.tail_budget = c->walk_tail_budget;
    }
    Counters *ctr = c->d_counters + COUNTERS_WALK;
    const bool pipelined = c->walk_mode == 3 || (c->walk_mode == 0 && CLV_PIPELINE_BY_DEFAULT);
    uint32_t *pipe_ctrl = nullptr;
    WalkPipe wp;
    memset(&wp, 0, sizeof wp);
    if (pipelined) {
        pipe_ctrl = r->d_pipe_ctrl + (size_t)slice * 8;
        wp.ctrl = pipe_ctrl;
        wp.trace_q = r->d_pipe_queues + (size_t)first * 11;
        wp.done_q = r->d_pipe_queues + n * 11 + (size_t)first * 12;
        wp.n0 = len;
        wp.first = first;
        /* every entry starts as EMPTY; a slot word written into it makes it valid */
        CK(cudaMemsetAsync(wp.trace_q, 0xff, (size_t)len * 11 * sizeof(uint32_t), st));
        CK(cudaMemsetAsync(wp.done_q, 0xff, (size_t)len * 12 * sizeof(uint32_t), st));
    }
    if (soa) k_walk_init<true><<<GRID(len), CLV_BLOCK, 0, st>>>(first, end, r->d_light_triangles, r->d_light_areas, r->d_light_tri_idx, r->light_count, r->d_materials, r->d_seeds, store, ws, wq, pipe_ctrl);
    else k_walk_init<false><<<GRID(len), CLV_BLOCK, 0, st>>>(first, end, r->d_light_triangles, r->d_light_areas, r->d_light_tri_idx, r->light_count, r->d_materials, r->d_seeds, store, ws, wq, pipe_ctrl);
    LAUNCHED(c);
    int rc = mark(r, st, TK_WALK_SHADE);
    if (rc != CLV_OK) return rc;
    if (pipelined) {
        const bool pl = r->accel->large != 0;
        /* persistent grid: every SM filled once with the resident blocks of the kernel (halved under slices, like the trace
         * grid), never more warps than samples */
#define PIPE(CT, SO, LG) k_walk_pipeline<CT, SO, LG>
#define PIPE_GO(CT, SO, LG) do { \
            if (!r->pipe_blocks_per_sm) { int per_sm = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, PIPE(CT, SO, LG), CLV_BLOCK, 0)); r->pipe_blocks_per_sm = per_sm > 0 ? per_sm : 1; } \
            const uint64_t resident = std::max<uint64_t>(1, (uint64_t)c->sm_count * (uint64_t)r->pipe_blocks_per_sm / (uint64_t)(sliced ? std::max(1, c->walk_grid_divisor) : 1)); \
            const unsigned pgrid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(resident, ((uint64_t)len + CLV_BLOCK - 1) / CLV_BLOCK)); \
            PIPE(CT, SO, LG)<<<pgrid, CLV_BLOCK, 0, st>>>(wp, ids, r->d_materials, r->d_camera, r->accel->view, r->d_seeds, store, ws, r->d_accum, ctr, c->walk_refill_below, c->walk_schedule, c->pipe_shade_every); } while (0)
        if (c->counting) { if (soa) { if (pl) PIPE_GO(true, true, 1); else PIPE_GO(true, true, 0); } else { if (pl) PIPE_GO(true, false, 1); else PIPE_GO(true, false, 0); } }
        else { if (soa) { if (pl) PIPE_GO(false, true, 1); else PIPE_GO(false, true, 0); } else { if (pl) PIPE_GO(false, false, 1); else PIPE_GO(false, false, 0); } }
#undef PIPE_GO
#undef PIPE
        LAUNCHED(c);
        return mark(r, st, TK_WALK_PIPELINE);
    }
    /* the trace kernel's flavour follows the scene (not the path layout or the counters): an instrumented run counts
     * the records of the very search the timed run performs */
    const bool tl = r->accel->large != 0;
    unsigned tgrid = 1;
    if (c->counting) rc = tl ? trace_grid(c, k_walk_trace<true, 1>, 1, len, sliced, &tgrid) : trace_grid(c, k_walk_trace<true, 0>, 0, len, sliced, &tgrid);
    else rc = tl ? trace_grid(c, k_walk_trace<false, 1>, 1, len, sliced, &tgrid) : trace_grid(c, k_walk_trace<false, 0>, 0, len, sliced, &tgrid);
    if (rc != CLV_OK) return rc;
    const unsigned sgrid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)c->sm_count * 16u, ((uint64_t)len + CLV_BLOCK - 1) / CLV_BLOCK));
    /* a sample whose searches were suspended k times needs k more rounds */
    const int rounds = wr.tail_budget > 0 ? CLV_WALK_LAUNCH_ROUNDS : CLV_WALK_ROUNDS;
    for (int round = 0; round < rounds; round++) {
        const uint32_t *qin = round == 0 ? nullptr : wq.q[round & 1];
        uint32_t *qout = wq.q[(round + 1) & 1];
        const uint32_t n_identity = round == 0 ? len : 0u;
        const V4 *rec_in = wr.half[round & 1];
        const uint32_t *rec_in_count = round > 0 && wr.tail_budget > 0 ? wq.pool + round - 1 : nullptr;
#define TRACE(CT, LG) k_walk_trace<CT, LG><<<tgrid, CLV_BLOCK, 0, st>>>(first, n_identity, qin, wq.count + round, wq.cursor + round, ws, r->accel->view, ctr, c->walk_refill_below, c->walk_schedule, rec_in, rec_in_count, wr.half[(round + 1) & 1], wq.pool + round, wr.capacity, wr.tail_budget)
        if (c->counting) { if (tl) TRACE(true, 1); else TRACE(true, 0); }
        else { if (tl) TRACE(false, 1); else TRACE(false, 0); }
#undef TRACE
        LAUNCHED(c);
        if ((rc = mark(r, st, TK_WALK_TRACE)) != CLV_OK) return rc;
#define SHADE(CT, SO) k_walk_shade<CT, SO><<<sgrid, CLV_BLOCK, 0, st>>>(first, n_identity, qin, wq.count + round, qout, wq.count + round + 1, ids, r->d_materials, r->d_camera, r->accel->view, r->d_seeds, store, ws, r->d_accum, ctr, rec_in, rec_in_count, wr.capacity)
        if (soa) { if (c->counting) SHADE(true, true); else SHADE(false, true); }
        else { if (c->counting) SHADE(true, false); else SHADE(false, false); }
#undef SHADE
        LAUNCHED(c);
        if ((rc = mark(r, st, TK_WALK_SHADE)) != CLV_OK) return rc;
    }
    return CLV_OK;
}

static int launch_pass(clv_renderer *r) {
    clv_ctx *c = r->ctx;
    const uint32_t *ids = r->sample_ids.empty() ? nullptr : r->d_sample_ids;
    uint32_t n = r->n_local;
    int rc = mark(r, c->stream, -1);
    if (rc != CLV_OK) return rc;
    PathStore store;
    memset(&store, 0, sizeof store);
    store.light_paths = r->d_light_paths;
    store.camera_paths = r->d_camera_paths;
    store.soa = r->d_soa;
    store.soa_len = r->d_soa_len;
    store.n = n;
    const bool soa = !r->path_records;
    LightSink sink;
    memset(&sink, 0, sizeof sink);
    sink.splat_image = r->d_light_image;
    sink.n_pixels = (int)r->n_pixels;
    const ConnectScratch cs_all = connect_scratch_at(r->d_connect_scratch, n);
    /* default configuration (ORDERED, SoA records, no counters) on a scene larger than the caches */
    const bool large = r->accel->large && r->mode == CLV_TRAVERSE_ORDERED && !c->counting && soa;
    /* stages 1-4 as wavefront rounds or as the megakernel.  The wavefront kernels exist for the ORDERED traversal only (the
     * reference-order mode is a parity instrument, not a product path).  Automatic: see clv_ctx_walk_mode */
    const bool wavefront = r->mode == CLV_TRAVERSE_ORDERED &&
                           (c->walk_mode >= 2 || (c->walk_mode == 0 && CLV_WAVEFRONT_BY_DEFAULT(r->accel->large) &&
                                                  n >= CLV_WAVEFRONT_MIN_SAMPLES));
    if (wavefront && (rc = ensure_walk_buffers(r)) != CLV_OK) return rc;
    if (wavefront && (c->walk_mode == 3 || (c->walk_mode == 0 && CLV_PIPELINE_BY_DEFAULT)) && (rc = ensure_pipe_buffers(r)) != CLV_OK) return rc;
    /* slices: whole blocks each; one slice (everything on the context's stream) when per-kernel events are wanted */
    int slices = c->pass_slices > 0 ? c->pass_slices : (n >= (uint32_t)c->sm_count * 1024u ? 3 : 1);
    if (r->timing) slices = 1;
    uint32_t per = (uint32_t)(((uint64_t)n + slices - 1) / slices);
    per = (per + CLV_BLOCK - 1) / CLV_BLOCK * CLV_BLOCK;
    if (per == 0) per = CLV_BLOCK;
    if (slices > 1) CK(cudaEventRecord(c->slice_fork, c->stream));
    for (int k = 0; k < slices; k++) {
        const uint32_t first = (uint32_t)std::min<uint64_t>((uint64_t)k * per, n), end = (uint32_t)std::min<uint64_t>((uint64_t)(k + 1) * per, n);
        if (first >= end) continue;
        cudaStream_t st = k == 0 ? c->stream : c->slice_stream[k];
        if (k > 0) CK(cudaStreamWaitEvent(st, c->slice_fork, 0));
        if (wavefront) {
            if ((rc = launch_wavefront_walks(r, st, k, first, end, ids, store, soa, large, slices > 1)) != CLV_OK) return rc;
        } else {
            if (large) {
                k_fused_walks<TRAVERSE_ORDERED, false, true, 1><<<GRID(end - first), CLV_BLOCK, 0, st>>>(first, end, ids, r->d_light_triangles, r->d_light_areas, r->d_light_tri_idx, r->light_count, r->d_materials, r->d_camera, r->accel->view, r->d_seeds, store, r->d_accum, c->d_counters);
            } else {
#define CALL(M, CT, SO) k_fused_walks<M, CT, SO><<<GRID(end - first), CLV_BLOCK, 0, st>>>(first, end, ids, r->d_light_triangles, r->d_light_areas, r->d_light_tri_idx, r->light_count, r->d_materials, r->d_camera, r->accel->view, r->d_seeds, store, r->d_accum, c->d_counters)
                DISPATCH_MCS(r->mode, c->counting, soa, CALL);
#undef CALL
            }
            LAUNCHED(c);
            if ((rc = mark(r, st, TK_FUSED_WALKS)) != CLV_OK) return rc;
        }
        ConnectScratch cs = cs_all;
        cs.task_count = cs_all.task_count + k;
        cs.cursor = cs_all.cursor + k;
        cs.tasks = cs_all.tasks + (size_t)first * 36;
        rc = launch_connect(c, st, r->mode, true, large, first, end, ids, store, soa, r->accel->view, r->d_materials,
                            r->d_camera, cs, r->d_aggs, nullptr, sink, r);
        if (rc != CLV_OK) return rc;
        if (k > 0) {
            CK(cudaEventRecord(c->slice_join[k], st));
            CK(cudaStreamWaitEvent(c->stream, c->slice_join[k], 0));
        }
    }
    k_fused_finalize<<<GRID(r->n_pixels), CLV_BLOCK, 0, c->stream>>>(r->n_pixels, r->d_aggs, r->d_camera, r->d_light_image, r->d_accum, r->d_pass_weight);
    LAUNCHED(c);
    if ((rc = mark(r, c->stream, TK_FINALIZE)) != CLV_OK) return rc;
    r->passes++;
    return CLV_OK;
}

/* per-kernel device times of the fused pass, from CUDA events on the launch stream */
extern "C" int clv_renderer_kernel_timing(clv_renderer *r, int enabled) {
    REQUIRE(r, "null");
    r->timing = enabled ? 1 : 0;
    return CLV_OK;
}
extern "C" int clv_renderer_kernel_times(clv_renderer *r, double ms_out[CLV_TIMED_KERNEL_SLOTS], uint64_t *passes_out, int reset) {
    REQUIRE(r && ms_out && passes_out, "null");
    static_assert(CLV_TIMED_KERNELS <= CLV_TIMED_KERNEL_SLOTS, "timing slots");
    CK(cudaSetDevice(r->ctx->device));
    CK(cudaStreamSynchronize(r->ctx->stream));
    for (size_t i = 0; i < r->events.size(); i++) {
        const int k = r->event_kernel[i];
        if (k < 0) { r->timed_passes++; continue; }
        float ms = 0.0f;
        CK(cudaEventElapsedTime(&ms, r->events[i - 1], r->events[i]));
        r->kernel_ms[k] += ms;
    }
    for (cudaEvent_t e : r->events) cudaEventDestroy(e);
    r->events.clear();
    r->event_kernel.clear();
    for (int k = 0; k < CLV_TIMED_KERNEL_SLOTS; k++) ms_out[k] = k < CLV_TIMED_KERNELS ? r->kernel_ms[k] : 0.0;
    *passes_out = r->timed_passes;
    if (reset) {
        for (int k = 0; k < CLV_PASS_KERNELS; k++) r->kernel_ms[k] = 0.0;
        r->timed_passes = 0;
    }
    return CLV_OK;
}

/* scheduling knobs of the two ray kernels (none changes a result); returns the previous value through *previous */
extern "C" int clv_ctx_set_knob(clv_ctx *c, const char *name, int value, int *previous) {
    REQUIRE(c && name, "null");
    int *slot = nullptr;
    if (!strcmp(name, "walk_refill_below")) slot = &c->walk_refill_below;
    else if (!strcmp(name, "walk_schedule")) slot = &c->walk_schedule;
    else if (!strcmp(name, "connect_refill_below")) slot = &c->refill_below;
    else if (!strcmp(name, "connect_schedule")) slot = &c->ray_schedule;
    else if (!strcmp(name, "connect_dynamic_fetch")) slot = &c->connect_dynamic_fetch;
    else if (!strcmp(name, "walk_grid_divisor")) slot = &c->walk_grid_divisor;
    else if (!strcmp(name, "finish_cooperative")) slot = &c->finish_cooperative;
    else if (!strcmp(name, "pipe_shade_every")) slot = &c->pipe_shade_every;
    else if (!strcmp(name, "walk_tail_budget")) slot = &c->walk_tail_budget;
    else return fail(CLV_ERR_INVALID, "clv_ctx_set_knob: unknown knob '%s'", name);
    if (previous) *previous = *slot;
    if (value >= 0 && slot == &c->ray_schedule) c->ray_schedule_large = value; /* one knob for both flavours */
    if (value >= 0) {
        if ((slot == &c->walk_refill_below || slot == &c->walk_schedule) && value < 1) value = 1;
        *slot = value;
    }
    return CLV_OK;
}

/* stages 1-4 of the fused pass: 0 = automatic, 1 = megakernel (k_fused_walks), 2 = wavefront (k_walk_init / k_walk_trace /
 * k_walk_shade).  < 0 only queries.  Results are bit-identical in both. */
extern "C" int clv_ctx_walk_mode(clv_ctx *c, int mode, int *current_out) {
    REQUIRE(c, "null ctx");
    REQUIRE(mode <= 3, "walk mode is 0 (automatic), 1 (megakernel), 2 (wavefront rounds) or 3 (wavefront, pipelined)");
    if (mode >= 0) c->walk_mode = mode;
    if (current_out) *current_out = c->walk_mode;
    return CLV_OK;
}

extern "C" int clv_renderer_run_sample(clv_renderer *r) {
    REQUIRE(r, "null");
    CK(cudaSetDevice(r->ctx->device));
    return launch_pass(r);
}

extern "C" int clv_renderer_run_sample_e2e(clv_renderer *r, const uint32_t *h_seeds, float *h_accum) {
    REQUIRE(r, "null");
    clv_ctx *c = r->ctx;
    CK(cudaSetDevice(c->device));
    if (h_seeds) {
        int rc = load_seeds(r, h_seeds);
        if (rc != CLV_OK) return rc;
    }
    int rc = launch_pass(r);
    if (rc != CLV_OK) return rc;
    if (h_accum) {
        CK(cudaMemcpyAsync(h_accum, r->d_accum, (size_t)r->n_pixels * sizeof(Accum), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return CLV_OK;
}

static int launch_tone_map(clv_renderer *r, float exposure) {
    clv_ctx *c = r->ctx;
    CK(cudaMemsetAsync(r->d_tone_sum, 0, sizeof(double), c->stream));
    k_tone_reduce<<<std::min<unsigned>((r->n_pixels + 255) / 256, 148u * 8u), 256, 0, c->stream>>>(r->n_pixels, r->d_accum, r->d_tone_sum);
    LAUNCHED(c);
    k_tone_map<<<(r->n_pixels + 255) / 256, 256, 0, c->stream>>>(r->n_pixels, r->d_accum, r->d_tone_sum, exposure, 1.0, r->d_tone_image);
    LAUNCHED(c);
    return CLV_OK;
}

/* Renderer.image on the device: tone-mapped uint8 BGR image [H][W][3] of the current estimate, downloaded to h_out */
extern "C" int clv_renderer_tone_map(clv_renderer *r, float exposure, uint8_t *h_out) {
    REQUIRE(r && h_out, "null");
    clv_ctx *c = r->ctx;
    CK(cudaSetDevice(c->device));
    int rc = launch_tone_map(r, exposure);
    if (rc != CLV_OK) return rc;
    CK(cudaMemcpyAsync(h_out, r->d_tone_image, (size_t)r->n_pixels * 3, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CLV_OK;
}

/* the interactive loop of render.py:31-37 in one call: seeds in (NULL = keep streams), one pass, preview image out */
extern "C" int clv_renderer_run_sample_preview(clv_renderer *r, const uint32_t *h_seeds, float exposure, uint8_t *h_image) {
    REQUIRE(r && h_image, "null");
    clv_ctx *c = r->ctx;
    CK(cudaSetDevice(c->device));
    int rc;
    if (h_seeds && (rc = load_seeds(r, h_seeds)) != CLV_OK) return rc;
    if ((rc = launch_pass(r)) != CLV_OK) return rc;
    if ((rc = launch_tone_map(r, exposure)) != CLV_OK) return rc;
    CK(cudaMemcpyAsync(h_image, r->d_tone_image, (size_t)r->n_pixels * 3, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CLV_OK;
}

extern "C" int clv_renderer_reset(clv_renderer *r, const uint32_t *h_seeds) {
    REQUIRE(r, "null");
    clv_ctx *c = r->ctx;
    CK(cudaSetDevice(c->device));
    CK(cudaMemsetAsync(r->d_accum, 0, (size_t)r->n_pixels * sizeof(Accum), c->stream));
    CK(cudaMemsetAsync(r->d_light_image, 0, (size_t)r->n_pixels * sizeof(V4), c->stream));
    CK(cudaMemsetAsync(r->d_aggs, 0, (size_t)r->n_pixels * sizeof(WeightAggregator), c->stream));
    CK(cudaMemsetAsync(r->d_pass_weight, 0, 16, c->stream));
    r->passes = 0;
    if (h_seeds) return load_seeds(r, h_seeds);
    return CLV_OK;
}

extern "C" int clv_renderer_device_ptr(clv_renderer *r, int buffer_id, void **d_ptr, size_t *nbytes) {
    REQUIRE(r && d_ptr && nbytes, "null");
    switch (buffer_id) {
    case CLV_BUF_ACCUM_PACKED: *d_ptr = r->d_accum; *nbytes = (size_t)r->n_pixels * sizeof(Accum); return CLV_OK;
    case CLV_BUF_SEEDS: *d_ptr = r->d_seeds; *nbytes = (size_t)r->n_local * 8; return CLV_OK;
    case CLV_BUF_LIGHT_PATHS:
    case CLV_BUF_CAMERA_PATHS:
        if (!r->path_records) return fail(CLV_ERR_INVALID, "path records are off: call clv_renderer_set_path_records(r, 1) before the pass");
        *d_ptr = buffer_id == CLV_BUF_LIGHT_PATHS ? r->d_light_paths : r->d_camera_paths;
        *nbytes = (size_t)r->n_local * sizeof(Path);
        return CLV_OK;
    case CLV_BUF_AGGREGATORS: *d_ptr = r->d_aggs; *nbytes = (size_t)r->n_pixels * sizeof(WeightAggregator); return CLV_OK;
    case CLV_BUF_PASS_WEIGHT: *d_ptr = r->d_pass_weight; *nbytes = 4; return CLV_OK;
    default: return fail(CLV_ERR_INVALID, "buffer %d has no device-side representation of its own", buffer_id);
    }
}

extern "C" int clv_renderer_read(clv_renderer *r, int buffer_id, void *h_dst, size_t nbytes) {
    REQUIRE(r && h_dst, "null");
    clv_ctx *c = r->ctx;
    CK(cudaSetDevice(c->device));
    if (r->d_pipe_ctrl) { /* the pipelined walk kernel's watchdog: a lost queue entry must not pass as a result */
        uint32_t ctrl[CLV_MAX_SLICES * 8];
        CK(cudaMemcpyAsync(ctrl, r->d_pipe_ctrl, sizeof ctrl, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int k = 0; k < CLV_MAX_SLICES; k++)
            if (ctrl[8 * k + PIPE_ERROR]) return fail(CLV_ERR_CUDA, "k_walk_pipeline gave up waiting for work with %u samples alive (slice %d)", ctrl[8 * k + PIPE_LIVE], k);
    }
    size_t np = r->n_pixels;
    if (buffer_id <= CLV_BUF_UNIDIRECTIONAL) {
        static const size_t per_pixel[4] = {12, 4, 4, 12};
        REQUIRE(nbytes == np * per_pixel[buffer_id], "size mismatch");
        std::vector<Accum> acc(np);
        CK(cudaMemcpyAsync(acc.data(), r->d_accum, np * sizeof(Accum), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        float *f = (float *)h_dst;
        int32_t *iv = (int32_t *)h_dst;
        for (size_t i = 0; i < np; i++) {
            switch (buffer_id) {
            case CLV_BUF_SUMMED_IMAGE: f[3 * i] = acc[i].image[0]; f[3 * i + 1] = acc[i].image[1]; f[3 * i + 2] = acc[i].image[2]; break;
            case CLV_BUF_SUMMED_SAMPLE_COUNTS: iv[i] = (int32_t)acc[i].count; break;
            case CLV_BUF_SUMMED_SAMPLE_WEIGHTS: f[i] = acc[i].weight; break;
            default: f[3 * i] = acc[i].uni[0]; f[3 * i + 1] = acc[i].uni[1]; f[3 * i + 2] = acc[i].uni[2]; break;
            }
        }
        return CLV_OK;
    }
    void *d;
    size_t sz;
    int rc = clv_renderer_device_ptr(r, buffer_id, &d, &sz);
    if (rc != CLV_OK) return rc;
    REQUIRE(nbytes == sz, "size mismatch");
    CK(cudaMemcpyAsync(h_dst, d, sz, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CLV_OK;
}

/* ---------------------------------------------------------------------------------------------------------------
 * section 6: parity / measurement hooks
 * ------------------------------------------------------------------------------------------------------------- */
extern "C" int clv_traverse(clv_ctx *c, const clv_accel *accel, const clv_Ray *h_rays, uint32_t n, int mode,
                            int32_t *h_tri, float *h_t, float *h_u, float *h_v) {
    REQUIRE(c && accel && (n == 0 || (h_rays && h_tri && h_t && h_u && h_v)), "null");
    REQUIRE(mode == CLV_TRAVERSE_REFERENCE_ORDER || mode == CLV_TRAVERSE_ORDERED, "bad mode");
    if (!n) return CLV_OK;
    CK(cudaSetDevice(c->device));
    Ray *d_rays = nullptr;
    int32_t *d_tri = nullptr;
    float *d_f = nullptr;
    CK(dev_alloc(c, &d_rays, (size_t)n * sizeof(Ray)));
    CK(dev_alloc(c, &d_tri, (size_t)n * 4));
    CK(dev_alloc(c, &d_f, (size_t)n * 12));
    CK(cudaMemcpyAsync(d_rays, h_rays, (size_t)n * sizeof(Ray), cudaMemcpyHostToDevice, c->stream));
#define CALL(M, CT) k_traverse<M, CT><<<GRID(n), CLV_BLOCK, 0, c->stream>>>(n, d_rays, accel->view, d_tri, d_f, d_f + n, d_f + 2 * (size_t)n, c->d_counters)
    DISPATCH_MC(mode, c->counting, CALL);
#undef CALL
    LAUNCHED(c);
    CK(cudaMemcpyAsync(h_tri, d_tri, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(h_t, d_f, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(h_u, d_f + n, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(h_v, d_f + 2 * (size_t)n, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    dev_free(c, d_rays); dev_free(c, d_tri); dev_free(c, d_f);
    return CLV_OK;
}

/* device-resident variant for timing: rays and outputs are device pointers (float outputs: t, u, v of n each) */
extern "C" int clv_traverse_device(clv_ctx *c, const clv_accel *accel, const clv_Ray *d_rays, uint32_t n, int mode,
                                   int32_t *d_tri, float *d_t, float *d_u, float *d_v) {
    REQUIRE(c && accel && d_rays && d_tri && d_t && d_u && d_v, "null");
    REQUIRE(mode == CLV_TRAVERSE_REFERENCE_ORDER || mode == CLV_TRAVERSE_ORDERED, "bad mode");
    if (!n) return CLV_OK;
    CK(cudaSetDevice(c->device));
#define CALL(M, CT) k_traverse<M, CT><<<GRID(n), CLV_BLOCK, 0, c->stream>>>(n, (const Ray *)d_rays, accel->view, d_tri, d_t, d_u, d_v, c->d_counters)
    DISPATCH_MC(mode, c->counting, CALL);
#undef CALL
    LAUNCHED(c);
    return CLV_OK;
}

extern "C" int clv_visibility(clv_ctx *c, const clv_accel *accel, const clv_Ray *h_a, const clv_Ray *h_b, uint32_t n,
                              int mode, int32_t *h_visible) {
    REQUIRE(c && accel && (n == 0 || (h_a && h_b && h_visible)), "null");
    REQUIRE(mode == CLV_TRAVERSE_REFERENCE_ORDER || mode == CLV_TRAVERSE_ORDERED, "bad mode");
    if (!n) return CLV_OK;
    CK(cudaSetDevice(c->device));
    Ray *d_a = nullptr, *d_b = nullptr;
    int32_t *d_v = nullptr;
    CK(dev_alloc(c, &d_a, (size_t)n * sizeof(Ray)));
    CK(dev_alloc(c, &d_b, (size_t)n * sizeof(Ray)));
    CK(dev_alloc(c, &d_v, (size_t)n * 4));
    CK(cudaMemcpyAsync(d_a, h_a, (size_t)n * sizeof(Ray), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_b, h_b, (size_t)n * sizeof(Ray), cudaMemcpyHostToDevice, c->stream));
#define CALL(M, CT) k_visibility<M, CT><<<GRID(n), CLV_BLOCK, 0, c->stream>>>(n, d_a, d_b, accel->view, d_v, c->d_counters)
    DISPATCH_MC(mode, c->counting, CALL);
#undef CALL
    LAUNCHED(c);
    CK(cudaMemcpyAsync(h_visible, d_v, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    dev_free(c, d_a); dev_free(c, d_b); dev_free(c, d_v);
    return CLV_OK;
}

/* streaming-read bandwidth of a buffer of `nbytes` read `iters` times back to back (first sweep untimed): with
 * nbytes well below the 126 MB L2 this is the L2 read peak the cache-resident configurations are bounded by, with
 * nbytes >> L2 the HBM read peak.  GB/s out. */
__global__ void k_stream_read(const uint4 *src, size_t n_vec, int sweeps, unsigned long long *sink) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned acc = 0;
    for (int sweep = 0; sweep < sweeps; sweep++) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n_vec; i += 4 * stride) {
            uint4 a = __ldcg(src + i), b = __ldcg(src + i + stride), c = __ldcg(src + i + 2 * stride), d = __ldcg(src + i + 3 * stride);
            acc += a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
        }
        for (; i < n_vec; i += stride) { uint4 a = __ldcg(src + i); acc += a.x ^ a.y ^ a.z ^ a.w; }
    }
    if (acc == 0x9e3779b9u) atomicAdd(sink, 1ull); /* keeps the loads alive */
}
extern "C" int clv_measure_read_bandwidth(clv_ctx *c, size_t nbytes, int iters, double *gbs_out) {
    REQUIRE(c && gbs_out && nbytes >= 1024 && iters > 0, "bad argument");
    CK(cudaSetDevice(c->device));
    uint4 *buf = nullptr;
    unsigned long long *sink = nullptr;
    CK(dev_alloc(c, &buf, nbytes));
    CK(dev_alloc(c, &sink, 8));
    CK(cudaMemsetAsync(buf, 1, nbytes, c->stream));
    CK(cudaMemsetAsync(sink, 0, 8, c->stream));
    const size_t n_vec = nbytes / 16;
    const unsigned grid = (unsigned)c->sm_count * 8u;
    /* enough sweeps per launch that launch latency does not matter (>= ~1 GB read per launch) */
    const int sweeps = (int)std::max<size_t>(1, ((size_t)1 << 30) / nbytes);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k_stream_read<<<grid, 256, 0, c->stream>>>(buf, n_vec, sweeps, sink);
    LAUNCHED(c);
    CK(cudaEventRecord(e0, c->stream));
    for (int k = 0; k < iters; k++) {
        k_stream_read<<<grid, 256, 0, c->stream>>>(buf, n_vec, sweeps, sink);
        LAUNCHED(c);
    }
    CK(cudaEventRecord(e1, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.0f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    *gbs_out = (double)n_vec * 16.0 * iters * sweeps / (ms * 1e-3) / 1e9;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    dev_free(c, buf);
    dev_free(c, sink);
    return CLV_OK;
}

extern "C" int clv_counters_enable(clv_ctx *c, int enabled) {
    REQUIRE(c, "null");
    c->counting = enabled ? 1 : 0;
    return CLV_OK;
}
extern "C" int clv_counters_read(clv_ctx *c, uint64_t out[10], int reset) {
    REQUIRE(c && out, "null");
    CK(cudaSetDevice(c->device));
    Counters h[COUNTERS_CLASSES];
    CK(cudaMemcpyAsync(h, c->d_counters, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < COUNTERS_CLASSES; k++) {
        uint64_t *o = out + CLV_COUNTER_WORDS * k;
        o[0] = h[k].rays_ref; o[1] = h[k].rays_traced; o[2] = h[k].node_fetches; o[3] = h[k].tri_fetches;
        o[4] = h[k].vertices;
    }
    if (reset) CK(cudaMemsetAsync(c->d_counters, 0, sizeof h, c->stream));
    return CLV_OK;
}
extern "C" int clv_ctx_pass_slices(clv_ctx *c, int slices, int *current_out) {
    REQUIRE(c, "null ctx");
    REQUIRE(slices <= CLV_MAX_SLICES, "at most 8 slices");
    if (slices >= 0) c->pass_slices = slices;
    if (current_out) *current_out = c->pass_slices;
    return CLV_OK;
}
extern "C" int clv_launch_count(clv_ctx *c, uint64_t *count) {
    REQUIRE(c && count, "null");
    *count = c->launches;
    return CLV_OK;
}
