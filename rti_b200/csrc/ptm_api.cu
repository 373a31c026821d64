// C ABI of libptm_b200.so (see include/ptm_b200.h).  Thin: argument checks,
// context bookkeeping and launches; no arithmetic lives here.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ptm_b200.h"
#include "ptm_device.cuh"
#include "ptm_hostpipe.h"
#include "ptm_internal.h"
#include "ptm_kernels.h"

namespace {

thread_local std::string g_error;

int fail(const char *what, const char *detail) {
    g_error = std::string(what) + ": " + detail;
    return 1;
}

#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess)                                     \
            return fail(#call, cudaGetErrorString(e_));            \
    } while (0)

#define REQUIRE(cond, msg)                                         \
    do {                                                           \
        if (!(cond))                                               \
            return fail("invalid argument", msg);                  \
    } while (0)

size_t sample_size(int sample_type) {
    return sample_type == PTMB_U8 ? 1 : sample_type == PTMB_U16 ? 2 : sample_type == PTMB_U32 ? 4 : 0;
}

}  // namespace

extern "C" {

const char *ptmb_last_error(void) { return g_error.c_str(); }

// used by ptm_api_host.cu
int ptmb_internal_fail(const char *what, const char *detail) { return fail(what, detail); }
int ptmb_internal_device(const ptmb_ctx_t *c) { return c->device; }
void *ptmb_internal_stream(const ptmb_ctx_t *c) { return (void *) c->stream; }
void *ptmb_stream(const ptmb_ctx_t *c) { return c ? (void *) c->stream : nullptr; }
int ptmb_device(const ptmb_ctx_t *c) { return c ? c->device : -1; }
int32_t *ptmb_internal_keys(ptmb_ctx_t *c) { return c->keys_dev; }
const float *ptmb_internal_matrix(const ptmb_ctx_t *c) { return c->M_dev; }

int ptmb_device_count(int *count) {
    REQUIRE(count, "count is NULL");
    CK(cudaGetDeviceCount(count));
    return 0;
}

int ptmb_create(int device, void *stream, ptmb_ctx_t **out) {
    REQUIRE(out, "ctx out pointer is NULL");
    *out = nullptr;
    int count = 0;
    CK(cudaGetDeviceCount(&count));
    if (count <= 0)
        return fail("ptmb_create", "no CUDA device (this library has no CPU path)");
    REQUIRE(device >= 0 && device < count, "device index out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail("ptmb_create", "kernels are built for sm_100a only");
    ptmb_ctx *c = new ptmb_ctx;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->stream = (cudaStream_t) stream;  // NULL = the CUDA default stream
    cudaError_t e = cudaMalloc(&c->keys_dev, PTMB_KEYS * sizeof(int));
    if (e == cudaSuccess)
        e = cudaMalloc(&c->sb_dev, sizeof(ptmb_scale_bias_t));
    if (e == cudaSuccess)
        e = cudaMalloc(&c->ticket_dev, 2 * sizeof(unsigned));   // [0] ticket, [1] dynamic tile counter
    if (e == cudaSuccess)
        e = cudaMemset(c->ticket_dev, 0, 2 * sizeof(unsigned));
    if (e == cudaSuccess)
        e = cudaMalloc(&c->mailbox_dev, ptm::XchgLayout::bytes);
    if (e == cudaSuccess)
        e = cudaMemset(c->mailbox_dev, 0, ptm::XchgLayout::bytes);
    if (e == cudaSuccess)
        e = ptm::launch_keys_init(c->keys_dev, c->stream);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        delete c;
        return fail("cudaMalloc", cudaGetErrorString(e));
    }
    *out = c;
    return 0;
}

static void encode_scratch_release(ptmb_ctx *c);
extern "C" void ptmb_internal_host_release(ptmb_ctx_t *c);
ptm::HostState **ptmb_internal_host(ptmb_ctx_t *c) { return &c->host; }
int ptmb_internal_host_threads(const ptmb_ctx_t *c) { return c->host_threads; }
void ptmb_internal_set_host_threads(ptmb_ctx_t *c, int n) { c->host_threads = n; }

void ptmb_destroy(ptmb_ctx_t *c) {
    if (!c)
        return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    encode_scratch_release(c);
    ptmb_internal_host_release(c);
    cudaFree(c->M_dev);
    cudaFree(c->digits_dev);
    cudaFree(c->keys_dev);
    cudaFree(c->sb_dev);
    cudaFree(c->ticket_dev);
    cudaFree(c->mailbox_dev);
    for (cudaEvent_t ev : c->probe_events)
        cudaEventDestroy(ev);
    cudaFree(c->light_dev);
    cudaFree(c->lq_dev);
    delete c;
}

int ptmb_set_stream(ptmb_ctx_t *c, void *stream) {
    REQUIRE(c, "ctx is NULL");
    c->stream = (cudaStream_t) stream;
    return 0;
}

int ptmb_synchronize(ptmb_ctx_t *c) {
    REQUIRE(c, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int ptmb_sm_count(const ptmb_ctx_t *c) { return c ? c->sm_count : 0; }

int ptmb_malloc(ptmb_ctx_t *c, size_t bytes, void **dev) {
    REQUIRE(c && dev, "NULL argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMalloc(dev, bytes ? bytes : 1));
    return 0;
}

int ptmb_free(ptmb_ctx_t *c, void *dev) {
    REQUIRE(c, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaFree(dev));
    return 0;
}

int ptmb_upload(ptmb_ctx_t *c, void *dev, const void *host, size_t bytes) {
    REQUIRE(c && (bytes == 0 || (dev && host)), "NULL argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->stream));
    return 0;
}

int ptmb_download(ptmb_ctx_t *c, void *host, const void *dev, size_t bytes) {
    REQUIRE(c && (bytes == 0 || (dev && host)), "NULL argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int ptmb_host_register(const void *host, size_t bytes) {
    CK(cudaHostRegister(const_cast<void *>(host), bytes, cudaHostRegisterDefault));
    return 0;
}

int ptmb_host_alloc(size_t bytes, int write_combined, void **host) {
    REQUIRE(host, "NULL argument");
    CK(cudaHostAlloc(host, bytes ? bytes : 1,
                     cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault)));
    return 0;
}

int ptmb_host_free(void *host) {
    CK(cudaFreeHost(host));
    return 0;
}

int ptmb_host_unregister(const void *host) {
    CK(cudaHostUnregister(const_cast<void *>(host)));
    return 0;
}

int ptmb_set_matrix(ptmb_ctx_t *c, const float *M, unsigned n_lights) {
    REQUIRE(c && M, "NULL argument");
    REQUIRE(n_lights >= 1 && n_lights <= 2048, "n_lights must be in 1..2048 (the pinv table lives in shared memory)");
    CK(cudaSetDevice(c->device));
    if (n_lights != c->n_lights || !c->M_dev) {
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaFree(c->M_dev));
        c->M_dev = nullptr;
        CK(cudaMalloc(&c->M_dev, (size_t) n_lights * PTMB_COEFFICIENTS * sizeof(float)));
    }
    // the previous matrix may still be in use by queued kernels reading M_host's copy: the
    // copy below is stream ordered, and pageable sources are staged before the call returns
    c->M_host.assign(M, M + (size_t) n_lights * PTMB_COEFFICIENTS);
    c->n_lights = n_lights;
    c->magic = n_lights == 1 ? 0u : (unsigned) ((1ull << 32) / n_lights) + 1u;
    CK(cudaMemcpyAsync(c->M_dev, c->M_host.data(), c->M_host.size() * sizeof(float), cudaMemcpyHostToDevice,
                       c->stream));
    // the same matrix as fixed-point digits for the tensor-core engine
    std::vector<int8_t> image;
    c->digits_ok = ptm::pack_mma_digits(c->M_host.data(), n_lights, image, c->digit_scale, &c->kpad);
    if (c->digits_ok) {
        if (image.size() > c->digits_cap) {
            CK(cudaStreamSynchronize(c->stream));
            CK(cudaFree(c->digits_dev));
            c->digits_dev = nullptr;
            c->digits_cap = 0;
            CK(cudaMalloc(&c->digits_dev, image.size()));
            c->digits_cap = image.size();
        }
        c->digits_host.swap(image);
        CK(cudaMemcpyAsync(c->digits_dev, c->digits_host.data(), c->digits_host.size(), cudaMemcpyHostToDevice,
                           c->stream));
    }
    return 0;
}

int ptmb_mma_pack_digits(const float *M, unsigned n_lights, int8_t *image, size_t image_bytes, float *scale,
                         unsigned *kpad) {
    REQUIRE(M && image && scale && kpad, "NULL argument");
    std::vector<int8_t> img;
    if (!ptm::pack_mma_digits(M, n_lights, img, scale, kpad))
        return fail("ptmb_mma_pack_digits", "matrix not representable (non-finite entries or more than 224 lights)");
    REQUIRE(image_bytes >= img.size(), "image buffer too small (needs 32 bytes per light rounded up to 32 lights)");
    memcpy(image, img.data(), img.size());
    return 0;
}

unsigned long long ptmb_mma_launches(void) { return ptm::mma_launch_count(); }

int ptmb_set_fit_engine(ptmb_ctx_t *c, int engine) {
    REQUIRE(c, "ctx is NULL");
    REQUIRE(engine >= PTMB_ENGINE_DEFAULT && engine <= PTMB_ENGINE_MMA, "unknown engine");
    c->engine = engine;
    return 0;
}

int ptmb_keys_init(ptmb_ctx_t *c, int32_t *keys_dev) {
    REQUIRE(c && keys_dev, "NULL argument");
    CK(cudaSetDevice(c->device));
    CK(ptm::launch_keys_init(keys_dev, c->stream));
    return 0;
}

static int make_fit_args(ptmb_ctx_t *c, const void *stack_dev, size_t image_stride, size_t pixels,
                         ptm::FitArgs &a) {
    REQUIRE(c, "ctx is NULL");
    REQUIRE(c->M_dev, "ptmb_set_matrix has not been called");
    REQUIRE(stack_dev || pixels == 0, "stack is NULL");
    a.stack = (const uint8_t *) stack_dev;
    a.image_stride = image_stride;
    a.pixels = pixels;
    a.n_lights = c->n_lights;
    a.magic = c->magic;
    a.M = c->M_dev;
    a.coeffs = nullptr;
    a.plane_stride = 0;
    a.rgb = nullptr;
    a.keys = nullptr;
    a.handoff = ptm::KeyHandoff{};
    a.digits = c->digits_ok ? c->digits_dev : nullptr;
    memcpy(a.digit_scale, c->digit_scale, sizeof(a.digit_scale));
    a.kpad = c->kpad;
    a.engine = c->engine;
    a.mode = 0;   // the LRGB-shaped entries set it from the context (ptmb_set_fit_mode)
    a.stack_stable = 0;
    a.tile_counter = c->ticket_dev ? c->ticket_dev + 1 : nullptr;
    static int store_hint = -1;
    if (store_hint < 0) {
        const char *v = getenv("PTM_K1_STORE_HINT");
        store_hint = v ? atoi(v) : 2;
    }
    a.store_hint = store_hint;
    return 0;
}

int ptmb_fit_lrgb(ptmb_ctx_t *c, const void *stack_dev, int sample_type, size_t image_stride, size_t pixels,
                  float *coeffs_dev, uint8_t *rgb_dev, int32_t *keys_dev) {
    ptm::FitArgs a;
    if (make_fit_args(c, stack_dev, image_stride, pixels, a))
        return 1;
    REQUIRE(sample_type == PTMB_U8 || sample_type == PTMB_U16, "LRGB fit takes uint8 or uint16 stacks");
    REQUIRE(c->n_lights <= 65536 || sample_type == PTMB_U8, "too many lights for 16-bit sums");
    REQUIRE(rgb_dev && keys_dev, "rgb / keys output is NULL");
    REQUIRE(image_stride >= pixels * 3 * sample_size(sample_type), "image_stride smaller than one image");
    a.coeffs = coeffs_dev;
    a.rgb = rgb_dev;
    a.keys = keys_dev;
    a.mode = c->fit_mode;
    REQUIRE(a.mode == 0 || sample_type == PTMB_U8, "PTM_FORMAT_LUM / RGB-input stacks are 8-bit paths");
    CK(cudaSetDevice(c->device));
    CK(ptm::launch_fit_lrgb(a, sample_type, c->sm_count, c->stream));
    return 0;
}

int ptmb_set_coeffs_scratch(ptmb_ctx_t *c, int scratch) {
    REQUIRE(c, "ctx is NULL");
    c->coeffs_scratch = scratch != 0;
    return 0;
}

int ptmb_set_stack_stable(ptmb_ctx_t *c, int stable) {
    REQUIRE(c, "ctx is NULL");
    c->stack_stable = stable != 0;
    return 0;
}

int ptmb_set_fit_mode(ptmb_ctx_t *c, int flags) {
    REQUIRE(c, "ctx is NULL");
    REQUIRE((flags & ~(PTMB_FIT_LUM | PTMB_FIT_RGB_INPUT)) == 0, "unknown fit mode flag");
    c->fit_mode = flags;
    return 0;
}

int ptmb_fit_rgb(ptmb_ctx_t *c, const void *stack_dev, int sample_type, size_t image_stride, size_t pixels,
                 float *coeffs_dev, size_t plane_stride, int32_t *keys_dev) {
    ptm::FitArgs a;
    if (make_fit_args(c, stack_dev, image_stride, pixels, a))
        return 1;
    REQUIRE(sample_type == PTMB_U8 || sample_type == PTMB_U16, "RGB fit takes uint8 or uint16 stacks");
    REQUIRE(coeffs_dev && keys_dev, "coeffs / keys output is NULL");
    REQUIRE(plane_stride >= pixels * PTMB_COEFFICIENTS, "plane_stride smaller than one plane");
    REQUIRE(image_stride >= pixels * 3 * sample_size(sample_type), "image_stride smaller than one image");
    a.coeffs = coeffs_dev;
    a.plane_stride = plane_stride;
    a.keys = keys_dev;
    CK(cudaSetDevice(c->device));
    CK(ptm::launch_fit_rgb(a, sample_type, c->sm_count, c->stream));
    return 0;
}

int ptmb_fit_channel(ptmb_ctx_t *c, const void *samples_dev, int sample_type, size_t pixel_stride,
                     size_t image_stride, size_t pixels, float *coeffs_dev) {
    REQUIRE(c && c->M_dev, "ptmb_set_matrix has not been called");
    REQUIRE(sample_size(sample_type) != 0, "unknown sample type");
    REQUIRE((samples_dev && coeffs_dev) || pixels == 0, "NULL argument");
    CK(cudaSetDevice(c->device));
    CK(ptm::launch_fit_channel(samples_dev, sample_type, pixel_stride, image_stride, pixels, c->n_lights,
                               c->M_dev, coeffs_dev, c->sm_count, c->stream));
    return 0;
}

int ptmb_chroma_avg(ptmb_ctx_t *c, const uint8_t *stack_dev, size_t image_stride, size_t pixels,
                    unsigned n_lights, uint8_t *avg_dev) {
    REQUIRE(c && ((stack_dev && avg_dev) || pixels == 0), "NULL argument");
    REQUIRE(n_lights >= 1, "n_lights must be >= 1");
    CK(cudaSetDevice(c->device));
    CK(ptm::launch_chroma_avg(stack_dev, image_stride, pixels, n_lights, avg_dev, c->sm_count, c->stream));
    return 0;
}

int ptmb_lrgb_colour_normalise(ptmb_ctx_t *c, uint8_t *ycc_dev, float *coeffs_dev, size_t pixels) {
    REQUIRE(c && ((ycc_dev && coeffs_dev) || pixels == 0), "NULL argument");
    CK(cudaSetDevice(c->device));
    CK(ptm::launch_lrgb_colour_normalise(ycc_dev, coeffs_dev, pixels, c->sm_count, c->stream));
    return 0;
}

int ptmb_minmax(ptmb_ctx_t *c, const float *coeffs_dev, size_t pixels, int32_t *keys_dev) {
    REQUIRE(c && keys_dev && (coeffs_dev || pixels == 0), "NULL argument");
    CK(cudaSetDevice(c->device));
    CK(ptm::launch_minmax(coeffs_dev, pixels, keys_dev, c->sm_count, c->stream));
    return 0;
}

int ptmb_quantise(ptmb_ctx_t *c, const float *coeffs_dev, size_t pixels, const int32_t *keys_dev,
                  uint8_t *bytes_dev, ptmb_scale_bias_t *sb_dev) {
    REQUIRE(c && keys_dev, "NULL argument");
    REQUIRE((coeffs_dev && bytes_dev) || pixels == 0, "NULL argument");
    CK(cudaSetDevice(c->device));
    CK(ptm::launch_quantise(coeffs_dev, pixels, ptm::KeySource{keys_dev, nullptr, 0, 0}, bytes_dev, sb_dev, false,
                            c->sm_count, c->stream));
    return 0;
}

int ptmb_quantise_consume(ptmb_ctx_t *c, float *coeffs_dev, size_t pixels, const int32_t *keys_dev,
                          uint8_t *bytes_dev, ptmb_scale_bias_t *sb_dev) {
    REQUIRE(c && keys_dev, "NULL argument");
    REQUIRE((coeffs_dev && bytes_dev) || pixels == 0, "NULL argument");
    CK(cudaSetDevice(c->device));
    CK(ptm::launch_quantise(coeffs_dev, pixels, ptm::KeySource{keys_dev, nullptr, 0, 0}, bytes_dev, sb_dev, true,
                            c->sm_count, c->stream));
    return 0;
}

// ---------------------------------------------------------------------------
// Whole device-resident LRGB encode of one slab in two launches: K1 hands the merged keys to the
// mailbox of every rank from its last block (no keys-init, no separate exchange kernels), K2
// picks them up in its prologue.
// ---------------------------------------------------------------------------
int ptmb_encode_lrgb_dev(ptmb_ctx_t *c, ptmb_xchg_t *x, const void *stack_dev, int sample_type, size_t image_stride,
                         size_t pixels, float *coeffs_dev, uint8_t *rgb_dev, uint8_t *coef_bytes_dev,
                         ptmb_scale_bias_t *sb_dev) {
    ptm::FitArgs a;
    if (make_fit_args(c, stack_dev, image_stride, pixels, a))
        return 1;
    REQUIRE(sample_type == PTMB_U8 || sample_type == PTMB_U16, "LRGB fit takes uint8 or uint16 stacks");
    REQUIRE(coeffs_dev && rgb_dev && coef_bytes_dev, "output pointer is NULL");
    REQUIRE(image_stride >= pixels * 3 * sample_size(sample_type), "image_stride smaller than one image");
    REQUIRE(!x || x->device == c->device, "exchange belongs to another device");
    CK(cudaSetDevice(c->device));
    a.coeffs = coeffs_dev;
    a.rgb = rgb_dev;
    a.keys = c->keys_dev;   // kept at the initial extrema between encodes (reset by the hand-off)
    a.mode = c->fit_mode;
    a.stack_stable = c->stack_stable;
    REQUIRE(a.mode == 0 || sample_type == PTMB_U8, "PTM_FORMAT_LUM / RGB-input stacks are 8-bit paths");
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (c->probe_next + 1 < c->probe_events.size()) {
        ev0 = c->probe_events[c->probe_next];
        ev1 = c->probe_events[c->probe_next + 1];
        c->probe_next += 2;
    }
    if (ev0)
        CK(cudaEventRecord(ev0, c->stream));

    // fused form: one streaming launch covers the whole slab
    ptm::KeySource src{};
    size_t done = 0;
    if (sample_type == PTMB_U8 && pixels > 0) {
        a.handoff.ticket = c->ticket_dev;
        a.handoff.tile_counter = a.tile_counter;
        a.handoff.world = x ? x->world : 1;
        a.handoff.rank = x ? x->rank : 0;
        a.handoff.epoch = x ? x->epoch + 1 : c->epoch + 1;
        for (int r = 0; r < a.handoff.world; ++r)
            a.handoff.mailbox[r] = x ? x->mailbox[r] : c->mailbox_dev;
        // the hand-off may only ride on a launch that covers every pixel of the slab
        const bool mma = a.mode == 0 && ptm::resolve_engine(a.engine, ptm::kDefaultEngineLrgbU8) == 2;
        if (mma && (pixels % 128) % 16 == 0)
            CK(ptm::launch_fit_mma(a, 1, c->sm_count, c->stream, &done));
        if (done == 0 && (pixels % 1024) % 16 == 0)
            CK(ptm::launch_fit_lrgb_tma(a, c->sm_count, c->stream, &done));
        if (done == pixels) {
            if (x)
                ++x->epoch;
            else
                ++c->epoch;
            src.mailbox = x ? x->local : c->mailbox_dev;
            src.world = a.handoff.world;
            src.epoch = a.handoff.epoch;
        } else {
            REQUIRE(done == 0, "internal: partial streaming launch with a hand-off");
        }
    }
    if (!src.mailbox) {
        // general form: any layout / sample type, keys merged by separate kernels
        a.handoff = ptm::KeyHandoff{};
        CK(ptm::launch_fit_lrgb(a, sample_type, c->sm_count, c->stream));
        if (x && ptmb_xchg_allreduce_max(c, x, c->keys_dev))
            return 1;
        src.keys = c->keys_dev;
    }
    if (ev1)
        CK(cudaEventRecord(ev1, c->stream));
    // scratch plane (ptmb_set_coeffs_scratch): K2 takes the L2-resident end of the plane first and discards its lines
    CK(ptm::launch_quantise(coeffs_dev, pixels, src, coef_bytes_dev, sb_dev, c->coeffs_scratch != 0, c->sm_count,
                            c->stream));
    if (!src.mailbox)
        CK(ptm::launch_keys_init(c->keys_dev, c->stream));   // leave the context keys reset, as the hand-off does
    return 0;
}

// ---------------------------------------------------------------------------
// The same for the RGB formats (three coefficient planes, rti-builder/ptm-encoder.c:272-284,406): the
// tensor-core fit kernel carries the key hand-off, one K2 launch covers the three contiguous planes.
// ---------------------------------------------------------------------------
int ptmb_encode_rgb_dev(ptmb_ctx_t *c, ptmb_xchg_t *x, const void *stack_dev, int sample_type, size_t image_stride,
                        size_t pixels, float *coeffs_dev, uint8_t *coef_bytes_dev, ptmb_scale_bias_t *sb_dev) {
    ptm::FitArgs a;
    if (make_fit_args(c, stack_dev, image_stride, pixels, a))
        return 1;
    REQUIRE(sample_type == PTMB_U8 || sample_type == PTMB_U16, "RGB fit takes uint8 or uint16 stacks");
    REQUIRE(coeffs_dev && coef_bytes_dev, "output pointer is NULL");
    REQUIRE(image_stride >= pixels * 3 * sample_size(sample_type), "image_stride smaller than one image");
    REQUIRE(!x || x->device == c->device, "exchange belongs to another device");
    CK(cudaSetDevice(c->device));
    a.coeffs = coeffs_dev;
    a.plane_stride = pixels * PTMB_COEFFICIENTS;   // contiguous planes: one K2 launch covers all three
    a.keys = c->keys_dev;
    ptm::KeySource src{};
    size_t done = 0;
    const int fmt_default = sample_type == PTMB_U8 ? ptm::kDefaultEngineRgbU8 : ptm::kDefaultEngineRgbU16;
    const unsigned unit_px = sample_type == PTMB_U8 ? 128 : 64;
    if (pixels > 0 && ptm::resolve_engine(a.engine, fmt_default) == 2 && (pixels % unit_px) % 16 == 0) {
        a.handoff.ticket = c->ticket_dev;
        a.handoff.tile_counter = a.tile_counter;
        a.handoff.world = x ? x->world : 1;
        a.handoff.rank = x ? x->rank : 0;
        a.handoff.epoch = x ? x->epoch + 1 : c->epoch + 1;
        for (int r = 0; r < a.handoff.world; ++r)
            a.handoff.mailbox[r] = x ? x->mailbox[r] : c->mailbox_dev;
        CK(ptm::launch_fit_mma(a, sample_type == PTMB_U8 ? 0 : 2, c->sm_count, c->stream, &done));
        if (done == pixels) {
            if (x)
                ++x->epoch;
            else
                ++c->epoch;
            src.mailbox = x ? x->local : c->mailbox_dev;
            src.world = a.handoff.world;
            src.epoch = a.handoff.epoch;
        } else {
            REQUIRE(done == 0, "internal: partial tensor-core launch with a hand-off");
        }
    }
    if (!src.mailbox) {
        a.handoff = ptm::KeyHandoff{};
        CK(ptm::launch_fit_rgb(a, sample_type, c->sm_count, c->stream));
        if (x && ptmb_xchg_allreduce_max(c, x, c->keys_dev))
            return 1;
        src.keys = c->keys_dev;
    }
    CK(ptm::launch_quantise(coeffs_dev, pixels * 3, src, coef_bytes_dev, sb_dev, c->coeffs_scratch != 0, c->sm_count,
                            c->stream));
    if (!src.mailbox)
        CK(ptm::launch_keys_init(c->keys_dev, c->stream));
    return 0;
}

// Timing probes: the next `pairs` calls of ptmb_encode_lrgb_dev record a CUDA event before and after
// their fit stage on the context's stream; ptmb_probe_read returns the elapsed milliseconds.
int ptmb_probe_arm(ptmb_ctx_t *c, unsigned pairs) {
    REQUIRE(c, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    for (cudaEvent_t ev : c->probe_events)
        cudaEventDestroy(ev);
    c->probe_events.clear();
    c->probe_next = 0;
    for (unsigned i = 0; i < 2 * pairs; ++i) {
        cudaEvent_t ev;
        CK(cudaEventCreate(&ev));
        c->probe_events.push_back(ev);
    }
    return 0;
}

int ptmb_probe_read(ptmb_ctx_t *c, float *ms_out, unsigned pairs, unsigned *recorded) {
    REQUIRE(c && ms_out && recorded, "NULL argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    unsigned n = (unsigned) (c->probe_next / 2);
    if (n > pairs)
        n = pairs;
    for (unsigned i = 0; i < n; ++i)
        CK(cudaEventElapsedTime(&ms_out[i], c->probe_events[2 * i], c->probe_events[2 * i + 1]));
    *recorded = n;
    return 0;
}

// ptm_normal (rti-builder/ptmlib.c:808-814) for every pixel: from a float coefficient plane ...
int ptmb_normal_map(ptmb_ctx_t *c, const float *coeffs_dev, size_t pixels, float *normals_dev) {
    REQUIRE(c && ((coeffs_dev && normals_dev) || pixels == 0), "NULL argument");
    CK(cudaSetDevice(c->device));
    CK(ptm::launch_normal_map(coeffs_dev, nullptr, ptm::NormalArgs{}, pixels, normals_dev, c->sm_count, c->stream));
    return 0;
}

// ... or from a quantised coefficient block, unscaled with the header's scale / bias first (:54)
int ptmb_normal_map_blocks(ptmb_ctx_t *c, const uint8_t *coef_dev, const float *scale, const int32_t *bias,
                           size_t pixels, float *normals_dev) {
    REQUIRE(c && scale && bias && ((coef_dev && normals_dev) || pixels == 0), "NULL argument");
    ptm::NormalArgs a;
    memcpy(a.scale, scale, sizeof(a.scale));
    memcpy(a.bias, bias, sizeof(a.bias));
    CK(cudaSetDevice(c->device));
    CK(ptm::launch_normal_map(nullptr, coef_dev, a, pixels, normals_dev, c->sm_count, c->stream));
    return 0;
}

// The disabled "L = R + G + B" branch of the reference encoder (rti-builder/ptm-encoder.c:356-400).
int ptmb_fit_lsum(ptmb_ctx_t *c, const uint8_t *stack_dev, size_t image_stride, size_t pixels, int chroma,
                  float *coeffs_dev, uint8_t *colour_dev, int32_t *keys_dev) {
    ptm::FitArgs a;
    if (make_fit_args(c, stack_dev, image_stride, pixels, a))
        return 1;
    REQUIRE(chroma == PTMB_CHROMA_MEAN || chroma == PTMB_CHROMA_MEDIAN, "chroma must be PTMB_CHROMA_MEAN or _MEDIAN");
    REQUIRE((coeffs_dev && colour_dev && keys_dev) || pixels == 0, "output pointer is NULL");
    REQUIRE(image_stride >= pixels * 3, "image_stride smaller than one image");
    a.coeffs = coeffs_dev;
    a.rgb = colour_dev;
    a.keys = keys_dev;
    CK(cudaSetDevice(c->device));
    CK(ptm::launch_fit_lsum(a, chroma == PTMB_CHROMA_MEDIAN, c->sm_count, c->stream));
    return 0;
}

static int relight_common(ptmb_ctx_t *c, ptm::RelightArgs &a, bool lrgb, size_t width, size_t height,
                          const float *scale, const int32_t *bias, const float *uv, unsigned n_dirs,
                          uint8_t *out_dev, bool lum = false) {
    REQUIRE(c && scale && bias, "NULL argument");
    REQUIRE((uv && out_dev) || n_dirs == 0, "NULL argument");
    CK(cudaSetDevice(c->device));
    if (n_dirs == 0 || width * height == 0)
        return 0;
    a.width = width;
    a.height = height;
    memcpy(a.scale, scale, sizeof(a.scale));
    memcpy(a.bias, bias, sizeof(a.bias));
    // light vector per direction (rti-builder/ptmlib.c:556-562)
    c->light_host.resize((size_t) n_dirs * PTMB_COEFFICIENTS);
    for (unsigned d = 0; d < n_dirs; ++d) {
        const float u = uv[2 * d], v = uv[2 * d + 1];
        float *l = &c->light_host[(size_t) d * PTMB_COEFFICIENTS];
        l[0] = u * u;
        l[1] = v * v;
        l[2] = u * v;
        l[3] = u;
        l[4] = v;
        l[5] = 1.0f;
    }
    if (c->light_cap < n_dirs) {
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaFree(c->light_dev));
        c->light_dev = nullptr;
        c->light_cap = 0;
        CK(cudaMalloc(&c->light_dev, (size_t) n_dirs * PTMB_COEFFICIENTS * sizeof(float)));
        c->light_cap = n_dirs;
    }
    CK(cudaMemcpyAsync(c->light_dev, c->light_host.data(), c->light_host.size() * sizeof(float),
                       cudaMemcpyHostToDevice, c->stream));
    static int env_mode = -1;
    if (env_mode < 0) {
        const char *v = getenv("PTM_RELIGHT");
        env_mode = !v ? PTMB_RELIGHT_AUTO : !strcmp(v, "exact") ? PTMB_RELIGHT_EXACT : !strcmp(v, "fast") ? PTMB_RELIGHT_FAST
                                                                                                       : PTMB_RELIGHT_AUTO;
    }
    const int mode = c->relight_mode != PTMB_RELIGHT_AUTO ? c->relight_mode : env_mode;
    // AUTO: a single direction is memory bound either way and stays bit-exact; batches take the tolerance kernels
    const bool fast = mode == PTMB_RELIGHT_FAST || (mode == PTMB_RELIGHT_AUTO && n_dirs >= 4);
    // RGB formats, tolerance form: the tensor-core kernel takes the sweep when every term scale_k * l_k(d) fits its
    // two-term f16 split (ptm_relight_mma.cu); PTM_RELIGHT_RGB_ENGINE=ffma keeps the FFMA kernel (tuning aid)
    float max_term = 0.f;
    const char *rgb_engine = getenv("PTM_RELIGHT_RGB_ENGINE");   // read per call: the tests compare both engines
    const bool rgb_mma = !(rgb_engine && !strcmp(rgb_engine, "ffma"));
    if (fast && !lum && !lrgb && rgb_mma)
        for (size_t i = 0; i < c->light_host.size(); ++i) {
            const float x = fabsf(scale[i % PTMB_COEFFICIENTS] * c->light_host[i]);
            max_term = x > max_term || x != x ? x : max_term;
        }
    const unsigned kBatch = 960;  // light table lives in shared memory (48 bytes per direction in the packed kernels)
    for (unsigned d0 = 0; d0 < n_dirs; d0 += kBatch) {
        a.n_dirs = n_dirs - d0 < kBatch ? n_dirs - d0 : kBatch;
        a.light = c->light_dev + (size_t) d0 * PTMB_COEFFICIENTS;
        a.out = out_dev + (size_t) d0 * width * height * 3;
        cudaError_t e = cudaErrorNotSupported;
        if (fast && !lum && !lrgb && rgb_mma)
            e = ptm::launch_relight_rgb_mma(a, max_term, c->sm_count, c->stream);
        if (e == cudaErrorNotSupported && fast && !lum)
            e = ptm::launch_relight_fast(a, lrgb, c->sm_count, c->stream);
        if (e == cudaErrorNotSupported)
            e = lum ? ptm::launch_relight_lum(a, c->sm_count, c->stream) : ptm::launch_relight(a, lrgb, c->sm_count, c->stream);
        CK(e);
    }
    return 0;
}

int ptmb_set_relight_mode(ptmb_ctx_t *c, int mode) {
    REQUIRE(c, "ctx is NULL");
    REQUIRE(mode >= PTMB_RELIGHT_AUTO && mode <= PTMB_RELIGHT_FAST, "unknown relight mode");
    c->relight_mode = mode;
    return 0;
}

int ptmb_relight_lrgb(ptmb_ctx_t *c, const uint8_t *coef_dev, const uint8_t *rgb_dev, size_t width,
                      size_t height, const float *scale, const int32_t *bias, const float *uv, unsigned n_dirs,
                      uint8_t *out_dev) {
    REQUIRE((coef_dev && rgb_dev) || width * height == 0, "NULL block");
    ptm::RelightArgs a{};
    a.plane[0] = coef_dev;
    a.plane[1] = rgb_dev;
    a.plane[2] = nullptr;
    return relight_common(c, a, true, width, height, scale, bias, uv, n_dirs, out_dev);
}

// PTM_FORMAT_LUM (rti-builder/ptmlib.c:601-609): rasters are (Y, Cb, Cr) triplets; crcb_dev holds 2-byte
// (cr, cb) records, the layout the reference's loop reads block 1 with.
int ptmb_relight_lum(ptmb_ctx_t *c, const uint8_t *coef_dev, const uint8_t *crcb_dev, size_t width, size_t height,
                     const float *scale, const int32_t *bias, const float *uv, unsigned n_dirs, uint8_t *out_dev) {
    REQUIRE((coef_dev && crcb_dev) || width * height == 0, "NULL block");
    ptm::RelightArgs a{};
    a.plane[0] = coef_dev;
    a.plane[1] = crcb_dev;
    a.plane[2] = nullptr;
    return relight_common(c, a, true, width, height, scale, bias, uv, n_dirs, out_dev, true);
}

int ptmb_relight_rgb(ptmb_ctx_t *c, const uint8_t *r_dev, const uint8_t *g_dev, const uint8_t *b_dev,
                     size_t width, size_t height, const float *scale, const int32_t *bias, const float *uv,
                     unsigned n_dirs, uint8_t *out_dev) {
    REQUIRE((r_dev && g_dev && b_dev) || width * height == 0, "NULL block");
    ptm::RelightArgs a{};
    a.plane[0] = r_dev;
    a.plane[1] = g_dev;
    a.plane[2] = b_dev;
    return relight_common(c, a, false, width, height, scale, bias, uv, n_dirs, out_dev);
}

int ptmb_selftest_div255(ptmb_ctx_t *c, uint64_t first_bits, uint64_t count, uint64_t *mismatches) {
    REQUIRE(c && mismatches, "NULL argument");
    CK(cudaSetDevice(c->device));
    unsigned long long *dev = nullptr;
    CK(cudaMalloc(&dev, sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(dev, 0, sizeof(unsigned long long), c->stream);
    if (e == cudaSuccess)
        e = ptm::launch_div255_selftest(first_bits, count, dev, c->sm_count, c->stream);
    unsigned long long host = 0;
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(&host, dev, sizeof(host), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    cudaFree(dev);
    if (e != cudaSuccess)
        return fail("ptmb_selftest_div255", cudaGetErrorString(e));
    *mismatches = host;
    return 0;
}

int ptmb_synth(ptmb_ctx_t *c, uint32_t seed, int mode, int sample_type, uint32_t pixel0, size_t pixels,
               unsigned n_lights, const int32_t *lights_q14, void *stack_dev, size_t image_stride) {
    REQUIRE(c && lights_q14 && (stack_dev || pixels == 0), "NULL argument");
    REQUIRE(sample_type == PTMB_U8 || sample_type == PTMB_U16, "synthetic stacks are uint8 or uint16");
    REQUIRE(mode == PTMB_SYNTH_YCBCR || mode == PTMB_SYNTH_RGB, "unknown synth mode");
    CK(cudaSetDevice(c->device));
    if (c->lq_cap < n_lights) {
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaFree(c->lq_dev));
        c->lq_dev = nullptr;
        c->lq_cap = 0;
        CK(cudaMalloc(&c->lq_dev, (size_t) n_lights * 3 * sizeof(int32_t)));
        c->lq_cap = n_lights;
    }
    CK(cudaMemcpyAsync(c->lq_dev, lights_q14, (size_t) n_lights * 3 * sizeof(int32_t), cudaMemcpyHostToDevice,
                       c->stream));
    CK(ptm::launch_synth(seed, mode, sample_type, pixel0, pixels, n_lights, c->lq_dev, stack_dev, image_stride,
                         c->sm_count, c->stream));
    return 0;
}

// ---------------------------------------------------------------------------
// Whole encode from host buffers: row slabs of the stack are copied to one of two
// device staging buffers on a copy stream while the fit kernel works on the other.
// Scratch lives in the context and is reused while the geometry stays the same.
// ---------------------------------------------------------------------------
static void encode_scratch_release(ptmb_ctx *c) {
    EncodeScratch &s = c->enc;
    for (int i = 0; i < 2; ++i) {
        cudaFree(s.stage[i]);
        s.stage[i] = nullptr;
        if (s.copied[i])
            cudaEventDestroy(s.copied[i]);
        if (s.consumed[i])
            cudaEventDestroy(s.consumed[i]);
        s.copied[i] = s.consumed[i] = nullptr;
    }
    cudaFree(s.coeffs);
    cudaFree(s.bytes);
    s.coeffs = nullptr;
    s.bytes = nullptr;
    if (s.copy)
        cudaStreamDestroy(s.copy);
    s.copy = nullptr;
    if (s.back)
        cudaStreamDestroy(s.back);
    s.back = nullptr;
    if (s.fitted)
        cudaEventDestroy(s.fitted);
    s.fitted = nullptr;
    if (s.copy_t0)
        cudaEventDestroy(s.copy_t0);
    if (s.copy_t1)
        cudaEventDestroy(s.copy_t1);
    s.copy_t0 = s.copy_t1 = nullptr;
    s.early_rgb = s.early_rgb_done = nullptr;
    s.stage_bytes = s.coeff_floats = s.out_bytes = 0;
    s.valid = false;
}

int ptmb_encode_host_early_rgb(ptmb_ctx_t *c, uint8_t *rgb_host) {
    REQUIRE(c, "ctx is NULL");
    c->enc.early_rgb = rgb_host;
    return 0;
}

// `host_image_bytes`: distance between two lights' images in the HOST stack (>= one slab image; a row slab of
// a larger image is passed with the pitch of the whole image -- the multi-GPU entry does that).
int ptmb_internal_encode_host_fit(ptmb_ctx_t *c, const void *stack_host, size_t host_image_bytes, int sample_type,
                                  size_t width, size_t height, unsigned n_lights, const float *M, int format_planes,
                                  int32_t *keys_dev) {
    REQUIRE(c && stack_host && M, "NULL argument");
    REQUIRE(format_planes == 1 || format_planes == 3, "format_planes must be 1 (LRGB) or 3 (RGB)");
    const size_t ss = sample_size(sample_type);
    REQUIRE(ss == 1 || ss == 2, "host encode takes uint8 or uint16 stacks");
    const size_t pixels = width * height;
    REQUIRE(pixels > 0, "empty image");
    if (ptmb_set_matrix(c, M, n_lights))
        return 1;
    CK(cudaSetDevice(c->device));

    const size_t image_bytes = host_image_bytes ? host_image_bytes : pixels * 3 * ss;
    REQUIRE(image_bytes >= pixels * 3 * ss, "host image pitch smaller than one image");
    // slab: whole rows, about 8 MiB per light and slab
    size_t slab_rows = (8u << 20) / (width * 3 * ss);
    if (slab_rows < 1)
        slab_rows = 1;
    if (slab_rows > height)
        slab_rows = height;
    // slab pixel counts a multiple of 4 keep every slab on the vector kernel's alignment
    while (slab_rows < height && (slab_rows * width) % 4 != 0)
        ++slab_rows;
    const size_t slab_pixels = slab_rows * width;
    const size_t slab_stride = (slab_pixels * 3 * ss + 255) & ~(size_t) 255;
    const size_t plane_floats = pixels * PTMB_COEFFICIENTS;
    const size_t out_bytes = pixels * PTMB_COEFFICIENTS * format_planes + (format_planes == 1 ? pixels * 3 : 0);

    EncodeScratch &s = c->enc;
    s.valid = false;
    if (!s.back) {
        CK(cudaStreamCreateWithFlags(&s.back, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&s.fitted, cudaEventDisableTiming));
    }
    uint8_t *const early = format_planes == 1 ? s.early_rgb : nullptr;
    s.early_rgb = nullptr;
    s.early_rgb_done = nullptr;
    if (!s.copy) {
        CK(cudaStreamCreateWithFlags(&s.copy, cudaStreamNonBlocking));
        CK(cudaEventCreate(&s.copy_t0));
        CK(cudaEventCreate(&s.copy_t1));
        for (int i = 0; i < 2; ++i) {
            CK(cudaEventCreateWithFlags(&s.copied[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&s.consumed[i], cudaEventDisableTiming));
        }
    }
    if (s.stage_bytes < slab_stride * n_lights) {
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaStreamSynchronize(s.copy));
        for (int i = 0; i < 2; ++i) {
            CK(cudaFree(s.stage[i]));
            s.stage[i] = nullptr;
        }
        s.stage_bytes = 0;
        for (int i = 0; i < 2; ++i)
            CK(cudaMalloc(&s.stage[i], slab_stride * n_lights));
        s.stage_bytes = slab_stride * n_lights;
    }
    if (s.coeff_floats < plane_floats * format_planes) {
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaFree(s.coeffs));
        s.coeffs = nullptr;
        s.coeff_floats = 0;
        CK(cudaMalloc(&s.coeffs, plane_floats * format_planes * sizeof(float)));
        s.coeff_floats = plane_floats * format_planes;
    }
    if (s.out_bytes < out_bytes) {
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaFree(s.bytes));
        s.bytes = nullptr;
        s.out_bytes = 0;
        CK(cudaMalloc(&s.bytes, out_bytes));
        s.out_bytes = out_bytes;
    }
    s.pixels = pixels;
    s.planes = format_planes;
    uint8_t *rgb_dev = s.bytes + pixels * PTMB_COEFFICIENTS * format_planes;

    CK(ptm::launch_keys_init(c->keys_dev, c->stream));
    // pageable operands go through the staging pipe (ptm_hostpipe.h); pinned ones are copied directly
    ptm::HostPipe *pipe = nullptr, *pipe_back = nullptr;
    if (!ptm::HostPipe::is_pinned(stack_host) || (early && !ptm::HostPipe::is_pinned(early))) {
        ptm::HostState *hs = ptm::host_state(c);
        if (hs) {
            pipe = ptm::HostPipe::is_pinned(stack_host) ? nullptr : &hs->pipe;
            pipe_back = (early && !ptm::HostPipe::is_pinned(early)) ? &hs->pipe : nullptr;
        }
    }
    int buf = 0;
    CK(cudaEventRecord(s.copy_t0, s.copy));
    for (size_t row = 0; row < height; row += slab_rows, buf ^= 1) {
        const size_t rows = height - row < slab_rows ? height - row : slab_rows;
        const size_t px0 = row * width, npx = rows * width;
        CK(cudaStreamWaitEvent(s.copy, s.consumed[buf], 0));
        // one strided copy per slab: n_lights rows of npx * 3 * ss bytes, host pitch = one image
        if (pipe)   // pageable stack: staged through the pinned ring by the host thread pool
            CK(pipe->h2d_2d(s.stage[buf], slab_stride, (const uint8_t *) stack_host + px0 * 3 * ss, image_bytes,
                            npx * 3 * ss, n_lights, s.copy));
        else
            CK(cudaMemcpy2DAsync(s.stage[buf], slab_stride, (const uint8_t *) stack_host + px0 * 3 * ss, image_bytes,
                                 npx * 3 * ss, n_lights, cudaMemcpyHostToDevice, s.copy));
        CK(cudaEventRecord(s.copied[buf], s.copy));
        CK(cudaStreamWaitEvent(c->stream, s.copied[buf], 0));
        ptm::FitArgs a;
        if (make_fit_args(c, s.stage[buf], slab_stride, npx, a))
            return 1;
        a.keys = c->keys_dev;
        a.coeffs = s.coeffs + px0 * PTMB_COEFFICIENTS;
        if (format_planes == 1) {
            a.rgb = rgb_dev + px0 * 3;
            a.mode = c->fit_mode;
            REQUIRE(a.mode == 0 || sample_type == PTMB_U8, "PTM_FORMAT_LUM / RGB-input stacks are 8-bit paths");
            CK(ptm::launch_fit_lrgb(a, sample_type, c->sm_count, c->stream));
        } else {
            a.plane_stride = plane_floats;
            CK(ptm::launch_fit_rgb(a, sample_type, c->sm_count, c->stream));
        }
        CK(cudaEventRecord(s.consumed[buf], c->stream));
        if (early) {
            // this slab's RGB bytes are final: send them home while the next slabs arrive
            CK(cudaStreamWaitEvent(s.back, s.consumed[buf], 0));
            if (pipe_back)
                CK(pipe_back->d2h(early + px0 * 3, rgb_dev + px0 * 3, npx * 3, s.back));
            else
                CK(cudaMemcpyAsync(early + px0 * 3, rgb_dev + px0 * 3, npx * 3, cudaMemcpyDeviceToHost, s.back));
        }
    }
    CK(cudaEventRecord(s.copy_t1, s.copy));
    if (early) {
        CK(cudaEventRecord(s.fitted, s.back));
        s.early_rgb_done = early;
    }
    if (keys_dev)
        CK(cudaMemcpyAsync(keys_dev, c->keys_dev, PTMB_KEYS * sizeof(int32_t), cudaMemcpyDeviceToDevice,
                           c->stream));
    s.valid = true;
    return 0;
}

// Milliseconds the input copies of the last ptmb_encode_host_fit took on the copy stream (first to last byte),
// valid once that work has completed; < 0 when unknown.  ptm_multi.cu balances its row slabs with it.
extern "C" float ptmb_internal_last_copy_ms(ptmb_ctx_t *c) {
    float ms = -1.f;
    if (!c || !c->enc.copy_t0 || cudaEventElapsedTime(&ms, c->enc.copy_t0, c->enc.copy_t1) != cudaSuccess) {
        cudaGetLastError();
        return -1.f;
    }
    return ms;
}

int ptmb_encode_host_fit(ptmb_ctx_t *c, const void *stack_host, int sample_type, size_t width, size_t height,
                         unsigned n_lights, const float *M, int format_planes, int32_t *keys_dev) {
    return ptmb_internal_encode_host_fit(c, stack_host, 0, sample_type, width, height, n_lights, M, format_planes,
                                         keys_dev);
}

// `coeffs_plane_floats`: distance between two coefficient planes in coeffs_host (0 = this slab's own size).
int ptmb_internal_encode_host_finish(ptmb_ctx_t *c, const int32_t *keys_dev, uint8_t *const *blocks_host,
                                     float *coeffs_host, size_t coeffs_plane_floats, ptmb_scale_bias_t *sb_host) {
    REQUIRE(c && blocks_host && sb_host, "NULL argument");
    EncodeScratch &s = c->enc;
    REQUIRE(s.valid, "ptmb_encode_host_fit has not been called");
    CK(cudaSetDevice(c->device));
    const int32_t *keys = keys_dev ? keys_dev : c->keys_dev;
    const size_t pixels = s.pixels, plane_floats = pixels * PTMB_COEFFICIENTS;
    const uint8_t *rgb_dev = s.bytes + pixels * PTMB_COEFFICIENTS * s.planes;
    for (int b = s.planes - 1; b >= 0; --b)   // last-written plane first: it is the one still in L2
        CK(ptm::launch_quantise(s.coeffs + b * plane_floats, pixels, ptm::KeySource{keys, nullptr, 0, 0},
                                s.bytes + (size_t) b * pixels * PTMB_COEFFICIENTS, b == 0 ? c->sb_dev : nullptr,
                                coeffs_host == nullptr, c->sm_count, c->stream));
    // device -> host: direct for pinned destinations, through the staging pipe for pageable ones
    ptm::HostState *hs = c->host;
    auto to_host = [&](void *dst, const void *src, size_t bytes) -> cudaError_t {
        if (!ptm::HostPipe::is_pinned(dst) && (hs || (hs = ptm::host_state(c))))
            return hs->pipe.d2h(dst, src, bytes, c->stream);
        return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream);
    };
    for (int b = 0; b < s.planes; ++b)
        CK(to_host(blocks_host[b], s.bytes + (size_t) b * pixels * PTMB_COEFFICIENTS, pixels * PTMB_COEFFICIENTS));
    if (s.planes == 1) {
        if (s.early_rgb_done && s.early_rgb_done == blocks_host[1])
            CK(cudaStreamWaitEvent(c->stream, s.fitted, 0));   // already on its way / there
        else
            CK(to_host(blocks_host[1], rgb_dev, pixels * 3));
    }
    if (coeffs_host)
        for (int b = 0; b < s.planes; ++b)
            CK(to_host(coeffs_host + (size_t) b * (coeffs_plane_floats ? coeffs_plane_floats : plane_floats),
                       s.coeffs + (size_t) b * plane_floats, plane_floats * sizeof(float)));
    CK(cudaMemcpyAsync(sb_host, c->sb_dev, sizeof(ptmb_scale_bias_t), cudaMemcpyDeviceToHost, c->stream));
    // invariant of the context: its own keys are at the initial extrema between calls (the fused device
    // encode merges into them without a keys-init launch of its own)
    CK(ptm::launch_keys_init(c->keys_dev, c->stream));
    if (c->host && c->host->pipe_ok)
        CK(c->host->pipe.finish());   // pageable destinations: the last staged chunks reach the caller's buffers
    if (s.early_rgb_done && s.back)
        CK(cudaStreamSynchronize(s.back));
    CK(cudaStreamSynchronize(c->stream));
    s.valid = false;
    s.early_rgb_done = nullptr;
    return 0;
}

int ptmb_encode_host_finish(ptmb_ctx_t *c, const int32_t *keys_dev, uint8_t *const *blocks_host,
                            float *coeffs_host, ptmb_scale_bias_t *sb_host) {
    return ptmb_internal_encode_host_finish(c, keys_dev, blocks_host, coeffs_host, 0, sb_host);
}

int ptmb_encode_host(ptmb_ctx_t *c, const void *stack_host, int sample_type, size_t width, size_t height,
                     unsigned n_lights, const float *M, int format_planes, uint8_t *const *blocks_host,
                     float *coeffs_host, ptmb_scale_bias_t *sb_host) {
    if (format_planes == 1 && blocks_host && ptmb_encode_host_early_rgb(c, blocks_host[1]))
        return 1;
    if (ptmb_encode_host_fit(c, stack_host, sample_type, width, height, n_lights, M, format_planes, nullptr))
        return 1;
    return ptmb_encode_host_finish(c, nullptr, blocks_host, coeffs_host, sb_host);
}

int ptmb_encode_host_release(ptmb_ctx_t *c) {
    REQUIRE(c, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    encode_scratch_release(c);
    return 0;
}

}  // extern "C"
