// JPEG in and out on the GPU (nvJPEG) -- the codec the reference takes from libjpeg.
//
// The reference decodes its inputs with libjpeg into the stack and writes relit images and the
// planes of the compressed PTM formats with libjpeg (rti-builder/ptm-encoder.c:181-190,248-264;
// rti-builder/ptmlib.c:371-412,473-520,565-627).  Entropy coding is outside the hot path; this file
// exists so that the command lines work on real JPEG files on a B200 box without libjpeg:
//
//   decode  nvjpegDecode.  For the LRGB / LUM formats the reference asks libjpeg for YCbCr
//           (out_color_space = JCS_YCbCr, rti-builder/ptm-encoder.c:188): no colour conversion, chroma
//           upsampled to full resolution.  nvJPEG hands back the planes at their stored resolution
//           (NVJPEG_OUTPUT_YUV); upsample_interleave_kernel restates libjpeg's default "fancy"
//           (triangle filter) upsampling for 4:2:2 and 4:2:0 and interleaves to [h][w][3]; for the
//           RGB formats it also applies libjpeg's integer YCbCr -> RGB.  The inverse DCT is nvJPEG's, so
//           samples can differ from libjpeg's by an LSB -- this is an I/O convenience, not a
//           parity surface; parity tests use the lossless raw container.
//   encode  nvjpegEncodeImage / nvjpegEncodeYUV, baseline, standard Huffman tables, 4:2:0 for
//           colour like libjpeg's defaults, quality from the PTM header.
#include <mutex>
#include <nvjpeg.h>

#include "../../include/ptm_b200.h"
#include "ptm_internal.h"

extern "C" int ptmb_internal_fail(const char *what, const char *detail);

namespace {

struct Jpeg {
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;
    nvjpegEncoderState_t enc_state = nullptr;
    nvjpegEncoderParams_t enc_params = nullptr;
    cudaStream_t enc_stream = nullptr;   // the stream the encoder objects were created against
    bool ok = false;
};

std::mutex g_mu;          // nvJPEG states are not shareable between threads: one call at a time
Jpeg g_jpeg[64];          // per device

const char *status_name(nvjpegStatus_t s) {
    switch (s) {
    case NVJPEG_STATUS_SUCCESS: return "success";
    case NVJPEG_STATUS_NOT_INITIALIZED: return "not initialized";
    case NVJPEG_STATUS_INVALID_PARAMETER: return "invalid parameter";
    case NVJPEG_STATUS_BAD_JPEG: return "bad jpeg";
    case NVJPEG_STATUS_JPEG_NOT_SUPPORTED: return "jpeg not supported";
    case NVJPEG_STATUS_ALLOCATOR_FAILURE: return "allocator failure";
    case NVJPEG_STATUS_EXECUTION_FAILED: return "execution failed";
    case NVJPEG_STATUS_ARCH_MISMATCH: return "arch mismatch";
    case NVJPEG_STATUS_INTERNAL_ERROR: return "internal error";
    case NVJPEG_STATUS_IMPLEMENTATION_NOT_SUPPORTED: return "implementation not supported";
    default: return "unknown nvjpeg status";
    }
}

#define NJ(call)                                                           \
    do {                                                                   \
        nvjpegStatus_t s_ = (call);                                        \
        if (s_ != NVJPEG_STATUS_SUCCESS)                                   \
            return ptmb_internal_fail(#call, status_name(s_));             \
    } while (0)

#define CJ(call)                                                           \
    do {                                                                   \
        cudaError_t e_ = (call);                                           \
        if (e_ != cudaSuccess)                                             \
            return ptmb_internal_fail(#call, cudaGetErrorString(e_));      \
    } while (0)

struct DevMem {
    void *p = nullptr;
    ~DevMem() { cudaFree(p); }
};

int jpeg_for(ptmb_ctx_t *c, Jpeg **out) {
    if (c->device < 0 || c->device >= 64)
        return ptmb_internal_fail("jpeg", "device index out of range");
    Jpeg &j = g_jpeg[c->device];
    if (!j.handle) {
        NJ(nvjpegCreateSimple(&j.handle));
        NJ(nvjpegJpegStateCreate(j.handle, &j.state));
    }
    if (!j.ok || j.enc_stream != c->stream) {
        // the encoder objects are tied to the stream they were created against: a context whose stream
        // changed (ptmb_set_stream), or a second context on this GPU, gets fresh ones
        if (j.ok) {
            CJ(cudaStreamSynchronize(j.enc_stream));
            nvjpegEncoderParamsDestroy(j.enc_params);
            nvjpegEncoderStateDestroy(j.enc_state);
            j.ok = false;
        }
        NJ(nvjpegEncoderStateCreate(j.handle, &j.enc_state, c->stream));
        NJ(nvjpegEncoderParamsCreate(j.handle, &j.enc_params, c->stream));
        j.enc_stream = c->stream;
        j.ok = true;
    }
    *out = &j;
    return 0;
}

// libjpeg's default upsampling (jdsample.c, do_fancy_upsampling = TRUE), restated per output sample.
//   h2v1: even x = 2i   -> i == 0      ? in[0]      : (3 in[i] + in[i-1] + 1) >> 2
//         odd  x = 2i+1 -> i == cw - 1 ? in[cw - 1] : (3 in[i] + in[i+1] + 2) >> 2
//   h2v2: s(j) = 3 near[j] + far[j], near = row y/2, far = the row above (y even) or below (y odd),
//         clamped at the image edge;
//         even x = 2j   -> j == 0      ? (4 s(0) + 8) >> 4 : (3 s(j) + s(j-1) + 8) >> 4
//         odd  x = 2j+1 -> j == cw - 1 ? (4 s(j) + 7) >> 4 : (3 s(j) + s(j+1) + 7) >> 4
__device__ __forceinline__ unsigned up_h2v1(const uint8_t *row, int cw, int x) {
    const int i = x >> 1;
    if ((x & 1) == 0)
        return i == 0 ? row[0] : (3u * row[i] + row[i - 1] + 1u) >> 2;
    return i == cw - 1 ? row[cw - 1] : (3u * row[i] + row[i + 1] + 2u) >> 2;
}

__device__ __forceinline__ unsigned up_h2v2(const uint8_t *plane, size_t pitch, int cw, int ch, int x, int y) {
    const int i = y >> 1;
    int f = (y & 1) ? i + 1 : i - 1;
    f = min(max(f, 0), ch - 1);
    const uint8_t *near = plane + (size_t) i * pitch, *far = plane + (size_t) f * pitch;
    const int j = x >> 1;
    const unsigned s = 3u * near[j] + far[j];
    if ((x & 1) == 0) {
        if (j == 0)
            return (4u * s + 8u) >> 4;
        return (3u * s + (3u * near[j - 1] + far[j - 1]) + 8u) >> 4;
    }
    if (j == cw - 1)
        return (4u * s + 7u) >> 4;
    return (3u * s + (3u * near[j + 1] + far[j + 1]) + 7u) >> 4;
}

// planes (Y full resolution, Cb / Cr at css resolution) -> interleaved [h][w][3]; with flip != 0 output
// row y goes to row h-1-y (the stored order of the reference's stack, rti-builder/ptm-encoder.c:252-256)
__global__ void upsample_interleave_kernel(const uint8_t *py, size_t pitch_y, const uint8_t *pcb, const uint8_t *pcr,
                                           size_t pitch_c, int w, int h, int cw, int ch, int css, int flip,
                                           int to_rgb, uint8_t *out) {
    const size_t n = (size_t) w * h;
    for (size_t o = (size_t) blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (size_t) gridDim.x * blockDim.x) {
        const int y = (int) (o / w), x = (int) (o - (size_t) y * w);
        unsigned cb, cr;
        if (css == NVJPEG_CSS_444) {
            cb = pcb[(size_t) y * pitch_c + x];
            cr = pcr[(size_t) y * pitch_c + x];
        } else if (css == NVJPEG_CSS_422) {
            cb = up_h2v1(pcb + (size_t) y * pitch_c, cw, x);
            cr = up_h2v1(pcr + (size_t) y * pitch_c, cw, x);
        } else {
            cb = up_h2v2(pcb, pitch_c, cw, ch, x, y);
            cr = up_h2v2(pcr, pitch_c, cw, ch, x, y);
        }
        uint8_t *d = out + ((size_t) (flip ? h - 1 - y : y) * w + x) * 3;
        const int yy = py[(size_t) y * pitch_y + x];
        if (!to_rgb) {
            d[0] = (uint8_t) yy;
            d[1] = (uint8_t) cb;
            d[2] = (uint8_t) cr;
        } else {
            // libjpeg's integer YCbCr -> RGB (jdcolor.c tables: FIX(x) = x * 65536 + 0.5, ONE_HALF = 32768)
            const int b0 = (int) cb - 128, r0 = (int) cr - 128;
            const int r = yy + ((91881 * r0 + 32768) >> 16);
            const int g = yy + ((-22554 * b0 + 32768 - 46802 * r0) >> 16);
            const int b = yy + ((116130 * b0 + 32768) >> 16);
            d[0] = (uint8_t) min(max(r, 0), 255);
            d[1] = (uint8_t) min(max(g, 0), 255);
            d[2] = (uint8_t) min(max(b, 0), 255);
        }
    }
}

__global__ void flip_rows_kernel(const uint8_t *in, size_t row_bytes, int h, uint8_t *out) {
    const size_t n = row_bytes * h;
    for (size_t o = (size_t) blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (size_t) gridDim.x * blockDim.x) {
        const size_t y = o / row_bytes, b = o - y * row_bytes;
        out[(size_t) (h - 1 - y) * row_bytes + b] = in[o];
    }
}

__global__ void deinterleave3_kernel(const uint8_t *in, size_t n, uint8_t *p0, uint8_t *p1, uint8_t *p2) {
    for (size_t o = (size_t) blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (size_t) gridDim.x * blockDim.x) {
        p0[o] = in[3 * o];
        p1[o] = in[3 * o + 1];
        p2[o] = in[3 * o + 2];
    }
}

unsigned grid_of(size_t n, int sm) {
    size_t g = (n + 255) / 256, cap = (size_t) sm * 8;
    return (unsigned) (g < 1 ? 1 : (g < cap ? g : cap));
}

}  // namespace

extern "C" {

int ptmb_jpeg_info(ptmb_ctx_t *c, const uint8_t *data, size_t len, unsigned *width, unsigned *height,
                   unsigned *components) {
    if (!c || !data || !width || !height || !components)
        return ptmb_internal_fail("ptmb_jpeg_info", "NULL argument");
    std::lock_guard<std::mutex> lock(g_mu);
    CJ(cudaSetDevice(c->device));
    Jpeg *j;
    if (jpeg_for(c, &j))
        return 1;
    int nc = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t css;
    NJ(nvjpegGetImageInfo(j->handle, data, len, &nc, &css, ws, hs));
    *width = (unsigned) ws[0];
    *height = (unsigned) hs[0];
    *components = nc == 1 ? 1u : 3u;
    return 0;
}

// Decode into an interleaved 8-bit image [height][width][components].  want_rgb = 0: YCbCr as
// libjpeg delivers it with out_color_space = JCS_YCbCr; 1: RGB.  Grayscale files give one component.
// `out` is a HOST pointer (out_is_device = 0) or a DEVICE pointer (1, e.g. one light's plane of the
// stack); flip != 0 stores the image bottom row first.  Synchronous.
int ptmb_jpeg_decode(ptmb_ctx_t *c, const uint8_t *data, size_t len, int want_rgb, int flip, void *out,
                     int out_is_device) {
    if (!c || !data || !out)
        return ptmb_internal_fail("ptmb_jpeg_decode", "NULL argument");
    std::lock_guard<std::mutex> lock(g_mu);
    CJ(cudaSetDevice(c->device));
    Jpeg *j;
    if (jpeg_for(c, &j))
        return 1;
    int nc = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t css;
    NJ(nvjpegGetImageInfo(j->handle, data, len, &nc, &css, ws, hs));
    const int w = ws[0], h = hs[0];
    const size_t pixels = (size_t) w * h;
    const int comps = nc == 1 ? 1 : 3;
    cudaStream_t st = c->stream;
    DevMem result;   // interleaved, final orientation (only when the destination is host memory)
    uint8_t *dst = (uint8_t *) out;
    if (!out_is_device) {
        CJ(cudaMalloc(&result.p, pixels * comps));
        dst = (uint8_t *) result.p;
    }
    nvjpegImage_t img{};
    if (comps == 1) {
        DevMem tmp;
        CJ(cudaMalloc(&tmp.p, pixels));
        img.channel[0] = (unsigned char *) tmp.p;
        img.pitch[0] = (size_t) w;
        NJ(nvjpegDecode(j->handle, j->state, data, len, NVJPEG_OUTPUT_Y, &img, st));
        if (flip)
            flip_rows_kernel<<<grid_of(pixels, c->sm_count), 256, 0, st>>>((const uint8_t *) tmp.p, (size_t) w, h, dst);
        else
            CJ(cudaMemcpyAsync(dst, tmp.p, pixels, cudaMemcpyDeviceToDevice, st));
        CJ(cudaGetLastError());
        CJ(cudaStreamSynchronize(st));
    } else {
        // planes at their stored resolution; upsampling and (for RGB) colour conversion restate libjpeg's
        if (css != NVJPEG_CSS_444 && css != NVJPEG_CSS_422 && css != NVJPEG_CSS_420)
            return ptmb_internal_fail("ptmb_jpeg_decode", "chroma subsampling other than 4:4:4, 4:2:2, 4:2:0");
        DevMem planes[3];
        for (int k = 0; k < 3; ++k) {
            CJ(cudaMalloc(&planes[k].p, (size_t) ws[k] * hs[k]));
            img.channel[k] = (unsigned char *) planes[k].p;
            img.pitch[k] = (size_t) ws[k];
        }
        NJ(nvjpegDecode(j->handle, j->state, data, len, NVJPEG_OUTPUT_YUV, &img, st));
        upsample_interleave_kernel<<<grid_of(pixels, c->sm_count), 256, 0, st>>>(
            (const uint8_t *) planes[0].p, (size_t) ws[0], (const uint8_t *) planes[1].p, (const uint8_t *) planes[2].p,
            (size_t) ws[1], w, h, ws[1], hs[1], (int) css, flip, want_rgb ? 1 : 0, dst);
        CJ(cudaGetLastError());
        CJ(cudaStreamSynchronize(st));
    }
    if (!out_is_device) {
        CJ(cudaMemcpyAsync(out, dst, pixels * comps, cudaMemcpyDeviceToHost, st));
        CJ(cudaStreamSynchronize(st));
    }
    return 0;
}

// Encode an interleaved 8-bit HOST image.  components = 1: grayscale; 3: RGB (input_is_ycbcr = 0) or
// YCbCr (1).  Colour is written 4:2:0 like libjpeg's defaults.  *jpeg_out is malloc'd; free() it.
int ptmb_jpeg_encode(ptmb_ctx_t *c, const uint8_t *pixels_host, unsigned width, unsigned height, unsigned components,
                     int input_is_ycbcr, int quality, uint8_t **jpeg_out, size_t *jpeg_len) {
    if (!c || !pixels_host || !jpeg_out || !jpeg_len || (components != 1 && components != 3) || !width || !height)
        return ptmb_internal_fail("ptmb_jpeg_encode", "bad argument");
    std::lock_guard<std::mutex> lock(g_mu);
    CJ(cudaSetDevice(c->device));
    Jpeg *j;
    if (jpeg_for(c, &j))
        return 1;
    cudaStream_t st = c->stream;
    const size_t pixels = (size_t) width * height;
    DevMem in;
    CJ(cudaMalloc(&in.p, pixels * components));
    CJ(cudaMemcpyAsync(in.p, pixels_host, pixels * components, cudaMemcpyHostToDevice, st));
    NJ(nvjpegEncoderParamsSetQuality(j->enc_params, quality < 1 ? 1 : (quality > 100 ? 100 : quality), st));
    NJ(nvjpegEncoderParamsSetEncoding(j->enc_params, NVJPEG_ENCODING_BASELINE_DCT, st));
    NJ(nvjpegEncoderParamsSetOptimizedHuffman(j->enc_params, 0, st));
    nvjpegImage_t img{};
    DevMem planes;
    if (components == 1) {
        NJ(nvjpegEncoderParamsSetSamplingFactors(j->enc_params, NVJPEG_CSS_GRAY, st));
        img.channel[0] = (unsigned char *) in.p;
        img.pitch[0] = width;
        NJ(nvjpegEncodeYUV(j->handle, j->enc_state, j->enc_params, &img, NVJPEG_CSS_GRAY, (int) width, (int) height, st));
    } else if (!input_is_ycbcr) {
        NJ(nvjpegEncoderParamsSetSamplingFactors(j->enc_params, NVJPEG_CSS_420, st));
        img.channel[0] = (unsigned char *) in.p;
        img.pitch[0] = (size_t) width * 3;
        NJ(nvjpegEncodeImage(j->handle, j->enc_state, j->enc_params, &img, NVJPEG_INPUT_RGBI, (int) width, (int) height, st));
    } else {
        NJ(nvjpegEncoderParamsSetSamplingFactors(j->enc_params, NVJPEG_CSS_420, st));
        CJ(cudaMalloc(&planes.p, pixels * 3));
        uint8_t *p0 = (uint8_t *) planes.p, *p1 = p0 + pixels, *p2 = p1 + pixels;
        deinterleave3_kernel<<<grid_of(pixels, c->sm_count), 256, 0, st>>>((const uint8_t *) in.p, pixels, p0, p1, p2);
        CJ(cudaGetLastError());
        img.channel[0] = p0;
        img.channel[1] = p1;
        img.channel[2] = p2;
        img.pitch[0] = img.pitch[1] = img.pitch[2] = width;
        NJ(nvjpegEncodeYUV(j->handle, j->enc_state, j->enc_params, &img, NVJPEG_CSS_444, (int) width, (int) height, st));
    }
    size_t len = 0;
    NJ(nvjpegEncodeRetrieveBitstream(j->handle, j->enc_state, nullptr, &len, st));
    CJ(cudaStreamSynchronize(st));
    uint8_t *buf = (uint8_t *) malloc(len ? len : 1);
    if (!buf)
        return ptmb_internal_fail("ptmb_jpeg_encode", "out of host memory");
    nvjpegStatus_t s = nvjpegEncodeRetrieveBitstream(j->handle, j->enc_state, buf, &len, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (s != NVJPEG_STATUS_SUCCESS || e != cudaSuccess) {
        free(buf);
        return ptmb_internal_fail("nvjpegEncodeRetrieveBitstream", s != NVJPEG_STATUS_SUCCESS ? status_name(s)
                                                                                               : cudaGetErrorString(e));
    }
    *jpeg_out = buf;
    *jpeg_len = len;
    return 0;
}

}  // extern "C"
