// Internal: context structures shared by the API translation units.
#pragma once

#include <vector>

#include "../../include/ptm_b200.h"
#include "ptm_kernels.h"

struct EncodeScratch {
    void *stage[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    float *coeffs = nullptr;
    size_t coeff_floats = 0;
    uint8_t *bytes = nullptr;  // planes * pixels * 6 (+ pixels * 3 for LRGB)
    size_t out_bytes = 0;
    cudaStream_t copy = nullptr;
    cudaStream_t back = nullptr;          // device -> host copies that overlap the fit (early RGB block)
    cudaEvent_t fitted = nullptr;
    cudaEvent_t copy_t0 = nullptr, copy_t1 = nullptr;   // first / last host -> device copy of the last fit (timed)
    uint8_t *early_rgb = nullptr;         // host destination set by ptmb_encode_host_early_rgb, consumed by the next fit
    uint8_t *early_rgb_done = nullptr;    // what the last fit already copied back
    cudaEvent_t copied[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
    size_t pixels = 0;
    int planes = 0;
    bool valid = false;
};

namespace ptm {
struct HostState;   // ptm_hostpipe.h: staging pipe + stack mirror of the step-by-step host entries
}

struct ptmb_ctx {
    EncodeScratch enc;
    ptm::HostState *host = nullptr;
    int host_threads = 0;                    // staging threads of the host entries (0 = all hardware threads)
    // key hand-off between K1 and K2 of the fused device encode
    unsigned *ticket_dev = nullptr;          // last-block ticket counter (0 between launches)
    int *mailbox_dev = nullptr;              // local mailbox used when no peer exchange is attached
    unsigned epoch = 0;                      // epoch of the local mailbox
    // optional timing probes around K1 (ptmb_probe_*)
    std::vector<cudaEvent_t> probe_events;
    size_t probe_next = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    std::vector<float> M_host;
    float *M_dev = nullptr;
    unsigned n_lights = 0;
    unsigned magic = 0;
    // tensor-core engine: B operand image of the current matrix (nullptr when not representable)
    std::vector<int8_t> digits_host;
    int8_t *digits_dev = nullptr;
    size_t digits_cap = 0;
    bool digits_ok = false;
    float digit_scale[6] = {0, 0, 0, 0, 0, 0};
    unsigned kpad = 0;
    int engine = 0;
    int relight_mode = 0;                    // PTMB_RELIGHT_*
    int fit_mode = 0;                        // PTMB_FIT_* flags of the LRGB-shaped entries
    int stack_stable = 0;                    // ptmb_set_stack_stable
    int coeffs_scratch = 0;                  // ptmb_set_coeffs_scratch
    int *keys_dev = nullptr;                 // scratch for the host-buffer entry points
    ptmb_scale_bias_t *sb_dev = nullptr;
    std::vector<float> light_host;
    float *light_dev = nullptr;
    size_t light_cap = 0;
    int32_t *lq_dev = nullptr;
    size_t lq_cap = 0;
};


struct ptmb_xchg {
    int device = 0;
    int rank = 0, world = 1;
    int *local = nullptr;
    int *mailbox[ptm::kMaxRanks] = {};
    bool opened[ptm::kMaxRanks] = {};
    unsigned epoch = 0;
};
