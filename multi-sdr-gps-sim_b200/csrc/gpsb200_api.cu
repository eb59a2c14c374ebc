// C ABI of libgpsb200.so (include/gpsb200.h): context, host-side exact carrier chain,
// parameter upload, kernel launches, result download.
//
// Host work per block and channel is what the reference's 10 Hz path hands to its
// sample loop (gps.c:2731-2765) plus ONE thing the loop carries implicitly: the
// carrier phase at the start of the block, which in the reference is simply
// whatever 300000 sequential FP64 additions left behind (gps.c:2821-2826). Here it
// is produced by the exact fast-forward of nco_exact.h, one host thread per group
// of channels, running ahead of the GPU batch by batch.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gpsb200.h"
#include "nco_exact.h"
#include "synth_kernels.h"
#include "synth_tables.h"

using namespace gpsb200;

namespace {

constexpr int kSubBatch = 128;   // blocks per pipeline stage of the host-destination path

struct ChainState {
    int prn = 0;
    double phase = 0.0;
};

}  // namespace

struct gpsb200_ctx {
    gpsb200_config_t cfg{};
    int nruns = 0, runs_per_cta = 0, ctas_per_block = 0;
    cudaStream_t s_compute = nullptr, s_copy = nullptr;
    cudaEvent_t ev[8]{};
    std::vector<cudaEvent_t> ev_done;      // one per sub-batch
    BlockChanDev *d_bc = nullptr, *h_bc = nullptr;
    RunCkpt *d_ck = nullptr;
    uint32_t *d_nav = nullptr, *h_nav = nullptr;
    uint32_t *d_chips = nullptr;
    double *d_carr_end = nullptr;
    void *d_out = nullptr;
    size_t out_bytes = 0;
    bool nav_dirty = true;
    // replay state
    SynthArgs last{};
    bool have_last = false;
    std::string err;
};

namespace {

int fail(gpsb200_ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg;
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(ctx, GPSB200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// Fill device-layout records for blocks [b0, b1) and advance the per-slot carrier chain.
// Returns GPSB200_OK or an argument/range error.
int prepare_blocks(gpsb200_ctx *ctx, const gpsb200_chan_t *chans, int b0, int b1, int nchan,
                   std::vector<ChainState> &chain, bool first_call_block) {
    const double delt = 1.0 / (double) GPSB200_SAMPLERATE;     // gps.c:2298
    const int nthreads = std::max(1, std::min(ctx->cfg.host_threads, nchan));
    std::vector<int> status(nchan, GPSB200_OK);

    auto work = [&](int c_lo, int c_hi) {
        for (int c = c_lo; c < c_hi; c++) {
            ChainState st = chain[c];
            for (int b = b0; b < b1; b++) {
                const gpsb200_chan_t &in = chans[(size_t) b * nchan + c];
                BlockChanDev &o = ctx->h_bc[(size_t) b * nchan + c];
                memset(&o, 0, sizeof o);
                if (in.prn <= 0) {
                    st.prn = 0;
                    continue;
                }
                if (in.prn > 32 || in.iword < 0 || in.iword >= GPSB200_NAV_WORDS || in.ibit < 0 || in.ibit >= 30 ||
                    in.icode < 0 || in.icode >= 20 || in.nav_frame < 0 || in.nav_frame >= ctx->cfg.max_nav_frames ||
                    !(in.code_phase >= 0.0 && in.code_phase < 1023.0) || !(in.f_code > 0.0) ||
                    !std::isfinite(in.f_carr) || !std::isfinite(in.gain)) {
                    status[c] = GPSB200_ERR_ARG;
                    return;
                }
                const bool fresh = (b == 0 && first_call_block) || st.prn != in.prn;
                if (fresh) {
                    if (!(in.carr_phase >= 0.0 && in.carr_phase < 1.0)) {
                        status[c] = GPSB200_ERR_ARG;
                        return;
                    }
                    st.prn = in.prn;
                    st.phase = in.carr_phase;
                }
                o.c_carr = in.f_carr * delt;                    // gps.c:2821
                o.c_code = in.f_code * delt;                    // gps.c:2789
                o.gain = in.gain;
                o.carr0 = st.phase;
                o.code0 = in.code_phase;
                o.prn = in.prn;
                o.nav0 = (uint32_t) in.iword | ((uint32_t) in.ibit << 8) | ((uint32_t) in.icode << 16);
                o.frame = in.nav_frame;
                int64_t dummy = 0;
                nco_advance<NCO_CARRIER>(st.phase, o.c_carr, GPSB200_BLOCK_SAMPLES, dummy);
            }
            chain[c] = st;
        }
    };
    if (nthreads == 1) {
        work(0, nchan);
    } else {
        std::vector<std::thread> th;
        const int per = (nchan + nthreads - 1) / nthreads;
        for (int t = 0; t < nthreads; t++) {
            const int lo = t * per, hi = std::min(nchan, lo + per);
            if (lo < hi) th.emplace_back(work, lo, hi);
        }
        for (auto &t : th) t.join();
    }
    for (int c = 0; c < nchan; c++)
        if (status[c] != GPSB200_OK) return fail(ctx, status[c], "invalid channel parameters in slot " + std::to_string(c));
    // The reference stores (short)i_acc (gps.c:2834); the packed I/Q accumulation is
    // exact as long as |acc| stays inside int16, which bounds the sum of amplitudes.
    for (int b = b0; b < b1; b++) {
        double amp = 0.0;
        for (int c = 0; c < nchan; c++) amp += std::fabs(ctx->h_bc[(size_t) b * nchan + c].gain) * 250.0;
        if (amp > 32767.0) return fail(ctx, GPSB200_ERR_RANGE, "sum of channel amplitudes exceeds int16 range");
    }
    return GPSB200_OK;
}

void fill_args(gpsb200_ctx *ctx, SynthArgs &a, int blk0, int nblk, int nchan, int sample_size, void *out) {
    a.bc = ctx->d_bc + (size_t) blk0 * nchan;
    a.ck = ctx->d_ck + (size_t) blk0 * ctx->nruns * nchan;
    a.nav = ctx->d_nav;
    a.chipbits = ctx->d_chips;
    a.carr_end = ctx->d_carr_end + (size_t) blk0 * nchan;
    a.out = out;
    a.nblk = nblk;
    a.nchan = nchan;
    a.nruns = ctx->nruns;
    a.run_samples = ctx->cfg.run_samples;
    a.iq16 = sample_size == GPSB200_SC16;
    // lanes per run follow the channel count; a CTA takes up to 24 warps' worth of runs
    const int grp = nchan > 16 ? 32 : (nchan > 8 ? 16 : 8);
    const int rpw = 32 / grp;
    int per_cta = 24 * rpw;
    int ctas = (ctx->nruns + per_cta - 1) / per_cta;
    per_cta = (ctx->nruns + ctas - 1) / ctas;
    a.runs_per_cta = per_cta;
    a.ctas_per_block = ctas;
}

int upload_nav(gpsb200_ctx *ctx, int nchan_stride, cudaStream_t s) {
    (void) nchan_stride;
    if (!ctx->nav_dirty) return GPSB200_OK;
    const size_t bytes = (size_t) ctx->cfg.max_nav_frames * ctx->cfg.max_chan * GPSB200_NAV_WORDS * 4;
    CU(cudaMemcpyAsync(ctx->d_nav, ctx->h_nav, bytes, cudaMemcpyHostToDevice, s));
    ctx->nav_dirty = false;
    return GPSB200_OK;
}

}  // namespace

extern "C" {

const char *gpsb200_version(void) { return "gpsb200 0.1 (sm_100a)"; }

const char *gpsb200_last_error(const gpsb200_ctx_t *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int gpsb200_codegen(int prn, uint8_t ca[GPSB200_CA_LEN]) { return ca_code(prn, ca) == 0 ? GPSB200_OK : GPSB200_ERR_ARG; }

double gpsb200_carrier_advance(double carr_phase, double f_carr, int64_t nsamples) {
    const double c = f_carr * (1.0 / (double) GPSB200_SAMPLERATE);
    int64_t dummy = 0;
    nco_advance<NCO_CARRIER>(carr_phase, c, nsamples, dummy);
    return carr_phase;
}

int gpsb200_carrier_chain(const gpsb200_chan_t *chans, int nblk, int nchan, const double *phase_in,
                          double *phase_out, int threads) {
    if (!chans || !phase_out || nblk < 0 || nchan < 1) return GPSB200_ERR_ARG;
    const double delt = 1.0 / (double) GPSB200_SAMPLERATE;
    auto work = [&](int lo, int hi) {
        for (int c = lo; c < hi; c++) {
            int prn = 0;
            double ph = phase_in ? phase_in[c] : 0.0;
            for (int b = 0; b < nblk; b++) {
                const gpsb200_chan_t &in = chans[(size_t) b * nchan + c];
                if (in.prn <= 0) {
                    prn = 0;
                    continue;
                }
                if ((b == 0 && !phase_in) || in.prn != prn) {
                    if (!(b == 0 && phase_in)) ph = in.carr_phase;
                    prn = in.prn;
                }
                int64_t dummy = 0;
                nco_advance<NCO_CARRIER>(ph, in.f_carr * delt, GPSB200_BLOCK_SAMPLES, dummy);
            }
            phase_out[c] = prn > 0 ? ph : 0.0;
        }
    };
    threads = std::max(1, std::min(threads, nchan));
    std::vector<std::thread> th;
    const int per = (nchan + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        const int lo = t * per, hi = std::min(nchan, lo + per);
        if (lo < hi) th.emplace_back(work, lo, hi);
    }
    for (auto &t : th) t.join();
    return GPSB200_OK;
}

int gpsb200_create(const gpsb200_config_t *cfg, gpsb200_ctx_t **out) {
    if (!cfg || !out) return GPSB200_ERR_ARG;
    gpsb200_ctx *ctx = new gpsb200_ctx();
    ctx->cfg = *cfg;
    gpsb200_config_t &c = ctx->cfg;
    if (c.run_samples == 0) c.run_samples = 2400;
    if (c.host_threads <= 0) c.host_threads = (int) std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (c.max_nav_frames <= 0) c.max_nav_frames = 1;
    if (c.max_chan < 1 || c.max_chan > GPSB200_MAX_CHAN || c.max_blocks < 1 || c.run_samples < 32 ||
        c.run_samples % 32 != 0 || GPSB200_BLOCK_SAMPLES % c.run_samples != 0) {
        delete ctx;
        return GPSB200_ERR_ARG;
    }
    ctx->nruns = GPSB200_BLOCK_SAMPLES / c.run_samples;
    *out = ctx;   // from here on errors are reported through the context
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(ctx, GPSB200_ERR_CUDA, "no CUDA device: gpsb200 has no CPU fallback");
    CU(cudaSetDevice(c.device));
    CU(cudaStreamCreateWithFlags(&ctx->s_compute, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ctx->s_copy, cudaStreamNonBlocking));
    for (auto &e : ctx->ev) CU(cudaEventCreate(&e));
    const int nsub = (c.max_blocks + kSubBatch - 1) / kSubBatch;
    ctx->ev_done.resize(nsub);
    for (auto &e : ctx->ev_done) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    const size_t nbc = (size_t) c.max_blocks * c.max_chan;
    CU(cudaMalloc(&ctx->d_bc, nbc * sizeof(BlockChanDev)));
    CU(cudaHostAlloc(&ctx->h_bc, nbc * sizeof(BlockChanDev), cudaHostAllocDefault));
    CU(cudaMalloc(&ctx->d_ck, nbc * ctx->nruns * sizeof(RunCkpt)));
    CU(cudaMalloc(&ctx->d_carr_end, nbc * sizeof(double)));
    const size_t navb = (size_t) c.max_nav_frames * c.max_chan * GPSB200_NAV_WORDS * 4;
    CU(cudaMalloc(&ctx->d_nav, navb));
    CU(cudaHostAlloc(&ctx->h_nav, navb, cudaHostAllocDefault));
    memset(ctx->h_nav, 0, navb);
    // packed C/A chips (ca[], gps.c:2817), periodically extended so that any 32-chip window
    // starting at chip 0..1022 is two consecutive words; row = prn
    std::vector<uint32_t> chips((size_t) 33 * kChipWords, 0);
    for (int prn = 1; prn <= 32; prn++) {
        uint8_t ca[GPSB200_CA_LEN];
        ca_code(prn, ca);
        for (int n = 0; n < kChipWords * 32; n++)
            if (ca[n % GPSB200_CA_LEN]) chips[(size_t) prn * kChipWords + (n >> 5)] |= 1u << (n & 31);
    }
    CU(cudaMalloc(&ctx->d_chips, chips.size() * 4));
    CU(cudaMemcpy(ctx->d_chips, chips.data(), chips.size() * 4, cudaMemcpyHostToDevice));
    return GPSB200_OK;
}

void gpsb200_destroy(gpsb200_ctx_t *ctx) {
    if (!ctx) return;
    if (ctx->s_compute) cudaStreamSynchronize(ctx->s_compute);
    if (ctx->s_copy) cudaStreamSynchronize(ctx->s_copy);
    cudaFree(ctx->d_bc);
    cudaFreeHost(ctx->h_bc);
    cudaFree(ctx->d_ck);
    cudaFree(ctx->d_carr_end);
    cudaFree(ctx->d_nav);
    cudaFreeHost(ctx->h_nav);
    cudaFree(ctx->d_chips);
    cudaFree(ctx->d_out);
    for (auto &e : ctx->ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : ctx->ev_done)
        if (e) cudaEventDestroy(e);
    if (ctx->s_compute) cudaStreamDestroy(ctx->s_compute);
    if (ctx->s_copy) cudaStreamDestroy(ctx->s_copy);
    delete ctx;
}

int gpsb200_set_nav(gpsb200_ctx_t *ctx, int frame, int chan, const uint32_t dwrd[GPSB200_NAV_WORDS]) {
    if (!ctx || !dwrd) return GPSB200_ERR_ARG;
    if (frame < 0 || frame >= ctx->cfg.max_nav_frames || chan < 0 || chan >= ctx->cfg.max_chan)
        return fail(ctx, GPSB200_ERR_ARG, "gpsb200_set_nav: frame/channel out of range");
    memcpy(ctx->h_nav + ((size_t) frame * ctx->cfg.max_chan + chan) * GPSB200_NAV_WORDS, dwrd, GPSB200_NAV_WORDS * 4);
    ctx->nav_dirty = true;
    return GPSB200_OK;
}

static int check_call(gpsb200_ctx *ctx, const gpsb200_chan_t *chans, int nblk, int nchan, int sample_size, void *dst) {
    if (!ctx) return GPSB200_ERR_ARG;
    if (!chans || !dst || nblk < 1 || nblk > ctx->cfg.max_blocks || nchan != ctx->cfg.max_chan ||
        (sample_size != GPSB200_SC08 && sample_size != GPSB200_SC16))
        return fail(ctx, GPSB200_ERR_ARG, "bad arguments (nchan must equal cfg.max_chan; 1 <= nblk <= cfg.max_blocks)");
    if (!ctx->s_compute) return fail(ctx, GPSB200_ERR_CUDA, "context has no CUDA device");
    return GPSB200_OK;
}

int gpsb200_synth_blocks_device(gpsb200_ctx_t *ctx, const gpsb200_chan_t *chans, int nblk, int nchan,
                                int sample_size, void *dst_device, void *stream_, double *carr_phase_out,
                                gpsb200_stats_t *stats) {
    int rc = check_call(ctx, chans, nblk, nchan, sample_size, dst_device);
    if (rc) return rc;
    cudaStream_t s = stream_ ? (cudaStream_t) stream_ : ctx->s_compute;
    gpsb200_stats_t st{};
    const double t0 = now_ms();
    std::vector<ChainState> chain(nchan);
    rc = prepare_blocks(ctx, chans, 0, nblk, nchan, chain, true);
    if (rc) return rc;
    st.host_chain_ms = now_ms() - t0;
    const size_t pbytes = (size_t) nblk * nchan * sizeof(BlockChanDev);
    CU(cudaEventRecord(ctx->ev[0], s));
    rc = upload_nav(ctx, nchan, s);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->d_bc, ctx->h_bc, pbytes, cudaMemcpyHostToDevice, s));
    CU(cudaEventRecord(ctx->ev[1], s));
    SynthArgs a{};
    fill_args(ctx, a, 0, nblk, nchan, sample_size, dst_device);
    CU(launch_checkpoints(a, s));
    CU(cudaEventRecord(ctx->ev[2], s));
    CU(launch_synth(a, s));
    CU(cudaEventRecord(ctx->ev[3], s));
    ctx->last = a;
    ctx->have_last = true;
    if (carr_phase_out)
        for (int c = 0; c < nchan; c++) carr_phase_out[c] = chain[c].prn > 0 ? chain[c].phase : 0.0;
    if (stats) {
        // statistics need the timings, so this variant synchronizes when asked for them
        CU(cudaEventSynchronize(ctx->ev[3]));
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
        st.h2d_ms = ms;
        cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]);
        st.checkpoint_kernel_ms = ms;
        cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
        st.synth_kernel_ms = ms;
        st.kernel_ms = st.checkpoint_kernel_ms + st.synth_kernel_ms;
        st.h2d_bytes = (int64_t) pbytes;
        st.launches = 2;
        *stats = st;
    }
    return GPSB200_OK;
}

int gpsb200_replay_device(gpsb200_ctx_t *ctx, void *dst_device, void *stream_, int kernel_mask) {
    if (!ctx || !ctx->have_last) return GPSB200_ERR_ARG;
    cudaStream_t s = stream_ ? (cudaStream_t) stream_ : ctx->s_compute;
    SynthArgs a = ctx->last;
    if (dst_device) a.out = dst_device;
    if (kernel_mask & 1) CU(launch_checkpoints(a, s));
    if (kernel_mask & 2) CU(launch_synth(a, s));
    return GPSB200_OK;
}

int gpsb200_synth_blocks(gpsb200_ctx_t *ctx, const gpsb200_chan_t *chans, int nblk, int nchan, int sample_size,
                         void *dst, double *carr_phase_out, gpsb200_stats_t *stats) {
    int rc = check_call(ctx, chans, nblk, nchan, sample_size, dst);
    if (rc) return rc;
    const size_t blk_bytes = (size_t) GPSB200_BLOCK_ELEMS * sample_size;
    const size_t need = (size_t) ctx->cfg.max_blocks * blk_bytes;
    if (ctx->out_bytes < need) {
        cudaFree(ctx->d_out);
        ctx->d_out = nullptr;
        ctx->out_bytes = 0;
        CU(cudaMalloc(&ctx->d_out, need));
        ctx->out_bytes = need;
    }
    gpsb200_stats_t st{};
    std::vector<ChainState> chain(nchan);
    rc = upload_nav(ctx, nchan, ctx->s_compute);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev[0], ctx->s_compute));
    // Pipeline over sub-batches: the host computes the carrier chain of sub-batch i+1
    // while the GPU synthesizes sub-batch i and the copy stream drains sub-batch i-1.
    int isub = 0;
    for (int b0 = 0; b0 < nblk; b0 += kSubBatch, isub++) {
        const int b1 = std::min(nblk, b0 + kSubBatch), nb = b1 - b0;
        const double t0 = now_ms();
        rc = prepare_blocks(ctx, chans, b0, b1, nchan, chain, b0 == 0);
        if (rc) {
            cudaStreamSynchronize(ctx->s_compute);
            cudaStreamSynchronize(ctx->s_copy);
            return rc;
        }
        st.host_chain_ms += now_ms() - t0;
        const size_t off = (size_t) b0 * nchan;
        const size_t pbytes = (size_t) nb * nchan * sizeof(BlockChanDev);
        CU(cudaMemcpyAsync(ctx->d_bc + off, ctx->h_bc + off, pbytes, cudaMemcpyHostToDevice, ctx->s_compute));
        st.h2d_bytes += (int64_t) pbytes;
        SynthArgs a{};
        char *dout = (char *) ctx->d_out + (size_t) b0 * blk_bytes;
        fill_args(ctx, a, b0, nb, nchan, sample_size, dout);
        CU(launch_checkpoints(a, ctx->s_compute));
        CU(launch_synth(a, ctx->s_compute));
        st.launches += 2;
        CU(cudaEventRecord(ctx->ev_done[isub], ctx->s_compute));
        CU(cudaStreamWaitEvent(ctx->s_copy, ctx->ev_done[isub], 0));
        CU(cudaMemcpyAsync((char *) dst + (size_t) b0 * blk_bytes, dout, (size_t) nb * blk_bytes,
                           cudaMemcpyDeviceToHost, ctx->s_copy));
        st.d2h_bytes += (int64_t) nb * (int64_t) blk_bytes;
    }
    CU(cudaEventRecord(ctx->ev[1], ctx->s_compute));
    CU(cudaStreamSynchronize(ctx->s_compute));
    CU(cudaStreamSynchronize(ctx->s_copy));
    ctx->have_last = false;
    if (carr_phase_out)
        for (int c = 0; c < nchan; c++) carr_phase_out[c] = chain[c].prn > 0 ? chain[c].phase : 0.0;
    if (stats) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
        st.kernel_ms = ms;   // span of the compute stream (uploads + both kernels of all sub-batches)
        *stats = st;
    }
    return GPSB200_OK;
}

}  // extern "C"
