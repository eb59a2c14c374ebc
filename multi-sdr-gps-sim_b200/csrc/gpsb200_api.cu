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

constexpr int kSynthChunk = 256;   // blocks per synthesis launch of the host-destination path (D2H overlap)

struct ChainState {
    int prn = 0;
    double phase = 0.0;
};

}  // namespace

struct gpsb200_ctx {
    gpsb200_config_t cfg{};
    int nruns = 0;
    cudaStream_t s_compute = nullptr, s_copy = nullptr;
    cudaEvent_t ev[8]{};
    std::vector<cudaEvent_t> ev_done;      // one per synthesis chunk
    BlockChanDev *d_bc = nullptr, *h_bc = nullptr;
    RunCkpt *d_ck = nullptr;
    uint32_t *d_nav = nullptr, *h_nav = nullptr;
    uint32_t *d_chips = nullptr;
    double *d_carr_end = nullptr;
    double *d_guess = nullptr, *h_guess = nullptr;     // speculative block-start phases
    double *d_carr0 = nullptr, *h_carr0 = nullptr;     // exact block-start phases
    CarrierProbe *d_probe = nullptr, *h_probe = nullptr;
    void *d_out = nullptr;
    size_t out_bytes = 0;
    bool nav_dirty = true;
    SynthArgs last{};                      // replay state
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

template <class F>
void parallel_channels(int nchan, int nthreads, F &&work) {
    nthreads = std::max(1, std::min(nthreads, nchan));
    if (nthreads == 1) {
        work(0, nchan);
        return;
    }
    std::vector<std::thread> th;
    const int per = (nchan + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        const int lo = t * per, hi = std::min(nchan, lo + per);
        if (lo < hi) th.emplace_back(work, lo, hi);
    }
    for (auto &t : th) t.join();
}

inline bool is_fresh(const gpsb200_chan_t *chans, int b, int c, int nchan) {
    // block 0 of a call, or a slot whose satellite changed: the caller's carr_phase applies
    return b == 0 || chans[(size_t) (b - 1) * nchan + c].prn != chans[(size_t) b * nchan + c].prn;
}

// Host pre-pass: validate, fill the device-layout records and GUESS every block's start
// carrier phase (closed form + expected rounding drift, long double accumulation).
int prepare_blocks(gpsb200_ctx *ctx, const gpsb200_chan_t *chans, int nblk, int nchan) {
    const double delt = 1.0 / (double) GPSB200_SAMPLERATE;     // gps.c:2298
    std::vector<int> status(nchan, GPSB200_OK);
    parallel_channels(nchan, ctx->cfg.host_threads, [&](int c_lo, int c_hi) {
        for (int c = c_lo; c < c_hi; c++) {
            long double acc = 0.0L;
            for (int b = 0; b < nblk; b++) {
                const gpsb200_chan_t &in = chans[(size_t) b * nchan + c];
                const size_t i = (size_t) b * nchan + c;
                BlockChanDev &o = ctx->h_bc[i];
                memset(&o, 0, sizeof o);
                ctx->h_guess[i] = 0.0;
                if (in.prn <= 0) continue;
                if (in.prn > 32 || in.iword < 0 || in.iword >= GPSB200_NAV_WORDS || in.ibit < 0 || in.ibit >= 30 ||
                    in.icode < 0 || in.icode >= 20 || in.nav_frame < 0 || in.nav_frame >= ctx->cfg.max_nav_frames ||
                    !(in.code_phase >= 0.0 && in.code_phase < 1023.0) || !(in.f_code > 0.0) ||
                    !std::isfinite(in.f_carr) || !std::isfinite(in.gain)) {
                    status[c] = GPSB200_ERR_ARG;
                    return;
                }
                if (is_fresh(chans, b, c, nchan)) {
                    if (!(in.carr_phase >= 0.0 && in.carr_phase < 1.0)) {
                        status[c] = GPSB200_ERR_ARG;
                        return;
                    }
                    acc = in.carr_phase;
                }
                o.c_carr = in.f_carr * delt;                    // gps.c:2821
                o.c_code = in.f_code * delt;                    // gps.c:2789
                o.gain = in.gain;
                o.code0 = in.code_phase;
                o.prn = in.prn;
                o.nav0 = (uint32_t) in.iword | ((uint32_t) in.ibit << 8) | ((uint32_t) in.icode << 16);
                o.frame = in.nav_frame;
                double g = (double) acc;
                if (!(g >= 0.0 && g < 1.0)) g = 0.0;
                ctx->h_guess[i] = g;
                acc += (long double) GPSB200_BLOCK_SAMPLES *
                       ((long double) o.c_carr + (long double) carrier_drift_per_step(o.c_carr));
                acc -= floorl(acc);
            }
        }
    });
    for (int c = 0; c < nchan; c++)
        if (status[c] != GPSB200_OK) return fail(ctx, status[c], "invalid channel parameters in slot " + std::to_string(c));
    // The reference stores (short)i_acc (gps.c:2834); the packed I/Q accumulation is
    // exact as long as |acc| stays inside int16, which bounds the sum of amplitudes.
    for (int b = 0; b < nblk; b++) {
        double amp = 0.0;
        for (int c = 0; c < nchan; c++) amp += std::fabs(ctx->h_bc[(size_t) b * nchan + c].gain) * 250.0;
        if (amp > 32767.0) return fail(ctx, GPSB200_ERR_RANGE, "sum of channel amplitudes exceeds int16 range");
    }
    return GPSB200_OK;
}

// Host fix-up scan: exact start phase of every block from the probes, serial over blocks
// per channel, parallel over channels. Returns the number of blocks that needed the
// sequential fallback walk.
int64_t resolve_chain(gpsb200_ctx *ctx, const gpsb200_chan_t *chans, int nblk, int nchan,
                      std::vector<ChainState> &chain) {
    std::vector<int64_t> fallbacks(nchan, 0);
    parallel_channels(nchan, ctx->cfg.host_threads, [&](int c_lo, int c_hi) {
        for (int c = c_lo; c < c_hi; c++) {
            ChainState st;
            for (int b = 0; b < nblk; b++) {
                const gpsb200_chan_t &in = chans[(size_t) b * nchan + c];
                const size_t i = (size_t) b * nchan + c;
                ctx->h_carr0[i] = 0.0;
                if (in.prn <= 0) {
                    st.prn = 0;
                    continue;
                }
                if (is_fresh(chans, b, c, nchan)) st.phase = in.carr_phase;
                st.prn = in.prn;
                ctx->h_carr0[i] = st.phase;
                const double cc = ctx->h_bc[i].c_carr;
                double xe;
                if (carrier_fixup(st.phase, cc, ctx->h_probe[i], xe)) {
                    st.phase = xe;
                } else {
                    int64_t dummy = 0;
                    nco_advance<NCO_CARRIER>(st.phase, cc, GPSB200_BLOCK_SAMPLES, dummy);
                    ++fallbacks[c];
                }
            }
            chain[c] = st;
        }
    });
    int64_t n = 0;
    for (auto f : fallbacks) n += f;
    return n;
}

void fill_args(gpsb200_ctx *ctx, SynthArgs &a, int blk0, int nblk, int nchan, int sample_size, void *out) {
    const size_t off = (size_t) blk0 * nchan;
    a.bc = ctx->d_bc + off;
    a.carr0 = ctx->d_carr0 + off;
    a.guess = ctx->d_guess + off;
    a.probe = ctx->d_probe + off;
    a.ck = ctx->d_ck + off * ctx->nruns;
    a.nav = ctx->d_nav;
    a.chipbits = ctx->d_chips;
    a.carr_end = ctx->d_carr_end + off;
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
    const int ctas = (ctx->nruns + per_cta - 1) / per_cta;
    per_cta = (ctx->nruns + ctas - 1) / ctas;
    a.runs_per_cta = per_cta;
    a.ctas_per_block = ctas;
}

int upload_nav(gpsb200_ctx *ctx, cudaStream_t s) {
    if (!ctx->nav_dirty) return GPSB200_OK;
    const size_t bytes = (size_t) ctx->cfg.max_nav_frames * ctx->cfg.max_chan * GPSB200_NAV_WORDS * 4;
    CU(cudaMemcpyAsync(ctx->d_nav, ctx->h_nav, bytes, cudaMemcpyHostToDevice, s));
    ctx->nav_dirty = false;
    return GPSB200_OK;
}

int check_call(gpsb200_ctx *ctx, const gpsb200_chan_t *chans, int nblk, int nchan, int sample_size, void *dst) {
    if (!ctx) return GPSB200_ERR_ARG;
    if (!chans || !dst || nblk < 1 || nblk > ctx->cfg.max_blocks || nchan != ctx->cfg.max_chan ||
        (sample_size != GPSB200_SC08 && sample_size != GPSB200_SC16))
        return fail(ctx, GPSB200_ERR_ARG, "bad arguments (nchan must equal cfg.max_chan; 1 <= nblk <= cfg.max_blocks)");
    if (!ctx->s_compute) return fail(ctx, GPSB200_ERR_CUDA, "context has no CUDA device");
    return GPSB200_OK;
}

// The whole path for nblk blocks. dst_host != NULL: results are copied to the host chunk by
// chunk while later chunks are still being synthesized; else they stay at dst_dev.
int run_pipeline(gpsb200_ctx *ctx, const gpsb200_chan_t *chans, int nblk, int nchan, int sample_size,
                 void *dst_dev, void *dst_host, cudaStream_t s, double *carr_phase_out, gpsb200_stats_t *stats) {
    gpsb200_stats_t st{};
    const size_t nbc = (size_t) nblk * nchan;
    const size_t blk_bytes = (size_t) GPSB200_BLOCK_ELEMS * sample_size;
    // 1. host pre-pass
    double t0 = now_ms();
    int rc = prepare_blocks(ctx, chans, nblk, nchan);
    if (rc) return rc;
    st.host_chain_ms = now_ms() - t0;
    // 2. parameters up, speculative carrier probe, probes down
    CU(cudaEventRecord(ctx->ev[0], s));
    rc = upload_nav(ctx, s);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->d_bc, ctx->h_bc, nbc * sizeof(BlockChanDev), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(ctx->d_guess, ctx->h_guess, nbc * sizeof(double), cudaMemcpyHostToDevice, s));
    CU(cudaEventRecord(ctx->ev[1], s));
    SynthArgs a{};
    fill_args(ctx, a, 0, nblk, nchan, sample_size, dst_dev);
    CU(launch_probe(a, s));
    CU(cudaEventRecord(ctx->ev[2], s));
    CU(cudaMemcpyAsync(ctx->h_probe, ctx->d_probe, nbc * sizeof(CarrierProbe), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    // 3. exact block-start phases (host, serial over blocks per channel, cheap)
    t0 = now_ms();
    std::vector<ChainState> chain(nchan);
    st.chain_fallbacks = (int32_t) resolve_chain(ctx, chans, nblk, nchan, chain);
    st.host_chain_ms += now_ms() - t0;
    // 4. start phases up, run checkpoints, synthesis (+ overlapped download)
    CU(cudaEventRecord(ctx->ev[3], s));
    CU(cudaMemcpyAsync(ctx->d_carr0, ctx->h_carr0, nbc * sizeof(double), cudaMemcpyHostToDevice, s));
    CU(launch_checkpoints(a, s));
    CU(cudaEventRecord(ctx->ev[4], s));
    st.launches = 2;
    st.h2d_bytes = (int64_t) (nbc * (sizeof(BlockChanDev) + 2 * sizeof(double)));
    st.d2h_bytes = (int64_t) (nbc * sizeof(CarrierProbe));
    if (!dst_host) {
        CU(launch_synth(a, s));
        st.launches += 1;
    } else {
        int ichunk = 0;
        for (int b0 = 0; b0 < nblk; b0 += kSynthChunk, ichunk++) {
            const int nb = std::min(kSynthChunk, nblk - b0);
            SynthArgs ac{};
            char *dout = (char *) dst_dev + (size_t) b0 * blk_bytes;
            fill_args(ctx, ac, b0, nb, nchan, sample_size, dout);
            CU(launch_synth(ac, s));
            st.launches += 1;
            CU(cudaEventRecord(ctx->ev_done[ichunk], s));
            CU(cudaStreamWaitEvent(ctx->s_copy, ctx->ev_done[ichunk], 0));
            CU(cudaMemcpyAsync((char *) dst_host + (size_t) b0 * blk_bytes, dout, (size_t) nb * blk_bytes,
                               cudaMemcpyDeviceToHost, ctx->s_copy));
            st.d2h_bytes += (int64_t) nb * (int64_t) blk_bytes;
        }
    }
    CU(cudaEventRecord(ctx->ev[5], s));
    ctx->last = a;
    ctx->have_last = true;
    if (carr_phase_out)
        for (int c = 0; c < nchan; c++) carr_phase_out[c] = chain[c].prn > 0 ? chain[c].phase : 0.0;
    if (dst_host) {
        CU(cudaStreamSynchronize(s));
        CU(cudaStreamSynchronize(ctx->s_copy));
    }
    if (stats) {
        CU(cudaEventSynchronize(ctx->ev[5]));
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
        st.h2d_ms = ms;
        cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]);
        st.probe_kernel_ms = ms;
        cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]);
        st.checkpoint_kernel_ms = ms;
        cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]);
        st.synth_kernel_ms = ms;
        st.kernel_ms = st.probe_kernel_ms + st.checkpoint_kernel_ms + st.synth_kernel_ms;
        *stats = st;
    }
    return GPSB200_OK;
}

}  // namespace

extern "C" {

const char *gpsb200_version(void) { return "gpsb200 0.2 (sm_100a)"; }

const char *gpsb200_last_error(const gpsb200_ctx_t *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int gpsb200_codegen(int prn, uint8_t ca[GPSB200_CA_LEN]) { return ca_code(prn, ca) == 0 ? GPSB200_OK : GPSB200_ERR_ARG; }

double gpsb200_carrier_advance(double carr_phase, double f_carr, int64_t nsamples) {
    const double c = f_carr * (1.0 / (double) GPSB200_SAMPLERATE);
    int64_t dummy = 0;
    nco_advance<NCO_CARRIER>(carr_phase, c, nsamples, dummy);
    return carr_phase;
}

int gpsb200_carrier_probe_fixup(double start, double guess, double f_carr, int64_t nsamples, double *end_out) {
    const double c = f_carr * (1.0 / (double) GPSB200_SAMPLERATE);
    CarrierProbe p;
    carrier_probe(guess, c, nsamples, p);
    double xe = 0.0;
    const bool ok = carrier_fixup(start, c, p, xe);
    if (ok && end_out) *end_out = xe;
    return ok ? 1 : 0;
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
    if (c.host_threads <= 0) c.host_threads = (int) std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
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
    const int nchunk = (c.max_blocks + kSynthChunk - 1) / kSynthChunk;
    ctx->ev_done.resize(nchunk);
    for (auto &e : ctx->ev_done) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    const size_t nbc = (size_t) c.max_blocks * c.max_chan;
    CU(cudaMalloc(&ctx->d_bc, nbc * sizeof(BlockChanDev)));
    CU(cudaHostAlloc(&ctx->h_bc, nbc * sizeof(BlockChanDev), cudaHostAllocDefault));
    CU(cudaMalloc(&ctx->d_ck, nbc * ctx->nruns * sizeof(RunCkpt)));
    CU(cudaMalloc(&ctx->d_carr_end, nbc * sizeof(double)));
    CU(cudaMalloc(&ctx->d_guess, nbc * sizeof(double)));
    CU(cudaHostAlloc(&ctx->h_guess, nbc * sizeof(double), cudaHostAllocDefault));
    CU(cudaMalloc(&ctx->d_carr0, nbc * sizeof(double)));
    CU(cudaHostAlloc(&ctx->h_carr0, nbc * sizeof(double), cudaHostAllocDefault));
    CU(cudaMalloc(&ctx->d_probe, nbc * sizeof(CarrierProbe)));
    CU(cudaHostAlloc(&ctx->h_probe, nbc * sizeof(CarrierProbe), cudaHostAllocDefault));
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
    cudaFree(ctx->d_guess);
    cudaFreeHost(ctx->h_guess);
    cudaFree(ctx->d_carr0);
    cudaFreeHost(ctx->h_carr0);
    cudaFree(ctx->d_probe);
    cudaFreeHost(ctx->h_probe);
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

int gpsb200_synth_blocks_device(gpsb200_ctx_t *ctx, const gpsb200_chan_t *chans, int nblk, int nchan,
                                int sample_size, void *dst_device, void *stream_, double *carr_phase_out,
                                gpsb200_stats_t *stats) {
    int rc = check_call(ctx, chans, nblk, nchan, sample_size, dst_device);
    if (rc) return rc;
    cudaStream_t s = stream_ ? (cudaStream_t) stream_ : ctx->s_compute;
    return run_pipeline(ctx, chans, nblk, nchan, sample_size, dst_device, nullptr, s, carr_phase_out, stats);
}

int gpsb200_replay_device(gpsb200_ctx_t *ctx, void *dst_device, void *stream_, int kernel_mask) {
    if (!ctx || !ctx->have_last) return GPSB200_ERR_ARG;
    cudaStream_t s = stream_ ? (cudaStream_t) stream_ : ctx->s_compute;
    SynthArgs a = ctx->last;
    if (dst_device) a.out = dst_device;
    if (kernel_mask & 4) CU(launch_probe(a, s));
    if (kernel_mask & 1) CU(launch_checkpoints(a, s));
    if (kernel_mask & 2) CU(launch_synth(a, s));
    return GPSB200_OK;
}

int gpsb200_synth_blocks(gpsb200_ctx_t *ctx, const gpsb200_chan_t *chans, int nblk, int nchan, int sample_size,
                         void *dst, double *carr_phase_out, gpsb200_stats_t *stats) {
    int rc = check_call(ctx, chans, nblk, nchan, sample_size, dst);
    if (rc) return rc;
    const size_t need = (size_t) ctx->cfg.max_blocks * GPSB200_BLOCK_ELEMS * sample_size;
    if (ctx->out_bytes < need) {
        cudaFree(ctx->d_out);
        ctx->d_out = nullptr;
        ctx->out_bytes = 0;
        CU(cudaMalloc(&ctx->d_out, need));
        ctx->out_bytes = need;
    }
    return run_pipeline(ctx, chans, nblk, nchan, sample_size, ctx->d_out, dst, ctx->s_compute, carr_phase_out, stats);
}

}  // extern "C"
