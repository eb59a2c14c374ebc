// C ABI of libgpsb200.so (include/gpsb200.h): context, pipeline, host share of the carrier chain.
//
// (k_tables writes the per-block carrier tables k_synth fetches by TMA; it depends on the parameters only.)
// Host work per block and channel is what the reference's 10 Hz path hands to its sample loop
// (gps.c:2731-2765) plus ONE thing the loop carries implicitly: the carrier phase at the start
// of the block, which in the reference is whatever 300000 sequential FP64 additions left behind
// (gps.c:2821-2826). That chain is resolved parallel in time (nco_exact.h): the host GUESSES
// every block's start phase, k_probe walks every block from its guess on the GPU, and a cheap
// sequential host scan (carrier_fixup) turns the probes into exact start phases -- falling back to
// the exact sequential walk for the rare block whose probe cannot be used. Then k_checkpoints and
// k_synth run, in chunks whose download overlaps the synthesis of later chunks.
#include <cuda_runtime.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
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
constexpr int kSegFirst = 256;     // first carrier-chain segment of the host-destination path: small, so
                                   // that the download can start early ...
constexpr int kSegBlocks = 1024;   // ... later ones larger (their probe kernels are latency bound)

struct ChainState {
    int prn = 0;
    double phase = 0.0;
};

// Small persistent worker pool: the per-segment host passes are sub-millisecond, so thread
// creation per pass would dominate them.
class WorkerPool {
public:
    explicit WorkerPool(int n) {
        for (int i = 0; i < n; i++) th_.emplace_back([this, i] { loop(i); });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    int size() const { return (int) th_.size(); }
    // run job(lo, hi) over [0, n) split into at most size() contiguous ranges; blocks until done
    void run(int n, const std::function<void(int, int)> &job) {
        const int parts = std::max(1, std::min(size(), n));
        if (parts <= 1 || th_.empty()) {
            job(0, n);
            return;
        }
        std::unique_lock<std::mutex> lk(mu_);
        job_ = &job;
        n_ = n;
        parts_ = parts;
        pending_ = parts;
        ++epoch_;
        cv_.notify_all();
        done_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }

private:
    void loop(int id) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int, int)> *job;
            int n, parts;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || epoch_ != seen; });
                if (stop_) return;
                seen = epoch_;
                job = job_;
                n = n_;
                parts = parts_;
            }
            if (id < parts) {
                const int per = (n + parts - 1) / parts, lo = id * per, hi = std::min(n, lo + per);
                if (lo < hi) (*job)(lo, hi);
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int, int)> *job_ = nullptr;
    int n_ = 0, parts_ = 0, pending_ = 0;
    uint64_t epoch_ = 0;
    bool stop_ = false;
};

}  // namespace

struct gpsb200_ctx {
    gpsb200_config_t cfg{};
    int nruns = 0;
    int units = 1, unit_samples = GPSB200_BLOCK_SAMPLES;   // carrier-chain units per block
    cudaStream_t s_compute = nullptr, s_copy = nullptr, s_pre = nullptr;
    cudaEvent_t ev[8]{};
    std::vector<cudaEvent_t> ev_done;      // one per synthesis chunk
    BlockChanDev *d_bc = nullptr, *h_bc = nullptr;
    RunCkpt *d_ck = nullptr;
    uint32_t *d_nav = nullptr, *h_nav = nullptr;
    uint32_t *d_chips = nullptr;
    int32_t *d_atab = nullptr;             // per-block carrier tables (k_tables -> k_synth)
    double *d_carr_end = nullptr;
    int *d_chain_errors = nullptr, *h_chain_errors = nullptr;   // device self-check of the carrier chain
    double *d_guess = nullptr, *h_guess = nullptr;     // speculative block-start phases
    double *d_carr0 = nullptr, *h_carr0 = nullptr;     // exact block-start phases
    CarrierProbe *d_probe = nullptr, *h_probe = nullptr;
    void *d_out = nullptr;
    size_t out_bytes = 0;
    bool nav_dirty = true;
    bool graded_chunks = true;             // GPSB200_GRADED_CHUNKS=0: uniform 256-block chunks (A/B knob)
    std::unique_ptr<WorkerPool> pool;      // host passes (guesses, fix-up scan)
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

// Host pre-pass: validate, fill the device-layout records and GUESS every block's start
// carrier phase (closed form + expected rounding drift, long double accumulation).
int prepare_blocks(gpsb200_ctx *ctx, const gpsb200_chan_t *chans, int b0, int b1, int nchan,
                   const std::vector<ChainState> &chain) {
    const double delt = 1.0 / (double) GPSB200_SAMPLERATE;     // gps.c:2298
    std::vector<int> status(nchan, GPSB200_OK);
    ctx->pool->run(nchan, [&](int c_lo, int c_hi) {
        for (int c = c_lo; c < c_hi; c++) {
            long double acc = chain[c].phase;               // exact phase after block b0-1 (if any)
            int prev_prn = chain[c].prn;                    // 0 at the start of a call: block 0 is "fresh"
            for (int b = b0; b < b1; b++) {
                const gpsb200_chan_t &in = chans[(size_t) b * nchan + c];
                const size_t i = (size_t) b * nchan + c;
                BlockChanDev &o = ctx->h_bc[i];
                memset(&o, 0, sizeof o);
                for (int u = 0; u < ctx->units; u++) ctx->h_guess[((size_t) b * ctx->units + u) * nchan + c] = 0.0;
                if (in.prn <= 0) {
                    prev_prn = 0;
                    continue;
                }
                if (in.prn > 32 || in.iword < 0 || in.iword >= GPSB200_NAV_WORDS || in.ibit < 0 || in.ibit >= 30 ||
                    in.icode < 0 || in.icode >= 20 || in.nav_frame < 0 || in.nav_frame >= ctx->cfg.max_nav_frames ||
                    !(in.code_phase >= 0.0 && in.code_phase < 1023.0) ||
                    // the per-lane chip window holds 24 chips per 64 samples: f_code <= 1.07 MHz (GPS: 1.023 MHz +- 4 Hz)
                    !(in.f_code > 0.0 && in.f_code <= 1.07e6) ||
                    // one wrap per step keeps the phase in [0,1) only for |f_carr * delt| < 1
                    !(std::isfinite(in.f_carr) && std::fabs(in.f_carr) < 2.9e6) || !std::isfinite(in.gain)) {
                    status[c] = GPSB200_ERR_ARG;
                    return;
                }
                // a slot whose satellite changed (or the first block of a call): the caller's carr_phase applies
                if (in.prn != prev_prn) {
                    if (!(in.carr_phase >= 0.0 && in.carr_phase < 1.0)) {
                        status[c] = GPSB200_ERR_ARG;
                        return;
                    }
                    acc = in.carr_phase;
                }
                prev_prn = in.prn;
                o.c_carr = in.f_carr * delt;                    // gps.c:2821
                o.c_code = in.f_code * delt;                    // gps.c:2789
                o.gain = in.gain;
                o.code0 = in.code_phase;
                o.prn = in.prn;
                o.nav0 = (uint32_t) in.iword | ((uint32_t) in.ibit << 8) | ((uint32_t) in.icode << 16);
                o.frame = in.nav_frame;
                const long double per_unit = (long double) ctx->unit_samples *
                                             ((long double) o.c_carr + (long double) carrier_drift_per_step(o.c_carr));
                for (int u = 0; u < ctx->units; u++) {
                    double g = (double) acc;
                    if (!(g >= 0.0 && g < 1.0)) g = 0.0;
                    ctx->h_guess[((size_t) b * ctx->units + u) * nchan + c] = g;
                    acc += per_unit;
                    acc -= floorl(acc);
                }
            }
        }
    });
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

// Host fix-up scan: exact start phase of every block from the probes, serial over blocks
// per channel, parallel over channels. Returns the number of blocks that needed the
// sequential fallback walk.
int64_t resolve_chain(gpsb200_ctx *ctx, const gpsb200_chan_t *chans, int b0, int b1, int nchan,
                      std::vector<ChainState> &chain) {
    std::vector<int64_t> fallbacks(nchan, 0);
    ctx->pool->run(nchan, [&](int c_lo, int c_hi) {
        for (int c = c_lo; c < c_hi; c++) {
            ChainState st = chain[c];
            for (int b = b0; b < b1; b++) {
                const gpsb200_chan_t &in = chans[(size_t) b * nchan + c];
                const size_t i = (size_t) b * nchan + c;
                for (int u = 0; u < ctx->units; u++) ctx->h_carr0[((size_t) b * ctx->units + u) * nchan + c] = 0.0;
                if (in.prn <= 0) {
                    st.prn = 0;
                    continue;
                }
                if (st.prn != in.prn) st.phase = in.carr_phase;
                st.prn = in.prn;
                const double cc = ctx->h_bc[i].c_carr;
                for (int u = 0; u < ctx->units; u++) {
                    const size_t iu = ((size_t) b * ctx->units + u) * nchan + c;
                    ctx->h_carr0[iu] = st.phase;
                    double xe;
                    if (carrier_fixup(st.phase, cc, ctx->h_probe[iu], xe)) {
                        st.phase = xe;
                    } else {
                        int64_t dummy = 0;
                        nco_advance<NCO_CARRIER>(st.phase, cc, ctx->unit_samples, dummy);
                        ++fallbacks[c];
                    }
                }
            }
            chain[c] = st;
        }
    });
    // fault injection for tests/test_gpu_parity.py::test_chain_self_check_catches_corruption: corrupt one
    // resolved start phase by one ulp; the device self-check in k_checkpoints must notice
    if (getenv("GPSB200_FAULT_INJECT_CHAIN") && b1 - b0 > 6 && ctx->h_bc[(size_t) (b0 + 5) * nchan].prn > 0 &&
        ctx->h_bc[(size_t) (b0 + 4) * nchan].prn == ctx->h_bc[(size_t) (b0 + 5) * nchan].prn) {
        double &v = ctx->h_carr0[(size_t) (b0 + 5) * ctx->units * nchan];
        v = bits_f64(f64_bits(v) ^ 1ull);
    }
    int64_t n = 0;
    for (auto f : fallbacks) n += f;
    return n;
}

void fill_args(gpsb200_ctx *ctx, SynthArgs &a, int blk0, int nblk, int nchan, int sample_size, void *out) {
    const size_t off = (size_t) blk0 * nchan;
    a.bc = ctx->d_bc + off;
    a.carr0 = ctx->d_carr0 + off * ctx->units;
    a.guess = ctx->d_guess + off * ctx->units;
    a.probe = ctx->d_probe + off * ctx->units;
    a.units = ctx->units;
    a.unit_samples = ctx->unit_samples;
    a.ck = ctx->d_ck + off * ctx->nruns;
    a.nav = ctx->d_nav;
    a.chipbits = ctx->d_chips;
    a.atab = ctx->d_atab + (size_t) blk0 * kAtabRows * 32;
    a.carr_end = ctx->d_carr_end + off;
    a.chain_errors = ctx->d_chain_errors;
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
    CU(cudaSetDevice(ctx->cfg.device));     // the caller may be a thread that never selected the context's device
    return GPSB200_OK;
}

// The whole path for nblk blocks. dst_host != NULL: the call is cut into segments; the
// carrier-chain resolution of a segment (parameters up -> probe -> probes down -> host fix-up
// -> start phases up -> run checkpoints) runs on its own stream and therefore CONCURRENTLY
// with the synthesis kernels (the probe/checkpoint kernels are latency bound and fit beside
// k_synth's CTAs) and the download of earlier segments; results are copied to the host chunk
// by chunk. Else one segment on the caller's stream, results stay at dst_dev.
int run_pipeline(gpsb200_ctx *ctx, const gpsb200_chan_t *chans, int nblk, int nchan, int sample_size,
                 void *dst_dev, void *dst_host, cudaStream_t s, double *carr_phase_out, gpsb200_stats_t *stats) {
    gpsb200_stats_t st{};
    const size_t blk_bytes = (size_t) GPSB200_BLOCK_ELEMS * sample_size;
    std::vector<ChainState> chain(nchan);
    int seg_blocks = dst_host ? kSegFirst : nblk;
    cudaStream_t sp = dst_host ? ctx->s_pre : s;        // stream of the pre-phase
    int rc = upload_nav(ctx, sp);
    if (rc) return rc;
    CU(cudaMemsetAsync(ctx->d_chain_errors, 0, sizeof(int), sp));
    CU(cudaEventRecord(ctx->ev[0], s));
    int ichunk = 0;
    for (int b0 = 0, b1 = 0; b0 < nblk; b0 = b1, seg_blocks = kSegBlocks) {
        b1 = std::min(nblk, b0 + seg_blocks);
        const int nb = b1 - b0;
        const size_t off = (size_t) b0 * nchan, cnt = (size_t) nb * nchan;
        // 1. host pre-pass: device records + guessed start phases
        double t0 = now_ms();
        rc = prepare_blocks(ctx, chans, b0, b1, nchan, chain);
        if (rc) {
            cudaStreamSynchronize(s);
            cudaStreamSynchronize(sp);
            cudaStreamSynchronize(ctx->s_copy);
            return rc;
        }
        st.host_chain_ms += now_ms() - t0;
        // 2. parameters up, speculative carrier probe, probes down
        CU(cudaMemcpyAsync(ctx->d_bc + off, ctx->h_bc + off, cnt * sizeof(BlockChanDev), cudaMemcpyHostToDevice, sp));
        const size_t offu = off * ctx->units, cntu = cnt * ctx->units;
        CU(cudaMemcpyAsync(ctx->d_guess + offu, ctx->h_guess + offu, cntu * sizeof(double), cudaMemcpyHostToDevice, sp));
        SynthArgs a{};
        fill_args(ctx, a, b0, nb, nchan, sample_size, (char *) dst_dev + (size_t) b0 * blk_bytes);
        CU(launch_tables(a, sp));                        // needs only the parameters: off the chain's critical path
        if (b0 == 0) CU(cudaEventRecord(ctx->ev[1], sp));
        CU(launch_probe(a, sp));
        if (b0 == 0) CU(cudaEventRecord(ctx->ev[2], sp));
        CU(cudaStreamSynchronize(sp));                   // probes are in (mapped) host memory now
        // 3. exact block-start phases (host, serial over blocks per channel, cheap)
        t0 = now_ms();
        st.chain_fallbacks += (int32_t) resolve_chain(ctx, chans, b0, b1, nchan, chain);
        st.host_chain_ms += now_ms() - t0;
        // 4. start phases up, run checkpoints, synthesis (+ overlapped download)
        if (b0 == 0) CU(cudaEventRecord(ctx->ev[3], sp));
        CU(cudaMemcpyAsync(ctx->d_carr0 + offu, ctx->h_carr0 + offu, cntu * sizeof(double), cudaMemcpyHostToDevice, sp));
        CU(launch_checkpoints(a, sp));
        if (b0 == 0) CU(cudaEventRecord(ctx->ev[4], sp));
        if (sp != s) {                                   // synthesis of this segment waits for its checkpoints
            CU(cudaEventRecord(ctx->ev_done[ichunk], sp));
            CU(cudaStreamWaitEvent(s, ctx->ev_done[ichunk], 0));
            ichunk++;
        }
        st.launches += 3;
        st.h2d_bytes += (int64_t) (cnt * sizeof(BlockChanDev) + cntu * 2 * sizeof(double));
        st.d2h_bytes += (int64_t) (cntu * sizeof(CarrierProbe));
        if (!dst_host) {
            // self-check result of k_checkpoints: fetched BEFORE the synthesis launch in stream order, so that the
            // host can look at it without waiting for the synthesis itself
            CU(cudaMemcpyAsync(ctx->h_chain_errors, ctx->d_chain_errors, sizeof(int), cudaMemcpyDeviceToHost, s));
            CU(cudaEventRecord(ctx->ev[6], s));
            CU(launch_synth(a, s));
            st.launches += 1;
        } else {
            // the very first chunks are short, so that the download (the long pole of this path) starts early
            for (int c0 = b0, nc = 0; c0 < b1; c0 += nc, ichunk++) {
                nc = kSynthChunk;
                if (ctx->graded_chunks) nc = c0 == 0 ? 32 : (c0 == 32 ? 96 : (c0 == 128 ? 128 : kSynthChunk));
                nc = std::min(nc, b1 - c0);
                SynthArgs ac{};
                char *dout = (char *) dst_dev + (size_t) c0 * blk_bytes;
                fill_args(ctx, ac, c0, nc, nchan, sample_size, dout);
                CU(launch_synth(ac, s));
                st.launches += 1;
                CU(cudaEventRecord(ctx->ev_done[ichunk], s));
                CU(cudaStreamWaitEvent(ctx->s_copy, ctx->ev_done[ichunk], 0));
                CU(cudaMemcpyAsync((char *) dst_host + (size_t) c0 * blk_bytes, dout, (size_t) nc * blk_bytes,
                                   cudaMemcpyDeviceToHost, ctx->s_copy));
                st.d2h_bytes += (int64_t) nc * (int64_t) blk_bytes;
            }
        }
    }
    CU(cudaEventRecord(ctx->ev[5], s));
    SynthArgs all{};
    fill_args(ctx, all, 0, nblk, nchan, sample_size, dst_dev);
    ctx->last = all;
    ctx->have_last = true;
    if (carr_phase_out)
        for (int c = 0; c < nchan; c++) carr_phase_out[c] = chain[c].prn > 0 ? chain[c].phase : 0.0;
    // the device self-check of the carrier chain is never skipped: a wrong start phase must not produce samples silently
    if (dst_host) {
        CU(cudaMemcpyAsync(ctx->h_chain_errors, ctx->d_chain_errors, sizeof(int), cudaMemcpyDeviceToHost, sp));
        CU(cudaStreamSynchronize(sp));
    } else {
        CU(cudaEventSynchronize(ctx->ev[6]));
    }
    if (*ctx->h_chain_errors != 0)
        return fail(ctx, GPSB200_ERR_INTERNAL, "carrier chain self-check failed on " +
                                                  std::to_string(*ctx->h_chain_errors) + " blocks");
    if (dst_host) {
        CU(cudaStreamSynchronize(s));
        CU(cudaStreamSynchronize(ctx->s_copy));
    }
    if (stats) {
        CU(cudaEventSynchronize(ctx->ev[5]));
        float ms = 0;
        // per-kernel times of the FIRST segment (the only one when dst_host == NULL) ...
        cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]);
        st.probe_kernel_ms = ms;
        cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]);
        st.checkpoint_kernel_ms = ms;
        if (!dst_host) {
            cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]);
            st.synth_kernel_ms = ms;
        }
        // ... and the whole span of the call's stream
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[5]);
        st.kernel_ms = ms;
        *stats = st;
    }
    return GPSB200_OK;
}

}  // namespace

extern "C" {

const char *gpsb200_version(void) { return "gpsb200 0.2 (sm_100a)"; }

int gpsb200_bind_numa(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int) sizeof bus, device) != cudaSuccess) {
        cudaGetLastError();
        return GPSB200_ERR_CUDA;
    }
    for (char *p = bus; *p; ++p) *p = (char) tolower((unsigned char) *p);
    int node = -1;
    if (FILE *f = fopen((std::string("/sys/bus/pci/devices/") + bus + "/numa_node").c_str(), "r")) {
        if (fscanf(f, "%d", &node) != 1) node = -1;
        fclose(f);
    }
    if (node < 0) return -1;
    char list[4096] = {0};
    FILE *f = fopen(("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist").c_str(), "r");
    if (!f) return -1;
    const bool got = fgets(list, sizeof list, f) != nullptr;
    fclose(f);
    if (!got) return -1;
    cpu_set_t cur, want;
    CPU_ZERO(&want);
    if (sched_getaffinity(0, sizeof cur, &cur) != 0) return -1;
    int picked = 0;
    for (const char *p = list; *p && *p != '\n';) {            // "0-31,64-95"
        char *end = nullptr;
        long a = strtol(p, &end, 10), b = a;
        if (end == p) break;
        if (*end == '-') b = strtol(end + 1, &end, 10);
        for (long c = a; c <= b && c < CPU_SETSIZE; c++)
            if (CPU_ISSET((int) c, &cur)) {
                CPU_SET((int) c, &want);
                ++picked;
            }
        p = *end == ',' ? end + 1 : end;
    }
    if (picked == 0 || sched_setaffinity(0, sizeof want, &want) != 0) return -1;
    return node;
}

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
    if (c.host_threads <= 0) c.host_threads = (int) std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (c.max_nav_frames <= 0) c.max_nav_frames = 1;
    if (c.max_chan < 1 || c.max_chan > GPSB200_MAX_CHAN || c.max_blocks < 1 || c.run_samples < 32 ||
        c.run_samples % 32 != 0 || GPSB200_BLOCK_SAMPLES % c.run_samples != 0) {
        delete ctx;
        return GPSB200_ERR_ARG;
    }
    ctx->nruns = GPSB200_BLOCK_SAMPLES / c.run_samples;
    // Carrier-chain units per block. 1 = whole blocks (default). Finer units (GPSB200_UNITS=5) give
    // the latency-bound walk kernels more, shorter threads, but measured on B200 the 5x larger
    // probe table written to mapped host memory and the 5x longer host scan cost more than they save
    // (k_probe 3.2 -> 6.7 ms, host 3.2 -> 8.9 ms at 32 channels), so this stays an experiment knob.
    if (const char *ev = getenv("GPSB200_UNITS")) {
        const int u = atoi(ev);
        if (u > 1 && GPSB200_BLOCK_SAMPLES % u == 0 && (GPSB200_BLOCK_SAMPLES / u) % c.run_samples == 0) {
            ctx->units = u;
            ctx->unit_samples = GPSB200_BLOCK_SAMPLES / u;
        }
    }
    if (const char *ev = getenv("GPSB200_GRADED_CHUNKS")) ctx->graded_chunks = atoi(ev) != 0;
    ctx->pool.reset(new WorkerPool(std::min(c.host_threads, c.max_chan)));
    *out = ctx;   // from here on errors are reported through the context
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(ctx, GPSB200_ERR_CUDA, "no CUDA device: gpsb200 has no CPU fallback");
    CU(cudaSetDevice(c.device));
    CU(cudaStreamCreateWithFlags(&ctx->s_compute, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ctx->s_copy, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ctx->s_pre, cudaStreamNonBlocking));
    for (auto &e : ctx->ev) CU(cudaEventCreate(&e));
    const int nchunk = (c.max_blocks + kSynthChunk - 1) / kSynthChunk + (c.max_blocks + kSegBlocks - 1) / kSegBlocks + 5;
    ctx->ev_done.resize(nchunk);
    for (auto &e : ctx->ev_done) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    const size_t nbc = (size_t) c.max_blocks * c.max_chan;
    CU(cudaMalloc(&ctx->d_bc, nbc * sizeof(BlockChanDev)));
    CU(cudaHostAlloc(&ctx->h_bc, nbc * sizeof(BlockChanDev), cudaHostAllocDefault));
    CU(cudaMalloc(&ctx->d_ck, nbc * ctx->nruns * sizeof(RunCkpt)));
    CU(cudaMalloc(&ctx->d_carr_end, nbc * sizeof(double)));
    CU(cudaMalloc(&ctx->d_atab, (size_t) c.max_blocks * kAtabRows * 32 * sizeof(int32_t)));
    CU(cudaMalloc(&ctx->d_chain_errors, sizeof(int)));
    CU(cudaHostAlloc(&ctx->h_chain_errors, sizeof(int), cudaHostAllocDefault));
    const size_t nbu = nbc * ctx->units;
    CU(cudaMalloc(&ctx->d_guess, nbu * sizeof(double)));
    CU(cudaHostAlloc(&ctx->h_guess, nbu * sizeof(double), cudaHostAllocDefault));
    CU(cudaMalloc(&ctx->d_carr0, nbu * sizeof(double)));
    CU(cudaHostAlloc(&ctx->h_carr0, nbu * sizeof(double), cudaHostAllocDefault));
    // probe results are written by the kernel straight into mapped pinned host memory: a
    // copy-engine download would queue behind the large result downloads of earlier segments
    CU(cudaHostAlloc(&ctx->h_probe, nbu * sizeof(CarrierProbe), cudaHostAllocMapped));
    CU(cudaHostGetDevicePointer((void **) &ctx->d_probe, ctx->h_probe, 0));
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
    if (ctx->s_compute) cudaSetDevice(ctx->cfg.device);
    if (ctx->s_compute) cudaStreamSynchronize(ctx->s_compute);
    if (ctx->s_copy) cudaStreamSynchronize(ctx->s_copy);
    if (ctx->s_pre) cudaStreamSynchronize(ctx->s_pre);
    cudaFree(ctx->d_bc);
    cudaFreeHost(ctx->h_bc);
    cudaFree(ctx->d_ck);
    cudaFree(ctx->d_carr_end);
    cudaFree(ctx->d_atab);
    cudaFree(ctx->d_chain_errors);
    cudaFreeHost(ctx->h_chain_errors);
    cudaFree(ctx->d_guess);
    cudaFreeHost(ctx->h_guess);
    cudaFree(ctx->d_carr0);
    cudaFreeHost(ctx->h_carr0);
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
    if (ctx->s_pre) cudaStreamDestroy(ctx->s_pre);
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

int gpsb200_carrier_chain_device(gpsb200_ctx_t *ctx, const gpsb200_chan_t *chans, int nblk, int nchan,
                                 const double *phase_in, double *phase_out) {
    if (!ctx || !chans || !phase_out || nblk < 0 || nchan != ctx->cfg.max_chan) return GPSB200_ERR_ARG;
    if (!ctx->s_compute) return fail(ctx, GPSB200_ERR_CUDA, "context has no CUDA device");
    CU(cudaSetDevice(ctx->cfg.device));
    cudaStream_t s = ctx->s_compute;
    std::vector<ChainState> chain(nchan);
    if (phase_in && nblk > 0)
        for (int c = 0; c < nchan; c++)
            if (chans[c].prn > 0) {
                chain[c].prn = chans[c].prn;
                chain[c].phase = phase_in[c];
            }
    for (int w0 = 0; w0 < nblk; w0 += ctx->cfg.max_blocks) {
        const int nw = std::min(ctx->cfg.max_blocks, nblk - w0);
        const gpsb200_chan_t *cw = chans + (size_t) w0 * nchan;
        int rc = prepare_blocks(ctx, cw, 0, nw, nchan, chain);
        if (rc) return rc;
        const size_t cnt = (size_t) nw * nchan, cntu = cnt * ctx->units;
        CU(cudaMemcpyAsync(ctx->d_bc, ctx->h_bc, cnt * sizeof(BlockChanDev), cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(ctx->d_guess, ctx->h_guess, cntu * sizeof(double), cudaMemcpyHostToDevice, s));
        SynthArgs a{};
        fill_args(ctx, a, 0, nw, nchan, GPSB200_SC08, nullptr);
        CU(launch_probe(a, s));
        CU(cudaStreamSynchronize(s));
        resolve_chain(ctx, cw, 0, nw, nchan, chain);
    }
    ctx->have_last = false;
    for (int c = 0; c < nchan; c++) phase_out[c] = chain[c].prn > 0 ? chain[c].phase : 0.0;
    return GPSB200_OK;
}

int gpsb200_replay_device(gpsb200_ctx_t *ctx, void *dst_device, void *stream_, int kernel_mask) {
    if (!ctx || !ctx->have_last) return GPSB200_ERR_ARG;
    CU(cudaSetDevice(ctx->cfg.device));
    cudaStream_t s = stream_ ? (cudaStream_t) stream_ : ctx->s_compute;
    SynthArgs a = ctx->last;
    if (dst_device) a.out = dst_device;
    if (kernel_mask & 8) CU(launch_tables(a, s));
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
