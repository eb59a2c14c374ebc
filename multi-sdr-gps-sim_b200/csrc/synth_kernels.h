// Device-side data layout and launchers of the GPS L1 C/A synthesis kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpsb200 {

constexpr int kBlockSamples = 300000;   // sdr.h:26
constexpr int kNavWords = 60;           // gps.h:52
constexpr int kChipWords = 33;          // 1023 chips, periodically extended to 33 x 32 bits
constexpr int kAtabRows = 513;          // carrier table rows; row 512 is never addressed (carr_phase < 1.0 always), a guard

// One record per (block, channel slot), written by the host, read by both kernels.
struct BlockChanDev {
    double c_carr;     // fl(f_carr * delt)  (gps.c:2821)
    double c_code;     // fl(f_code * delt)  (gps.c:2789)
    double gain;       // gps.c:2756
    double reserved0;  // (start phases travel in SynthArgs::carr0)
    double code0;      // code phase at the first sample (computeCodePhase, gps.c:2049)
    int32_t prn;       // 0 = slot unused
    uint32_t nav0;     // iword | ibit << 8 | icode << 16 at the first sample
    int32_t frame;     // NAV frame index
    int32_t pad[3];
};
static_assert(sizeof(BlockChanDev) == 64, "BlockChanDev layout");

// Exact NCO state at the first sample of a run (run = run_samples consecutive samples).
struct RunCkpt {
    double x;          // carrier phase
    double y;          // code phase
    uint32_t nav;      // iword | ibit << 8 | icode << 16
    uint32_t pad;
};
static_assert(sizeof(RunCkpt) == 24, "RunCkpt layout");

struct CarrierProbe;          // nco_exact.h

struct SynthArgs {
    const BlockChanDev *bc;   // [nblk][nchan]
    // The carrier chain is resolved in UNITS: a block is cut into `units` pieces of
    // `unit_samples` samples (a whole number of runs) so that the latency-bound walks of
    // k_probe / k_checkpoints get `units` times more threads that are `units` times shorter.
    const double *carr0;      // [nblk][units][nchan] exact carrier phase at the first sample of each unit
    const double *guess;      // [nblk][units][nchan] guessed start phases for the speculative probe
    CarrierProbe *probe;      // [nblk][units][nchan] probe results
    RunCkpt *ck;              // [nblk][nruns][nchan]
    const uint32_t *nav;      // [frames][nchan][60]
    const uint32_t *chipbits; // [33][33] packed C/A chips per PRN (bit n = ca[n mod 1023]), row 0 unused
    int32_t *atab;            // [nblk][513][32] gain-scaled carrier table per block: I + (Q << 16), column = lane
    double *carr_end;         // [nblk][nchan] carrier phase after the block
    int *chain_errors;        // self-check counter: blocks whose walked end phase != the next block's start phase
    void *out;                // nblk * 600000 int8 or int16
    int nblk, nchan, nruns, run_samples, runs_per_cta, ctas_per_block, iq16;
    int units, unit_samples;
};

// Gain-scaled carrier tables of every block (gps.c:2781-2782), fetched by k_synth with TMA bulk copies.
cudaError_t launch_tables(const SynthArgs &a, cudaStream_t s);
// Speculative carrier walk of every (block, channel) from a guessed start phase (nco_exact.h).
cudaError_t launch_probe(const SynthArgs &a, cudaStream_t s);
// Run-start checkpoints for every (block, channel): exact walk, O(#binade crossings).
cudaError_t launch_checkpoints(const SynthArgs &a, cudaStream_t s);
// The per-sample synthesis (gps.c:2767-2857): lanes = channels, warp-sum over channels.
cudaError_t launch_synth(const SynthArgs &a, cudaStream_t s);
// Threads per CTA and dynamic shared memory the synthesis launch will use (for reporting).
void synth_launch_shape(const SynthArgs &a, int *ctas, int *threads, size_t *smem);

}  // namespace gpsb200
